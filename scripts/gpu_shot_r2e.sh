#!/bin/bash
# round 2, GPU call 5: 8-bit reduction digits (two-level digit sums) vs 5-bit; GPU tests; the new bench line
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.txt 2>&1; tail -5 gpurun_out/r2e_pytest.txt
{
for vb in 8 5 7 6; do
  echo "== MGB_DEBUG_VB=$vb"
  for cfg in "20 bls12-377" "16 bls12-377" "18 pallas" "22 bls12-377"; do MGB_DEBUG_VB=$vb timeout 60 python scripts/quick_time.py $cfg; done
done
} > gpurun_out/r2e_vb.txt 2>&1
cat gpurun_out/r2e_vb.txt
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -3 gpurun_out/r2e_bench.err; cat gpurun_out/r2e_bench.json | cut -c1-3000
timeout 300 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r2e_bench_ref.json 2>> gpurun_out/r2e_bench.err; cat gpurun_out/r2e_bench_ref.json | cut -c1-600
