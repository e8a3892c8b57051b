#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.txt 2>&1; tail -3 gpurun_out/r2h_pytest.txt
{ for cfg in "20 bls12-377" "20 bls12-377" "16 bls12-377" "18 pallas" "18 ed-on-bls12-377" "22 bls12-377" "14 bls12-377"; do timeout 60 python scripts/quick_time.py $cfg; done
echo "== MGB_DEBUG_TREE_ROUNDS=1 (separate tree-round launches)"; MGB_DEBUG_TREE_ROUNDS=1 timeout 60 python scripts/quick_time.py 20; } > gpurun_out/r2h_times.txt 2>&1
cat gpurun_out/r2h_times.txt
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2h_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2h_ncu_bench.log 2>&1
python scripts/summarize_ncu.py launches gpurun_out/r2h_launches.csv
