#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest.txt 2>&1; tail -4 gpurun_out/r2o_pytest.txt
{ for cfg in "20 bls12-377" "20 bls12-377" "20 bls12-377" "18 pallas" "18 ed-on-bls12-377" "16 bls12-377" "20 bls12-381"; do timeout 60 python scripts/quick_time.py $cfg; done; } > gpurun_out/r2o_times.txt 2>&1
cat gpurun_out/r2o_times.txt
python scripts/microbench.py 2>&1 | tail -12
