#!/bin/bash
set -u
mkdir -p gpurun_out
{ for cfg in "20 bls12-377" "20 bls12-377" "16 bls12-377" "18 pallas" "18 ed-on-bls12-377" "22 bls12-377" "14 bls12-377"; do timeout 60 python scripts/quick_time.py $cfg; done; } > gpurun_out/r2g_times.txt 2>&1
cat gpurun_out/r2g_times.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.txt 2>&1; tail -3 gpurun_out/r2g_pytest.txt
