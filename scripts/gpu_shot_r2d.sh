#!/bin/bash
# round 2, GPU call 4: resident blocks per SM (4 / 5 / 6) of k_batch_add with the single x buffer; new GPU tests
set -u
mkdir -p gpurun_out
{
for v in "" _m5 _m6; do
  export MGB_LIB=$PWD/montgomery_b200/libmontgomery_b200$v.so
  echo "== lib$v"
  for i in 1 2; do timeout 60 python scripts/quick_time.py 20; done
  MGB_DEBUG_ROUNDS=1 timeout 60 python scripts/quick_time.py 20 2>&1 | grep round | tail -5
  timeout 60 python scripts/quick_time.py 16; timeout 60 python scripts/quick_time.py 18 pallas; timeout 60 python scripts/quick_time.py 22
done
} > gpurun_out/r2d_minb.txt 2>&1
cat gpurun_out/r2d_minb.txt
unset MGB_LIB
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.txt 2>&1; tail -15 gpurun_out/r2d_pytest.txt
