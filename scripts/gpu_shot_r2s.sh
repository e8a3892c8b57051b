#!/bin/bash
# A/B: long-lived values of k_batch_add's passes in staging slots (default) against the register version (-DMGB_BWD_REGS)
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "prefetch or closed_form or batch or parity" > gpurun_out/r2s_pytest.txt 2>&1; tail -3 gpurun_out/r2s_pytest.txt
qt() { timeout 120 python scripts/quick_time.py "$@" 2>&1 | tail -1; }
for rep in 1; do
for lib in "" montgomery_b200/libmontgomery_b200_bwdregs.so; do
  echo "== lib=${lib:-default} rep $rep"
  for cfg in "20 bls12-377" "16 bls12-377" "18 pallas" "20 pallas" "20 bls12-381" "22 bls12-377"; do
    MGB_LIB=$lib qt $cfg
  done
done
done 2>&1 | tee gpurun_out/r2s_ab.txt

