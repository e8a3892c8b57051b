#!/bin/bash
# ncu --set full captures of the latency-bound launches of the 2^16 BLS12-377 MSM (4 tree rounds per MSM): round 0 and round 3 of the
# second MSM, and the reduction kernels.  The reports are reduced to CSV on the box (raw + source pages); outputs: gpurun_out/s3_*.csv
set -u
mkdir -p gpurun_out
cap() { # name regex skip
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$2 --launch-skip $3 --launch-count 1 -f -o /tmp/s3_$1 python scripts/profile_msm.py 16 2 bls12-377 > gpurun_out/s3_$1.log 2>&1
  ncu -i /tmp/s3_$1.ncu-rep --page raw --csv > gpurun_out/s3_$1_raw.csv 2>/dev/null
  ncu -i /tmp/s3_$1.ncu-rep --page source --csv > gpurun_out/s3_$1_source.csv 2>/dev/null
  rm -f /tmp/s3_$1.ncu-rep
}
cap r0 k_batch_add 4
cap r3 k_batch_add 7
cap finish k_bucket_finish 1
cap tailquad k_tree_tail_quad 1
cap dsums k_digit_sums 1
cap wasm k_window_assemble 1
ls -la gpurun_out/
