#!/bin/bash
# Last short GPU call of the round: compute-sanitizer over the kernels that changed (shared-memory staging of
# k_batch_add, the fused Horner doublings) and a fresh ncu --set full capture of round 0 of k_batch_add.
set -u
out=gpurun_out/sanitize.txt
mkdir -p gpurun_out; : > $out
run() {  # tool logn curve
  echo "== $1 $3 2^$2" >> $out
  timeout 45 compute-sanitizer --tool $1 --error-exitcode 9 python scripts/profile_msm.py $2 1 $3 2>&1 | grep -v "^{" | tail -6 >> $out
  echo "exit=${PIPESTATUS[0]}" >> $out
}
run racecheck 12 bls12-377
run memcheck 13 bls12-377
timeout 60 ncu --set full --clock-control none --import-source on -k regex:k_batch_add --launch-skip 5 --launch-count 1 -f \
    -o gpurun_out/prof_batch_add_round0 python scripts/profile_msm.py 20 2 > gpurun_out/prof_round0.log 2>&1
timeout 30 python scripts/summarize_ncu.py full gpurun_out/prof_batch_add_round0.ncu-rep > gpurun_out/ncu_full_round0.csv 2>> gpurun_out/prof_round0.log
run racecheck 12 ed-on-bls12-377
run memcheck 13 ed-on-bls12-377
cat $out
