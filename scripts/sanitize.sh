#!/bin/bash
# compute-sanitizer over small MSMs of every curve (memcheck) and the shared-memory kernels (racecheck).
# Run on a GPU box: gpurun -- scripts/sanitize.sh ; output in gpurun_out/sanitize.txt
out=gpurun_out/sanitize.txt
mkdir -p gpurun_out; : > $out
for cv in bls12-377 pallas ed-on-bls12-377 bls12-381; do
  for logn in 4 13; do
    echo "== memcheck $cv 2^$logn" >> $out
    compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/profile_msm.py $logn 1 $cv 2>&1 | grep -v "^{" | tail -4 >> $out
    echo "exit=$?" >> $out
  done
done
echo "== racecheck bls12-377 2^12" >> $out
compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/profile_msm.py 12 1 bls12-377 2>&1 | grep -v "^{" | tail -6 >> $out
echo "== racecheck ed-on-bls12-377 2^12" >> $out
compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/profile_msm.py 12 1 ed-on-bls12-377 2>&1 | grep -v "^{" | tail -6 >> $out
echo "== synccheck bls12-377 2^12" >> $out
compute-sanitizer --tool synccheck --error-exitcode 9 python scripts/profile_msm.py 12 1 bls12-377 2>&1 | grep -v "^{" | tail -4 >> $out
