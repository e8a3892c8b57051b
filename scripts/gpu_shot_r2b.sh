#!/bin/bash
# round 2, second GPU call: tests on the new defaults (lane-parallel inversion, one-warp Horner, balanced tiles),
# finer tile sweep, depth / window sweeps, EMAX=128 build, ncu captures of the shipped round-0 and round-3 kernels
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.txt 2>&1; tail -2 gpurun_out/r2b_pytest.txt
qt() { "$@" timeout 60 python scripts/quick_time.py ${LOGN:-20} ${LABEL:-bls12-377} ${C:-} 2>&1 | grep -E "round|accumulate" | tail -${NL:-6} | sed -E "s/.*('accumulate': [0-9.]+).*('reduce': [0-9.]+).*('final_sum': [0-9.]+).*('total': [0-9.]+).*('rounds': [0-9]+).*/\1 \2 \3 \4 \5/"; }
{
echo "== default"; NL=1 qt env
echo "== default with per-round times"; qt env MGB_DEBUG_ROUNDS=1
for cfg in "48,28,14,14,7" "52,28,14,14,7" "54,28,14,14,7" "58,28,14,14,7" "60,28,14,14,7" "56,56,14,14,7" "56,19,14,14,7" "56,28,28,14,7" "56,28,10,14,7" "56,28,14,7,7" "56,28,14,14,4" "56,28,14,14,14"; do
  echo "== E=$cfg"; qt env MGB_DEBUG_E=$cfg MGB_DEBUG_ROUNDS=1
done
for nr in 4 6 7; do echo "== NROUNDS=$nr"; NL=1 qt env MGB_DEBUG_NROUNDS=$nr; done
for c in 15 17 18; do echo "== c=$c"; NL=1 C=$c qt env; done
echo "== EMAX=128 build, E=111,56,28,14,7"; qt env MGB_LIB=$PWD/montgomery_b200/libmontgomery_b200_e128.so MGB_DEBUG_E=111,56,28,14,7 MGB_DEBUG_ROUNDS=1
echo "== EMAX=128 build, E=111,111,56,28,14"; qt env MGB_LIB=$PWD/montgomery_b200/libmontgomery_b200_e128.so MGB_DEBUG_E=111,111,56,28,14 MGB_DEBUG_ROUNDS=1
echo "== EMAX=128 build, E=74,56,28,14,7 (3 tiles per 2 warps)"; qt env MGB_LIB=$PWD/montgomery_b200/libmontgomery_b200_e128.so MGB_DEBUG_E=74,56,28,14,7 MGB_DEBUG_ROUNDS=1
} > gpurun_out/r2b_sweep.txt 2>&1
{ for cfg in "16 bls12-377" "18 bls12-377" "20 bls12-377" "22 bls12-377" "18 pallas" "20 pallas" "18 ed-on-bls12-377" "20 ed-on-bls12-377" "20 bls12-381"; do timeout 120 python scripts/quick_time.py $cfg; done; } > gpurun_out/r2b_other_configs.txt 2>&1
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2b_ncu_bench.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_batch_add --launch-skip 5 --launch-count 1 -f -o gpurun_out/r2b_prof_round0 \
    python scripts/profile_msm.py 20 2 > gpurun_out/r2b_prof0.log 2>&1
timeout 200 ncu --set full --clock-control none -k regex:k_batch_add --launch-skip 8 --launch-count 1 -f -o gpurun_out/r2b_prof_round3 \
    python scripts/profile_msm.py 20 2 > gpurun_out/r2b_prof3.log 2>&1
cat gpurun_out/r2b_sweep.txt; cat gpurun_out/r2b_other_configs.txt; ls -la gpurun_out/*.ncu-rep
