#!/bin/bash
# round 2, GPU call 14: tests; ncu captures of the shipped kernels (round 0 with source, round 3), launch list
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.txt 2>&1; tail -5 gpurun_out/r2n_pytest.txt
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2n_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2n_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_batch_add --launch-skip 5 --launch-count 1 -f -o gpurun_out/r2n_prof_round0 \
    python scripts/profile_msm.py 20 2 > gpurun_out/r2n_prof0.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:k_batch_add --launch-skip 8 --launch-count 1 -f -o gpurun_out/r2n_prof_round3 \
    python scripts/profile_msm.py 20 2 > gpurun_out/r2n_prof3.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:k_group_partial --launch-skip 1 --launch-count 1 -f -o gpurun_out/r2n_prof_group_partial \
    python scripts/profile_msm.py 20 2 > gpurun_out/r2n_prof_gp.log 2>&1
ls -la gpurun_out/r2n*.ncu-rep
