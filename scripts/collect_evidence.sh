#!/bin/bash
# Evidence run on one B200 (via gpurun): bench lines of both arms, microbenchmarks, ncu launch list of the
# bench command, one ncu --set full capture of round 0 of k_batch_add
# (5 k_batch_add launches per MSM at 2^20: skipping 5 lands on round 0 of the second MSM).  Outputs land in gpurun_out/.
set -u
tag=${1:-r01}
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_${tag}_reference.json 2>> gpurun_out/bench_$tag.err
python scripts/microbench.py > gpurun_out/microbench_$tag.jsonl 2>&1
{ for cfg in "16 bls12-377" "18 bls12-377" "20 bls12-377" "22 bls12-377" "18 pallas" "20 pallas" "18 ed-on-bls12-377" "20 ed-on-bls12-377" "20 bls12-381"; do python scripts/quick_time.py $cfg; done; } > gpurun_out/other_configs_$tag.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_batch_add --launch-skip 5 --launch-count 1 -f -o gpurun_out/prof_batch_add_$tag \
    python scripts/profile_msm.py 20 2 > gpurun_out/prof_$tag.log 2>&1
tail -c 600 gpurun_out/bench_$tag.json; echo; cat gpurun_out/bench_${tag}_reference.json | cut -c1-400
