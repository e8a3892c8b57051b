#!/bin/bash
# Round-2 evidence run on one B200 (via gpurun): GPU tests, bench lines of both arms, other configs, microbenchmarks,
# ncu launch list of the bench command, ncu --set full captures of the shipped round-0 (with source) and round-3
# launches of k_batch_add, compute-sanitizer.  Outputs land in gpurun_out/ (r02_*); scripts/summarize_ncu.py and
# scripts/ncu_regions.py turn the raw ncu files into the summaries under profiles/.
set -u
tag=r02
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.txt 2>&1; tail -3 gpurun_out/${tag}_pytest.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err
python scripts/microbench.py > gpurun_out/${tag}_microbench.jsonl 2>&1
{ for cfg in "14 bls12-377" "16 bls12-377" "18 bls12-377" "20 bls12-377" "22 bls12-377" "24 bls12-377" "16 pallas" "18 pallas" "20 pallas" "16 ed-on-bls12-377" "18 ed-on-bls12-377" "20 ed-on-bls12-377" "20 bls12-381"; do timeout 120 python scripts/quick_time.py $cfg; done
  echo "per-round times (synchronising, MGB_DEBUG_ROUNDS=1):"; MGB_DEBUG_ROUNDS=1 timeout 60 python scripts/quick_time.py 20 2>&1 | tail -6; } > gpurun_out/${tag}_other_configs.txt 2>&1
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_batch_add --launch-skip 5 --launch-count 1 -f -o gpurun_out/${tag}_prof_round0 \
    python scripts/profile_msm.py 20 2 > gpurun_out/${tag}_prof0.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:k_batch_add --launch-skip 8 --launch-count 1 -f -o gpurun_out/${tag}_prof_round3 \
    python scripts/profile_msm.py 20 2 > gpurun_out/${tag}_prof3.log 2>&1
timeout 600 bash scripts/sanitize.sh; cp gpurun_out/sanitize.txt gpurun_out/${tag}_sanitizer.txt
tail -c 400 gpurun_out/${tag}_bench.json; echo; cut -c1-300 gpurun_out/${tag}_bench_reference.json; cat gpurun_out/${tag}_other_configs.txt; grep -c "exit=0" gpurun_out/${tag}_sanitizer.txt; grep -i "error\|hazard" gpurun_out/${tag}_sanitizer.txt | head
