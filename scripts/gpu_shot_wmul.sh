#!/bin/bash
# One short GPU call: A/B of the warp-cooperative Horner build against the product library (parity + phase times),
# a round-0 ncu --set full capture of k_batch_add, then the bench line and the GPU tests with the winner.
set -u
mkdir -p gpurun_out
BASE=montgomery_b200/libmontgomery_b200.so
EXP=montgomery_b200/libmontgomery_b200_wmul.so
timeout 150 python scripts/ab_variant.py $BASE 20 > gpurun_out/ab_base.json 2> gpurun_out/ab_base.err
timeout 100 python scripts/ab_variant.py $EXP 20 16 > gpurun_out/ab_wmul.json 2> gpurun_out/ab_wmul.err
WIN=$(python - <<'PY'
import json
try:
    b = json.load(open("gpurun_out/ab_base.json")); w = json.load(open("gpurun_out/ab_wmul.json"))
    ok = all(v for k, v in w.items() if k.startswith("closed_form") or k.endswith("closed_form") or k.endswith("_ok") or k.startswith("projective"))
    print("wmul" if ok and w["2^20"]["total"] < b["2^20"]["total"] else "base")
except Exception as e:
    print("base")
PY
)
echo "winner: $WIN" | tee gpurun_out/winner.txt
if [ "$WIN" = wmul ]; then export MGB_LIB=$PWD/$EXP; fi
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_batch_add --launch-skip 5 --launch-count 1 -f \
    -o gpurun_out/prof_batch_add_round0 python scripts/profile_msm.py 20 2 > gpurun_out/prof_round0.log 2>&1
timeout 60 python scripts/summarize_ncu.py full gpurun_out/prof_batch_add_round0.ncu-rep > gpurun_out/ncu_full_round0.csv 2>> gpurun_out/prof_round0.log
timeout 200 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_$WIN.json 2> gpurun_out/bench_$WIN.err
timeout 60 python scripts/microbench.py > gpurun_out/microbench_$WIN.jsonl 2>&1
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$WIN.txt 2>&1
tail -3 gpurun_out/pytest_gpu_$WIN.txt
cat gpurun_out/ab_base.json gpurun_out/ab_wmul.json
tail -c 1500 gpurun_out/bench_$WIN.json
