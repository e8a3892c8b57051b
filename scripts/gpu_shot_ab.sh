#!/bin/bash
# One short GPU call: A/B of experiment builds (montgomery_b200/libmontgomery_b200_<name>.so) against the product
# library: parity (closed forms, msmProjective) + best-of-6 phase times; then GPU tests + bench line with the winner.
#   bash scripts/gpu_shot_ab.sh name1 name2 ...
set -u
mkdir -p gpurun_out
timeout 100 python scripts/ab_variant.py montgomery_b200/libmontgomery_b200.so 20 16 > gpurun_out/ab_base.json 2> gpurun_out/ab_base.err
for v in "$@"; do
  timeout 60 python scripts/ab_variant.py montgomery_b200/libmontgomery_b200_$v.so 20 16 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
done
WIN=$(python - "$@" <<'PY'
import json, sys
best, bt = "base", json.load(open("gpurun_out/ab_base.json"))["2^20"]["total"]
for v in sys.argv[1:]:
    try:
        w = json.load(open("gpurun_out/ab_%s.json" % v))
        ok = all(val for k, val in w.items() if "closed_form" in k or k.endswith("_ok") or k.startswith("projective"))
        if ok and w["2^20"]["total"] < bt:
            best, bt = v, w["2^20"]["total"]
    except Exception:
        pass
print(best)
PY
)
echo "winner: $WIN" | tee gpurun_out/winner.txt
if [ "$WIN" != base ]; then export MGB_LIB=$PWD/montgomery_b200/libmontgomery_b200_$WIN.so; fi
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$WIN.txt 2>&1
if [ "${QUICK:-0}" = 1 ]; then tail -3 gpurun_out/pytest_gpu_$WIN.txt; cat gpurun_out/ab_*.json; exit 0; fi
timeout 200 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_$WIN.json 2> gpurun_out/bench_$WIN.err
# launch list of the bench command under ncu (shares only: the times are cold-cache and serialised)
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$WIN.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_$WIN.log 2>&1
for f in gpurun_out/ab_*.json; do python - $f <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
bad = [k for k, v in d.items() if ("closed_form" in k or k.endswith("_ok") or k.startswith("projective")) and not v]
print(d["lib"], "2^20", d["2^20"], "2^16 total", d.get("2^16", {}).get("total"), "FAILED: %s" % bad if bad else "parity ok")
PY
done
tail -3 gpurun_out/pytest_gpu_$WIN.txt
