#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== default"; timeout 60 python scripts/quick_time.py 20; timeout 60 python scripts/quick_time.py 20
for nb in 1.95 1.9 1.8 1.6; do echo "== NBIG=$nb"; MGB_DEBUG_NBIG=$nb timeout 60 python scripts/quick_time.py 20; done
echo "== per-round"; MGB_DEBUG_ROUNDS=1 timeout 60 python scripts/quick_time.py 20 2>&1 | tail -6
} > gpurun_out/r2p_times.txt 2>&1
cat gpurun_out/r2p_times.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; tail -2 gpurun_out/r2p_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2p_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "msm_ms_device", "parity_ok")}, "e2e", d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"])
print(d["phases_ms"]); print(d["roofline"]["dominant_kernel"])
PY
