#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.txt 2>&1; tail -12 gpurun_out/r2m_pytest.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; tail -3 gpurun_out/r2m_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2m_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "msm_ms_device", "parity_ok")}, "e2e", d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"])
print(d["phases_ms"])
for k, v in d.get("configs", {}).items(): print(k, v["ms_device"], v["roofline_frac"], v["parity_ok"], v.get("cpu_port_ms"))
print(d.get("strong_2p24"))
PY
