#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -2 gpurun_out/r02_bench.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_bench.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "msm_ms_device", "parity_ok")}, "e2e", d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"])
print(d["phases_ms"])
for k, v in d["configs"].items(): print(k, round(v["ms_device"], 3), round(v["roofline_frac"], 3), v["parity_ok"], v.get("cpu_port_ms"), round(v["e2e_ms_per_step"], 3))
print(d["strong_2p24"]["ms_device_max"], d["strong_2p24"]["roofline_frac"], d["strong_2p24"]["parity_ok"])
PY
