#!/bin/bash
# GPU tests + the driver's bench command (pipelined end-to-end block included)
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2t_pytest.txt 2>&1; tail -3 gpurun_out/r2t_pytest.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; tail -3 gpurun_out/r2t_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2t_bench.json").read().strip().splitlines()[-1])
p = d["e2e"]["pipelined"]
print("value %.2f M wall %.3f dev %.3f | e2e %.3f | pipelined %.3f | frac %.4f parity %s" % (d["value"] / 1e6, d["ms_per_step"], d["msm_ms_device"], d["e2e"]["ms_per_step"], p["ms_per_step"], d["roofline"]["frac"], d["parity_ok"]))
for k, v in d["configs"].items():
    print(k, "dev %.3f wall %.3f e2e %.3f pipe %.3f frac %.3f" % (v["ms_device"], v["ms_per_step"], v["e2e_ms_per_step"], v["e2e_pipelined_ms_per_step"], v["roofline_frac"]))
s = d["strong_2p24"]
print("2^24: wall %.2f e2e %.2f pipe %.2f frac %.3f" % (s["ms_per_step"], s["e2e_ms_per_step"], s["e2e_pipelined_ms_per_step"], s["roofline_frac"]))
PY
