#!/bin/bash
# 4-GPU bench (weak 2^20 per GPU), pipelined end-to-end path: upload started at once against upload started behind round 0
set -u
mkdir -p gpurun_out
n=4
for mode in now defer; do
  if [ $mode = defer ]; then export MGB_PF_DEFER=1; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus $n --steps 10 --warmup 3 --no-extras > gpurun_out/r2u_bench${n}_$mode.json 2> gpurun_out/r2u_bench${n}_$mode.err
  tail -2 gpurun_out/r2u_bench${n}_$mode.err
  python - $n $mode <<'PY'
import json, sys
d = json.loads(open("gpurun_out/r2u_bench%s_%s.json" % (sys.argv[1], sys.argv[2])).read().strip().splitlines()[-1])
p = d["e2e"]["pipelined"]
print(sys.argv[2], "N=%s value %.1f M wall %.3f dev %.3f | e2e %.3f | pipelined %.3f parity %s" % (sys.argv[1], d["value"] / 1e6, d["ms_per_step"], d["msm_ms_device"], d["e2e"]["ms_per_step"], p["ms_per_step"], d["parity_ok"]))
r = lambda x: {k: round(v, 3) for k, v in x.items()}
print(" dev  ", r(d["phases_ms"])); print(" e2e  ", r(d["e2e"]["phases_ms"])); print(" pipe ", r(p["phases_ms"]))
PY
done
