#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest.txt 2>&1; tail -3 gpurun_out/r2r_pytest.txt
{ for cfg in "20 bls12-377" "20 bls12-377" "16 bls12-377" "18 pallas" "18 ed-on-bls12-377" "20 ed-on-bls12-377"; do timeout 60 python scripts/quick_time.py $cfg; done; } > gpurun_out/r2r_times.txt 2>&1
cat gpurun_out/r2r_times.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2r_bench2.json 2> gpurun_out/r2r_bench2.err
tail -2 gpurun_out/r2r_bench2.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2r_bench2.json"))
print({k: d[k] for k in ("n_gpus", "value", "ms_per_step", "msm_ms_device", "parity_ok")}, "e2e", d["e2e"]["ms_per_step"], d.get("strong_2p24"))
PY
