#!/bin/bash
# round 2, GPU call 6 (2 GPUs): all GPU tests incl. the two multi-GPU ones; timings of the current build; TE Horner
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2f_gpus.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.txt 2>&1; tail -8 gpurun_out/r2f_pytest.txt
{ for cfg in "20 bls12-377" "20 bls12-377" "16 bls12-377" "18 pallas" "18 ed-on-bls12-377" "20 ed-on-bls12-377" "20 bls12-381"; do timeout 60 python scripts/quick_time.py $cfg; done; } > gpurun_out/r2f_times.txt 2>&1
cat gpurun_out/r2f_times.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2f_bench2.json 2> gpurun_out/r2f_bench2.err
tail -3 gpurun_out/r2f_bench2.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r2f_bench2.json"))
print({k: d[k] for k in ("value", "ms_per_step", "msm_ms_device", "parity_ok", "n_gpus")}, d["e2e"]["ms_per_step"], d.get("strong_2p24"))
PY
