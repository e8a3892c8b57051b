#!/bin/bash
# round 2, 8-GPU call: the driver's bench command at 8 and 4 GPUs (weak 2^20 per GPU + the strong 2^24 block), multi-GPU tests
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2q_topo.txt 2>&1
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2q_bench$n.json 2> gpurun_out/r2q_bench$n.err
  tail -2 gpurun_out/r2q_bench$n.err
  python - $n <<'PY'
import json, sys
d = json.load(open("gpurun_out/r2q_bench%s.json" % sys.argv[1]))
print({k: d[k] for k in ("n_gpus", "value", "ms_per_step", "msm_ms_device", "parity_ok")}, "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"], d.get("host_affinity"))
print(d.get("strong_2p24"))
PY
done
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "multi" > gpurun_out/r2q_pytest_multi.txt 2>&1; tail -3 gpurun_out/r2q_pytest_multi.txt
