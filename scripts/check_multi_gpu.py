"""torchrun --nproc-per-node G scripts/check_multi_gpu.py [logn]: the sharded MSM (one process per GPU, NCCL
communicator owned by the context, mgb_msm_sharded) == the closed form over all ranks' shards, on three curves; then a
call in which only rank 0 has pairs (the other shards are empty).  Launched by tests/test_gpu_round2.py when the box has
>= 2 GPUs; prints one "multi-gpu ok" line per curve."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import montgomery_b200 as m
from montgomery_b200 import inputs
from montgomery_b200.distributed import ShardedMsm
from tests.helpers import OracleCurve

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for label in ("bls12-377", "pallas", "ed-on-bls12-377"):
    cv = m.curves.BY_LABEL[label]
    O = OracleCurve(label)
    n = 1 << logn
    sm = ShardedMsm(cv, local, n)
    info = sm.engine.comm_info()
    assert (info["rank"], info["world"]) == (rank, world) and info["nccl_version"] > 0
    sm.random_points(n, seed=500)
    sc = inputs.random_scalars(cv.q, n, 900 + rank)
    res, tm = sm.msm(torch.from_numpy(sc).pin_memory(), n)
    # closed form over all ranks: sum_r sum_i s_{r,i} * a_{r,i}
    k = 0
    for r in range(world):
        k += inputs.dot_known_dlogs(inputs.random_scalars(cv.q, n, 900 + r), inputs.known_dlogs(500 + r, n))
    exp = O.result_of(O.scale(k % O.q, O.G))
    assert res == exp, (label, rank)
    # scalars already on the device
    d = torch.from_numpy(sc).cuda()
    torch.cuda.synchronize()
    assert sm.msm(d, n, on_device=True)[0] == exp, (label, rank, "device scalars")
    # fewer pairs than ranks: only rank 0 has work, every rank still gets the result
    res0, _ = sm.msm(torch.from_numpy(sc[:5].copy()), 5 if rank == 0 else 0)
    s0 = inputs.random_scalars(cv.q, n, 900)[:5]
    exp0 = O.result_of(O.scale(inputs.dot_known_dlogs(s0, inputs.known_dlogs(500, 5)) % O.q, O.G))
    assert res0 == exp0, (label, rank, "empty shards")
    if rank == 0:
        print("multi-gpu ok:", label, "world", world, "n/rank 2^%d" % logn, "ms", round(tm["total"], 3), "nccl", info["nccl_version"])
    sm.close()
if rank == 0:
    print("empty shards ok")
dist.destroy_process_group()
