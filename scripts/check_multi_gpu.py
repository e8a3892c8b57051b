"""torchrun --nproc-per-node G scripts/check_multi_gpu.py [logn]: sharded MSM == closed form, and
== the single-GPU result of rank 0 over the concatenated shards when they fit."""
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
    sm.random_points(n, seed=500)
    sc = inputs.random_scalars(cv.q, n, 900 + rank)
    res, tm = sm.msm(torch.from_numpy(sc).pin_memory(), n)
    # closed form over all ranks: sum_r sum_i s_{r,i} * a_{r,i}
    k = 0
    for r in range(world):
        a = inputs.known_dlogs(500 + r, n)
        s = inputs.scalars_to_ints(inputs.random_scalars(cv.q, n, 900 + r))
        k += sum(si * int(ai) for si, ai in zip(s, a))
    exp = O.result_of(O.scale(k % O.q, O.G))
    assert res == exp, (label, rank)
    if rank == 0:
        print("multi-gpu ok:", label, "world", world, "n/rank 2^%d" % logn, "ms", round(tm["total"], 3))
    sm.close()
dist.destroy_process_group()
