"""World-size-2 gloo test (CPU) of the host-side sharding logic: contiguous shard ranges, the
all-gather of partial sums in rank order, that combining per-shard partial MSMs gives the full
MSM, and the hand-over of the NCCL communicator id (mgb_comm_unique_id) to the other ranks.  The per-shard partials are computed by the oracle here (there is no GPU in this test); the GPU
version of the same flow is scripts/check_multi_gpu.py, launched under torchrun by tests/test_gpu_round2.py on >= 2 GPUs."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, out_q):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from montgomery_b200 import inputs
    from montgomery_b200.distributed import all_gather_partials, shard_range
    from tests.helpers import OracleCurve
    O = OracleCurve("bls12-377")
    a = inputs.known_dlogs(11, n)
    sc = inputs.scalars_to_ints(inputs.random_scalars(O.q, n, 12))
    lo, hi = shard_range(n, rank, world)
    # partial sum of this rank's shard (oracle stands in for the device kernel in this CPU test)
    k = sum(s * int(ai) for s, ai in zip(sc[lo:hi], a[lo:hi])) % O.q
    part = O.P.scale(k, O.P.one)
    limbs = []
    for coord in part:
        limbs += [(coord >> (32 * i)) & 0xFFFFFFFF for i in range(12)]
    t = torch.tensor(np.array(limbs, dtype=np.int64), dtype=torch.int64)
    g = all_gather_partials(t)
    assert g.shape == (world, 36)
    total = O.P.zero
    for r in range(world):
        row = [int(v) for v in g[r]]
        pt = tuple(sum(row[12 * j + i] << (32 * i) for i in range(12)) for j in range(3))
        total = O.P.add(total, pt)
    full = sum(s * int(ai) for s, ai in zip(sc, a)) % O.q
    ok = O.P.to_affine(total) == O.P.to_affine(O.P.scale(full, O.P.one))
    # the communicator id of the in-library collective travels over the same group (mgb_comm_unique_id needs no GPU):
    # every rank must end up with rank 0's 128 bytes
    from montgomery_b200.distributed import exchange_comm_id
    cid = exchange_comm_id(rank)
    ids = [None] * world
    dist.all_gather_object(ids, cid)
    ok = ok and len(cid) == 128 and any(cid) and all(i == ids[0] for i in ids)
    out_q.put((rank, lo, hi, ok))
    dist.destroy_process_group()


def test_shard_ranges_cover():
    sys.path.insert(0, ROOT)
    from montgomery_b200.distributed import shard_range
    for n in (0, 1, 7, 8, 1000, 1 << 20):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                assert lo == prev and hi >= lo
                prev = hi
            assert prev == n


def test_two_rank_gloo_allgather_and_combine():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, n = 2, 101
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1:3] == (0, 51) and res[1][1:3] == (51, 101)
    assert all(r[3] for r in res)
