"""Multi-GPU MSM: one process per GPU, points sharded contiguously, one small all-gather.

Replaces the reference's only parallel mechanism, the SPMD thread pool of src/threads/*.ts: there
every worker runs the same msm() on a `range(N)` slice of the inputs (threads.ts:354-359) and the
main thread adds the per-thread partial sums (msm-batched-affine.ts:311-320).  Here rank r owns the
pairs [lo, hi) of `shard_range`, computes the partial sum of its shard on its own B200, and the G
un-normalised partial accumulators (192 bytes each for BLS12-377) are exchanged with ONE all-gather
over NCCL / NVLink; every rank then adds them and normalises.  MSM has no other exchange step.
"""
import torch
import torch.distributed as dist

from .api import MsmEngine


def shard_range(n: int, rank: int, world: int):
    """Contiguous split, the rule of the reference's `range()` (src/threads/threads.ts:354-359)."""
    per = -(-n // world)
    lo = min(n, per * rank)
    return lo, min(n, lo + per)


def all_gather_partials(partial: torch.Tensor) -> torch.Tensor:
    """partial: 1-D tensor (any device / backend) -> (world, len) tensor in rank order."""
    world = dist.get_world_size()
    out = torch.empty(world * partial.numel(), dtype=partial.dtype, device=partial.device)
    dist.all_gather_into_tensor(out, partial.contiguous())
    return out.view(world, partial.numel())


class ShardedMsm:
    """The sharded engine.  Every rank calls the same methods with its own shard."""

    def __init__(self, curve, local_device: int, max_points_per_rank: int):
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.engine = MsmEngine(curve, local_device, max_points_per_rank)
        nwords = self.engine.partial_bytes // 4
        self._partial = torch.zeros(nwords, dtype=torch.int32, device=torch.device("cuda", local_device))

    def set_points(self, xy_bytes_shard, is_zero=None):
        return self.engine.set_points(xy_bytes_shard, is_zero)

    def random_points(self, n_local: int, seed: int):
        """Known-dlog points; rank r uses seed + r so the global set is the union of the shards."""
        self.engine.random_points(n_local, seed + self.rank)

    def msm(self, scalars, n_local: int, on_device: bool = False, c=None):
        """scalars: pinned/any host tensor or numpy array (on_device=False) or a CUDA uint8 tensor."""
        if self.world == 1:
            if on_device:
                return self.engine.msm(None, n=n_local, c=c, device_ptr=scalars.data_ptr())
            arr = scalars.numpy() if isinstance(scalars, torch.Tensor) else scalars
            return self.engine.msm(arr, n=n_local, c=c)
        ptr = scalars.data_ptr() if isinstance(scalars, torch.Tensor) else scalars.ctypes.data
        tm = self.engine.msm_partial(ptr, on_device, n_local, self._partial.data_ptr(), c=c)
        gathered = all_gather_partials(self._partial)
        torch.cuda.current_stream().synchronize()
        res = self.engine.combine_partials(gathered.data_ptr(), self.world)
        tm["n_launches"] += 2          # the all-gather and the combine/normalise kernel
        return res, tm

    def close(self):
        self.engine.close()
