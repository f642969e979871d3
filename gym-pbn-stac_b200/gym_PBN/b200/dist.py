"""Multi-GPU plumbing: one process per GPU (torch.distributed), envs sharded by contiguous ranges of GLOBAL env id.

Envs never interact, so there is no data-path collective; the only exchanges are the SUM all-reduce of the SSD
histogram (uint64 [2^g], 1 KiB at g = 7) and of the small episode-statistics vector — NCCL on GPUs, gloo in CPU tests.
Because every env's Philox stream is keyed by its global id, an N-rank run is bit-identical to a 1-rank run.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(total: int, rank: int = None, world_size: int = None, align: int = 1):
    """Contiguous [start, stop) of global env ids owned by `rank` (first ranks take the remainder).  With align > 1
    every boundary is a multiple of `align` (the SSD perturbation stream is shared by groups of 32 consecutive ids)."""
    if rank is None or world_size is None:
        rank, world_size = world()
    units = -(-int(total) // align)
    base, rem = divmod(units, int(world_size))
    start = rank * base + min(rank, rem)
    stop = start + base + (1 if rank < rem else 0)
    return min(start * align, int(total)), min(stop * align, int(total))


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    """In-place SUM over ranks (no-op when not distributed)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


class EpisodeStats:
    """Running episode statistics of a VectorEnv: [episodes, return sum, length sum, successes, cap hits, env steps, env steps
    whose action was out of range (the intervention is ignored on the device and counted here), reserved]."""

    FIELDS = ("episodes", "return_sum", "length_sum", "successes", "cap_hits", "env_steps", "invalid_actions", "reserved")

    def __init__(self, device):
        self.v = torch.zeros(len(self.FIELDS), dtype=torch.int64, device=device)

    def reduced(self):
        t = self.v.clone()
        allreduce_sum_(t)
        return dict(zip(self.FIELDS, t.tolist()))
