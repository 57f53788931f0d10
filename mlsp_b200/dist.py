"""Multi-GPU plumbing for the hot path: the batch of clouds is sharded across ranks, one process per GPU.
Every op is independent per cloud, so there is NO collective on the data path; the
only exchanges are the timing reduction of bench.py and -- in a full training step -- DDP's gradient
all-reduce over NCCL/NVLink, which belongs to the torch layers outside this package."""
from __future__ import annotations

import os

import torch


def env_rank():
    """(rank, world_size, local_rank) from the torchrun environment (1-process defaults)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_bounds(global_batch: int, rank: int, world: int):
    """Contiguous, balanced slice [lo, hi) of a global batch owned by `rank` (strong-scaling layout)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(global_batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Max of a host scalar over all ranks (the time every multi-GPU number is reported with)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def gather_counts(local_count: int, device=None, group=None):
    """All ranks' processed-unit counts (so throughput = sum(units) / max(time))."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [int(local_count)]
    t = torch.tensor([local_count], dtype=torch.int64, device=device or "cpu")
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, t, group=group)
    return [int(o.item()) for o in out]
