"""Multi-GPU plumbing.  Frames are independent units, so the batch is split contiguously across
ranks (one process per GPU, torch.distributed for the rendezvous only); the data path has no
collective.  Results are gathered on rank 0 in frame order.  """
from __future__ import annotations

from typing import List, Tuple


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, end) slice of `total` frames owned by `rank`; sizes differ by at most 1."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_counts(local_counts, rank: int, world: int):
    """Frame-ordered concatenation of the per-rank face counts on rank 0 (None elsewhere)."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(local_counts, dtype=torch.int32)
    if world == 1:
        return t
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([t.numel()], dtype=torch.int64))
    mx = int(max(int(s) for s in sizes))
    pad = torch.zeros(mx, dtype=torch.int32)
    pad[:t.numel()] = t
    bufs = [torch.zeros(mx, dtype=torch.int32) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, bufs, dst=0)
    if rank != 0:
        return None
    return torch.cat([b[:int(s)] for b, s in zip(bufs, sizes)])


def max_over_ranks(value: float, world: int, device=None) -> float:
    """Step time of the job = the slowest rank's device time."""
    if world == 1:
        return value
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
