"""Multi-GPU plumbing: one process per GPU, contiguous block partition of the batch, no data-path collective.

Trajectories are independent (each is a pure function of its own parameter record, Circle.cpp:30-94), so the batch
is sharded by index with tgx_shard_range and every rank evaluates its own shard into its own HBM.  The only
exchange the path has is optional: an all-gather of the 1-byte feasibility flags (BASELINE.json configs[4]).  On the
GPUs it is issued by libtgx itself (engine.Comm -> tgx_gather_flags -> ncclAllGather, include/tgx.h); the
torch.distributed version below carries the same partition rule on any backend and is what the CPU tests run over gloo.
"""
from __future__ import annotations

from typing import Optional

from .engine import shard_range


def shard_bounds(n: int, rank: int, world: int):
    """[lo, hi) owned by `rank` (floor(rank*n/world) boundaries, the rule tgx_shard_range implements)."""
    return shard_range(n, rank, world)


def shard_sizes(n: int, world: int):
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def gather_flags(local_flags, n_total: int, group=None):
    """All-gather per-rank uint8 flag vectors (shards may differ in length by one) into the full [n_total] vector.

    local_flags: torch.uint8 tensor holding this rank's shard, on the device the process group works on.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(n_total, world)
    assert local_flags.dtype == torch.uint8 and local_flags.numel() == sizes[rank], (local_flags.shape, sizes[rank])
    width = max(sizes)
    padded = torch.zeros(width, dtype=torch.uint8, device=local_flags.device)
    padded[: sizes[rank]] = local_flags
    gathered = torch.empty(world * width, dtype=torch.uint8, device=local_flags.device)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    parts = [gathered[r * width: r * width + sizes[r]] for r in range(world)]
    return torch.cat(parts)


def count_feasible(local_flags, group=None) -> int:
    """Global number of feasible trajectories (sum over ranks)."""
    import torch
    import torch.distributed as dist

    t = local_flags.sum(dtype=torch.int64).reshape(1)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, group=group)
    return int(t.item())
