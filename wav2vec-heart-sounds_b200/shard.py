"""Index sharding of recordings over the GPUs of one node.

Every function on the hot path is per row, so the only multi-GPU structure is: rank ``r`` of ``G`` owns the
contiguous block of recordings ``[r * ceil(B / G), ...)`` and runs the same kernels on it -- no collective on the data
path.  The one optional exchange is an all-gather of the equally sized window tensors into a data-parallel
trainer, issued on ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests) and kept off
the timed path.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block ``[lo, hi)`` of ``n_items`` owned by ``rank``; blocks differ in size by at most one chunk
    (the last ranks may be short or empty when ``world`` does not divide ``n_items``)."""
    if not (0 <= rank < world):
        raise ValueError("rank must be in [0, world)")
    per = -(-n_items // world)
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)


def local_shard(x: torch.Tensor, rank: int | None = None, world: int | None = None) -> torch.Tensor:
    """The slice of a batch-first tensor this rank owns (defaults: the initialised process group)."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


def gather_windows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather the per-rank window tensors ``[b_r, ...]`` into ``[n_total, ...]`` on every rank.  Ranks pad to
    the common block size so a plain ``all_gather_into_tensor`` (no ``AllGatherV``) suffices."""
    world = dist.get_world_size(group)
    per = -(-n_total // world)
    pad = per - local.shape[0]
    if pad < 0:
        raise ValueError("local shard is larger than its block")
    block = local if pad == 0 else torch.cat([local, local.new_zeros((pad,) + tuple(local.shape[1:]))], dim=0)
    out = local.new_empty((per * world,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, block.contiguous(), group=group)
    return out[:n_total]


def sharded_apply(fn, x: torch.Tensor, *, gather: bool = False, group=None):
    """Run ``fn`` on this rank's block of ``x``; with ``gather`` return the all-gathered result."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_bounds(x.shape[0], rank, world)
    out = fn(x[lo:hi])
    return gather_windows(out, x.shape[0], group) if gather else out
