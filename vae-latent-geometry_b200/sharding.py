"""Pair-list sharding over the GPUs of one box.

Curves are independent (per-curve loss, element-wise Adam, draws keyed on the global curve id),
so the only multi-GPU step is the partition of the pair list and one final gather of
(omega_optimized, energy) -- no collective on the step path.  Works with any torch.distributed
backend (NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced shard [lo, hi) of n curves for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n: int, world: int) -> List[int]:
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def gather_results(local: torch.Tensor, n_total: int, dst: int = 0):
    """Gather per-shard results (first dim = curves of this rank, contiguous shards in rank order)
    on rank `dst`; returns the full tensor there and None elsewhere.  Shards may be ragged."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = shard_sizes(n_total, world)
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    if rank == dst:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.gather(buf, parts, dst=dst)
        return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)
    dist.gather(buf, None, dst=dst)
    return None
