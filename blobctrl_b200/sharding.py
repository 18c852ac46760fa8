"""Batch sharding by image across the GPUs of one box.

Every image is independent in all three stages, so the hot path needs NO collective: rank r of G
renders images [r*N/G, (r+1)*N/G).  ``gather_maps`` is the optional final NCCL/gloo all-gather for a
consumer that needs every map on every rank; it is never part of the throughput figure.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_images: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first ``n % world`` ranks take one extra image."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_images, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_blobs(blobs: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    """Slice every per-image tensor ([N, ...]) of a blob dict to this rank's images."""
    n = blobs["covs"].shape[0]
    lo, hi = shard_bounds(n, rank, world)
    return {k: (v[lo:hi] if torch.is_tensor(v) and v.ndim >= 1 and v.shape[0] == n else v) for k, v in blobs.items()}


def gather_maps(local: torch.Tensor, n_images: int, group=None) -> torch.Tensor:
    """All-gather per-rank maps [N_r, ...] back to [N, ...] on every rank (ragged shards padded)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_bounds(n_images, r, world)[1] - shard_bounds(n_images, r, world)[0] for r in range(world)]
    cap = max(sizes)
    pad = local
    if local.shape[0] < cap:
        pad = torch.cat([local, local.new_zeros((cap - local.shape[0],) + tuple(local.shape[1:]))], 0)
    out = local.new_empty((world * cap,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    if all(s == cap for s in sizes):
        return out
    return torch.cat([out[r * cap: r * cap + sizes[r]] for r in range(world)], 0)
