"""Multi-GPU plumbing for the ASM path: one process per GPU, contiguous batch shards, NO collective on the hot
path (every sample's propagation depends only on its own field and distance -- SURVEY.md section 8e).
NCCL (or gloo on CPU, for tests) is used only for the optional final gather of the holograms.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of `batch` samples: the first `batch % world` ranks get one extra sample."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(t: torch.Tensor, rank: Optional[int] = None, world: Optional[int] = None) -> torch.Tensor:
    """This rank's slice of a batch-major tensor."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


def gather_batch(local: torch.Tensor, batch: int, group=None) -> torch.Tensor:
    """Optional, off the hot path: all ranks end up with the full [batch, ...] tensor assembled from the
    per-rank shards produced with `shard_bounds` (ragged shards are padded to the largest one)."""
    world = dist.get_world_size(group)
    sizes = [shard_bounds(batch, r, world) for r in range(world)]
    biggest = max(hi - lo for lo, hi in sizes)
    if local.shape[0] < biggest:
        pad = torch.zeros((biggest - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    out = torch.empty((world * biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    parts = [out[r * biggest: r * biggest + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(parts, dim=0)
