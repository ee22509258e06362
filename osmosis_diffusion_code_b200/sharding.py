"""Batch-index sharding of the sampling path across GPUs (one process per GPU).

Images are independent (per-image loss norm, per-image phi), so rank r of R samples images [lo, hi) with a full model
replica and there is NO collective inside the loop.  The only communication is optional and happens once, after the
last step: gathering the finished samples / phi / losses (NCCL over NVLink on GPUs, gloo in the CPU tests), and the
max-over-ranks reduction of benchmark timings.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_images: int, rank: int, world: int):
    """Contiguous, balanced partition of range(n_images): the first n % world ranks take one extra image."""
    base, extra = divmod(n_images, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device="cpu") -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def gather_images(local: torch.Tensor, n_images: int) -> torch.Tensor:
    """All-gather per-rank results [n_local, ...] into [n_images, ...] in image order (ragged shards allowed)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_images, r, world) for r in range(world)]
    n_max = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)
