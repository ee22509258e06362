"""Batch-index sharding of the sampling path across GPUs (one process per GPU).

Images are independent (per-image loss norm, per-image phi), so every rank samples its own images with a full model
replica and there is NO collective inside the loop.  ONE partition is used everywhere (the input loader
`osmosis_utils.data.ShardedImageLoader`, `sampling.run_sampling`, the final gather): image i of the run goes to rank
i mod world, and a rank holds its images in increasing i.  The only communication is optional and happens once, after
the last step: gathering the finished samples / phi / losses (NCCL over NVLink on GPUs, gloo in the CPU tests), and the
max-over-ranks reduction of benchmark timings.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_indices(n_images: int, rank: int, world: int):
    """Indices of the images rank `rank` of `world` owns: rank, rank + world, ... (balanced to within one image)."""
    return list(range(rank, n_images, world))


def max_over_ranks(value: float, device="cpu") -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def gather_images(local: torch.Tensor, n_images: int) -> torch.Tensor:
    """All-gather per-rank results [n_local, ...] (rows in `shard_indices` order) into [n_images, ...] in image order.
    Ragged shards (n_images % world != 0) are padded for the collective and trimmed afterwards."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    owned = [shard_indices(n_images, r, world) for r in range(world)]
    if local.shape[0] != len(owned[rank]):
        raise ValueError(f"rank {rank} holds {local.shape[0]} rows but owns {len(owned[rank])} of {n_images} images")
    n_max = max(len(o) for o in owned)
    pad = torch.zeros((n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    full = torch.empty((n_images,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for o, idx in zip(out, owned):
        if idx:
            full[torch.as_tensor(idx, device=local.device)] = o[: len(idx)]
    return full
