"""Image-parallel sharding across the GPUs of one box.

The editing loop has no exchange step: images are independent work items, every rank holds a full weight replica and the
CFG / source-target rows of one image stay batched on one GPU (the controllers couple them in all 32 attention layers).
`torch.distributed` (NCCL over NVLink on the GPU box, gloo in the CPU tests) is used for exactly two things: distributing
the input latents and gathering the edited latents.  Nothing here touches the data path of a step.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split with the remainder spread over the first ranks (700 images on 8 GPUs -> 88,88,88,88,87,87,87,87)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n_items: int, world: int) -> List[int]:
    return [shard_bounds(n_items, r, world)[1] - shard_bounds(n_items, r, world)[0] for r in range(world)]


def scatter_items(items: Optional[torch.Tensor], n_items: int, item_shape: Sequence[int], dtype, device, dist=None,
                  src: int = 0) -> torch.Tensor:
    """Rank `src` holds `items` (n_items, *item_shape); every rank returns its own contiguous shard.
    Implemented as one broadcast (KB-MB messages: latency-bound, never on the critical path)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        assert items is not None
        return items.to(device)
    rank, world = dist.get_rank(), dist.get_world_size()
    buf = torch.empty((n_items, *item_shape), dtype=dtype, device=device)
    if rank == src:
        buf.copy_(items.to(device))
    dist.broadcast(buf, src=src)
    lo, hi = shard_bounds(n_items, rank, world)
    return buf[lo:hi].clone()


def gather_items(local: torch.Tensor, n_items: int, dist=None) -> torch.Tensor:
    """All ranks return the full (n_items, ...) tensor in global item order (all_gather of padded shards)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = shard_sizes(n_items, world)
    mx = max(sizes)
    pad = torch.zeros((mx, *local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)
