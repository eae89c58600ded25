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


class EditLanes:
    """Several images in flight on ONE GPU.

    One image's 650 UNet calls are a serial chain of ~360 short kernels each: between two kernels of a chain the SMs idle
    for the launch latency, and the B=1 inversion third of the chain cannot fill the chip at all.  A second, independent
    image on its own CUDA stream (own engine handle = own activation arena, CUDA graph and controller state; the weights
    are replicated, 1.7 GB per lane) fills those bubbles.  Each lane is driven by its own host thread (the C ABI calls
    release the GIL) under its own `torch.cuda.Stream`; results come back in job order, and the caller's current stream
    waits for every lane before `run` returns, so ordinary stream semantics hold for the caller.

    `make_editor()` must build a fresh model + editor pair (e.g. `P2PEditor([...], device, model=FusedModel(...))`).
    """

    def __init__(self, make_editor, lanes: int = 2, device=None):
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.editors = [make_editor() for _ in range(max(1, lanes))]
        # a non-CUDA device keeps the lane / thread / ordering logic testable on a CPU-only machine (no streams)
        self.streams = ([torch.cuda.Stream(device=self.device) for _ in self.editors]
                        if self.device.type == "cuda" else [None] * len(self.editors))

    def __len__(self):
        return len(self.editors)

    def run(self, jobs):
        """jobs: callables `job(editor) -> result`; job i runs on lane i % lanes, jobs of one lane in order."""
        import contextlib
        import threading

        n = len(self.editors)
        jobs = list(jobs)
        results = [None] * len(jobs)
        errors = []
        cuda = self.device.type == "cuda"
        caller = torch.cuda.current_stream(self.device) if cuda else None

        def worker(lane):
            try:
                if cuda:
                    torch.cuda.set_device(self.device)
                    self.streams[lane].wait_stream(caller)  # inputs produced on the caller's stream are visible
                with (torch.cuda.stream(self.streams[lane]) if cuda else contextlib.nullcontext()):
                    for i in range(lane, len(jobs), n):
                        results[i] = jobs[i](self.editors[lane])
            except BaseException as e:  # surfaced in the caller's thread
                errors.append(e)

        threads = [threading.Thread(target=worker, args=(lane,)) for lane in range(min(n, len(jobs)))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if cuda:
            for s in self.streams:
                caller.wait_stream(s)
        if errors:
            raise errors[0]
        return results
