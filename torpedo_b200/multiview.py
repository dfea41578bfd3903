"""View sharding for multi-view batches across the GPUs of one box (BASELINE.json config 5).

The Gaussian set is replicated (one broadcast), independent camera views are partitioned round-robin
(view v -> rank v mod world), every rank renders its own views with its own GaussianEngine, and the frames are
gathered on rank 0. There is no data-path collective inside a frame: the path does not shard within a frame
(global scan + global sort), it shards across views. The functions here only move bytes with torch.distributed —
NCCL on the GPU box, gloo in the CPU tests — and never render anything themselves.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch
import torch.distributed as dist


def views_of_rank(n_views: int, rank: int, world: int) -> list[int]:
    """Round-robin ownership: rank r renders views r, r + world, r + 2*world, ..."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_views, world))


def owner_of_view(view: int, world: int) -> tuple[int, int]:
    """(rank, slot in that rank's local batch) of a view."""
    return view % world, view // world


def broadcast_scene(records: torch.Tensor | None, n: int, device: torch.device, src: int = 0) -> torch.Tensor:
    """Replicate the (n, 60) float32 GaussianPoint records from `src` to every rank (one collective, once per scene)."""
    buf = torch.empty((n, 60), dtype=torch.float32, device=device)
    if dist.get_rank() == src:
        if records is None or tuple(records.shape) != (n, 60):
            raise ValueError("source rank must provide the (n, 60) records")
        buf.copy_(records)
    dist.broadcast(buf, src=src)
    return buf


def gather_frames(local_frames: torch.Tensor, n_views: int, dst: int = 0) -> torch.Tensor | None:
    """Collect the per-rank frame batches on `dst` and put them back in view order.

    local_frames: (slots, H, W, 4) uint8 with slots = ceil(n_views / world) on EVERY rank (ranks that own one view fewer
    leave their last slot unused) so that one fixed-size gather suffices. Returns (n_views, H, W, 4) on dst, None elsewhere.
    """
    world, rank = dist.get_world_size(), dist.get_rank()
    slots = -(-n_views // world)
    if local_frames.shape[0] != slots:
        raise ValueError(f"every rank must pass {slots} slots, got {local_frames.shape[0]}")
    if world == 1:
        return local_frames[:n_views]
    parts = [torch.empty_like(local_frames) for _ in range(world)] if rank == dst else None
    dist.gather(local_frames, parts, dst=dst)
    if rank != dst:
        return None
    out = torch.empty((n_views,) + tuple(local_frames.shape[1:]), dtype=local_frames.dtype, device=local_frames.device)
    for v in range(n_views):
        r, s = owner_of_view(v, world)
        out[v] = parts[r][s]
    return out


def render_views(render_batch: Callable[[Sequence[int], torch.Tensor], None], n_views: int, height: int, width: int,
                 device: torch.device, dst: int = 0, chunk: int = 0) -> torch.Tensor | None:
    """Shard `n_views` over the ranks, let `render_batch(view_ids, out)` fill this rank's slots (out[k] <- view_ids[k]),
    then gather. `render_batch` is the only place pixels are produced (GaussianEngine.raster_views on the GPU box).

    chunk > 0: render `chunk` slots at a time and start the gather of each chunk asynchronously, so that the transfer of
    one chunk overlaps the rendering of the next; the result is the same tensor."""
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = views_of_rank(n_views, rank, world)
    slots = -(-n_views // world)
    local = torch.zeros((slots, height, width, 4), dtype=torch.uint8, device=device)
    if chunk <= 0 or world == 1:
        if mine:
            render_batch(mine, local[: len(mine)])
        return gather_frames(local, n_views, dst)
    parts = [torch.empty_like(local) for _ in range(world)] if rank == dst else None
    pending = []
    for s0 in range(0, slots, chunk):
        s1 = min(s0 + chunk, slots)
        ids = mine[s0:s1]
        if ids:
            render_batch(ids, local[s0:s0 + len(ids)])
        pending.append(dist.gather(local[s0:s1], [p[s0:s1] for p in parts] if rank == dst else None, dst=dst, async_op=True))
    for work in pending:
        work.wait()
    if rank != dst:
        return None
    out = torch.empty((n_views,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for v in range(n_views):
        r, s = owner_of_view(v, world)
        out[v] = parts[r][s]
    return out
