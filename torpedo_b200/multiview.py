"""View sharding for multi-view batches across the GPUs of one box (BASELINE.json config 5).

The Gaussian set is replicated (one broadcast), independent camera views are partitioned round-robin
(view v -> rank v mod world), every rank renders its own views with its own GaussianEngine, and the frames are
collected on rank 0 — pushed or stored into one frame array in rank 0's HBM that every process has mapped
(`SharedFrames`, the fast path on a GPU box) or gathered with the process group's collectives (`render_views`, the
generic path: NCCL without peer mappings, gloo in the CPU tests). There is no data-path collective inside a frame: the
path does not shard within a frame (global scan + global sort), it shards across views. Nothing here renders.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch
import torch.distributed as dist


def _world_rank() -> tuple[int, int]:
    """(world, rank) of the default process group; a single process without one is world 1."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


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
    world, rank = _world_rank()
    buf = torch.empty((n, 60), dtype=torch.float32, device=device)
    if rank == src:
        if records is None or tuple(records.shape) != (n, 60):
            raise ValueError("source rank must provide the (n, 60) records")
        buf.copy_(records)
    if world > 1:
        dist.broadcast(buf, src=src)
    return buf


def gather_frames(local_frames: torch.Tensor, n_views: int, dst: int = 0) -> torch.Tensor | None:
    """Collect the per-rank frame batches on `dst`, in view order.

    local_frames: (slots, H, W, 4) uint8 with slots = ceil(n_views / world) on EVERY rank (ranks that own one view fewer
    leave their last slot unused) so that fixed-size gathers suffice. Returns (n_views, H, W, 4) on dst, None elsewhere.
    """
    return _gather_slots(local_frames, n_views, dst, 0, local_frames.shape[0], None)


def _gather_slots(local_frames, n_views, dst, s0, s1, out, pending=None):
    """Gather slots [s0, s1) straight into view order: view v = slot * world + owner, so the frame rank r holds in slot s
    belongs at out[s * world + r] and one gather per slot with that list of destinations needs no re-ordering copy
    afterwards (a gather of whole per-rank batches followed by 64 device copies on `dst` cost a third of the 8-GPU batch).
    With `pending` the gathers are started asynchronously and appended to it."""
    world, rank = _world_rank()
    slots = -(-n_views // world)
    if local_frames.shape[0] != slots:
        raise ValueError(f"every rank must pass {slots} slots, got {local_frames.shape[0]}")
    if world == 1:
        return local_frames[:n_views]
    if out is None and rank == dst:
        out = torch.empty((slots * world,) + tuple(local_frames.shape[1:]), dtype=local_frames.dtype, device=local_frames.device)
    for s in range(s0, s1):
        dests = [out[s * world + r] for r in range(world)] if rank == dst else None
        work = dist.gather(local_frames[s], dests, dst=dst, async_op=pending is not None)
        if pending is not None:
            pending.append(work)
    return out


def render_views(render_batch: Callable[[Sequence[int], torch.Tensor], None], n_views: int, height: int, width: int,
                 device: torch.device, dst: int = 0, chunk: int = 0) -> torch.Tensor | None:
    """Shard `n_views` over the ranks, let `render_batch(view_ids, out)` fill this rank's slots (out[k] <- view_ids[k]),
    and gather the frames on `dst` in view order. `render_batch` is the only place pixels are produced (the GaussianEngine
    on the GPU box); it may return before the frames are finished as long as the work is ordered on the current stream.

    chunk > 0: render `chunk` slots at a time and start the gathers of each chunk asynchronously, so that the transfer of
    one chunk overlaps the rendering of the next; the result is the same tensor."""
    world, rank = _world_rank()
    mine = views_of_rank(n_views, rank, world)
    slots = -(-n_views // world)
    local = torch.zeros((slots, height, width, 4), dtype=torch.uint8, device=device)
    if world == 1:
        if mine:
            render_batch(mine, local[: len(mine)])
        return local[:n_views]
    step = chunk if chunk > 0 else max(slots, 1)
    out = None
    pending = [] if chunk > 0 else None
    for s0 in range(0, slots, step):
        s1 = min(s0 + step, slots)
        ids = mine[s0:s1]
        if ids:
            render_batch(ids, local[s0:s0 + len(ids)])
        out = _gather_slots(local, n_views, dst, s0, s1, out, pending)
    for work in pending or []:
        work.wait()
    return out[:n_views] if rank == dst else None


class SharedFramesUnavailable(RuntimeError):
    """CUDA IPC could not map the frame array into every rank (raised on EVERY rank alike): collect with `render_views`."""


class SharedFrames:
    """The batch's (n_views, H, W, 4) uint8 frame array, living in `dst`'s HBM and mapped into every rank of the box.

    The gather above costs the collecting GPU twice: NCCL's receive kernels take SMs away from its own views, and the last
    chunk's transfer trails the last frame. Here nothing is gathered: rank r writes its view v into slot v of this array
    (`ptr_of_view`) — either by pushing the finished frame with `tpdcu_read_frame_async` (a copy engine moves it over NVLink /
    NVSwitch while the next view renders: the fastest at every N measured) or by binding the slot as the render
    target, so that the blend kernel's pixel stores cross the link as they are produced (as fast up to 4 GPUs; at 8 the seven
    senders stall on their fine-grained stores) — and the only collective left is the completion fence. Set up once per
    (batch shape, process group) — an IPC handle exchange and a peer mapping — like the scene broadcast;
    `include/tpdcu.h` (tpdcu_ipc_frames_*) is the ABI underneath. Raises SharedFramesUnavailable on every rank alike when
    the mapping cannot be made (no peer access, IPC forbidden in the container).
    """

    def __init__(self, n_views: int, height: int, width: int, device_index: int, dst: int = 0):
        import ctypes as C

        from ._lib import TpdError, check, tpdcu
        self._lib, self._check = tpdcu(), check
        self.world, self.rank = _world_rank()
        self.n_views, self.height, self.width, self.dst, self.device_index = n_views, height, width, dst, device_index
        self.frame_bytes = height * width * 4
        self.owner = self.rank == dst
        self._ptr = C.c_void_p()
        self._fence = None
        handle = (C.c_ubyte * 64)()
        # Every rank goes through the same collectives whether its own step worked or not, and all of them agree on the outcome:
        # either every rank holds a mapping or none does and all raise SharedFramesUnavailable (callers fall back to the gathers).
        problem = ""
        if self.owner:
            try:
                check(self._lib.tpdcu_ipc_frames_create(device_index, n_views * self.frame_bytes, C.byref(self._ptr), handle))
            except TpdError as e:
                problem = f"rank {self.rank}: {e}"
        box = [bytes(handle), problem]
        if self.world > 1:
            dist.broadcast_object_list(box, src=dst)
        problem = box[1]
        if not self.owner and not problem:
            raw = (C.c_ubyte * 64).from_buffer_copy(box[0])
            try:
                check(self._lib.tpdcu_ipc_frames_open(device_index, raw, C.byref(self._ptr)))
            except TpdError as e:
                problem = f"rank {self.rank}: {e}"
        if self.world > 1:
            problems = [None] * self.world
            dist.all_gather_object(problems, problem)
            problem = next((p for p in problems if p), "")
        if problem:
            self._release()
            raise SharedFramesUnavailable(problem)

    def _release(self) -> None:
        if self._ptr.value:
            if self.owner:
                self._check(self._lib.tpdcu_ipc_frames_destroy(self.device_index, self._ptr))
            else:
                self._check(self._lib.tpdcu_ipc_frames_close(self.device_index, self._ptr))
            self._ptr.value = None

    def ptr_of_view(self, view: int) -> int:
        if not (0 <= view < self.n_views):
            raise IndexError(view)
        return self._ptr.value + view * self.frame_bytes

    def fence(self) -> None:
        """Returns (stream-ordered on NCCL, host-ordered otherwise) once every rank's frames of the batch are in the array."""
        if self.world == 1:
            return
        if dist.get_backend() == "nccl":
            if self._fence is None:
                self._fence = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", self.device_index))
            dist.all_reduce(self._fence)     # queued behind this rank's frames on the current stream
        else:
            torch.cuda.synchronize(self.device_index)
            dist.barrier()

    def tensor(self) -> torch.Tensor:
        """The frames as a torch tensor (collecting rank only; the other ranks hold a peer mapping, not local memory)."""
        if not self.owner:
            raise RuntimeError("only the collecting rank can read the frames")
        shape = (self.n_views, self.height, self.width, 4)
        iface = {"shape": shape, "typestr": "|u1", "data": (self._ptr.value, False), "version": 3, "strides": None}
        holder = type("_Frames", (), {"__cuda_array_interface__": iface, "_keep": self})()
        return torch.as_tensor(holder, device=torch.device("cuda", self.device_index))

    def close(self) -> None:
        if self._ptr.value:
            if self.world > 1:
                self.fence()
                torch.cuda.synchronize(self.device_index)
                if dist.get_backend() == "nccl":
                    dist.barrier()
            self._release()


def render_views_direct(render_to: Callable[[Sequence[int], Sequence[int]], None], shared: SharedFrames) -> None:
    """Shard the views like `render_views`, but let `render_to(view_ids, target_ptrs)` deliver each view straight into its slot
    of the shared frame array (push the finished frame with tpdcu_read_frame_async, or bind the slot as the render target);
    ends with the completion fence. On return (stream-ordered on NCCL) `shared.tensor()` on the collecting rank holds the
    batch in view order. A frame whose pair buffers overflowed is only repeated by the engine's next finish(), into the
    engine's own target: callers that cannot rule that out by a warm-up over their views check `frames_repeated()` after the
    batch and deliver those views again."""
    mine = views_of_rank(shared.n_views, shared.rank, shared.world)
    if mine:
        render_to(mine, [shared.ptr_of_view(v) for v in mine])
    shared.fence()
