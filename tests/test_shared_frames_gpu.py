"""Views sharded over processes with the frames written straight into the collecting process's memory (SURVEY.md §8e;
include/tpdcu.h tpdcu_ipc_frames_*; torpedo_b200.multiview.SharedFrames / render_views_direct — the path bench.py's 64-view
batch takes on N GPUs).

The driver's GPU box for the tests has one GPU, so the two ranks here are two PROCESSES on cuda:0 joined by gloo: the CUDA IPC
mapping, the binding of a foreign pointer as render target and the completion fence are exactly those of the N-GPU run; only
the wire the stores travel on differs. Rank 0 then renders every view itself and compares the bytes."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, H, N_VIEWS = 320, 192, 5


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ubos(E):
    out = []
    for k in range(N_VIEWS):
        cam = E.PerspectiveCamera(W, H)
        cam.look_at(E.to_cartesian(float(np.float32(2.0 * np.pi * k / N_VIEWS)), 0.9, 5.0), (0, 0, 0), (0, 0, 1))
        out.append(cam.pack())
    return np.stack(out)


def _worker(rank, world, port, result_path):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from torpedo_b200 import engine as E
        from torpedo_b200 import multiview as mv
        from torpedo_b200 import scenes
        from torpedo_b200._lib import check, tpdcu

        lib = tpdcu()
        torch.zeros(1, device="cuda:0")
        g = scenes.garden(30000, seed=5, log_scale_mean=-3.6)
        scene = E.Scene()
        scene.add_group(g)
        eng = E.GaussianEngine(W, H, device=0)
        eng.compile(scene, E.Settings(3))
        ubos = _ubos(E)
        stream = torch.cuda.current_stream().cuda_stream
        shared = mv.SharedFrames(N_VIEWS, H, W, 0)
        rendered = []

        def render_to(view_ids, ptrs):       # the blend kernel's stores go straight into the (foreign) array
            for v, p in zip(view_ids, ptrs):
                rendered.append(v)
                check(lib.tpdcu_bind_output_device_ptr(eng.ctx, p, W * 4))
                eng.raster_ubo(ubos[v], 3, stream)
            eng.finish()

        def push_to(view_ids, ptrs):         # rendered locally, pushed by the copy engine (tpdcu_read_frame_async to device memory)
            check(lib.tpdcu_bind_output_device_ptr(eng.ctx, None, 0))
            for v, p in zip(view_ids, ptrs):
                rendered.append(v)
                eng.raster_ubo(ubos[v], 3, stream)
                check(lib.tpdcu_read_frame_async(eng.ctx, p, W * 4, stream))
            eng.finish()

        want = None
        for fn in (render_to, push_to):
            rendered.clear()
            for _ in range(2):   # the array is reused batch after batch
                mv.render_views_direct(fn, shared)
            assert rendered == mv.views_of_rank(N_VIEWS, rank, world) * 2
            if rank == 0:
                got = shared.tensor().cpu().numpy()
                if want is None:
                    check(lib.tpdcu_bind_output_device_ptr(eng.ctx, None, 0))
                    want = []
                    for v in range(N_VIEWS):
                        eng.raster_ubo(ubos[v], 3, stream)
                        want.append(eng.draw().copy())
                for v in range(N_VIEWS):
                    assert want[v][..., :3].max() > 0
                    assert np.array_equal(got[v], want[v]), f"{fn.__name__}: view {v} (rendered by rank {v % world})"
                shared.tensor().zero_()
                torch.cuda.synchronize()
            dist.barrier()
        if rank == 0:
            with pytest.raises(IndexError):
                shared.ptr_of_view(N_VIEWS)
            np.save(result_path, got)
        else:
            with pytest.raises(RuntimeError):
                shared.tensor()
        shared.close()
        eng.close()
    finally:
        dist.destroy_process_group()


def test_two_processes_render_into_one_frame_array(tmp_path, built_libs):
    import torch.multiprocessing as mp

    out = str(tmp_path / "frames.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    frames = np.load(out)
    assert frames.shape == (N_VIEWS, H, W, 4) and (frames[..., 3] == 255).all()


def test_ipc_frames_argument_errors(built_libs):
    import ctypes as C

    from torpedo_b200._lib import tpdcu

    lib = tpdcu()
    p = C.c_void_p()
    handle = (C.c_ubyte * 64)()
    assert lib.tpdcu_ipc_frames_create(0, 0, C.byref(p), handle) == -1
    assert lib.tpdcu_ipc_frames_create(0, 4096, None, handle) == -1
    assert lib.tpdcu_ipc_frames_create(0, 4096, C.byref(p), handle) == 0 and p.value
    q = C.c_void_p()
    assert lib.tpdcu_ipc_frames_open(0, handle, C.byref(q)) != 0   # a process cannot open its own handle
    assert b"cudaIpcOpenMemHandle" in lib.tpdcu_last_error()
    assert lib.tpdcu_ipc_frames_destroy(0, p) == 0
    assert lib.tpdcu_ipc_frames_close(0, None) == 0 and lib.tpdcu_ipc_frames_destroy(0, None) == 0
