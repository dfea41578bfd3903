"""world_size-2 gloo test (CPU) of the N > 1 host logic: scene broadcast, round-robin view sharding, frame gather.
Rendering is replaced by a stand-in that stamps the view id into the frame: what is tested is who renders what and
that rank 0 ends up with every frame in view order — the same code bench.py / the NCCL path runs on the GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from torpedo_b200 import multiview as mv


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_views, result_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dev = torch.device("cpu")
        n = 1000
        src = torch.arange(n * 60, dtype=torch.float32).reshape(n, 60) if rank == 0 else None
        scene = mv.broadcast_scene(src, n, dev)
        assert torch.equal(scene, torch.arange(n * 60, dtype=torch.float32).reshape(n, 60))

        rendered = []

        def render_batch(view_ids, out):
            rendered.extend(view_ids)
            for k, v in enumerate(view_ids):
                out[k] = v + 1  # stand-in for GaussianEngine.raster_views
                out[k, 0, 0, 0] = rank

        frames = mv.render_views(render_batch, n_views, 4, 6, dev)
        assert rendered == mv.views_of_rank(n_views, rank, world)
        # chunked: every chunk's gather is started asynchronously behind its rendering; same result, same render order
        rendered.clear()
        chunked = mv.render_views(render_batch, n_views, 4, 6, dev, chunk=2)
        assert rendered == mv.views_of_rank(n_views, rank, world)
        assert (chunked is None) == (frames is None) and (frames is None or torch.equal(chunked, frames))
        # the shared frame array needs CUDA IPC: without a GPU the owner's allocation fails, EVERY rank learns it through the
        # same collectives and raises the same exception, and the caller falls back to the gathers above
        with pytest.raises(mv.SharedFramesUnavailable) as unavailable:
            mv.SharedFrames(n_views, 4, 6, 0)
        assert "rank 0" in str(unavailable.value)
        if rank == 0:
            assert frames.shape == (n_views, 4, 6, 4)
            for v in range(n_views):
                assert int(frames[v, 1, 1, 1]) == v + 1
                assert int(frames[v, 0, 0, 0]) == v % world
            np.save(result_path, frames.numpy())
        else:
            assert frames is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_views", [1, 5, 8])
def test_two_rank_sharding_and_gather(tmp_path, n_views):
    port = _free_port()
    out = str(tmp_path / "frames.npy")
    mp.spawn(_worker, args=(2, port, n_views, out), nprocs=2, join=True)
    frames = np.load(out)
    assert frames.shape[0] == n_views


def test_view_ownership_is_a_partition():
    for world in (1, 2, 4, 8):
        for n_views in (0, 1, 7, 64):
            seen = sorted(v for r in range(world) for v in mv.views_of_rank(n_views, r, world))
            assert seen == list(range(n_views))
            for v in range(n_views):
                r, s = mv.owner_of_view(v, world)
                assert mv.views_of_rank(n_views, r, world)[s] == v
    with pytest.raises(ValueError):
        mv.views_of_rank(4, 2, 2)
