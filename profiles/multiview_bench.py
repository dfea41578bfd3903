"""BASELINE config 5 through torpedo_b200.multiview on N GPUs of one box: 64 views of a 3 M-Gaussian scene at 1080p, views
sharded round-robin, scene replicated with one NCCL broadcast, frames gathered on rank 0 in asynchronous chunks.
torchrun --nproc-per-node N profiles/multiview_bench.py   (N = 1 works too: python profiles/multiview_bench.py)"""
import json
import os
import statistics
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from torpedo_b200 import engine as E  # noqa: E402
from torpedo_b200 import multiview as mv  # noqa: E402
from torpedo_b200 import scenes  # noqa: E402

N, W, H, DEG, VIEWS, RADIUS, CHUNK = 3_000_000, 1920, 1080, 3, 64, 5.0, 4
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if not dist.is_initialized():
    if world == 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

g = scenes.garden(N, 5, log_scale_mean=bench.LOG_SCALE_MEAN) if rank == 0 else None
recs = mv.broadcast_scene(torch.from_numpy(g).to(dev) if rank == 0 else None, N, dev)
eng = E.GaussianEngine(W, H, device=local_rank)
stream = torch.cuda.current_stream().cuda_stream
eng.compile_device(recs.data_ptr(), N, E.Settings(DEG), stream)
torch.cuda.synchronize()
del recs
ubos = []
for k in range(VIEWS):
    cam = E.PerspectiveCamera(W, H)
    cam.look_at(E.to_cartesian(2.0 * np.pi * k / VIEWS, 0.9, RADIUS), (0, 0, 0), (0, 0, 1))
    ubos.append(cam.pack())
ubos = np.stack(ubos)


def render_batch(view_ids, out):
    eng.raster_views(ubos[list(view_ids)], out.data_ptr(), H * W * 4, DEG, stream)


times, frames = [], None
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    frames = mv.render_views(render_batch, VIEWS, H, W, dev, chunk=CHUNK)
    e1.record()
    eng.finish()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rep >= 2:
        times.append(float(t.item()))
if rank == 0:
    from oracle import oracle as O  # checker only
    ms = statistics.median(times)
    out = frames.cpu().numpy()
    worst, off = 0, 0
    checked = [0, 21, 42, 63]
    for k in checked:
        ref = O.render(g, ubos[k], W, H, DEG)
        d = np.abs(out[k].astype(np.int32) - ref.rgba.astype(np.int32))
        worst, off = max(worst, int(d[..., :3].max())), off + int((d[..., :3].max(axis=-1) > 0).sum())
    print(json.dumps({"config": "5 64 views x 3M SH3 1080p via torpedo_b200.multiview", "n_gpus": world, "views": VIEWS, "ms_per_batch": round(ms, 3),
                      "ms_per_view": round(ms / VIEWS, 4), "views_per_s": round(VIEWS / ms * 1e3, 1), "gather_chunk": CHUNK,
                      "parity": {"views_checked": checked, "max_abs_rgb_lsb": worst, "pixels_off_by_one": off}}))
eng.close()
dist.destroy_process_group()
