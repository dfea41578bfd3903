import sys, os, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from torpedo_b200 import engine as E
W, H = bench.WIDTH, bench.HEIGHT
g = bench.scene_cached(bench.N_GAUSSIANS)
scene = E.Scene(); scene.add_group(g)
eng = E.GaussianEngine(W, H); eng.compile(scene, E.Settings(3))
cams = []
for v in range(64):
    cam = E.PerspectiveCamera(W, H); cam.look_at(E.to_cartesian(*bench.ring_camera_params(v)), (0,0,0), (0,0,1)); cams.append(cam)
for v in range(0, 64, 4):
    eng.raster_frame(cams[v]); eng.finish()
hosts = [torch.empty((H, W, 4), dtype=torch.uint8).pin_memory() for _ in range(4)]
hn = [h.numpy() for h in hosts]
evs = [torch.cuda.Event() for _ in range(4)]
def sync(): torch.cuda.synchronize()
for sname, s in (("legacy", None), ("torch_stream", torch.cuda.Stream())):
    st = None if s is None else s.cuda_stream
    ctx = torch.cuda.stream(s) if s is not None else torch.cuda.stream(torch.cuda.default_stream())
    with ctx:
        K = 24
        sync(); t0 = time.perf_counter()
        for k in range(K): eng.raster_frame(cams[k], st)
        sync(); a = (time.perf_counter()-t0)*1e3/K
        sync(); t0 = time.perf_counter()
        for k in range(K):
            eng.raster_frame(cams[k], st); eng.draw_async(hn[k & 3], st)
        sync(); b = (time.perf_counter()-t0)*1e3/K
        sync(); t0 = time.perf_counter()
        for k in range(K):
            eng.raster_frame(cams[k], st); eng.draw_async(hn[k & 1], st); evs[k & 1].record()
            if k > 0: evs[(k-1) & 1].synchronize()
        sync(); c = (time.perf_counter()-t0)*1e3/K
        sync(); t0 = time.perf_counter()
        for k in range(K):
            eng.raster_frame(cams[k], st); eng.draw_async(hn[k % 3], st); evs[k % 3].record()
            if k > 1: evs[(k-2) % 3].synchronize()
        sync(); d = (time.perf_counter()-t0)*1e3/K
        print(f"{sname}: raster only {a:.3f} | +draw_async no waits {b:.3f} | wait for k-1 {c:.3f} | wait for k-2 {d:.3f} ms/frame", flush=True)
