"""Experiment: two frames in flight (two engines with private buffers on two streams) vs one. Throughput only."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from torpedo_b200 import engine as E

n = int(sys.argv[1]) if len(sys.argv) > 1 else bench.N_GAUSSIANS
g = bench.scene_cached(n)
engs, streams = [], []
for k in range(2):
    scene = E.Scene(); scene.add_group(g)
    e = E.GaussianEngine(bench.WIDTH, bench.HEIGHT); e.compile(scene, E.Settings(3))
    engs.append(e); streams.append(torch.cuda.Stream())
ubos = []
for v in range(64):
    cam = E.PerspectiveCamera(bench.WIDTH, bench.HEIGHT)
    cam.look_at(E.to_cartesian(*bench.ring_camera_params(v)), (0, 0, 0), (0, 0, 1)); ubos.append(cam.pack())
for e in engs:
    for v in range(0, 64, 4):
        e.raster_ubo(ubos[v], 3, None); e.finish()
torch.cuda.synchronize()
def run(k_frames, two):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for f in range(k_frames):
        i = f & 1 if two else 0
        engs[i].raster_ubo(ubos[f % 64], 3, streams[i].cuda_stream)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / k_frames
for rep in range(3):
    print("one in flight: %.4f ms/frame   two in flight: %.4f ms/frame" % (run(40, False), run(40, True)), flush=True)
