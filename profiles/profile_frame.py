"""Small driver for ncu captures: the headline scene (6 M Gaussians, 1080p, SH3), 3 warm-up frames, then `frames` frames.
Usage (under gpurun):  ncu ... python profiles/profile_frame.py [frames] [n_gaussians] [width] [height]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from torpedo_b200 import engine as E  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else bench.N_GAUSSIANS
w = int(sys.argv[3]) if len(sys.argv) > 3 else bench.WIDTH
h = int(sys.argv[4]) if len(sys.argv) > 4 else bench.HEIGHT
g = bench.scene_cached(n)
scene = E.Scene()
scene.add_group(g)
eng = E.GaussianEngine(w, h)
eng.compile(scene, E.Settings(3))
cam = E.PerspectiveCamera(w, h)
cam.look_at(E.to_cartesian(*bench.ring_camera_params(0)), (0, 0, 0), (0, 0, 1))
for _ in range(3):
    eng.raster_frame(cam)
    eng.finish()
for _ in range(frames):
    eng.raster_frame(cam)
print("pairs", eng.finish(), "visible", eng.counts()[1])
