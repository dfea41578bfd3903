"""Per-stage CUDA-event times of the headline frame (median of `reps` frames). python profiles/stage_bench.py [reps] [n] [w] [h]"""
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from torpedo_b200 import engine as E  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = int(sys.argv[2]) if len(sys.argv) > 2 else bench.N_GAUSSIANS
w = int(sys.argv[3]) if len(sys.argv) > 3 else bench.WIDTH
h = int(sys.argv[4]) if len(sys.argv) > 4 else bench.HEIGHT
g = bench.scene_cached(n)
scene = E.Scene()
scene.add_group(g)
eng = E.GaussianEngine(w, h)
eng.compile(scene, E.Settings(3))
cams = []
for v in range(8):
    cam = E.PerspectiveCamera(w, h)
    cam.look_at(E.to_cartesian(*bench.ring_camera_params(v * 8)), (0, 0, 0), (0, 0, 1))
    cams.append(cam)
for cam in cams:
    eng.raster_frame(cam)
    eng.finish()
eng.enable_stage_timing(True)
runs = []
for r in range(reps):
    eng.raster_frame(cams[r % 8])
    runs.append(eng.stage_times_ms())
med = {k: round(statistics.median(x[k] for x in runs), 4) for k in runs[0]}
print(json.dumps({"n": n, "w": w, "h": h, "pairs": eng.counts()[0], "stages_ms": med}))
