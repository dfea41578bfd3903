"""A/B driver for kernel variants built by profiles/build_variant.sh: checks the frame's sorted pairs against torch's stable
sort of the unsorted ones, then prints per-stage CUDA-event medians and the pipelined ms/frame.
python profiles/variant_bench.py <label> [reps] [n] [w] [h]"""
import json
import os
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from torpedo_b200 import engine as E  # noqa: E402

label = sys.argv[1] if len(sys.argv) > 1 else "?"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
n = int(sys.argv[3]) if len(sys.argv) > 3 else bench.N_GAUSSIANS
w = int(sys.argv[4]) if len(sys.argv) > 4 else bench.WIDTH
h = int(sys.argv[5]) if len(sys.argv) > 5 else bench.HEIGHT
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
# correctness of both sort levels on the last view
eng.raster_frame(cams[0])
pairs = eng.finish()
uk, uv = eng.read_unsorted()
sk, sv = eng.read_sorted()
keys0 = torch.from_numpy(uk.view(np.int64)).cuda()
order = torch.sort(keys0, stable=True).indices.cpu().numpy()
ok = bool((uk[order] == sk).all() and (uv[order] == sv).all())
ranges = eng.read_ranges()
tiles = (sk >> np.uint64(32)).astype(np.int64)
starts = np.searchsorted(tiles, np.arange(ranges.shape[0]), "left")
ends = np.searchsorted(tiles, np.arange(ranges.shape[0]), "right")
nonempty = ends > starts
ok_ranges = bool((ranges[nonempty, 0] == starts[nonempty]).all() and (ranges[nonempty, 1] == ends[nonempty]).all() and (ranges[~nonempty] == 0).all())
eng.enable_stage_timing(True)
runs = []
for r in range(reps):
    eng.raster_frame(cams[r % 8])
    runs.append(eng.stage_times_ms())
med = {k: round(statistics.median(x[k] for x in runs), 4) for k in runs[0]}
eng.enable_stage_timing(False)
# pipelined: K frames back to back
K = 48
for r in range(8):
    eng.raster_frame(cams[r % 8])
eng.finish()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for r in range(K):
    eng.raster_frame(cams[r % 8])
e1.record()
eng.finish()
torch.cuda.synchronize()
print(json.dumps({"label": label, "sorted_ok": ok, "ranges_ok": ok_ranges, "pairs": pairs, "pipelined_ms": round(e0.elapsed_time(e1) / K, 4), "stages_ms": med}))
