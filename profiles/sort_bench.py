"""Standalone timing of the onesweep sort on the headline frame's OWN pairs (tile|depth keys as emitted by the
duplication stage), verified against torch's stable sort. Usage: python profiles/sort_bench.py [repeats] [n_gaussians]"""
import ctypes as C
import json
import os
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from torpedo_b200 import engine as E  # noqa: E402
from torpedo_b200._lib import check, tpdcu  # noqa: E402

repeats = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n = int(sys.argv[2]) if len(sys.argv) > 2 else bench.N_GAUSSIANS
w, h = bench.WIDTH, bench.HEIGHT
g = bench.scene_cached(n)
scene = E.Scene()
scene.add_group(g)
eng = E.GaussianEngine(w, h)
eng.compile(scene, E.Settings(3))
cam = E.PerspectiveCamera(w, h)
cam.look_at(E.to_cartesian(*bench.ring_camera_params(0)), (0, 0, 0), (0, 0, 1))
eng.raster_frame(cam)
eng.finish()
eng.raster_frame(cam)
pairs = eng.finish()
uk, uv = eng.read_unsorted()
sk, sv = eng.read_sorted()
keys0 = torch.from_numpy(uk.view(np.int64)).cuda()
vals0 = torch.from_numpy(uv.view(np.int32)).cuda()
order = torch.sort(keys0, stable=True).indices
assert torch.equal(keys0[order].cpu(), torch.from_numpy(sk.view(np.int64))) and torch.equal(vals0[order].cpu(), torch.from_numpy(sv.view(np.int32)))
lib = tpdcu()
end_bit = 32 + (((w + 15) // 16) * ((h + 15) // 16) - 1).bit_length()
times = []
for r in range(repeats + 2):
    k, v = keys0.clone(), vals0.clone()
    torch.cuda.synchronize()
    check(lib.tpdcu_sort_pairs_device(eng.ctx, k.data_ptr(), v.data_ptr(), pairs, end_bit, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ms, passes = C.c_float(0), C.c_uint32(0)
    check(lib.tpdcu_sort_last_ms(eng.ctx, C.byref(ms), C.byref(passes)))
    if r >= 2:
        times.append(ms.value)
assert torch.equal(k, keys0[order]) and torch.equal(v, vals0[order])
ms = statistics.median(times)
peak = bench.measured_peak_hbm()[0]
model_bytes = pairs * (passes.value * 24 + 8)
print(json.dumps({"pairs": pairs, "end_bit": end_bit, "passes": passes.value, "sort_ms_median": ms, "sort_ms_min": min(times),
                  "gkeys_per_s": pairs / ms / 1e6, "model_gbs": model_bytes / ms / 1e6, "frac_of_hbm_peak": model_bytes / ms / 1e6 / peak,
                  "verified": True}))
