"""Freeze SHA-256 hashes of the oracle's integer outputs for the FULL-SIZE BASELINE configs (tests/full_size_cases.py) into
tests/golden/full_size.json. The camera blocks come from the reference's own Camera.cpp / PerspectiveCamera.cpp
(oracle/_ref, compiled in place), so the script needs /root/reference; the fixture it writes does not.
    python tests/golden/make_full_size_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests.full_size_cases import cases  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


out = {}
for name, c in cases().items():
    g = c["gen"]()
    eye = c["eye"]
    if eye[0] == "cart":
        eye = [float(x) for x in O.ref_to_cartesian(*eye[1:])]
    ubo = O.ref_camera_ubo(c["w"], c["h"], eye, c.get("center", (0.0, 0.0, 0.0)), c.get("up", (0.0, 0.0, 1.0)))
    model = c.get("model")
    ref = O.render(g, ubo, c["w"], c["h"], c["deg"], models=None if model is None else np.asarray(model, np.float32).reshape(1, 16))
    lens = ref.ranges[:, 1].astype(np.int64) - ref.ranges[:, 0]
    out[name] = {"n": int(g.shape[0]), "ubo": ubo.tobytes().hex(), "pairs": int(ref.pairs), "visible": int((ref.tiles > 0).sum()),
                 "keys_sha": sha(ref.keys), "vals_sha": sha(ref.vals), "ranges_sha": sha(ref.ranges), "offsets_sha": sha(ref.splats[:, 3]),
                 "image_mean": [float(x) for x in ref.rgba[..., :3].reshape(-1, 3).mean(axis=0)],
                 "tile_list": {"mean": float(lens.mean()), "max": int(lens.max())}}
    print(name, out[name]["pairs"], out[name]["visible"], flush=True)
with open(os.path.join(ROOT, "tests", "golden", "full_size.json"), "w") as f:
    json.dump(out, f, indent=1)
