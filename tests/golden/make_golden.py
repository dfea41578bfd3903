"""Regenerates tests/golden/*.json. Run HERE (the container that has /root/reference), never on the GPU box.

cameras.json   — camera blocks produced by the REFERENCE's own Camera.cpp / PerspectiveCamera.cpp / math headers
                 (oracle/_ref/libtpdref.so, built in place from /root/reference by oracle/Makefile). These pin the
                 host layer (include/torpedo_b200/Camera.hpp) bit-for-bit to real reference code.
frames.json    — SHA-256 of the integer outputs (tile counts, offsets, unsorted/sorted pairs, ranges) of the CPU
                 oracle on seeded synthetic scenes + a coarse image signature. The reference ships no golden vectors
                 for the shader path ("parity unpinned"); these freeze the reviewed oracle so that
                 neither it nor the scene generators can drift unnoticed.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from torpedo_b200 import scenes as S  # noqa: E402
from tests.cases import frame_cases  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def hexf(a):
    return np.ascontiguousarray(a, dtype=np.float32).tobytes().hex()


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def cameras():
    cases = []

    def add(name, w, h, eye, center, up, fov=0.0, near=0.0, far=0.0):
        eye = [float(np.float32(x)) for x in eye]
        center = [float(np.float32(x)) for x in center]
        up = [float(np.float32(x)) for x in up]
        ubo = O.ref_camera_ubo(w, h, eye, center, up, fov, near, far)
        cases.append(dict(name=name, w=w, h=h, eye=eye, center=center, up=up, fov=fov, near=near, far=far, ubo=hexf(ubo)))

    hello_eye = O.ref_to_cartesian(0.785, 0.9, 8.0)
    add("hello_1280x720", 1280, 720, hello_eye, (0, 0, 0), (0, 0, 1))
    for (w, h) in [(1920, 1080), (3840, 2160), (1280, 720), (256, 144), (100, 70), (64, 64), (33, 17)]:
        add(f"garden_{w}x{h}", w, h, S.CAMERAS["garden"]["eye"], (0, 0, 0), (0, 0, 1))
    add("volume_1280x720", 1280, 720, (-2, -1, 0), (0, 0, 0), (0, -1, 0))
    add("volume_256x144", 256, 144, (-2, -1, 0), (0, 0, 0), (0, -1, 0))
    for k, (theta, phi, r) in enumerate(S.ring_angles(64)):
        eye = O.ref_to_cartesian(float(theta), float(phi), float(r))
        add(f"ring64_{k:02d}_1920x1080", 1920, 1080, eye, (0, 0, 0), (0, 0, 1))
    add("fov45_near_far", 1280, 720, (1, 2, 3), (0, 0, 0), (0, 0, 1), 45.0, 0.05, 50.0)
    rng = np.random.default_rng(7)
    for k in range(32):
        w, h = int(rng.integers(16, 4096)), int(rng.integers(16, 2400))
        add(f"random_{k:02d}", w, h, rng.normal(size=3) * rng.uniform(0.1, 20), rng.normal(size=3), rng.normal(size=3))
    tc = [dict(theta=t, phi=p, radius=r, xyz=hexf(O.ref_to_cartesian(t, p, r))) for (t, p, r) in
          [(0.785, 0.9, 8.0), (0.0, 0.9, 5.0), (3.0, 0.2, 1.0), (6.1850104331970215, 0.8999999761581421, 5.0)]]
    with open(os.path.join(HERE, "cameras.json"), "w") as f:
        json.dump(dict(source="oracle/_ref/libtpdref.so (reference Camera.cpp, PerspectiveCamera.cpp, math/*.h)", cases=cases,
                       to_cartesian=tc), f, indent=1)
    return {c["name"]: np.frombuffer(bytes.fromhex(c["ubo"]), dtype=np.float32) for c in cases}


def frames(cams):
    out = {}
    for name, (gen, cam, w, h, deg, model) in frame_cases().items():
        g = gen()
        ubo = cams[cam].copy()
        fr = O.render(g, ubo, w, h, deg, models=None if model is None else model.reshape(1, 16), want_float=True)
        out[name] = dict(
            n=int(g.shape[0]), w=w, h=h, sh_degree=deg, camera=cam, pairs=fr.pairs, visible=int((fr.tiles > 0).sum()),
            gaussians_sha=sha(g), tiles_sha=sha(fr.tiles), offsets_sha=sha(fr.splats[:, 3]), unsorted_keys_sha=sha(fr.unsorted_keys),
            unsorted_vals_sha=sha(fr.unsorted_vals), keys_sha=sha(fr.keys), vals_sha=sha(fr.vals), ranges_sha=sha(fr.ranges),
            image_mean=[float(x) for x in fr.rgba[..., :3].reshape(-1, 3).mean(axis=0)],
            image_8x8=fr.rgba[..., :3].reshape(8, h // 8, 8, w // 8, 3).mean(axis=(1, 3)).round(2).tolist() if h % 8 == 0 and w % 8 == 0 else None,
        )
        print(name, "P", fr.pairs, "visible", out[name]["visible"])
    with open(os.path.join(HERE, "frames.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    if not O.ref_available():
        raise SystemExit("oracle/_ref is missing: this script needs /root/reference")
    frames(cameras())
