"""Shared definitions of the seeded parity cases (used by tests/ and tests/golden/make_golden.py)."""
import json
import os

import numpy as np

from torpedo_b200 import scenes as S

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def frame_cases():
    """name -> (scene thunk, golden camera name, width, height, SH degree, model matrix or None)"""
    return {
        "hello_8193_sh0_256x144": (lambda: S.hello_gaussian(8192, seed=1), "hello_1280x720", 256, 144, 0, None),
        "garden_20k_sh3_256x144": (lambda: S.garden(20000, seed=2, log_scale_mean=-3.6), "garden_256x144", 256, 144, 3, None),
        "garden_20k_sh1_100x70": (lambda: S.garden(20000, seed=2, log_scale_mean=-3.6), "garden_100x70", 100, 70, 1, None),
        "garden_5k_sh2_33x17": (lambda: S.garden(5000, seed=6, log_scale_mean=-3.0), "garden_33x17", 33, 17, 2, None),
        "volume_30k_sh2_256x144": (lambda: S.dense_volume(30000, seed=4, log_scale_mean=-3.2), "volume_256x144", 256, 144, 2,
                                   S.VOLUME_TRANSFORM),
    }


def golden_cameras():
    with open(os.path.join(GOLDEN, "cameras.json")) as f:
        data = json.load(f)
    cams = {c["name"]: np.frombuffer(bytes.fromhex(c["ubo"]), dtype=np.float32).copy() for c in data["cases"]}
    return data, cams


def golden_frames():
    with open(os.path.join(GOLDEN, "frames.json")) as f:
        return json.load(f)
