"""The BASELINE.json configs at FULL size, shared by tests/test_full_size_gpu.py and
tests/golden/make_full_size_golden.py. Scene generation is deterministic (counter-based, bit-identical on every machine)."""
import numpy as np

from torpedo_b200 import scenes as S

LOG_SCALE_MEAN = -5.2   # bench.py's calibration: P/N = 2.64 at 1080p


def _garden6m():
    import bench
    return bench.scene_cached(6_000_000)


def cases():
    """name -> dict(scene thunk, width, height, sh degree, eye, center, up, model)"""
    hello_eye = ("cart", 0.785, 0.9, 8.0)
    garden_eye = (2.8, 2.8, 2.6)
    return {
        "1a_hello_8193_sh0_720p": dict(gen=lambda: S.hello_gaussian(8192, seed=1), w=1280, h=720, deg=0, eye=hello_eye),
        "1b_hello_100k_sh3_720p": dict(gen=lambda: S.hello_gaussian(100000, seed=1), w=1280, h=720, deg=3, eye=hello_eye),
        "2_garden_1m_sh3_1080p": dict(gen=lambda: S.garden(1_000_000, 2, log_scale_mean=LOG_SCALE_MEAN), w=1920, h=1080, deg=3, eye=garden_eye),
        "3a_garden_6m_sh3_1080p": dict(gen=_garden6m, w=1920, h=1080, deg=3, eye=garden_eye),
        "3b_garden_6m_sh3_2160p": dict(gen=_garden6m, w=3840, h=2160, deg=3, eye=garden_eye),
        "4_volume_2m_sh2_720p": dict(gen=lambda: S.dense_volume(2_000_000, seed=4), w=1280, h=720, deg=2, eye=(-2.0, -1.0, 0.0),
                                     up=(0.0, -1.0, 0.0), model=S.VOLUME_TRANSFORM),
        "5_ring_3m_sh3_1080p_view21": dict(gen=lambda: S.garden(3_000_000, 5, log_scale_mean=LOG_SCALE_MEAN), w=1920, h=1080, deg=3,
                                           eye=("cart", float(np.float32(2.0 * np.pi * 21 / 64)), 0.9, 5.0)),
    }
