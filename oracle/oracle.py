"""ctypes front-end of oracle/libtpd_oracle.so — TEST INFRASTRUCTURE ONLY.

The product (torpedo_b200/) never imports this module. Every function maps 1:1 onto a function of
oracle/tpd_oracle.h, which cites the reference shader lines it restates.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtpd_oracle.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libtpdref.so")

GAUSSIAN_FLOATS = 60
SPLAT_WORDS = 12
TILE = 16


def build(force: bool = False) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference is present). Building is not using."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "tpd_oracle.c")):
        subprocess.run(["make", "-C", _HERE, "libtpd_oracle.so"], check=True, capture_output=True)
    ref_stale = os.path.exists(_REF_PATH) and os.path.getmtime(_REF_PATH) < os.path.getmtime(os.path.join(_HERE, "ref_shim.cpp"))
    if os.path.isdir("/root/reference/torpedo") and (force or ref_stale or not os.path.exists(_REF_PATH)):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        u32, vp = C.c_uint32, C.c_void_p
        _lib.tpdo_higher_msb.restype = u32
        _lib.tpdo_higher_msb.argtypes = [u32]
        _lib.tpdo_radix_pass_count.restype = u32
        _lib.tpdo_radix_pass_count.argtypes = [u32, u32]
        _lib.tpdo_project.restype = None
        _lib.tpdo_project.argtypes = [vp, u32, vp, vp, u32, vp, u32, u32, u32, vp]
        _lib.tpdo_prefix.restype = u32
        _lib.tpdo_prefix.argtypes = [vp, u32]
        _lib.tpdo_keygen.restype = None
        _lib.tpdo_keygen.argtypes = [vp, u32, u32, u32, vp, vp]
        _lib.tpdo_sort.restype = None
        _lib.tpdo_sort.argtypes = [vp, vp, u32, vp, vp, u32]
        _lib.tpdo_range.restype = None
        _lib.tpdo_range.argtypes = [vp, u32, vp, u32]
        _lib.tpdo_blend.restype = None
        _lib.tpdo_blend.argtypes = [vp, vp, vp, u32, u32, vp, vp, vp]
        _lib.tpdo_frame.restype = u32
        _lib.tpdo_frame.argtypes = [vp, u32, vp, vp, u32, vp, u32, u32, u32, vp, vp, vp, vp, vp, u32, vp, vp, vp]
        _lib.tpdo_from_model_fields.restype = None
        _lib.tpdo_from_model_fields.argtypes = [vp, u32, vp]
        _lib.tpdo_num_threads.restype = C.c_int
        _lib.tpdo_set_num_threads.restype = None
        _lib.tpdo_set_num_threads.argtypes = [C.c_int]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def grid(width: int, height: int) -> tuple[int, int]:
    return (width + TILE - 1) // TILE, (height + TILE - 1) // TILE


def radix_pass_count(width: int, height: int) -> int:
    return int(lib().tpdo_radix_pass_count(width, height))


def num_threads() -> int:
    return int(lib().tpdo_num_threads())


def use_all_cores() -> int:
    """Make the oracle use every core this process may run on (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    lib().tpdo_set_num_threads(cores)
    return num_threads()


@dataclass
class Frame:
    splats: np.ndarray          # (n, 12) uint32, reference Splat layout, tiles = exclusive offset
    tiles: np.ndarray           # (n,) uint32 per-Gaussian tile counts (before the scan)
    pairs: int                  # P
    unsorted_keys: np.ndarray
    unsorted_vals: np.ndarray
    keys: np.ndarray            # sorted
    vals: np.ndarray
    ranges: np.ndarray          # (tiles, 2) uint32
    rgba: np.ndarray | None     # (h, w, 4) uint8
    rgbf: np.ndarray | None     # (h, w, 3) float32
    evals: np.ndarray | None    # (h, w) uint32


def _check_inputs(gaussians, entity_idx, models):
    g = np.ascontiguousarray(gaussians, dtype=np.float32).reshape(-1, GAUSSIAN_FLOATS)
    e = None if entity_idx is None else np.ascontiguousarray(entity_idx, dtype=np.uint32)
    m = None if models is None else np.ascontiguousarray(models, dtype=np.float32).reshape(-1, 16)
    return g, e, m


def project(gaussians, camera34, width, height, sh_degree=3, entity_idx=None, models=None, splats=None):
    g, e, m = _check_inputs(gaussians, entity_idx, models)
    n = g.shape[0]
    cam = np.ascontiguousarray(camera34, dtype=np.float32).reshape(34)
    if splats is None:
        splats = np.zeros((n, SPLAT_WORDS), dtype=np.uint32)
    lib().tpdo_project(_p(g), n, _p(e), _p(m), 0 if m is None else m.shape[0], _p(cam), width, height, sh_degree, _p(splats))
    return splats


def render(gaussians, camera34, width, height, sh_degree=3, entity_idx=None, models=None, blend=True, want_float=False,
           want_evals=False) -> Frame:
    """The whole reference frame, stage by stage (GaussianEngine.cpp:621-712), keeping every intermediate."""
    L = lib()
    splats = project(gaussians, camera34, width, height, sh_degree, entity_idx, models)
    n = splats.shape[0]
    tiles = splats[:, 3].copy()
    pairs = int(L.tpdo_prefix(_p(splats), n))
    keys = np.zeros(max(pairs, 1), dtype=np.uint64)
    vals = np.zeros(max(pairs, 1), dtype=np.uint32)
    L.tpdo_keygen(_p(splats), n, width, height, _p(keys), _p(vals))
    keys, vals = keys[:pairs], vals[:pairs]
    uk, uv = keys.copy(), vals.copy()
    tk, tv = np.empty_like(keys), np.empty_like(vals)
    L.tpdo_sort(_p(keys), _p(vals), pairs, _p(tk), _p(tv), radix_pass_count(width, height))
    gx, gy = grid(width, height)
    ranges = np.zeros((gx * gy, 2), dtype=np.uint32)
    L.tpdo_range(_p(keys), pairs, _p(ranges), gx * gy)
    rgba = rgbf = evals = None
    if blend:
        rgba = np.zeros((height, width, 4), dtype=np.uint8)
        rgbf = np.zeros((height, width, 3), dtype=np.float32) if want_float else None
        evals = np.zeros((height, width), dtype=np.uint32) if want_evals else None
        L.tpdo_blend(_p(splats), _p(vals), _p(ranges), width, height, _p(rgba), _p(rgbf), _p(evals))
    return Frame(splats, tiles, pairs, uk, uv, keys, vals, ranges, rgba, rgbf, evals)


def sort_pairs(keys, vals, pass_count):
    keys = np.ascontiguousarray(keys, dtype=np.uint64).copy()
    vals = np.ascontiguousarray(vals, dtype=np.uint32).copy()
    tk, tv = np.empty_like(keys), np.empty_like(vals)
    lib().tpdo_sort(_p(keys), _p(vals), keys.shape[0], _p(tk), _p(tv), pass_count)
    return keys, vals


class FrameTimer:
    """Re-usable buffers around tpdo_frame for the CPU baseline (bench.py only)."""

    def __init__(self, gaussians, width, height, sh_degree=3, entity_idx=None, models=None, capacity=None):
        self.g, self.e, self.m = _check_inputs(gaussians, entity_idx, models)
        self.n = self.g.shape[0]
        self.w, self.h, self.deg = width, height, sh_degree
        self.capacity = capacity or 4 * self.n + 1024
        self._alloc()

    def _alloc(self):
        gx, gy = grid(self.w, self.h)
        self.splats = np.zeros((self.n, SPLAT_WORDS), dtype=np.uint32)
        self.keys = np.zeros(self.capacity, dtype=np.uint64)
        self.vals = np.zeros(self.capacity, dtype=np.uint32)
        self.tk = np.zeros(self.capacity, dtype=np.uint64)
        self.tv = np.zeros(self.capacity, dtype=np.uint32)
        self.ranges = np.zeros((gx * gy, 2), dtype=np.uint32)
        self.rgba = np.zeros((self.h, self.w, 4), dtype=np.uint8)

    def frame(self, camera34):
        cam = np.ascontiguousarray(camera34, dtype=np.float32).reshape(34)
        ms = (C.c_double * 7)()
        while True:
            p = lib().tpdo_frame(_p(self.g), self.n, _p(self.e), _p(self.m), 0 if self.m is None else self.m.shape[0], _p(cam),
                                 self.w, self.h, self.deg, _p(self.splats), _p(self.keys), _p(self.vals), _p(self.tk), _p(self.tv),
                                 self.capacity, _p(self.ranges), _p(self.rgba), ms)
            if p != 0xFFFFFFFF:
                break
            self.capacity *= 2
            self._alloc()
        names = ["project", "prefix", "keygen", "sort", "range", "blend", "total"]
        return int(p), dict(zip(names, [float(x) for x in ms]))


# ---- oracle/_ref: the reference's own camera code (only where /root/reference was present at build time)

_ref = None


def ref_available() -> bool:
    return os.path.exists(_REF_PATH)


def ref_lib() -> C.CDLL:
    global _ref
    if _ref is None:
        build()
        _ref = C.CDLL(_REF_PATH)
        f3 = C.c_float * 3
        _ref.tpdref_camera_ubo.restype = None
        _ref.tpdref_camera_ubo.argtypes = [C.c_uint32, C.c_uint32, f3, f3, f3, C.c_float, C.c_float, C.c_float, C.c_void_p]
        _ref.tpdref_to_cartesian.restype = None
        _ref.tpdref_to_cartesian.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p]
        _ref.tpdref_rgb2sh.restype = None
        _ref.tpdref_rgb2sh.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p]
        _ref.tpdref_sizeof_gaussian_point.restype = C.c_uint32
        _ref.tpdref_normalize4.restype = None
        _ref.tpdref_normalize4.argtypes = [C.c_void_p, C.c_void_p]
        _ref.tpdref_from_model.restype = C.c_int64
        _ref.tpdref_from_model.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64]
    return _ref


def ref_from_model(ply_path: str) -> np.ndarray:
    """The reference's own GaussianPoint::fromModel (GaussianGeometry.cpp:59-127 + miniply.cpp, compiled in place)."""
    n = ref_lib().tpdref_from_model(ply_path.encode(), None, 0)
    if n < 0:
        raise RuntimeError(f"the reference's fromModel threw on {ply_path}")
    out = np.zeros((n, GAUSSIAN_FLOATS), dtype=np.float32)
    ref_lib().tpdref_from_model(ply_path.encode(), _p(out), n)
    return out


def ref_camera_ubo(width, height, eye, center, up, fov_deg=0.0, near=0.0, far=0.0) -> np.ndarray:
    f3 = C.c_float * 3
    out = np.zeros(34, dtype=np.float32)
    ref_lib().tpdref_camera_ubo(width, height, f3(*eye), f3(*center), f3(*up), fov_deg, near, far, _p(out))
    return out


def ref_to_cartesian(theta, phi, radius) -> np.ndarray:
    out = np.zeros(3, dtype=np.float32)
    ref_lib().tpdref_to_cartesian(theta, phi, radius, _p(out))
    return out
