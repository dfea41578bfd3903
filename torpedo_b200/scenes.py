"""Deterministic synthetic Gaussian scenes for the BASELINE.json configs.

The reference's only generator, ``GaussianPoint::random`` (volumetric/src/GaussianGeometry.cpp:8-32), is
seeded from ``std::random_device`` and so is not reproducible; these generators are counter-based
(``splitmix64(seed, field, index)`` -> 24-bit uniforms) and use only exactly-rounded IEEE operations
(+ - * / sqrt, ldexp, comparisons) in float64 before the final cast to float32, so every machine
produces bit-identical 240-byte records — which is what makes frozen hashes of integer outputs
(tests/golden/) meaningful.

Record layout (tpd::GaussianPoint, GaussianGeometry.h:10-18 == splat.slang:24-31), 60 floats:
  [0:3] position  [3] opacity  [4:8] quaternion (x,y,z,w)  [8:12] scale (x,y,z,modifier)  [12:60] sh
SH layout: sh[0:3] = DC rgb, then 15 R, 15 G, 15 B coefficients (splat/common.slang:25-31).
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

GAUSSIAN_FLOATS = 60
SH_C0 = np.float32(0.28209479177387814)

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix(z: np.ndarray) -> np.ndarray:
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def uniform(seed: int, field: int, n: int, start: int = 0) -> np.ndarray:
    """n uniforms in [0,1) with 24 random bits each (exact in float32), as float64."""
    with np.errstate(over="ignore"):
        idx = np.arange(start, start + n, dtype=np.uint64)
        key = _mix(np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(field) * np.uint64(0xD1B54A32D192ED03))
        z = _mix(idx * np.uint64(0x9E3779B97F4A7C15) + key)
    return (z >> np.uint64(40)).astype(np.float64) * (1.0 / 16777216.0)


def normalish(seed: int, field: int, n: int) -> np.ndarray:
    """Unit-variance, zero-mean bell (Irwin-Hall of 4 uniforms): exact ops only, no log/cos."""
    s = uniform(seed, field * 4 + 0, n) + uniform(seed, field * 4 + 1, n) + uniform(seed, field * 4 + 2, n) + uniform(seed, field * 4 + 3, n)
    return (s - 2.0) * 1.7320508075688772


_LN2_HI = 0.6931471803691238
_LN2_LO = 1.9082149292705877e-10
_EXP_COEF = [1.0 / 479001600, 1.0 / 39916800, 1.0 / 3628800, 1.0 / 362880, 1.0 / 40320, 1.0 / 5040, 1.0 / 720, 1.0 / 120,
             1.0 / 24, 1.0 / 6, 0.5, 1.0, 1.0]


def exp_det(x: np.ndarray) -> np.ndarray:
    """exp(x) from exactly-rounded float64 operations only (machine-independent); ~1e-13 relative."""
    x = np.asarray(x, dtype=np.float64)
    k = np.floor(x * 1.4426950408889634 + 0.5)
    r = (x - k * _LN2_HI) - k * _LN2_LO
    p = np.full_like(r, _EXP_COEF[0])
    for c in _EXP_COEF[1:]:
        p = p * r + c
    return np.ldexp(p, k.astype(np.int32))


def sigmoid_det(x: np.ndarray) -> np.ndarray:
    return 1.0 / (1.0 + exp_det(-np.asarray(x, dtype=np.float64)))


def rgb2sh(c: np.ndarray) -> np.ndarray:
    """utils::rgb2sh (GaussianGeometry.h:37-45): (c - 0.5f) / C0 in float32."""
    return ((c.astype(np.float32) - np.float32(0.5)) / SH_C0).astype(np.float32)


def _unit_dirs(seed: int, field: int, n: int) -> np.ndarray:
    v = np.stack([normalish(seed, field + k, n) for k in range(3)], axis=1)
    norm = np.sqrt(v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1] + v[:, 2] * v[:, 2])
    norm = np.where(norm < 1e-6, 1.0, norm)
    return v / norm[:, None]


def _unit_quats(seed: int, field: int, n: int) -> np.ndarray:
    q = np.stack([normalish(seed, field + k, n) for k in range(4)], axis=1)
    norm = np.sqrt(q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1] + q[:, 2] * q[:, 2] + q[:, 3] * q[:, 3])
    bad = norm < 1e-6
    q[bad] = np.array([0.0, 0.0, 0.0, 1.0])
    norm = np.where(bad, 1.0, norm)
    return q / norm[:, None]


def hello_gaussian(count: int = 8192, seed: int = 1, radius: float = 10.0, min_scale: float = 0.005, max_scale: float = 0.2,
                   min_opacity: float = 0.1, max_opacity: float = 1.0, with_center: bool = True) -> np.ndarray:
    """demo/HelloGaussian/main.cpp:28-37: GaussianPoint::random(count, 10, {0,0,0}, 0.005, 0.2) in one group plus one
    white Gaussian of scale 2 at the origin (a single entity, which lands last)."""
    n = count + (1 if with_center else 0)
    g = np.zeros((n, GAUSSIAN_FLOATS), dtype=np.float32)
    f = np.float32

    def u(field):  # float32 arithmetic from here on, mirroring tpd::GaussianPoint::random in GaussianGeometry.hpp
        return uniform(seed, field, count).astype(np.float32)

    for k in range(3):
        g[:count, k] = (u(1 + k) * f(2.0) - f(1.0)) * f(radius) + f(0.0)
    g[:count, 3] = f(min_opacity) + u(4) * (f(max_opacity) - f(min_opacity))
    g[:count, 7] = 1.0
    for k in range(3):
        g[:count, 8 + k] = f(min_scale) + u(5 + k) * (f(max_scale) - f(min_scale))
    g[:count, 11] = 1.0
    for k in range(3):
        g[:count, 12 + k] = rgb2sh(u(8 + k))
    if with_center:
        c = g[count]
        c[3] = 1.0
        c[7] = 1.0
        c[8:12] = [2.0, 2.0, 2.0, 1.0]
        c[12:15] = rgb2sh(np.ones(3))
    return g


def garden(n: int, seed: int, log_scale_mean: float = -5.1, log_scale_std: float = 0.7, sh_rest_std: float = 0.05) -> np.ndarray:
    """Generator G of SURVEY.md §8d: 70 % of the points in a ball r = 1.5, 30 % in a shell r in [3, 12]; per-axis
    log-normal-like scales; random rotations; opacity = sigmoid(N(0.5, 2)); DC colour U(0,1); small non-zero higher-order
    SH so that degree 3 and the band-3 quirk (splat/common.slang:69) are exercised."""
    g = np.zeros((n, GAUSSIAN_FLOATS), dtype=np.float32)

    def positions():
        dirs = _unit_dirs(seed, 10, n)
        in_ball = uniform(seed, 1, n) < 0.7
        r_ball = 1.5 * np.maximum(np.maximum(uniform(seed, 2, n), uniform(seed, 3, n)), uniform(seed, 4, n))  # pdf ~ r^2
        r_shell = 3.0 + 9.0 * uniform(seed, 5, n)
        r = np.where(in_ball, r_ball, r_shell)
        g[:, 0:3] = (dirs * r[:, None]).astype(np.float32)

    def opacity():
        g[:, 3] = sigmoid_det(0.5 + 2.0 * normalish(seed, 20, n)).astype(np.float32)

    def quats():
        g[:, 4:8] = _unit_quats(seed, 30, n).astype(np.float32)

    def scale(k):
        g[:, 8 + k] = exp_det(log_scale_mean + log_scale_std * normalish(seed, 40 + k, n)).astype(np.float32)

    def dc(k):
        g[:, 12 + k] = rgb2sh(uniform(seed, 50 + k, n))

    def rest(k):
        g[:, 15 + k] = (sh_rest_std * normalish(seed, 60 + k, n)).astype(np.float32)

    jobs = [positions, opacity, quats] + [lambda k=k: scale(k) for k in range(3)] + [lambda k=k: dc(k) for k in range(3)] + \
           [lambda k=k: rest(k) for k in range(45)]
    g[:, 11] = 1.0
    _run_jobs(jobs, n)
    return g


def _run_jobs(jobs, n):
    """Columns are independent and numpy releases the GIL: fill them in parallel for multi-million-point scenes."""
    if n < 200_000:
        for j in jobs:
            j()
        return
    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as pool:
        for f in [pool.submit(j) for j in jobs]:
            f.result()


def dense_volume(n: int, seed: int = 4, log_scale_mean: float = -4.8, log_scale_std: float = 0.6) -> np.ndarray:
    """VolumeSplatting stand-in: n Gaussians in a unit ball, opacity U(0.5,1): high depth
    complexity, stresses early termination. Rendered with SH degree 2 and the demo's model matrix."""
    g = np.zeros((n, GAUSSIAN_FLOATS), dtype=np.float32)
    dirs = _unit_dirs(seed, 10, n)
    r = np.maximum(np.maximum(uniform(seed, 2, n), uniform(seed, 3, n)), uniform(seed, 4, n))
    g[:, 0:3] = (dirs * r[:, None]).astype(np.float32)
    g[:, 3] = (0.5 + 0.5 * uniform(seed, 20, n)).astype(np.float32)
    g[:, 4:8] = _unit_quats(seed, 30, n).astype(np.float32)
    for k in range(3):
        g[:, 8 + k] = exp_det(log_scale_mean + log_scale_std * normalish(seed, 40 + k, n)).astype(np.float32)
    g[:, 11] = 1.0
    for k in range(3):
        g[:, 12 + k] = rgb2sh(uniform(seed, 50 + k, n))
    for k in range(45):
        g[:, 15 + k] = (0.05 * normalish(seed, 60 + k, n)).astype(np.float32)
    return g


# demo/VolumeSplatting/main.cpp:14-19
VOLUME_TRANSFORM = np.array([1, 0, 0, 0, 0, 1, 0, -1, 0, 0, 1, -0.5, 0, 0, 0, 1], dtype=np.float32)

# Named cameras of the BASELINE configs (eye, center, up); the matrices themselves come from the host
# layer's Camera/PerspectiveCamera (validated bit-exactly against tests/golden/cameras.json).
CAMERAS = {
    "hello": dict(theta=0.785, phi=0.9, radius=8.0, center=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0)),  # OrbitControl.h:21,32-34
    "garden": dict(eye=(2.8, 2.8, 2.6), center=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0)),
    "volume": dict(eye=(-2.0, -1.0, 0.0), center=(0.0, 0.0, 0.0), up=(0.0, -1.0, 0.0)),        # VolumeSplatting/main.cpp:47
}


def ring_angles(views: int = 64):
    """Config 5: eye = to_cartesian(2*pi*k/views, 0.9, 5), target 0, up +z."""
    return [(np.float32(2.0 * np.pi * k / views), np.float32(0.9), np.float32(5.0)) for k in range(views)]
