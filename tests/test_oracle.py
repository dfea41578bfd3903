"""CPU tests of the oracle itself: hand-checkable micro-scenes, structural properties, frozen hashes.

The reference has no tests or golden vectors for the shader path, so the pins are created here:
 (1) micro-scenes whose expected numbers are re-derived below in float64 straight from the 3DGS/EWA formulas,
 (2) properties every frame must satisfy, (3) tests/golden/frames.json (SHA-256 of the integer outputs).
"""
import hashlib
import math

import numpy as np
import pytest

from tests.cases import frame_cases, golden_cameras, golden_frames

SH_C0 = 0.28209479177387814


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def identity_camera(w, h):
    """view = identity (x right, y down, z forward); reversed-z 60-degree projection (PerspectiveCamera.cpp:9-23)."""
    fy = np.float32(math.sqrt(3.0))
    fx = np.float32(fy / np.float32(np.float32(w) / np.float32(h)))
    n, f = np.float32(0.01), np.float32(100.0)
    P = np.array([[fx, 0, 0, 0], [0, fy, 0, 0], [0, 0, n / (n - f), n * f / (f - n)], [0, 0, 1, 0]], dtype=np.float32)
    ubo = np.zeros(34, dtype=np.float32)
    ubo[:16] = np.eye(4, dtype=np.float32).reshape(-1)
    ubo[16:32] = P.reshape(-1)  # P * I
    ubo[32], ubo[33] = fx, fy
    return ubo


def gaussian(pos, scale, opacity=1.0, rgb=(1.0, 1.0, 1.0), quat=(0, 0, 0, 1)):
    g = np.zeros(60, dtype=np.float32)
    g[0:3] = pos
    g[3] = opacity
    g[4:8] = quat
    g[8:11] = scale
    g[11] = 1.0
    g[12:15] = [(c - 0.5) / SH_C0 for c in rgb]
    return g


def f32(words):
    return np.asarray(words, dtype=np.uint32).view(np.float32)


# ---------------------------------------------------------------------------------------------------
# (1) micro-scenes
# ---------------------------------------------------------------------------------------------------

def test_single_isotropic_gaussian_on_axis(oracle):
    w = h = 64
    s, z = 0.1, 5.0
    fr = oracle.render(gaussian((0, 0, z), (s, s, s), rgb=(1.0, 0.5, 0.25)).reshape(1, 60), identity_camera(w, h), w, h, 0, want_float=True)
    # float64 re-derivation (EWA, 3DGS eq. 2-3)
    focal = 0.5 * w * math.sqrt(3.0)
    var = (focal * s / z) ** 2 + 0.3
    lam = var + math.sqrt(max(0.1, 0.0))
    radius = math.ceil(3.0 * math.sqrt(lam))
    assert radius == 5
    sp = fr.splats[0]
    assert f32(sp[7:8])[0] == radius
    assert f32(sp[4:6]).tolist() == [31.5, 31.5]          # ndc2pix: ((0+1)*64-1)/2
    assert f32(sp[6:7])[0] == z                           # view depth
    np.testing.assert_allclose(f32(sp[8:11]), [1 / var, 0.0, 1 / var], rtol=1e-5, atol=1e-7)
    # straddles the corner of tiles (1,1),(2,1),(1,2),(2,2): rect [1,3) x [1,3)
    assert fr.tiles[0] == 4 and fr.pairs == 4
    assert [int(k >> 32) for k in fr.keys] == [1 * 4 + 1, 1 * 4 + 2, 2 * 4 + 1, 2 * 4 + 2]
    assert all(int(k & 0xFFFFFFFF) == int(np.float32(z).view(np.uint32)) for k in fr.keys)
    assert fr.ranges.tolist()[5] == [0, 1] and fr.ranges.tolist()[0] == [0, 0]
    alpha = min(0.99, math.exp(-0.5 * (1 / var) * 0.5))   # pixel (31,31): d = (0.5, 0.5)
    np.testing.assert_allclose(fr.rgbf[31, 31], [alpha * 1.0, alpha * 0.5, alpha * 0.25], rtol=1e-5)
    assert fr.rgba[31, 31].tolist() == [round(alpha * 255), round(alpha * 0.5 * 255), round(alpha * 0.25 * 255), 255]
    assert fr.rgba[0, 0].tolist() == [0, 0, 0, 255]        # background black, alpha always 1.0 (blend.slang:104)
    assert fr.rgba[32, 32].tolist() == fr.rgba[31, 31].tolist()  # symmetric about the corner


def test_equal_depth_pairs_keep_emission_order(oracle):
    w = h = 32
    g = np.stack([gaussian((0, 0, 4), (0.3, 0.3, 0.3), opacity=0.5, rgb=(1, 0, 0)),
                  gaussian((0.01, 0, 4), (0.3, 0.3, 0.3), opacity=0.5, rgb=(0, 1, 0)),
                  gaussian((0, 0.01, 4), (0.3, 0.3, 0.3), opacity=0.5, rgb=(0, 0, 1))])
    fr = oracle.render(g, identity_camera(w, h), w, h, 0, want_float=True)
    assert len(set(fr.keys.tolist())) == 4                 # 2x2 tiles, identical depth bits for all three
    for t in range(4):
        s, e = fr.ranges[t]
        assert fr.vals[s:e].tolist() == [0, 1, 2]          # stable: ties stay in ascending Gaussian index
    # front-to-back compositing in that order at the centre pixel
    px = fr.rgbf[16, 16]
    assert px[0] > px[1] > px[2] > 0


def test_culling_cases(oracle):
    w = h = 64
    g = np.stack([
        gaussian((0, 0, -1), (0.1, 0.1, 0.1)),       # behind the camera: viewPos.z <= 0 (splat/common.slang:103)
        gaussian((0, 0, 0.005), (0.001,) * 3),       # nearer than the near plane: clip.z > clip.w (:114)
        gaussian((0, 0, 150), (0.1, 0.1, 0.1)),      # beyond the far plane: clip.z < 0 (:114)
        gaussian((50, 0, 5), (0.1, 0.1, 0.1)),       # outside 1.3x the x frustum (:109)
        gaussian((0, -50, 5), (0.1, 0.1, 0.1)),      # outside 1.3x the y frustum (:110)
        gaussian((0, 0, 5), (0.1, 0.1, 0.1)),        # visible
        gaussian((3.4, 0, 5), (0.01, 0.01, 0.01)),   # inside the 1.3x guard band but its rect misses every tile
    ])
    stale = np.full((7, 12), 0xDEADBEEF, dtype=np.uint32)
    splats = oracle.project(g, identity_camera(w, h), w, h, 0, splats=stale.copy())
    assert splats[:, 3].tolist() == [0, 0, 0, 0, 0, 4, 0]
    assert f32(splats[:, 7]).tolist() == [0, 0, 0, 0, 0, 5, 0]
    culled = [0, 1, 2, 3, 4, 6]
    # culled Gaussians keep stale colour/conic/xy: only radius and tiles are reset (project.slang:34-35)
    assert (splats[culled][:, [0, 1, 2, 4, 5, 6, 8, 9, 10, 11]] == 0xDEADBEEF).all()


def test_sh_band3_plus_quirk(oracle):
    """Coefficient 12 enters as '+ feature' instead of '* feature' (splat/common.slang:69)."""
    w = h = 64
    g = gaussian((0.5, -0.25, 5), (0.1, 0.1, 0.1), rgb=(0.5, 0.5, 0.5))
    sh = g[12:]
    for c, v in enumerate((0.2, -0.1, 0.05)):
        sh[3 + c * 15 + (12 - 1)] = v
    splats = oracle.project(g.reshape(1, 60), identity_camera(w, h), w, h, 3)
    d = np.array([0.5, -0.25, 5.0]) / np.linalg.norm([0.5, -0.25, 5.0])   # camera at the origin
    x, y, z = d
    scalar = 0.3731763325901154 * z * (2 * z * z - 3 * x * x - 3 * y * y)
    expected = [0.0 + scalar + v + 0.5 for v in (0.2, -0.1, 0.05)]
    np.testing.assert_allclose(f32(splats[0, 0:3]), expected, rtol=2e-6)
    # with degree 2 the coefficient is ignored altogether
    splats2 = oracle.project(g.reshape(1, 60), identity_camera(w, h), w, h, 2)
    np.testing.assert_allclose(f32(splats2[0, 0:3]), [0.5, 0.5, 0.5], rtol=1e-6)


def test_anisotropic_rotated_covariance(oracle):
    """cov3D = R S S^T R^T with the quaternion stored (x,y,z,w) and NOT re-normalised (splat/volume.slang:20-41)."""
    w, h = 128, 64
    q = np.array([0.3, -0.2, 0.5, 0.7])          # deliberately not unit length
    s = np.array([0.2, 0.05, 0.1])
    g = gaussian((0.3, 0.2, 4.0), s, quat=q)
    cam = identity_camera(w, h)
    sp = oracle.project(g.reshape(1, 60), cam, w, h, 0)[0]
    x, y, z, ww = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - ww * z), 2 * (x * z + ww * y)],
                  [2 * (x * y + ww * z), 1 - 2 * (x * x + z * z), 2 * (y * z - ww * x)],
                  [2 * (x * z - ww * y), 2 * (y * z + ww * x), 1 - 2 * (x * x + y * y)]])
    cov3 = R @ np.diag(s * s) @ R.T
    focal = 0.5 * np.array([w, h]) * cam[32:34].astype(np.float64)
    vx, vy, vz = 0.3, 0.2, 4.0
    J = np.array([[focal[0] / vz, 0, -focal[0] * vx / vz ** 2], [0, focal[1] / vz, -focal[1] * vy / vz ** 2]])
    cov2 = J @ cov3 @ J.T + 0.3 * np.eye(2)
    conic = np.linalg.inv(cov2)
    np.testing.assert_allclose(f32(sp[8:11]), [conic[0, 0], conic[0, 1], conic[1, 1]], rtol=2e-4)
    lam = np.linalg.eigvalsh(cov2).max()
    assert f32(sp[7:8])[0] == math.ceil(3 * math.sqrt(lam))


def test_pass_count_matches_reference_formula(oracle):
    # GaussianEngine.cpp:351-357 with getHigherMSB (GaussianEngine.h:234-245); values from SURVEY.md §8
    assert oracle.radix_pass_count(1280, 720) == 22
    assert oracle.radix_pass_count(1920, 1080) == 23
    assert oracle.radix_pass_count(3840, 2160) == 24
    for n in [1, 2, 3, 4, 255, 256, 4095, 4096, 8160, 32400, 2 ** 20 + 1]:
        assert oracle.lib().tpdo_higher_msb(n) == n.bit_length()


def test_unorm8_store_saturates_and_rounds(oracle):
    w = h = 16
    # colour far above 1 (huge DC) saturates at 255; the store rounds to nearest
    g = gaussian((0, 0, 2), (1, 1, 1), opacity=0.6, rgb=(5.0, 0.5, 0.003))
    fr = oracle.render(g.reshape(1, 60), identity_camera(w, h), w, h, 0, want_float=True)
    px, pf = fr.rgba[8, 8], fr.rgbf[8, 8]
    assert px[0] == 255 and pf[0] > 1.0
    assert px[1] == int(np.rint(np.float32(pf[1]) * np.float32(255)))
    assert px[3] == 255


# ---------------------------------------------------------------------------------------------------
# (2) properties of whole frames, (3) frozen hashes
# ---------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("name", list(frame_cases()))
def test_frame_properties_and_frozen_hashes(oracle, name):
    gen, cam_name, w, h, deg, model = frame_cases()[name]
    _, cams = golden_cameras()
    g = gen()
    fr = oracle.render(g, cams[cam_name], w, h, deg, models=None if model is None else model.reshape(1, 16), want_evals=True)
    gx, gy = oracle.grid(w, h)
    # scan / duplication
    assert int(fr.tiles.sum()) == fr.pairs == len(fr.keys)
    off = fr.splats[:, 3]
    assert (np.diff(off.astype(np.int64)) == fr.tiles[:-1]).all() and off[0] == 0
    # every pair lies inside its Gaussian's tile rect and carries its depth bits
    tile = (fr.unsorted_keys >> np.uint64(32)).astype(np.int64)
    ty, tx = tile // gx, tile % gx
    sp = fr.splats[fr.unsorted_vals]
    px, py, rad = f32(sp[:, 4]), f32(sp[:, 5]), f32(sp[:, 7])
    assert (tx >= 0).all() and (tx < gx).all() and (ty < gy).all()
    assert ((tx + 1) * 16 > px - rad - 1).all() and (tx * 16 <= px + rad).all()
    assert ((fr.unsorted_keys & np.uint64(0xFFFFFFFF)).astype(np.uint32) == sp[:, 6]).all()
    # sort: non-decreasing keys, stable (ties ascending in value), a permutation of the unsorted pairs
    assert (np.diff(fr.keys.astype(np.uint64)) >= 0).all() if fr.pairs > 1 else True
    ties = fr.keys[1:] == fr.keys[:-1]
    assert (fr.vals[1:][ties] > fr.vals[:-1][ties]).all()
    order = np.lexsort((fr.unsorted_vals, fr.unsorted_keys))
    assert (fr.unsorted_keys[order] == fr.keys).all() and (fr.unsorted_vals[order] == fr.vals).all()
    # ranges partition [0, P) in tile order; empty tiles are (0, 0)
    nonempty = fr.ranges[:, 1] > fr.ranges[:, 0]
    r = fr.ranges[nonempty]
    assert r[0, 0] == 0 and r[-1, 1] == fr.pairs and (r[1:, 0] == r[:-1, 1]).all()
    assert (fr.ranges[~nonempty] == 0).all()
    counts = np.bincount((fr.keys >> np.uint64(32)).astype(np.int64), minlength=gx * gy)
    assert ((fr.ranges[:, 1] - fr.ranges[:, 0]) == counts).all()
    # blend
    assert (fr.rgba[..., 3] == 255).all()
    per_pixel_len = np.repeat(np.repeat(counts.reshape(gy, gx), 16, axis=0), 16, axis=1)[:h, :w]
    assert (fr.evals <= per_pixel_len).all()
    # frozen
    gold = golden_frames()[name]
    assert gold["n"] == g.shape[0] and gold["pairs"] == fr.pairs and gold["visible"] == int((fr.tiles > 0).sum())
    assert gold["gaussians_sha"] == sha(g), "scene generator drifted"
    for key, arr in [("tiles_sha", fr.tiles), ("offsets_sha", fr.splats[:, 3]), ("unsorted_keys_sha", fr.unsorted_keys),
                     ("unsorted_vals_sha", fr.unsorted_vals), ("keys_sha", fr.keys), ("vals_sha", fr.vals), ("ranges_sha", fr.ranges)]:
        assert gold[key] == sha(arr), key
    np.testing.assert_allclose(fr.rgba[..., :3].reshape(-1, 3).mean(axis=0), gold["image_mean"], atol=0.02)


def test_sort_is_stable_on_masked_bits_only(oracle):
    rng = np.random.default_rng(3)
    keys = rng.integers(0, 2 ** 46, size=50000, dtype=np.uint64) | (rng.integers(0, 4, size=50000, dtype=np.uint64) << np.uint64(60))
    keys[::7] = keys[0]
    vals = np.arange(50000, dtype=np.uint32)
    k, v = oracle.sort_pairs(keys, vals, 23)           # bits [0,46): the top bits must NOT take part
    low = keys & np.uint64((1 << 46) - 1)
    order = np.argsort(low, kind="stable")
    assert (k == keys[order]).all() and (v == vals[order]).all()
    k0, v0 = oracle.sort_pairs(keys[:0], vals[:0], 23)
    assert len(k0) == 0 and len(v0) == 0


def test_from_model_field_transforms(oracle):
    """GaussianGeometry.cpp:110-117: sigmoid opacity, (rot_1,rot_2,rot_3,rot_0) normalised, exp scale."""
    rng = np.random.default_rng(5)
    raw = rng.normal(size=(100, 59)).astype(np.float32)
    out = np.zeros((100, 60), dtype=np.float32)
    oracle.lib().tpdo_from_model_fields(raw.ctypes.data, 100, out.ctypes.data)
    np.testing.assert_allclose(out[:, 3], 1 / (1 + np.exp(-raw[:, 10].astype(np.float64))), rtol=1e-6)
    q = raw[:, [4, 5, 6, 3]].astype(np.float64)
    np.testing.assert_allclose(out[:, 4:8], q / np.linalg.norm(q, axis=1, keepdims=True), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(out[:, 8:11], np.exp(raw[:, 7:10].astype(np.float64)), rtol=1e-6)
    assert (out[:, 11] == 1).all() and (out[:, 12:] == raw[:, 11:]).all() and (out[:, :3] == raw[:, :3]).all()
    if oracle.ref_available():  # the quaternion normalisation is pinned to the reference's own math::normalize(vec4)
        q_in = np.ascontiguousarray(raw[:, [4, 5, 6, 3]])
        q_ref = np.zeros_like(q_in)
        for i in range(len(q_in)):
            oracle.ref_lib().tpdref_normalize4(q_in[i].ctypes.data, q_ref[i].ctypes.data)
        assert (q_ref.view(np.uint32) == np.ascontiguousarray(out[:, 4:8]).view(np.uint32)).all()
