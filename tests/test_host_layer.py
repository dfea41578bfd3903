"""CPU tests of the host layer (include/torpedo_b200/*.hpp through lib/libtpdhost.so): camera blocks are bit-identical
to the ones the REFERENCE's own camera/math sources produce (tests/golden/cameras.json, generated from oracle/_ref)."""
import numpy as np
import pytest

from tests.cases import golden_cameras


@pytest.fixture(scope="module")
def E(built_libs):
    from torpedo_b200 import engine
    engine.tpdhost()
    return engine


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_camera_blocks_match_reference_bit_for_bit(E):
    data, cams = golden_cameras()
    assert len(data["cases"]) >= 100
    for c in data["cases"]:
        cam = E.PerspectiveCamera(c["w"], c["h"])
        if c["near"] > 0:
            cam.set_near(c["near"])
        if c["far"] > 0:
            cam.set_far(c["far"])
        if c["fov"] > 0 or c["near"] > 0 or c["far"] > 0:
            cam.set_vertical_fov(c["fov"] if c["fov"] > 0 else 60.0)
        cam.look_at(c["eye"], c["center"], c["up"])
        assert (bits(cam.pack()) == bits(cams[c["name"]])).all(), c["name"]


def test_to_cartesian_matches_reference(E):
    data, _ = golden_cameras()
    for t in data["to_cartesian"]:
        got = E.to_cartesian(t["theta"], t["phi"], t["radius"])
        assert got.tobytes().hex() == t["xyz"]


def test_golden_cameras_still_match_oracle_ref(oracle):
    """Only where /root/reference was available to build oracle/_ref (this container, not the GPU box)."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    data, cams = golden_cameras()
    for c in data["cases"][:40]:
        ref = oracle.ref_camera_ubo(c["w"], c["h"], c["eye"], c["center"], c["up"], c["fov"], c["near"], c["far"])
        assert (bits(ref) == bits(cams[c["name"]])).all(), c["name"]
    assert oracle.ref_lib().tpdref_sizeof_gaussian_point() == 240


def test_camera_view_matrix_conventions(E):
    """z forward, x right, y down; translation -dot(axis, eye) (rendering/src/Camera.cpp:3-16)."""
    cam = E.PerspectiveCamera(1280, 720)
    cam.look_at((0, 0, -5), (0, 0, 0), (0, -1, 0))
    V = cam.pack()[:16].reshape(4, 4)
    np.testing.assert_allclose(V[2, :3], [0, 0, 1], atol=1e-7)      # forward = +z
    np.testing.assert_allclose(V[:3, 3], [0, 0, 5], atol=1e-6)      # eye lands at the origin of view space
    np.testing.assert_allclose(V[:3, :3] @ V[:3, :3].T, np.eye(3), atol=1e-6)
    assert V[3].tolist() == [0, 0, 0, 1]
    P = cam.pack()[16:32].reshape(4, 4) @ np.linalg.inv(V)
    np.testing.assert_allclose(P[3], [0, 0, 1, 0], atol=1e-6)       # clip.w = view z
    assert abs(cam.pack()[33] - np.sqrt(3)) < 1e-6 and abs(cam.pack()[32] - np.sqrt(3) * 720 / 1280) < 1e-6
    cam.on_image_size_change(1000, 1000)
    assert abs(cam.pack()[32] - cam.pack()[33]) < 1e-7


def test_scene_layout_groups_first_then_singles(E):
    s = E.Scene()
    a = np.zeros((3, 60), dtype=np.float32)
    b = np.ones((2, 60), dtype=np.float32)
    e_single = s.add(np.full(60, 7, dtype=np.float32))
    e_a = s.add_group(a)
    e_b = s.add_group(b)
    assert s.count_all() == 6
    assert len({e_single, e_a, e_b}) == 3
    assert E.tpdhost().tpdh_sizeof_gaussian_point() == 240
    with pytest.raises(E.TpdError):
        s.add(np.zeros((2, 60), dtype=np.float32))
    with pytest.raises(E.TpdError):
        s.add_group(np.zeros(61, dtype=np.float32))


def test_rgb2sh_and_random_points(E):
    out = np.zeros(48, dtype=np.float32)
    E.tpdhost().tpdh_rgb2sh(1.0, 0.5, 0.0, out.ctypes.data)
    c0 = np.float32(0.28209479177387814)
    assert out[0] == np.float32(0.5) / c0 and out[1] == 0 and out[2] == np.float32(-0.5) / c0 and (out[3:] == 0).all()
    pts = np.zeros((1000, 60), dtype=np.float32)
    assert E.tpdhost().tpdh_random_points(1000, 10.0, 0.005, 0.2, 0.1, 1.0, 1, pts.ctypes.data) == 0
    assert np.abs(pts[:, :3]).max() <= 10 and (pts[:, 3] >= 0.1).all() and (pts[:, 3] <= 1.0).all()
    assert (pts[:, 4:8] == [0, 0, 0, 1]).all() and (pts[:, 8:11] >= 0.005).all() and (pts[:, 8:11] <= 0.2).all()
    assert (pts[:, 15:] == 0).all()
    # same recipe as the numpy generator used for the benchmark scenes: bit-identical clouds
    from torpedo_b200 import scenes
    assert (scenes.hello_gaussian(1000, seed=1, with_center=False).view(np.uint32) == pts.view(np.uint32)).all()


def _write_ply(path, raw, rest_count, binary=True, extra_first=False):
    """A 3DGS-style PLY: x y z nx ny nz f_dc_0..2 f_rest_* opacity scale_0..2 rot_0..3 (the order 3DGS trainers write)."""
    names = ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"] + [f"f_rest_{k}" for k in range(rest_count)] + \
            ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    n = raw["x"].shape[0]
    cols = [raw.get(name, np.zeros(n, np.float32)).astype(np.float32) for name in names]
    header = "ply\nformat %s 1.0\ncomment test\nelement vertex %d\n" % ("binary_little_endian" if binary else "ascii", n)
    header += "".join(f"property float {name}\n" for name in names) + "end_header\n"
    with open(path, "wb") as f:
        f.write(header.encode())
        table = np.stack(cols, axis=1)
        if binary:
            f.write(table.astype("<f4").tobytes())
        else:
            for row in table:
                f.write((" ".join(repr(float(v)) for v in row) + "\n").encode())


@pytest.mark.parametrize("rest_count,binary", [(45, True), (24, True), (0, True), (45, False)])
def test_from_model_matches_reference_field_transforms(E, oracle, tmp_path, rest_count, binary):
    """fromModel (GaussianGeometry.cpp:59-127): PLY parsing + sigmoid / quaternion reorder+normalise / exp, bit for bit
    against the oracle's restatement of :110-117."""
    rng = np.random.default_rng(rest_count + 1)
    n = 257
    raw = {name: rng.normal(size=n).astype(np.float32) for name in
           ["x", "y", "z", "f_dc_0", "f_dc_1", "f_dc_2", "opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]}
    for k in range(rest_count):
        raw[f"f_rest_{k}"] = rng.normal(size=n).astype(np.float32)
    path = str(tmp_path / "cloud.ply")
    _write_ply(path, raw, rest_count, binary)
    got = E.from_model(path)
    assert got.shape == (n, 60)
    raw59 = np.zeros((n, 59), dtype=np.float32)
    raw59[:, 0:3] = np.stack([raw["x"], raw["y"], raw["z"]], axis=1)
    raw59[:, 3:7] = np.stack([raw[f"rot_{k}"] for k in range(4)], axis=1)
    raw59[:, 7:10] = np.stack([raw[f"scale_{k}"] for k in range(3)], axis=1)
    raw59[:, 10] = raw["opacity"]
    raw59[:, 11:14] = np.stack([raw[f"f_dc_{k}"] for k in range(3)], axis=1)
    for k in range(rest_count):
        raw59[:, 14 + k] = raw[f"f_rest_{k}"]
    want = np.zeros((n, 60), dtype=np.float32)
    oracle.lib().tpdo_from_model_fields(raw59.ctypes.data, n, want.ctypes.data)
    if binary:
        assert (got.view(np.uint32) == want.view(np.uint32)).all()
    else:
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-7)
    if oracle.ref_available():  # this container: the reference's own fromModel + miniply, compiled in place (oracle/_ref)
        ref = oracle.ref_from_model(path)
        assert (got.view(np.uint32) == ref.view(np.uint32)).all()
    with pytest.raises(E.TpdError):
        E.from_model(str(tmp_path / "missing.ply"))


PLY_FIXTURES = ["sh3_binary", "sh2_binary", "sh0_binary", "sh3_ascii", "sh1_mixed_types", "sh2_shuffled"]


@pytest.mark.parametrize("name", PLY_FIXTURES)
def test_from_model_matches_the_reference_on_committed_ply_fixtures(E, name):
    """tests/golden/ply/<name>.ref.npy holds what the REFERENCE's own GaussianPoint::fromModel (GaussianGeometry.cpp:59-127
    with its miniply reader, compiled from /root/reference into oracle/_ref) returned for <name>.ply
    (tests/golden/make_ply_golden.py). The drop-in's own reader must produce the same 240-byte records bit for bit:
    binary and ASCII bodies, double / uchar columns, a face element after the vertices, CRLF headers, shuffled property
    order (the reference takes the properties FOLLOWING f_dc_2 as the rest coefficients, :89-92)."""
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ply")
    want = np.load(os.path.join(here, name + ".ref.npy"))
    got = E.from_model(os.path.join(here, name + ".ply"))
    assert got.shape == want.shape == (64, 60)
    assert (got.view(np.uint32) == want.view(np.uint32)).all()
    assert (got[:, 11] == 1.0).all()                                   # scale modifier (:115)
    np.testing.assert_allclose(np.linalg.norm(got[:, 4:8], axis=1), 1.0, atol=1e-6)


def test_ply_fixtures_still_match_oracle_ref(oracle):
    """Only where /root/reference is present: the committed .ref.npy files are what the reference returns today."""
    import os
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ply")
    for name in PLY_FIXTURES:
        ref = oracle.ref_from_model(os.path.join(here, name + ".ply"))
        assert (ref.view(np.uint32) == np.load(os.path.join(here, name + ".ref.npy")).view(np.uint32)).all(), name
