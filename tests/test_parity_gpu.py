"""GPU parity: the CUDA path (through the C++ host layer and the C ABI) against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): per-Gaussian tile counts/offsets, P, duplicated pairs, sorted keys and values, per-tile
ranges BIT-EXACT; RGB within 1/255 max-abs per channel (PSNR >= 50 dB) and alpha == 255.
"""
import hashlib

import numpy as np
import pytest

from tests.cases import frame_cases, golden_cameras, golden_frames

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def E(built_libs):
    from torpedo_b200 import engine
    return engine


def f32(a):
    return np.ascontiguousarray(a, dtype=np.uint32).view(np.float32)


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2) / 255.0 ** 2
    return 99.0 if mse == 0 else -10.0 * np.log10(mse)


def render_both(E, oracle, g, ubo, w, h, deg, model=None, groups=None):
    scene = E.Scene()
    entity = scene.add_group(g)
    eng = E.GaussianEngine(w, h)
    eng.compile(scene, E.Settings(deg))
    if model is not None:
        eng.transform(entity, model)
    eng.raster_ubo(ubo, deg)
    img = eng.draw()
    ref = oracle.render(g, ubo, w, h, deg, models=None if model is None else np.asarray(model, dtype=np.float32).reshape(1, 16))
    return eng, img, ref


def assert_frame_parity(eng, img, ref, n):
    pairs, visible = eng.counts()
    assert pairs == ref.pairs
    assert visible == int((ref.tiles > 0).sum())
    splats = eng.read_splats(n)
    vis = ref.tiles > 0
    # offsets (the reference's in-place exclusive scan) for every Gaussian, culled ones included
    assert (splats[:, 3] == ref.splats[:, 3]).all()
    # bit-critical per-Gaussian fields: pixel centre, depth, radius (only where visible: culled are stale in the reference)
    for col in (4, 5, 6, 7):
        assert (splats[vis, col] == ref.splats[vis, col]).all(), f"splat word {col}"
    assert (f32(splats[~vis, 7]) == 0).all()
    assert (splats[vis, 11] == ref.splats[vis, 11]).all()  # opacity is copied
    # tolerance-checked fields: conic (FMA-free but 1/det path identical => should still be exact) and colour
    np.testing.assert_allclose(f32(splats[vis][:, 8:11]), f32(ref.splats[vis][:, 8:11]), rtol=1e-6, atol=1e-30)
    np.testing.assert_allclose(f32(splats[vis][:, 0:3]), f32(ref.splats[vis][:, 0:3]), rtol=2e-5, atol=2e-6)
    # duplication: same pairs in the same (keygen.slang) order
    uk, uv = eng.read_unsorted()
    assert (uk == ref.unsorted_keys).all() and (uv == ref.unsorted_vals).all()
    # sort + ranges
    keys, vals = eng.read_sorted()
    assert (keys == ref.keys).all()
    assert (vals == ref.vals).all()
    assert (eng.read_ranges() == ref.ranges).all()
    # image
    assert (img[..., 3] == 255).all()
    diff = np.abs(img[..., :3].astype(np.int32) - ref.rgba[..., :3].astype(np.int32))
    assert diff.max() <= 1, f"max abs diff {diff.max()} LSB"
    assert psnr(img[..., :3], ref.rgba[..., :3]) >= 50.0
    return keys, vals


@pytest.mark.parametrize("name", list(frame_cases()))
def test_seeded_scenes_match_oracle_and_golden(E, oracle, name):
    gen, cam_name, w, h, deg, model = frame_cases()[name]
    _, cams = golden_cameras()
    g = gen()
    eng, img, ref = render_both(E, oracle, g, cams[cam_name], w, h, deg, model)
    keys, vals = assert_frame_parity(eng, img, ref, g.shape[0])
    gold = golden_frames()[name]  # the committed fixtures, independent of the oracle built on this box
    assert gold["pairs"] == len(keys)
    assert gold["keys_sha"] == sha(keys) and gold["vals_sha"] == sha(vals)
    assert gold["ranges_sha"] == sha(eng.read_ranges())
    uk, uv = eng.read_unsorted()
    assert gold["unsorted_keys_sha"] == sha(uk) and gold["unsorted_vals_sha"] == sha(uv)
    assert gold["offsets_sha"] == sha(eng.read_splats(g.shape[0])[:, 3])
    np.testing.assert_allclose(img[..., :3].reshape(-1, 3).mean(axis=0), gold["image_mean"], atol=0.05)
    eng.close()


def test_reference_api_flow_hello_gaussian(E, oracle):
    """demo/HelloGaussian/main.cpp:13-58 line by line: group + single entity, SH degree 0, lookAt, rasterFrame, draw."""
    from torpedo_b200 import scenes
    w, h = 320, 180
    pts = scenes.hello_gaussian(2048, seed=11)
    scene = E.Scene()
    scene.add_group(pts[:-1])
    scene.add(pts[-1])
    engine = E.GaussianEngine(w, h)
    engine.compile(scene, E.Settings(spherical_harmonics_degree=0))
    camera = E.PerspectiveCamera(w, h)
    camera.look_at(E.to_cartesian(0.785, 0.9, 8.0), (0, 0, 0), (0, 0, 1))
    engine.raster_frame(camera)
    img = engine.draw()
    # two entities (group, single) with identity transforms
    ref = oracle.render(pts, camera.pack(), w, h, 0, entity_idx=np.r_[np.zeros(2048, np.uint32), np.uint32(1)], models=np.tile(np.eye(4, dtype=np.float32).reshape(1, 16), (2, 1)))
    assert engine.counts()[0] == ref.pairs
    k, v = engine.read_sorted()
    assert (k == ref.keys).all() and (v == ref.vals).all()
    assert np.abs(img.astype(np.int32) - ref.rgba.astype(np.int32)).max() <= 1
    assert img[..., :3].max() > 200  # the big white Gaussian is there
    engine.close()


def test_multi_entity_transforms(E, oracle):
    """TransformHost semantics: model = M[transformIndices[i]] (project.slang:43-44), effective next frame."""
    from torpedo_b200 import scenes
    w, h = 256, 144
    a = scenes.garden(6000, seed=21, log_scale_mean=-3.4)
    b = scenes.garden(4000, seed=22, log_scale_mean=-3.4)
    scene = E.Scene()
    ea, eb = scene.add_group(a), scene.add_group(b)
    eng = E.GaussianEngine(w, h)
    eng.compile(scene)
    cam = E.PerspectiveCamera(w, h)
    cam.look_at((2.8, 2.8, 2.6), (0, 0, 0), (0, 0, 1))
    Ma = np.array([1, 0, 0, 0.5, 0, 1, 0, 0, 0, 0, 1, -0.25, 0, 0, 0, 1], dtype=np.float32)
    c, s = np.float32(np.cos(0.7)), np.float32(np.sin(0.7))
    Mb = np.array([c, -s, 0, 0, s, c, 0, 0.3, 0, 0, 1.5, 0, 0, 0, 0, 1], dtype=np.float32)
    g = np.concatenate([a, b])
    idx = np.r_[np.zeros(len(a), np.uint32), np.ones(len(b), np.uint32)]
    for models in ([np.eye(4, dtype=np.float32).reshape(-1)] * 2, [Ma, Mb]):
        eng.transform(ea, models[0])
        eng.transform(eb, models[1])
        eng.raster_frame(cam)
        img = eng.draw()
        ref = oracle.render(g, cam.pack(), w, h, 3, entity_idx=idx, models=np.stack(models))
        assert_frame_parity(eng, img, ref, len(g))
    eng.transform(12345, Ma)  # unknown entities are ignored (TransformHost.cpp:4-6)
    eng.close()


def test_edge_cases(E, oracle):
    from torpedo_b200 import scenes
    _, cams = golden_cameras()
    # nothing visible: all Gaussians behind the camera -> P == 0, black frame, all ranges (0,0)
    g = scenes.garden(3000, seed=31, log_scale_mean=-3.5)
    g[:, 0:3] += np.float32(100.0)
    eng, img, ref = render_both(E, oracle, g, cams["garden_256x144"], 256, 144, 3)
    assert eng.counts() == (0, 0) and ref.pairs == 0
    assert (img[..., :3] == 0).all() and (img[..., 3] == 255).all() and (eng.read_ranges() == 0).all()
    eng.close()
    # a single Gaussian
    g1 = scenes.hello_gaussian(0, seed=1)  # just the big white one
    eng, img, ref = render_both(E, oracle, g1, cams["garden_100x70"], 100, 70, 0)
    assert_frame_parity(eng, img, ref, 1)
    eng.close()
    # an empty scene is a no-op (GaussianEngine.cpp:362-365) and rasterFrame then records nothing
    scene = E.Scene()
    eng = E.GaussianEngine(64, 64)
    eng.compile(scene)
    eng.raster_frame(E.PerspectiveCamera(64, 64))
    with pytest.raises(E.TpdError):
        eng.draw()
    eng.close()
    # SH degree is clamped to 3 (GaussianEngine.cpp:366-370)
    g = scenes.garden(2000, seed=32, log_scale_mean=-3.2)
    eng, img, ref = render_both(E, oracle, g, cams["garden_100x70"], 100, 70, 7)
    assert_frame_parity(eng, img, ref, len(g))
    eng.close()


def test_huge_splats_and_partial_tiles(E, oracle):
    """Gaussians covering the whole screen (rect clamped to the grid) at a size that is not a multiple of 16."""
    from torpedo_b200 import scenes
    _, cams = golden_cameras()
    g = scenes.garden(1500, seed=41, log_scale_mean=-0.7)  # scales ~0.5: radii of hundreds of pixels
    eng, img, ref = render_both(E, oracle, g, cams["garden_100x70"], 100, 70, 3)
    assert ref.tiles.max() == 7 * 5  # at least one covers the whole 7x5 grid
    assert_frame_parity(eng, img, ref, len(g))
    eng.close()


def test_capacity_growth_resize_and_rerender(E, oracle):
    """P is never read back mid-frame: an overflowing frame is re-rendered after the buffers grow (grow-only)."""
    from torpedo_b200 import scenes
    g = scenes.garden(30000, seed=51, log_scale_mean=-3.6)
    scene = E.Scene()
    scene.add_group(g)
    eng = E.GaussianEngine(128, 72)
    eng.compile(scene)
    cam = E.PerspectiveCamera(128, 72)
    cam.look_at((2.8, 2.8, 2.6), (0, 0, 0), (0, 0, 1))
    caps = []
    for (w, h) in [(128, 72), (512, 288), (1024, 576), (128, 72)]:
        eng.resize(w, h)
        cam.on_image_size_change(w, h)
        eng.raster_frame(cam)
        img = eng.draw()
        ref = oracle.render(g, cam.pack(), w, h, 3)
        assert_frame_parity(eng, img, ref, len(g))
        caps.append(eng.capacity())
        assert caps[-1] >= ref.pairs
    assert caps == sorted(caps) and caps[2] > caps[0]
    # back-to-back frames without any host sync in between give the same result as a single one
    for _ in range(5):
        eng.raster_frame(cam)
    img2 = eng.draw()
    assert (img2 == img).all()
    eng.close()


def test_raster_views_batch(E, oracle):
    """Independent views of one scene in one call; frames land in a caller-owned device buffer."""
    import torch
    from torpedo_b200 import scenes
    w, h = 256, 144
    g = scenes.garden(20000, seed=61, log_scale_mean=-3.6)
    scene = E.Scene()
    scene.add_group(g)
    eng = E.GaussianEngine(w, h)
    eng.compile(scene)
    ubos = []
    for k in range(6):
        cam = E.PerspectiveCamera(w, h)
        cam.look_at(E.to_cartesian(2 * np.pi * k / 6, 0.9, 5.0), (0, 0, 0), (0, 0, 1))
        ubos.append(cam.pack())
    frames = torch.zeros((6, h, w, 4), dtype=torch.uint8, device="cuda")
    eng.raster_views(np.stack(ubos), frames.data_ptr(), h * w * 4, 3, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    out = frames.cpu().numpy()
    for k in range(6):
        ref = oracle.render(g, ubos[k], w, h, 3)
        assert np.abs(out[k].astype(np.int32) - ref.rgba.astype(np.int32)).max() <= 1, k
    eng.close()


def test_two_level_sort_info_and_equal_depths(E, oracle):
    """A frame sorts the visible Gaussians by the bits its depth range occupies, then the pairs by tile bits; Gaussians at
    EXACTLY the same depth must keep index order (the reference's stable sort of emission-ordered pairs)."""
    from torpedo_b200 import scenes
    _, cams = golden_cameras()
    g = scenes.garden(20000, seed=71, log_scale_mean=-3.6)
    g[1::2, 0:3] = g[0::2, 0:3]  # every odd Gaussian sits exactly on its even neighbour: identical view depth
    ref = oracle.render(g, cams["garden_256x144"], 256, 144, 3)
    scene = E.Scene()
    scene.add_group(g)
    eng = E.GaussianEngine(256, 144)
    eng.compile(scene)
    eng.raster_ubo(cams["garden_256x144"], 3)
    img = eng.draw()
    info = eng.sort_info()
    assert info["tile_bits"] == 8 and info["tile_passes"] == 1      # 16 x 9 tiles
    assert 1 <= info["depth_bits"] <= 32 and info["depth_passes"] <= (info["depth_bits"] + 7) // 8
    keys, vals = assert_frame_parity(eng, img, ref, len(g))
    same = keys[1:] == keys[:-1]
    assert same.any() and (vals[1:][same] > vals[:-1][same]).all()
    eng.close()


def test_depth_split_corner_cases(E, oracle):
    """The cut of tile | depth between the two sorts (DepthSplit, csrc/common.cuh): every Gaussian at the SAME depth (zero
    depth bits: the depth sort has nothing to do), and a tile grid whose id fills whole digits (no room for depth bits)."""
    from torpedo_b200 import scenes
    _, cams = golden_cameras()
    g = scenes.garden(3000, seed=5, log_scale_mean=-3.2)
    g[:, 0:3] = g[0, 0:3]  # one position, 3000 different shapes / colours / opacities
    for (w, h, cam) in [(256, 144, "garden_256x144")]:
        ref = oracle.render(g, cams[cam], w, h, 3)
        scene = E.Scene()
        scene.add_group(g)
        eng = E.GaussianEngine(w, h)
        eng.compile(scene)
        eng.raster_ubo(cams[cam], 3)
        img = eng.draw()
        info = eng.sort_info()
        assert info["depth_bits"] == 0 and info["depth_passes"] == 0
        if ref.pairs:
            assert_frame_parity(eng, img, ref, len(g))
        eng.close()
    # 16 x 16 tiles = exactly 8 tile bits: top depth bits cannot ride in the pair key, the depth sort takes them all
    g = scenes.garden(20000, seed=9, log_scale_mean=-3.6)
    cam = E.PerspectiveCamera(256, 256)
    cam.look_at((2.8, 2.8, 2.6), (0, 0, 0), (0, 0, 1))
    ref = oracle.render(g, cam.pack(), 256, 256, 2)
    scene = E.Scene()
    scene.add_group(g)
    eng = E.GaussianEngine(256, 256)
    eng.compile(scene, E.Settings(2))
    eng.raster_frame(cam)
    img = eng.draw()
    info = eng.sort_info()
    assert info["tile_bits"] == 8 and info["tile_passes"] == 1 and info["depth_passes"] == (info["depth_bits"] + 7) // 8
    assert_frame_parity(eng, img, ref, len(g))
    eng.close()


@pytest.mark.parametrize("size", [(16, 16), (5, 3), (17, 33), (4096, 16), (7680, 4320)])
def test_odd_framebuffer_sizes(E, oracle, size):
    """One tile (no tile bits: the tile sort has nothing to do), partial tiles, a one-tile-high strip, and 8K (19 tile
    bits: three tile-sort passes): sorted keys, values and ranges stay bit-exact."""
    from torpedo_b200 import scenes
    w, h = size
    g = scenes.garden(5000, seed=11, log_scale_mean=-3.0)
    cam = E.PerspectiveCamera(w, h)
    cam.look_at((2.8, 2.8, 2.6), (0, 0, 0), (0, 0, 1))
    scene = E.Scene()
    scene.add_group(g)
    eng = E.GaussianEngine(w, h)
    eng.compile(scene, E.Settings(1))
    for _ in range(2):  # the second frame runs with grown buffers and a replayed graph
        eng.raster_frame(cam)
        img = eng.draw()
    ref = oracle.render(g, cam.pack(), w, h, 1)
    keys, vals = eng.read_sorted()
    assert eng.counts()[0] == ref.pairs
    assert (keys == ref.keys).all() and (vals == ref.vals).all() and (eng.read_ranges() == ref.ranges).all()
    assert np.abs(img.astype(np.int32) - ref.rgba.astype(np.int32)).max() <= 1
    tiles = ((w + 15) // 16) * ((h + 15) // 16)
    info = eng.sort_info()
    assert info["tile_passes"] == ((max(tiles - 1, 0).bit_length() + 7) // 8)
    eng.close()


def test_cpp_hello_gaussian_demo(E, oracle, built_libs):
    """The reference's HelloGaussian demo compiled against the header-only C++ drop-in gives the same frame as the oracle."""
    import os
    import subprocess
    from torpedo_b200 import scenes
    exe = os.path.join(built_libs.LIB_DIR, "hello_gaussian")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    words = out.stdout.split()
    pairs, visible, checksum = int(words[1]), int(words[3]), int(words[5])
    cam = E.PerspectiveCamera(1280, 720)
    cam.look_at(E.to_cartesian(0.785, 0.9, 8.0), (0, 0, 0), (0, 0, 1))
    g = scenes.hello_gaussian(8192, seed=1)
    ref = oracle.render(g, cam.pack(), 1280, 720, 0, entity_idx=np.r_[np.zeros(8192, np.uint32), np.uint32(1)],
                        models=np.tile(np.eye(4, dtype=np.float32).reshape(1, 16), (2, 1)))
    assert pairs == ref.pairs and visible == int((ref.tiles > 0).sum())
    assert abs(checksum - int(ref.rgba[..., :3].astype(np.int64).sum())) <= 1280 * 720 * 3 * 0.001


def test_graph_replay_matches_direct_launches(E, oracle):
    """The frame's invariant middle section is replayed from a CUDA graph; results are identical to direct launches, and
    the graph is re-captured when the buffers it bakes in change."""
    from torpedo_b200 import scenes
    g = scenes.garden(20000, seed=81, log_scale_mean=-3.6)
    scene = E.Scene()
    scene.add_group(g)
    eng = E.GaussianEngine(256, 144)
    eng.compile(scene)
    cams = []
    for k in range(4):
        cam = E.PerspectiveCamera(256, 144)
        cam.look_at(E.to_cartesian(0.5 + k, 0.9, 5.0), (0, 0, 0), (0, 0, 1))
        cams.append(cam)
    eng.graph_replay(0)
    direct = []
    for cam in cams:
        eng.raster_frame(cam)
        direct.append((eng.draw().copy(), eng.read_sorted(), eng.read_ranges().copy()))
    eng.graph_replay(1)
    c0, l0 = eng.graph_replay()
    for cam, (img, (keys, vals), ranges) in zip(cams, direct):
        eng.raster_frame(cam)
        assert (eng.draw() == img).all()
        k2, v2 = eng.read_sorted()
        assert (k2 == keys).all() and (v2 == vals).all() and (eng.read_ranges() == ranges).all()
    c1, l1 = eng.graph_replay()
    assert l1 - l0 >= 4 and c1 - c0 >= 1
    ref = oracle.render(g, cams[-1].pack(), 256, 144, 3)
    assert (direct[-1][1][0] == ref.keys).all()
    eng.resize(320, 180)  # new target size -> new zero region -> re-capture
    cam = E.PerspectiveCamera(320, 180)
    cam.look_at((2.8, 2.8, 2.6), (0, 0, 0), (0, 0, 1))
    eng.raster_frame(cam)
    eng.raster_frame(cam)
    img = eng.draw()
    ref = oracle.render(g, cam.pack(), 320, 180, 3)
    assert np.abs(img.astype(np.int32) - ref.rgba.astype(np.int32)).max() <= 1
    assert eng.graph_replay()[0] > c1
    eng.close()


def test_two_frames_in_flight_match_serial_rendering(E, oracle):
    """Consecutive raster calls alternate between two frame slots on private streams (the reference keeps two Frame objects
    in flight); every frame still equals what strictly serial rendering gives, whether frames share a target or not."""
    import torch
    from torpedo_b200 import scenes
    from torpedo_b200._lib import check, tpdcu
    w, h = 256, 144
    g = scenes.garden(20000, seed=91, log_scale_mean=-3.6)
    scene = E.Scene()
    scene.add_group(g)
    eng = E.GaussianEngine(w, h)
    eng.compile(scene)
    ubos = []
    for k in range(7):
        cam = E.PerspectiveCamera(w, h)
        cam.look_at(E.to_cartesian(0.3 + 0.9 * k, 0.9, 4.0 + 0.3 * k), (0, 0, 0), (0, 0, 1))
        ubos.append(cam.pack())
    eng.set_frames_in_flight(1)
    serial = []
    for u in ubos:
        eng.raster_ubo(u, 3)
        serial.append(eng.draw().copy())
    ref = oracle.render(g, ubos[3], w, h, 3)
    assert np.abs(serial[3].astype(np.int32) - ref.rgba.astype(np.int32)).max() <= 1
    eng.set_frames_in_flight(3)
    # (a) distinct targets, no host sync between frames
    frames = torch.zeros((7, h, w, 4), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    for k, u in enumerate(ubos):
        check(tpdcu().tpdcu_bind_output_device_ptr(eng.ctx, frames[k].data_ptr(), w * 4))
        eng.raster_ubo(u, 3, stream)
    eng.finish()  # the one host-side check: the second slot's buffers were still small, its first frame is repeated here
    torch.cuda.synchronize()
    out = frames.cpu().numpy()
    for k in range(7):
        assert (out[k] == serial[k]).all(), k
    # (b) one shared target: the caller sees the newest frame, and stream order protects every intermediate consumer
    check(tpdcu().tpdcu_bind_output_device_ptr(eng.ctx, None, 0))
    copies = []
    shared = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    check(tpdcu().tpdcu_bind_output_device_ptr(eng.ctx, shared.data_ptr(), w * 4))
    for u in ubos:
        eng.raster_ubo(u, 3, stream)
        copies.append(shared.clone())       # enqueued on the same stream right after the frame
    torch.cuda.synchronize()
    for k in range(7):
        assert (copies[k].cpu().numpy() == serial[k]).all(), k
    assert (eng.draw() == serial[-1]).all()
    eng.close()


def test_blend_dispatch_order_never_changes_the_image(E, oracle):
    """The blend processes tiles by decreasing expected cost (list length, then the splats it consumed on that tile in an
    earlier frame): a scheduling hint. The same view rendered with no hints, with its own hints and with another view's
    hints gives the same bytes, equal to a fresh context's and within tolerance of the oracle; a resize drops the hints."""
    from torpedo_b200 import scenes
    w, h = 400, 240  # 25 x 15 tiles: dense centre, sparse border, so the order really differs from image order
    g = scenes.garden(30000, seed=17, log_scale_mean=-3.2)
    cams = []
    for theta, radius in ((0.4, 4.0), (2.9, 2.5)):
        cam = E.PerspectiveCamera(w, h)
        cam.look_at(E.to_cartesian(theta, 0.9, radius), (0, 0, 0), (0, 0, 1))
        cams.append(cam.pack())
    scene = E.Scene()
    scene.add_group(g)
    eng = E.GaussianEngine(w, h)
    eng.compile(scene)
    eng.set_frames_in_flight(1)
    imgs = []
    for u in (cams[0], cams[0], cams[1], cams[0]):  # no hints | own hints | (other view) | the other view's hints
        eng.raster_ubo(u, 3)
        imgs.append(eng.draw().copy())
    assert (imgs[0] == imgs[1]).all() and (imgs[0] == imgs[3]).all()
    ref = oracle.render(g, cams[0], w, h, 3)
    assert (eng.read_ranges() == ref.ranges).all()
    assert np.abs(imgs[0].astype(np.int32) - ref.rgba.astype(np.int32)).max() <= 1
    assert (imgs[0][..., 3] == 255).all()
    eng.resize(w // 2, h // 2)  # other tile grid: hints of the old one must not be used
    eng.raster_ubo(cams[1], 3)
    small = eng.draw().copy()
    eng.close()
    fresh = E.GaussianEngine(w // 2, h // 2)
    fresh.compile(scene)
    fresh.raster_ubo(cams[1], 3)
    assert (fresh.draw() == small).all()
    fresh.close()


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_every_sh_degree_on_demand_colour(E, oracle, deg):
    """The frame evaluates the SH colour inside the blend, for the splats it stages; tpdcu_read_splats evaluates it for
    every visible Gaussian with a kernel of its own. Both must agree with the oracle at every degree (the planes of a
    row a degree does not use hold non-zero coefficients here)."""
    from torpedo_b200 import scenes
    w, h = 320, 200
    g = scenes.garden(12000, seed=40 + deg, log_scale_mean=-3.4, sh_rest_std=0.2)
    cam = E.PerspectiveCamera(w, h)
    cam.look_at((2.4, -2.9, 2.2), (0, 0, 0), (0, 0, 1))
    eng, img, ref = render_both(E, oracle, g, cam.pack(), w, h, deg)
    assert np.abs(img.astype(np.int32) - ref.rgba.astype(np.int32)).max() <= 1
    assert psnr(img[..., :3], ref.rgba[..., :3]) >= 50.0
    splats = eng.read_splats(len(g))
    vis = ref.tiles > 0
    got, want = f32(splats[vis, 0:3]), f32(ref.splats[vis, 0:3])
    assert np.allclose(got, want, rtol=2e-6, atol=2e-6)
    # a second frame after the export still renders the same image (the export's colour kernel shares the slot's arrays)
    eng.raster_ubo(cam.pack(), deg)
    assert (eng.draw() == img).all()
    eng.close()


def test_many_frames_in_flight_keep_every_tile(E, oracle):
    """Regression: the blend's dispatch order is built from cost hints that the blends of OTHER frames in flight rewrite
    at the same time. Built from two different reads of a hint, the order was once no permutation: a tile rendered twice,
    another never (stale pixels of an older frame). 1080p (8160 tiles), 128 frames over four very different views, three
    in flight, each into its own buffer pre-filled with a sentinel: every frame must equal the serial rendering of its
    view."""
    import torch
    from torpedo_b200 import scenes
    from torpedo_b200._lib import check, tpdcu
    w, h = 1920, 1080
    g = scenes.garden(150000, seed=23, log_scale_mean=-4.2)
    scene = E.Scene()
    scene.add_group(g)
    eng = E.GaussianEngine(w, h)
    eng.compile(scene)
    ubos = []
    for theta, phi, radius in ((0.2, 0.9, 4.5), (1.9, 0.5, 2.2), (3.6, 1.3, 7.0), (5.1, 0.9, 1.4)):
        cam = E.PerspectiveCamera(w, h)
        cam.look_at(E.to_cartesian(theta, phi, radius), (0, 0, 0), (0, 0, 1))
        ubos.append(cam.pack())
    eng.set_frames_in_flight(1)
    serial = []
    for u in ubos:
        eng.raster_ubo(u, 3)
        serial.append(torch.from_numpy(eng.draw().copy()).cuda())
    eng.set_frames_in_flight(3)
    n_slots, n_frames = 16, 128
    frames = torch.empty((n_slots, h, w, 4), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    for base in range(0, n_frames, n_slots):
        frames.fill_(0x5a)
        for j in range(n_slots):
            check(tpdcu().tpdcu_bind_output_device_ptr(eng.ctx, frames[j].data_ptr(), w * 4))
            eng.raster_ubo(ubos[((base + j) * 3) % 4], 3, stream)
        eng.finish()
        torch.cuda.synchronize()
        for j in range(n_slots):
            assert torch.equal(frames[j], serial[((base + j) * 3) % 4]), base + j
    check(tpdcu().tpdcu_bind_output_device_ptr(eng.ctx, None, 0))
    eng.close()


def test_overflowing_older_frame_leaves_the_newest_frame_newest(E, oracle):
    """Three frames in flight; of two frames enqueued back to back only the FIRST overflows the grow-only pair buffers. The
    repeat of that frame (after growth) must not become what draw / read_sorted / read_ranges describe: they refer to the
    newest frame, whose intermediate buffers may have been re-allocated by the growth (it is rendered again too)."""
    from torpedo_b200 import scenes
    w, h = 320, 192
    g = scenes.garden(30000, seed=71, log_scale_mean=-3.4)
    scene = E.Scene()
    scene.add_group(g)
    near, far = E.PerspectiveCamera(w, h), E.PerspectiveCamera(w, h)
    near.look_at(E.to_cartesian(0.3, 0.9, 2.0), (0, 0, 0), (0, 0, 1))
    far.look_at(E.to_cartesian(0.3, 0.9, 14.0), (0, 0, 0), (0, 0, 1))
    ref_near = oracle.render(g, near.pack(), w, h, 3)
    ref_far = oracle.render(g, far.pack(), w, h, 3)
    eng = E.GaussianEngine(w, h)
    eng.compile(scene)
    eng.set_frames_in_flight(3)
    for _ in range(3):  # every slot sized by the far view only
        eng.raster_frame(far)
    eng.finish()
    cap = eng.capacity()
    assert ref_far.pairs <= cap < ref_near.pairs, (ref_far.pairs, cap, ref_near.pairs)
    before = eng.frames_repeated()
    eng.raster_frame(near)  # overflows
    eng.raster_frame(far)   # does not; this is the newest frame
    img = eng.draw()
    assert eng.frames_repeated() > before
    assert eng.counts()[0] == ref_far.pairs
    keys, vals = eng.read_sorted()
    assert (keys == ref_far.keys).all() and (vals == ref_far.vals).all()
    assert (eng.read_ranges() == ref_far.ranges).all()
    assert np.abs(img.astype(np.int32) - ref_far.rgba.astype(np.int32)).max() <= 1
    # and the near view itself, now that the buffers have grown
    eng.raster_frame(near)
    img = eng.draw()
    keys, vals = eng.read_sorted()
    assert (keys == ref_near.keys).all() and (vals == ref_near.vals).all()
    assert np.abs(img.astype(np.int32) - ref_near.rgba.astype(np.int32)).max() <= 1
    eng.close()


def test_reserve_pairs_keeps_the_newest_frame_readable(E, oracle):
    """tpdcu_reserve_pairs re-allocates the pair buffers: the newest frame is rendered again so that read_* still describe it."""
    from torpedo_b200 import scenes
    from torpedo_b200._lib import check, tpdcu
    w, h = 256, 144
    g = scenes.garden(20000, seed=72, log_scale_mean=-3.6)
    scene = E.Scene()
    scene.add_group(g)
    cam = E.PerspectiveCamera(w, h)
    cam.look_at((2.8, 2.8, 2.6), (0, 0, 0), (0, 0, 1))
    ref = oracle.render(g, cam.pack(), w, h, 3)
    eng = E.GaussianEngine(w, h)
    eng.compile(scene)
    eng.raster_frame(cam)
    eng.finish()
    check(tpdcu().tpdcu_reserve_pairs(eng._ctx, eng.capacity() * 4))
    keys, vals = eng.read_sorted()
    assert (keys == ref.keys).all() and (vals == ref.vals).all()
    assert (eng.read_ranges() == ref.ranges).all()
    assert np.abs(eng.draw().astype(np.int32) - ref.rgba.astype(np.int32)).max() <= 1
    eng.close()


@pytest.mark.parametrize("name,deg", [("sh3_binary", 3), ("sh2_shuffled", 2), ("sh1_mixed_types", 1)])
def test_ply_file_to_rendered_frame(E, oracle, name, deg):
    """PLY ingest to pixels: GaussianPoint::fromModel (the drop-in's reader, held bit-exact to the reference's by
    tests/test_host_layer.py) -> Scene -> compile -> rasterFrame -> draw, against the oracle fed with the records the
    REFERENCE's fromModel produced for the same file (tests/golden/ply/<name>.ref.npy)."""
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ply")
    g = E.from_model(os.path.join(here, name + ".ply"))
    want = np.load(os.path.join(here, name + ".ref.npy"))
    assert (g.view(np.uint32) == want.view(np.uint32)).all()
    w, h = 320, 200
    cam = E.PerspectiveCamera(w, h)
    cam.look_at((0.0, -6.0, 1.5), (0, 0, 0), (0, 0, 1))
    scene = E.Scene()
    scene.add_group(g)
    eng = E.GaussianEngine(w, h)
    eng.compile(scene, E.Settings(deg))
    eng.raster_frame(cam)
    img = eng.draw()
    ref = oracle.render(want, cam.pack(), w, h, deg)
    assert ref.pairs > 0
    assert_frame_parity(eng, img, ref, len(g))
    eng.close()


@pytest.mark.parametrize("n,scale", [(700, -2.6), (3000, -3.0), (12000, -3.4), (60000, -4.0), (250000, -4.6)])
def test_sort_sizes_around_tile_and_chain_boundaries(E, oracle, n, scale):
    """The onesweep passes cut their input into look-back chains (segments of whole tiles for a sort's first pass, runs of
    previous-pass bins for the others) and persistent CTAs walk the tiles by ticket: pair counts from a fraction of one tile
    to dozens of tiles, so that chains are empty, hold one partial tile, or end in one — sorted keys, values and ranges
    bit-exact, image within 1 LSB."""
    from torpedo_b200 import scenes
    w, h = 320, 180
    g = scenes.garden(n, seed=100 + n % 97, log_scale_mean=scale)
    cam = E.PerspectiveCamera(w, h)
    cam.look_at((2.8, 2.8, 2.6), (0, 0, 0), (0, 0, 1))
    scene = E.Scene()
    scene.add_group(g)
    eng = E.GaussianEngine(w, h)
    eng.compile(scene, E.Settings(3))
    eng.raster_frame(cam)
    img = eng.draw()
    ref = oracle.render(g, cam.pack(), w, h, 3)
    assert eng.counts() == (ref.pairs, int((ref.tiles > 0).sum()))
    keys, vals = eng.read_sorted()
    assert (keys == ref.keys).all() and (vals == ref.vals).all()
    assert (eng.read_ranges() == ref.ranges).all()
    assert np.abs(img.astype(np.int32) - ref.rgba.astype(np.int32)).max() <= 1
    # a second frame through the CUDA graph of the slot, same view: same bytes
    eng.raster_frame(cam)
    assert (eng.draw() == img).all()
    eng.close()
