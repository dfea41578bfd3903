"""Python mirror of the reference's public renderer API for the Gaussian path, over lib/libtpdhost.so.

Same vocabulary as the reference (demo/HelloGaussian/main.cpp:13-58):

    scene = Scene(); cloud = scene.add_group(points); scene.add(point)
    engine = GaussianEngine(1280, 720); engine.compile(scene, Settings(spherical_harmonics_degree=0))
    camera = PerspectiveCamera(1280, 720); camera.look_at(eye, target, up)
    engine.raster_frame(camera); image = engine.draw()

Every call goes C++ host layer (include/torpedo_b200/*.hpp) -> C ABI (include/tpdcu.h) -> CUDA kernels. Nothing here
computes: there is no NumPy/PyTorch fallback, and a missing library or GPU raises TpdError.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import TpdError, check, tpdcu

u32, vp, sz, i32, f32 = C.c_uint32, C.c_void_p, C.c_size_t, C.c_int, C.c_float

TPDH_SYMBOLS = {
    "tpdh_last_error": (C.c_char_p, []),
    "tpdh_camera_create": (vp, [u32, u32]),
    "tpdh_camera_destroy": (None, [vp]),
    "tpdh_camera_look_at": (None, [vp, vp, vp, vp]),
    "tpdh_camera_look_at_rt": (None, [vp, vp, vp]),
    "tpdh_camera_set_near": (None, [vp, f32]),
    "tpdh_camera_set_far": (None, [vp, f32]),
    "tpdh_camera_set_vertical_fov": (None, [vp, f32]),
    "tpdh_camera_on_image_size_change": (None, [vp, u32, u32]),
    "tpdh_camera_pack": (None, [vp, vp]),
    "tpdh_to_cartesian": (None, [f32, f32, f32, vp]),
    "tpdh_rgb2sh": (None, [f32, f32, f32, vp]),
    "tpdh_sizeof_gaussian_point": (u32, []),
    "tpdh_random_points": (i32, [u32, f32, f32, f32, f32, f32, C.c_uint64, vp]),
    "tpdh_model_load": (C.c_int64, [C.c_char_p]),
    "tpdh_model_take": (None, [vp]),
    "tpdh_scene_create": (vp, []),
    "tpdh_scene_destroy": (None, [vp]),
    "tpdh_scene_add_group": (u32, [vp, vp, u32]),
    "tpdh_scene_add_point": (u32, [vp, vp]),
    "tpdh_scene_count_all": (u32, [vp]),
    "tpdh_engine_create": (vp, [u32, u32, i32]),
    "tpdh_engine_destroy": (None, [vp]),
    "tpdh_engine_compile": (i32, [vp, vp, u32]),
    "tpdh_engine_compile_device": (i32, [vp, vp, u32, u32, vp]),
    "tpdh_engine_transform": (i32, [vp, u32, vp]),
    "tpdh_engine_raster_frame": (i32, [vp, vp, vp]),
    "tpdh_engine_draw": (i32, [vp, vp, sz]),
    "tpdh_engine_draw_async": (i32, [vp, vp, sz, vp]),
    "tpdh_engine_resize": (i32, [vp, u32, u32]),
    "tpdh_engine_wait_idle": (i32, [vp]),
    "tpdh_engine_handle": (vp, [vp]),
}

_host = None


def tpdhost() -> C.CDLL:
    global _host
    if _host is None:
        tpdcu()  # load the CUDA library first (and fail loudly if it is missing)
        if not os.path.exists(_lib.TPDHOST_PATH):
            raise TpdError(f"{_lib.TPDHOST_PATH} is missing: run __graft_entry__.build()")
        lib = C.CDLL(_lib.TPDHOST_PATH)
        for name, (res, args) in TPDH_SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _host = lib
    return _host


def _hcheck(status: int) -> None:
    if status != 0:
        raise TpdError(tpdhost().tpdh_last_error().decode())


def _f3(v):
    return (f32 * 3)(*[float(x) for x in v])


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(vp)


def to_cartesian(theta: float, phi: float, radius: float = 1.0) -> np.ndarray:
    """math::to_cartesian (torpedo/math/include/torpedo/math/transform.h:10-16)"""
    out = np.zeros(3, dtype=np.float32)
    tpdhost().tpdh_to_cartesian(theta, phi, radius, _ptr(out))
    return out


def from_model(ply_file: str) -> np.ndarray:
    """GaussianPoint::fromModel (volumetric/src/GaussianGeometry.cpp:59-127): load a 3DGS point cloud as (n, 60) float32."""
    n = tpdhost().tpdh_model_load(os.fsencode(ply_file))
    if n < 0:
        raise TpdError(tpdhost().tpdh_last_error().decode())
    out = np.zeros((n, 60), dtype=np.float32)
    if n:
        tpdhost().tpdh_model_take(_ptr(out))
    return out


class Camera:
    """tpd::Camera (rendering/include/torpedo/rendering/Camera.h:11-38) — abstract in the reference too."""

    def look_at(self, eye, center, up) -> None:
        raise NotImplementedError

    def pack(self) -> np.ndarray:
        raise NotImplementedError


class PerspectiveCamera(Camera):
    """tpd::PerspectiveCamera (extension/include/torpedo/extension/PerspectiveCamera.h:9-44)."""

    def __init__(self, image_width: int, image_height: int):
        self._h = tpdhost().tpdh_camera_create(image_width, image_height)

    def __del__(self):
        if getattr(self, "_h", None) and _host is not None:
            _host.tpdh_camera_destroy(self._h)
            self._h = None

    def look_at(self, eye, center, up) -> None:
        tpdhost().tpdh_camera_look_at(self._h, _f3(eye), _f3(center), _f3(up))

    def look_at_rt(self, R, t) -> None:
        r = np.ascontiguousarray(R, dtype=np.float32).reshape(9)
        tpdhost().tpdh_camera_look_at_rt(self._h, _ptr(r), _f3(t))

    def set_near(self, near: float) -> None:
        tpdhost().tpdh_camera_set_near(self._h, near)

    def set_far(self, far: float) -> None:
        tpdhost().tpdh_camera_set_far(self._h, far)

    def set_vertical_fov(self, degrees: float) -> None:
        tpdhost().tpdh_camera_set_vertical_fov(self._h, degrees)

    def on_image_size_change(self, w: int, h: int) -> None:
        tpdhost().tpdh_camera_on_image_size_change(self._h, w, h)

    def pack(self) -> np.ndarray:
        """The 136-byte camera block of GaussianEngine::updateCameraBuffer (GaussianEngine.cpp:764-775)."""
        out = np.zeros(_lib.CAMERA_FLOATS, dtype=np.float32)
        tpdhost().tpdh_camera_pack(self._h, _ptr(out))
        return out


class Scene:
    """tpd::Scene (rendering/include/torpedo/rendering/Scene.h:17-54), GaussianPoint components only."""

    def __init__(self):
        self._h = tpdhost().tpdh_scene_create()
        self._keepalive = []  # groups are borrowed spans in the reference; keep the arrays alive

    def __del__(self):
        if getattr(self, "_h", None) and _host is not None:
            _host.tpdh_scene_destroy(self._h)
            self._h = None

    @staticmethod
    def _as_points(points) -> np.ndarray:
        a = np.ascontiguousarray(points, dtype=np.float32)
        if a.size % 60:
            raise TpdError("GaussianPoint arrays must hold 60 floats (240 bytes) per point")
        return a.reshape(-1, 60)

    def add_group(self, points) -> int:
        """scene.add(tpd::ent::group(points)) — one entity (one transform) for the whole cloud."""
        a = self._as_points(points)
        self._keepalive.append(a)
        return int(tpdhost().tpdh_scene_add_group(self._h, _ptr(a), a.shape[0]))

    def add(self, point) -> int:
        """scene.add(GaussianPoint{...}) — a single-point entity."""
        a = self._as_points(point)
        if a.shape[0] != 1:
            raise TpdError("Scene.add takes exactly one GaussianPoint; use add_group for clouds")
        return int(tpdhost().tpdh_scene_add_point(self._h, _ptr(a)))

    def count_all(self) -> int:
        return int(tpdhost().tpdh_scene_count_all(self._h))


@dataclass
class Settings:
    """GaussianEngine::Settings (GaussianEngine.h:17-21)"""
    spherical_harmonics_degree: int = 3


class GaussianEngine:
    """tpd::GaussianEngine (volumetric/include/torpedo/volumetric/GaussianEngine.h:15-29) over CUDA."""

    def __init__(self, width: int, height: int, device: int = 0):
        self.width, self.height = width, height
        self._h = tpdhost().tpdh_engine_create(width, height, device)
        if not self._h:
            raise TpdError(tpdhost().tpdh_last_error().decode())
        self._ctx = tpdhost().tpdh_engine_handle(self._h)

    def close(self) -> None:
        if getattr(self, "_h", None) and _host is not None:
            _host.tpdh_engine_destroy(self._h)
            self._h = None

    __del__ = close

    # ---- the reference's public surface -------------------------------------------------------------
    def compile(self, scene: Scene, settings: Settings | None = None) -> None:
        settings = settings or Settings()
        _hcheck(tpdhost().tpdh_engine_compile(self._h, scene._h, settings.spherical_harmonics_degree))

    def compile_device(self, d_records240: int, count: int, settings: Settings | None = None, stream: int | None = None) -> None:
        """compile() for a cloud already resident on this GPU as 240-byte records (after the NCCL scene broadcast)."""
        settings = settings or Settings()
        _hcheck(tpdhost().tpdh_engine_compile_device(self._h, d_records240, count, settings.spherical_harmonics_degree, stream))

    def transform(self, entity: int, matrix) -> None:
        """engine->getTransformHost()->transform(entity, mat4) (rendering/src/TransformHost.cpp:3-11)"""
        m = np.ascontiguousarray(matrix, dtype=np.float32).reshape(16)
        _hcheck(tpdhost().tpdh_engine_transform(self._h, entity, _ptr(m)))

    def raster_frame(self, camera: Camera, stream: int | None = None) -> None:
        _hcheck(tpdhost().tpdh_engine_raster_frame(self._h, camera._h, stream))

    def draw(self, out: np.ndarray | None = None) -> np.ndarray:
        """Wait for the frame and copy the RGBA8 target to host memory (height, width, 4)."""
        if out is None:
            out = np.empty((self.height, self.width, 4), dtype=np.uint8)
        _hcheck(tpdhost().tpdh_engine_draw(self._h, _ptr(out), out.strides[0]))
        return out

    def draw_async(self, out: np.ndarray, stream: int | None = None) -> None:
        """Enqueue the copy of the newest frame into (pinned) host memory on `stream`; the caller waits on the stream."""
        _hcheck(tpdhost().tpdh_engine_draw_async(self._h, _ptr(out), out.strides[0], stream))

    def frames_repeated(self) -> int:
        n = u32(0)
        check(tpdcu().tpdcu_frames_repeated(self._ctx, C.byref(n)))
        return n.value

    def resize(self, width: int, height: int) -> None:
        _hcheck(tpdhost().tpdh_engine_resize(self._h, width, height))
        self.width, self.height = width, height

    def wait_idle(self) -> None:
        _hcheck(tpdhost().tpdh_engine_wait_idle(self._h))

    # ---- introspection through the C ABI (parity tests, bench) ----------------------------------------
    @property
    def ctx(self):
        return self._ctx

    def raster_ubo(self, ubo34, sh_degree: int = 3, stream: int | None = None) -> None:
        u = np.ascontiguousarray(ubo34, dtype=np.float32).reshape(34)
        check(tpdcu().tpdcu_raster(self._ctx, _ptr(u), sh_degree, stream))

    def raster_views(self, ubos, d_frames: int, frame_stride: int, sh_degree: int = 3, stream: int | None = None) -> None:
        u = np.ascontiguousarray(ubos, dtype=np.float32).reshape(-1, 34)
        check(tpdcu().tpdcu_raster_views(self._ctx, _ptr(u), u.shape[0], sh_degree, d_frames, frame_stride, stream))

    def finish(self) -> int:
        p = u32(0)
        check(tpdcu().tpdcu_finish(self._ctx, C.byref(p)))
        return p.value

    def counts(self) -> tuple[int, int]:
        p, v = u32(0), u32(0)
        check(tpdcu().tpdcu_get_counts(self._ctx, C.byref(p), C.byref(v)))
        return p.value, v.value

    def read_splats(self, n: int) -> np.ndarray:
        out = np.zeros((n, 12), dtype=np.uint32)
        check(tpdcu().tpdcu_read_splats(self._ctx, _ptr(out), n))
        return out

    def read_sorted(self) -> tuple[np.ndarray, np.ndarray]:
        p, _ = self.counts()
        keys, vals = np.zeros(p, dtype=np.uint64), np.zeros(p, dtype=np.uint32)
        if p:
            check(tpdcu().tpdcu_read_keys(self._ctx, _ptr(keys), p))
            check(tpdcu().tpdcu_read_values(self._ctx, _ptr(vals), p))
        return keys, vals

    def read_unsorted(self) -> tuple[np.ndarray, np.ndarray]:
        p, _ = self.counts()
        keys, vals = np.zeros(p, dtype=np.uint64), np.zeros(p, dtype=np.uint32)
        if p:
            check(tpdcu().tpdcu_read_unsorted(self._ctx, _ptr(keys), _ptr(vals), p))
        return keys, vals

    def read_emitted(self) -> np.ndarray:
        """The pair words emit_kernel wrote for the newest frame (tpdcu_read_emitted), before the tile sort."""
        p = self.counts()[0]
        words = np.zeros(p, dtype=np.uint64)
        if p:
            check(tpdcu().tpdcu_read_emitted(self._ctx, _ptr(words), p))
        return words

    def read_ranges(self) -> np.ndarray:
        tiles = ((self.width + 15) // 16) * ((self.height + 15) // 16)
        out = np.zeros((tiles, 2), dtype=np.uint32)
        check(tpdcu().tpdcu_read_ranges(self._ctx, _ptr(out), tiles))
        return out

    def enable_stage_timing(self, enable: bool = True) -> None:
        check(tpdcu().tpdcu_enable_stage_timing(self._ctx, int(enable)))

    def stage_times_ms(self) -> dict:
        t = np.zeros(_lib.NUM_STAGES, dtype=np.float32)
        check(tpdcu().tpdcu_stage_times_ms(self._ctx, _ptr(t)))
        names = ["setup", "preprocess", "depth_sort", "duplicate", "tile_sort", "ranges", "blend", "frame",
                 "depth_passes_run", "tile_passes_run", "tile_sort_hist_plan"]
        return dict(zip(names, [float(x) for x in t]))

    def sort_info(self) -> dict:
        d, dp, t, tp = u32(0), u32(0), u32(0), u32(0)
        check(tpdcu().tpdcu_get_sort_info(self._ctx, C.byref(d), C.byref(dp), C.byref(t), C.byref(tp)))
        return {"depth_bits": d.value, "depth_passes": dp.value, "tile_bits": t.value, "tile_passes": tp.value}

    def graph_replay(self, enable: int = -1) -> tuple[int, int]:
        """Switch CUDA-graph replay of the frame on (1) / off (0) or just query (-1); returns (captures, launches)."""
        cap, lau = u32(0), u32(0)
        check(tpdcu().tpdcu_set_graph_replay(self._ctx, enable, C.byref(cap), C.byref(lau)))
        return cap.value, lau.value

    def set_frames_in_flight(self, frames: int) -> None:
        check(tpdcu().tpdcu_set_frames_in_flight(self._ctx, frames))

    def capacity(self) -> int:
        c = u32(0)
        check(tpdcu().tpdcu_get_capacity(self._ctx, C.byref(c)))
        return c.value

    def device_info(self) -> tuple[str, int]:
        buf = C.create_string_buffer(256)
        sms = i32(0)
        check(tpdcu().tpdcu_device_info(self._ctx, buf, 256, C.byref(sms)))
        return buf.value.decode(), sms.value
