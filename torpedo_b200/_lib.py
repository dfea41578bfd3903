"""ctypes loader for lib/libtpdcu.so (the C ABI declared in include/tpdcu.h). Fails loudly when the library is missing."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_HERE, "lib")
TPDCU_PATH = os.path.join(LIB_DIR, "libtpdcu.so")
TPDHOST_PATH = os.path.join(LIB_DIR, "libtpdhost.so")

CAMERA_FLOATS = 34
GAUSSIAN_BYTES = 240
SPLAT_BYTES = 48
NUM_STAGES = 11

u32, vp, sz, i32 = C.c_uint32, C.c_void_p, C.c_size_t, C.c_int

# name -> (restype, argtypes); kept in sync with include/tpdcu.h by tests/test_abi.py
TPDCU_SYMBOLS = {
    "tpdcu_create": (i32, [i32, C.POINTER(vp)]),
    "tpdcu_destroy": (None, [vp]),
    "tpdcu_last_error": (C.c_char_p, []),
    "tpdcu_device_info": (i32, [vp, C.c_char_p, sz, C.POINTER(i32)]),
    "tpdcu_upload_gaussians": (i32, [vp, vp, u32, vp, u32]),
    "tpdcu_upload_gaussians_device": (i32, [vp, vp, u32, vp, u32, vp]),
    "tpdcu_set_transform": (i32, [vp, u32, vp]),
    "tpdcu_resize": (i32, [vp, u32, u32]),
    "tpdcu_bind_output_device_ptr": (i32, [vp, vp, sz]),
    "tpdcu_bind_output_fd": (i32, [vp, i32, sz]),
    "tpdcu_ipc_frames_create": (i32, [i32, sz, C.POINTER(vp), vp]),
    "tpdcu_ipc_frames_open": (i32, [i32, vp, C.POINTER(vp)]),
    "tpdcu_ipc_frames_close": (i32, [i32, vp]),
    "tpdcu_ipc_frames_destroy": (i32, [i32, vp]),
    "tpdcu_raster": (i32, [vp, vp, u32, vp]),
    "tpdcu_raster_views": (i32, [vp, vp, u32, u32, vp, sz, vp]),
    "tpdcu_finish": (i32, [vp, C.POINTER(u32)]),
    "tpdcu_read_frame": (i32, [vp, vp, sz]),
    "tpdcu_read_frame_async": (i32, [vp, vp, sz, vp]),
    "tpdcu_frames_repeated": (i32, [vp, C.POINTER(u32)]),
    "tpdcu_get_counts": (i32, [vp, C.POINTER(u32), C.POINTER(u32)]),
    "tpdcu_read_splats": (i32, [vp, vp, u32]),
    "tpdcu_read_keys": (i32, [vp, vp, u32]),
    "tpdcu_read_values": (i32, [vp, vp, u32]),
    "tpdcu_read_ranges": (i32, [vp, vp, u32]),
    "tpdcu_read_unsorted": (i32, [vp, vp, vp, u32]),
    "tpdcu_read_emitted": (i32, [vp, vp, u32]),
    "tpdcu_enable_stage_timing": (i32, [vp, i32]),
    "tpdcu_stage_times_ms": (i32, [vp, vp]),
    "tpdcu_get_sort_info": (i32, [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]),
    "tpdcu_set_graph_replay": (i32, [vp, i32, C.POINTER(u32), C.POINTER(u32)]),
    "tpdcu_set_frames_in_flight": (i32, [vp, i32]),
    "tpdcu_get_capacity": (i32, [vp, C.POINTER(u32)]),
    "tpdcu_reserve_pairs": (i32, [vp, u32]),
    "tpdcu_sort_pairs_device": (i32, [vp, vp, vp, u32, u32, vp]),
    "tpdcu_sort_last_ms": (i32, [vp, C.POINTER(C.c_float), C.POINTER(u32)]),
}

_tpdcu = None


class TpdError(RuntimeError):
    """Raised for every non-zero status of the C ABI; the reference throws std::runtime_error / vk::SystemError."""


def tpdcu() -> C.CDLL:
    global _tpdcu
    if _tpdcu is None:
        if not os.path.exists(TPDCU_PATH):
            raise TpdError(f"{TPDCU_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU or PyTorch fallback for the rasterizer)")
        lib = C.CDLL(TPDCU_PATH)
        for name, (res, args) in TPDCU_SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _tpdcu = lib
    return _tpdcu


def check(status: int) -> None:
    if status != 0:
        msg = tpdcu().tpdcu_last_error()
        raise TpdError(f"tpdcu status {status}: {msg.decode() if msg else '?'}")
