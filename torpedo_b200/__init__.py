"""torpedo_b200 — B200-native (sm_100a) implementation of ndming/torpedo's Gaussian-splatting forward rasterizer.

Product layers:
  include/tpdcu.h + torpedo_b200/csrc/     hand-written CUDA kernels behind a C ABI (lib/libtpdcu.so)
  include/torpedo_b200/*.hpp               Vulkan-free C++ drop-in of tpd::GaussianEngine & friends
  torpedo_b200/{_lib,engine}.py            ctypes mirror of the same interface (tests, bench)
There is no CPU fallback: importing works anywhere, but every call needs the built library and an sm_100 GPU.
"""
__all__ = ["Camera", "GaussianEngine", "PerspectiveCamera", "Scene", "Settings", "TpdError"]


def __getattr__(name):
    # resolved on first use: `import torpedo_b200.scenes` (pure numpy input generators, also used by the benchmark's CPU
    # reference arm) must not pull in the ctypes bindings of the native libraries
    if name in __all__:
        from . import engine
        return getattr(engine, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
