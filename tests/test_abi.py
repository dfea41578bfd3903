"""The C-ABI library loads and exports exactly what include/tpdcu.h declares (no compute calls: no GPU here)."""
import ctypes as C
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tpdcu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tpdcu_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(built_libs):
    lib = C.CDLL(built_libs.TPDCU_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in include/tpdcu.h but not exported by libtpdcu.so"


def test_python_prototypes_cover_the_header(built_libs):
    assert sorted(built_libs.TPDCU_SYMBOLS) == declared_symbols()


def test_every_declaration_cites_the_reference():
    text = open(os.path.join(ROOT, "include", "tpdcu.h")).read()
    assert len(re.findall(r"GaussianEngine\.cpp:\d+|:\d+-\d+", text)) >= 15 and ".slang" in text


def test_library_is_sm100a_only(built_libs):
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", built_libs.TPDCU_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_oracle_in_product():
    """The product must not import, link or dlopen anything under oracle/."""
    bad = []
    for base in ("torpedo_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                    src = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"^\s*(from|import)\s+oracle\b|#\s*include\s*[\"<][^\">]*oracle|libtpd_oracle|libtpdref", src, flags=re.M):
                        bad.append(os.path.join(dirpath, f))
    out = subprocess.run(["ldd", os.path.join(ROOT, "torpedo_b200", "lib", "libtpdcu.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out
    assert not bad, bad


def test_create_fails_loudly_without_gpu(built_libs):
    """No CPU fallback: without an sm_100 device the very first call reports an error."""
    import torch
    if torch.cuda.is_available():
        return
    lib = built_libs.tpdcu()
    h = C.c_void_p()
    rc = lib.tpdcu_create(0, C.byref(h))
    assert rc != 0 and not h.value
    assert lib.tpdcu_last_error()
