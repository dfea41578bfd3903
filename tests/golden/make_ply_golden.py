"""Regenerates tests/golden/ply/*: small 3DGS-style .ply fixtures and, beside each, the records the REFERENCE's own
GaussianPoint::fromModel (volumetric/src/GaussianGeometry.cpp:59-127 + src/miniply.cpp, compiled in place into
oracle/_ref/libtpdref.so by oracle/Makefile) produces for it. Run HERE (the container that has /root/reference), never on
the GPU box. tests/test_host_layer.py then holds include/torpedo_b200/PlyLoader.hpp to these bytes everywhere.

    python tests/golden/make_ply_golden.py
"""
import ctypes as C
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ply")
N = 64
BASE = ["x", "y", "z", "f_dc_0", "f_dc_1", "f_dc_2", "opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]


def ref_from_model(path, capacity=1 << 20):
    lib = C.CDLL(O._REF_PATH)
    lib.tpdref_from_model.restype = C.c_int64
    lib.tpdref_from_model.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64]
    n = lib.tpdref_from_model(path.encode(), None, 0)
    if n < 0:
        raise RuntimeError(f"the reference's fromModel threw on {path}")
    out = np.zeros((n, 60), dtype=np.float32)
    lib.tpdref_from_model(path.encode(), out.ctypes.data, n)
    return out


def columns(seed, rest_count):
    rng = np.random.default_rng(seed)
    raw = {name: rng.normal(size=N).astype(np.float32) for name in BASE}
    raw["opacity"] = (raw["opacity"] * 3).astype(np.float32)          # sigmoid over a useful range
    raw["scale_0"] = (raw["scale_0"] - 4).astype(np.float32)           # log-scales like a trained cloud
    raw["scale_1"] = (raw["scale_1"] - 4).astype(np.float32)
    raw["scale_2"] = (raw["scale_2"] - 4).astype(np.float32)
    for k in range(rest_count):
        raw[f"f_rest_{k}"] = (rng.normal(size=N) * 0.05).astype(np.float32)
    return raw


def trainer_order(rest_count):
    """x y z nx ny nz f_dc_* f_rest_* opacity scale_* rot_*: the order 3DGS trainers write."""
    return (["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"] + [f"f_rest_{k}" for k in range(rest_count)] +
            ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"])


def write_ply(path, raw, names, fmt="binary_little_endian", types=None, tail_element=False, crlf=False):
    types = types or {}
    nl = "\r\n" if crlf else "\n"
    header = f"ply{nl}format {fmt} 1.0{nl}comment torpedo-b200 fixture{nl}element vertex {N}{nl}"
    header += "".join(f"property {types.get(name, 'float')} {name}{nl}" for name in names)
    if tail_element:
        header += f"element face 2{nl}property list uchar int vertex_indices{nl}"
    header += f"end_header{nl}"
    code = {"float": "f", "double": "d", "uchar": "B", "int": "i", "short": "h"}
    with open(path, "wb") as f:
        f.write(header.encode())
        for i in range(N):
            vals = [raw.get(name, np.zeros(N, np.float32))[i] for name in names]
            if fmt == "ascii":
                f.write((" ".join(repr(float(v)) if types.get(name, "float") in ("float", "double") else str(int(v))
                                  for name, v in zip(names, vals)) + "\n").encode())
            else:
                for name, v in zip(names, vals):
                    t = types.get(name, "float")
                    f.write(struct.pack("<" + code[t], float(v) if t in ("float", "double") else int(v)))
        if tail_element:
            if fmt == "ascii":
                f.write(b"3 0 1 2\n3 1 2 3\n")
            else:
                f.write(struct.pack("<Biii", 3, 0, 1, 2) + struct.pack("<Biii", 3, 1, 2, 3))


def fixtures():
    out = {}
    out["sh3_binary"] = dict(raw=columns(1, 45), names=trainer_order(45))
    out["sh2_binary"] = dict(raw=columns(2, 24), names=trainer_order(24))
    out["sh0_binary"] = dict(raw=columns(3, 0), names=trainer_order(0))
    out["sh3_ascii"] = dict(raw=columns(4, 45), names=trainer_order(45), fmt="ascii")
    # doubles, an integer column, a face element after the vertices, CRLF header: what exporters other than the trainer write
    raw = columns(5, 9)
    raw["label"] = np.arange(N).astype(np.float32) % 200
    names = ["label"] + trainer_order(9)
    out["sh1_mixed_types"] = dict(raw=raw, names=names, types={"x": "double", "y": "double", "z": "double", "label": "uchar", "opacity": "double"},
                                  tail_element=True, crlf=True)
    # properties shuffled: rotation and scale first; the 24 properties after f_dc_2 are what the reference takes as the rest
    names = ["rot_0", "rot_1", "rot_2", "rot_3", "scale_0", "scale_1", "scale_2", "opacity", "x", "y", "z", "f_dc_0", "f_dc_1", "f_dc_2"] + \
            [f"f_rest_{k}" for k in range(24)]
    out["sh2_shuffled"] = dict(raw=columns(6, 24), names=names)
    return out


def main():
    os.makedirs(HERE, exist_ok=True)
    for name, spec in fixtures().items():
        path = os.path.join(HERE, name + ".ply")
        write_ply(path, spec["raw"], spec["names"], spec.get("fmt", "binary_little_endian"), spec.get("types"),
                  spec.get("tail_element", False), spec.get("crlf", False))
        recs = ref_from_model(path)
        assert recs.shape == (N, 60), recs.shape
        np.save(os.path.join(HERE, name + ".ref.npy"), recs)
        print(name, os.path.getsize(path), "bytes ->", recs.shape)


if __name__ == "__main__":
    main()
