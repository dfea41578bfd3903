"""GPU parity of the standalone onesweep sort (tpdcu_sort_pairs_device) against numpy / the oracle's stable sort."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(built_libs):
    lib = built_libs.tpdcu()
    h = C.c_void_p()
    built_libs.check(lib.tpdcu_create(0, C.byref(h)))
    yield lib, h, built_libs
    lib.tpdcu_destroy(h)


def gpu_sort(ctx, keys, vals, end_bit):
    import torch
    lib, h, L = ctx
    dk = torch.from_numpy(keys.view(np.int64).copy()).cuda()
    dv = torch.from_numpy(vals.view(np.int32).copy()).cuda()
    torch.cuda.synchronize()
    L.check(lib.tpdcu_sort_pairs_device(h, dk.data_ptr(), dv.data_ptr(), len(keys), end_bit, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ms, passes = C.c_float(0), C.c_uint32(0)
    L.check(lib.tpdcu_sort_last_ms(h, C.byref(ms), C.byref(passes)))
    return dk.cpu().numpy().view(np.uint64), dv.cpu().numpy().view(np.uint32), ms.value, passes.value


def expect(keys, vals, end_bit):
    masked = keys if end_bit >= 64 else keys & np.uint64((1 << end_bit) - 1)
    order = np.argsort(masked, kind="stable")
    return keys[order], vals[order]


@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 33, 4095, 4096, 4097, 8192, 100_003, 1_000_000])
def test_random_keys_all_sizes(ctx, n):
    rng = np.random.default_rng(n + 1)
    keys = rng.integers(0, 2 ** 45, size=n, dtype=np.uint64)
    vals = np.arange(n, dtype=np.uint32)
    k, v, _, _ = gpu_sort(ctx, keys, vals, 45)
    ek, ev = expect(keys, vals, 45)
    assert (k == ek).all() and (v == ev).all()


@pytest.mark.parametrize("end_bit", [1, 7, 8, 9, 16, 33, 45, 46, 48, 64])
def test_bit_ranges_and_stability(ctx, end_bit):
    rng = np.random.default_rng(end_bit)
    n = 200_000
    keys = rng.integers(0, 2 ** 63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    keys[::3] = keys[0]            # many exact duplicates: ties must keep their input order
    vals = rng.integers(0, 2 ** 32, size=n, dtype=np.uint32)
    k, v, _, _ = gpu_sort(ctx, keys, vals, end_bit)
    ek, ev = expect(keys, vals, end_bit)
    assert (k == ek).all() and (v == ev).all()


def test_skewed_and_degenerate_distributions(ctx):
    n = 300_000
    vals = np.arange(n, dtype=np.uint32)
    rng = np.random.default_rng(9)
    cases = {
        "all_equal": np.full(n, 0x1234_5678_9ABC, dtype=np.uint64),
        "sorted": np.arange(n, dtype=np.uint64) << np.uint64(7),
        "reversed": (np.arange(n, dtype=np.uint64)[::-1].copy()) << np.uint64(3),
        "two_values": np.where(rng.random(n) < 0.5, np.uint64(5), np.uint64(5 << 40)).astype(np.uint64),
        "one_hot_bin": (rng.integers(0, 2, size=n, dtype=np.uint64) << np.uint64(44)) | np.uint64(0xFF),
        "tile_like": (rng.integers(0, 8160, size=n, dtype=np.uint64) << np.uint64(32)) | rng.normal(4.0, 1.0, n).astype(np.float32).view(np.uint32).astype(np.uint64),
    }
    for name, keys in cases.items():
        k, v, _, passes = gpu_sort(ctx, keys, vals, 48)
        ek, ev = expect(keys, vals, 48)
        assert (k == ek).all() and (v == ev).all(), name
        if name == "all_equal":
            assert passes == 0       # every pass is an identity permutation and is skipped
        if name == "two_values":
            assert passes == 2


def test_full_size_properties_16M(ctx):
    """BASELINE-size sort (~16 M pairs): sortedness + stability + permutation checksums (size-independent properties)."""
    rng = np.random.default_rng(2026)
    n = 16_000_000
    tiles = rng.integers(0, 8160, size=n, dtype=np.uint64)
    depth = rng.uniform(0.2, 30.0, size=n).astype(np.float32).view(np.uint32).astype(np.uint64)
    keys = (tiles << np.uint64(32)) | depth
    keys[::5] = keys[::5] & np.uint64(0xFFFFFFFFFFFF0000)   # force plenty of ties
    vals = np.arange(n, dtype=np.uint32)
    k, v, ms, passes = gpu_sort(ctx, keys, vals, 45)
    assert (k[1:] >= k[:-1]).all()
    ties = k[1:] == k[:-1]
    assert (v[1:][ties] > v[:-1][ties]).all()
    assert (keys[v] == k).all()                                   # each value still carries its own key
    assert np.bitwise_xor.reduce(v) == np.bitwise_xor.reduce(vals) and int(v.astype(np.uint64).sum()) == int(vals.astype(np.uint64).sum())
    assert passes == 6 and ms > 0
