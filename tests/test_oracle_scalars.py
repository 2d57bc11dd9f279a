"""The oracle's scalar building blocks against independent numpy formulations / known answers."""
import ctypes as C

import numpy as np


def test_f16_to_f32_exhaustive(O):
    L = O.lib()
    h = np.arange(65536, dtype=np.uint32)
    got = np.array([L.gso_f16tof32(int(v)) for v in h], np.float32)
    want = h.astype(np.uint16).view(np.float16).astype(np.float32)
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(got), nan)      # (NaN payloads do not survive the float->double->float ctypes trip)
    assert np.array_equal(got.view(np.uint32)[~nan], want.view(np.uint32)[~nan])


def test_f32_to_f16_is_round_to_nearest_even(O):
    L = O.lib()
    rng = np.random.default_rng(0)
    vals = np.concatenate([
        rng.standard_normal(20000).astype(np.float32) * 10.0 ** rng.integers(-9, 6, 20000),
        np.array([0.0, -0.0, 1.0, 65504.0, 65519.9, 65520.0, 1e6, -1e6, np.inf, -np.inf, 5.96e-8, 2.98e-8, 2.9802322e-8, 6.1e-5,
                  6.0975552e-05, 1.0009765625, 1.00048828125, 1.00146484375], np.float32),
        (np.arange(1, 2049, dtype=np.float32) + 0.5) / 1024.0,   # exact ties
    ])
    got = np.array([L.gso_f32tof16(float(v)) for v in vals], np.uint32)
    with np.errstate(over="ignore"):
        want = vals.astype(np.float16).view(np.uint16).astype(np.uint32)
    assert np.array_equal(got, want)
    assert L.gso_f32tof16(float("nan")) & 0x7C00 == 0x7C00 and L.gso_f32tof16(float("nan")) & 0x3FF != 0


def test_exp_neg_accuracy_and_range(O):
    L = O.lib()
    xs = np.linspace(-8.0, 0.0, 20001).astype(np.float32)
    got = np.array([L.gso_exp_neg(float(x)) for x in xs], np.float64)
    want = np.exp(xs.astype(np.float64))
    assert np.max(np.abs(got / want - 1.0)) < 8e-7        # polynomial 2.2e-7 + rounding of x*log2e: the class of exp() = ex2.approx(x*log2e)
    assert 1.0 <= L.gso_exp_neg(0.0) <= 1.0000002
    assert L.gso_exp_neg(-200.0) >= 0.0 and L.gso_exp_neg(-200.0) < 1e-37


def test_float_to_sortable_uint_is_order_preserving(O):
    L = O.lib()
    rng = np.random.default_rng(1)
    f = np.concatenate([rng.standard_normal(5000).astype(np.float32) * 100, np.array([0.0, 1e-40, -1e-40, np.inf, -np.inf], np.float32)])
    k = np.array([L.gso_float_to_sortable_uint(float(v)) for v in f], np.uint64)
    o = np.argsort(f, kind="stable")
    assert np.all(np.diff(k[o].astype(np.int64)) >= 0)
    assert L.gso_float_to_sortable_uint(-0.0) < L.gso_float_to_sortable_uint(0.0)   # -0 sorts just below +0
    assert L.gso_float_to_sortable_uint(1.0) == 0xBF800000 and L.gso_float_to_sortable_uint(-1.0) == 0x407FFFFF


def test_inv_square_centered01_inverts_the_importers_transform(O):
    L = O.lib()

    def square_centered01(x):   # R/GaussianUtils.cs:25-30
        x = np.float32(x) - np.float32(0.5)
        x = x * (x * np.sign(x))
        return np.float32(x * np.float32(2.0) + np.float32(0.5))

    for v in np.linspace(0, 1, 101):
        assert abs(L.gso_inv_square_centered01(float(square_centered01(v))) - v) < 2e-4


def test_texture_swizzle_matches_the_packer_and_is_a_bijection(O, g):
    from unitygaussiansplatting_b200 import _native as N
    L, P = O.lib(), N.asset_lib()
    idx = np.concatenate([np.arange(0, 70000), np.arange(6_000_000, 6_000_600)])
    lin = []
    for i in idx:
        x, y = C.c_uint32(), C.c_uint32()
        t = L.gso_splat_index_to_pixel_index(int(i), C.byref(x), C.byref(y))
        assert t == y.value * 2048 + x.value == P.gsa_splat_index_to_texture_index(int(i))
        lin.append(t)
    assert len(set(lin)) == len(lin)
    # a 256-splat chunk covers exactly one 16x16 texel block
    blk = {(t % 2048) // 16 + ((t // 2048) // 16) * 128 for t in lin[:256]}
    assert len(blk) == 1
