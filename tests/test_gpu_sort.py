"""Stand-alone sorter (GpuSorting.Dispatch replacement): bit-exact, stable, all edge sizes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check(ctx, O, keys):
    n = keys.size
    payload = np.arange(n, dtype=np.uint32)
    k1, p1 = keys.copy(), payload.copy()
    ctx.sort_pairs(k1, p1)
    order = np.argsort(keys, kind="stable").astype(np.uint32)
    assert np.array_equal(p1, order), "payload order differs from a stable sort"
    assert np.array_equal(k1, keys[order])
    k2, p2 = keys.copy(), payload.copy()
    O.sort_pairs(k2, p2, threads=4)
    assert np.array_equal(k1, k2) and np.array_equal(p1, p2)


@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 33, 255, 4095, 4096, 4097, 8192, 12289, 100003, (1 << 20) + 3])
def test_random_keys(ctx, O, n):
    rng = np.random.default_rng(n)
    _check(ctx, O, rng.integers(0, 2**32, n, dtype=np.uint32))


@pytest.mark.parametrize("kind", ["equal", "sorted", "reversed", "few", "lowbits", "highbits", "ffff"])
def test_adversarial_keys(ctx, O, kind):
    n = 300000
    rng = np.random.default_rng(7)
    keys = {"equal": np.full(n, 0xDEADBEEF, np.uint32), "sorted": np.arange(n, dtype=np.uint32) * 7,
            "reversed": (np.arange(n, dtype=np.uint32) * 11)[::-1].copy(), "few": rng.integers(0, 3, n, dtype=np.uint32) * 0x01010101,
            "lowbits": rng.integers(0, 256, n, dtype=np.uint32), "highbits": rng.integers(0, 256, n, dtype=np.uint32) << 24,
            "ffff": np.where(rng.random(n) < 0.5, 0xFFFFFFFF, rng.integers(0, 2**32, n, dtype=np.uint32)).astype(np.uint32)}[kind]
    _check(ctx, O, keys)


def test_full_size_6m(ctx, O):
    """BASELINE configs[1] size: 6,131,954 pairs, depth-key-like distribution."""
    n = 6_131_954
    rng = np.random.default_rng(11)
    z = rng.normal(8.0, 6.0, n).astype(np.float32)
    u = z.view(np.uint32)
    keys = (u ^ np.where(u >> 31, 0xFFFFFFFF, 0x80000000).astype(np.uint32)).astype(np.uint32)
    payload = np.arange(n, dtype=np.uint32)
    k1, p1 = keys.copy(), payload.copy()
    ctx.sort_pairs(k1, p1)
    k2, p2 = keys.copy(), payload.copy()
    O.sort_pairs(k2, p2, threads=O.max_threads())
    assert np.array_equal(k1, k2) and np.array_equal(p1, p2)
    assert np.all(k1[1:] >= k1[:-1])
