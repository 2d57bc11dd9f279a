"""Committed golden fixtures (tests/golden/*.npz, generator: tests/golden/make_golden.py).
CPU: the oracle and the packer still reproduce them.  GPU: the CUDA path reproduces them."""
import sys
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).parent / "golden"
sys.path.insert(0, str(GOLD))
import make_golden  # noqa: E402


@pytest.mark.parametrize("name", list(make_golden.CASES))
def test_oracle_and_packer_reproduce_golden(name):
    want = np.load(GOLD / (name + ".npz"))
    _asset, _cam, _sh, got = make_golden.build(name)
    for k in want.files:
        assert np.array_equal(got[k], want[k]), "%s: %s drifted from the committed fixture" % (name, k)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(make_golden.CASES))
def test_cuda_path_reproduces_golden(g, ctx, name):
    want = np.load(GOLD / (name + ".npz"))
    kind, n, seed, quality, w, h, pos, sh = make_golden.CASES[name]
    from util import camera
    asset = g.synthetic_asset(kind, n, seed, quality)
    r = g.GaussianSplatRenderer(asset, ctx)
    r.m_SHOrder = sh
    rt = np.zeros((h, w, 4), np.float16)
    tgt = np.full((h, w, 4), 0.25, np.float32)
    r.SortAndRenderSplats(camera(g, w, h, pos=pos), rt=rt, camera_target=tgt)
    assert np.array_equal(r.readback_keys(), want["keys"])
    assert np.array_equal(r.readback_order(), want["order"])
    r.CalcViewData(camera(g, w, h, pos=pos))
    assert np.array_equal(r.readback_view(), want["view"])
    assert np.abs(rt.astype(np.float32) - want["rt"].astype(np.float32)).max() <= 1e-3
    assert np.array_equal(rt, want["rt"])
    assert np.array_equal(tgt, want["composite"])
    r.Dispose()
