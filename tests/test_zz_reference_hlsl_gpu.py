"""The CUDA path against the reference's own shader code (oracle/_ref/libref_hlsl.so, prebuilt here and shipped with the
tree: the GPU box has no /root/reference).  Same comparisons as tests/test_reference_hlsl.py, with the C-ABI library's
output in place of the oracle's.  Named test_zz_* so that it runs after the parity tests proper."""
import numpy as np
import pytest

from test_reference_hlsl import _cov, _unsortable
from util import camera, view_fields

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("quality", ["Medium", "VeryHigh"])
def test_cuda_view_data_and_keys_match_the_reference_shader_code(g, O, ctx, quality):
    if O.ref_hlsl() is None:
        pytest.skip("oracle/_ref/libref_hlsl.so did not travel")
    n = 20000
    asset = g.synthetic_asset(g.SCENE_CLUSTERED, n, 0x5EED0091, quality)
    cam = camera(g, 320, 240)
    r = g.GaussianSplatRenderer(asset, ctx)
    r.SortPoints(cam)
    r.CalcViewData(cam)
    got, keys, order = view_fields(r.readback_view()), r.readback_keys(), r.readback_order()
    fp, _keep = g.make_frame_params(cam, r.localToWorldMatrix, r.m_SplatScale, r.m_OpacityScale, r.m_SHOrder, r.m_SHOnly)
    ref = view_fields(O.ref_calc_view(asset, fp))
    assert np.array_equal(ref["pos"][:, 3] <= 0, got["pos"][:, 3] <= 0)
    assert (np.abs(ref["pos"] - got["pos"]).max(1) <= 2e-6 * (1 + np.abs(got["pos"]).max(1))).all()
    vis = got["pos"][:, 3] > 0
    for ch in "rgb":
        d = np.abs(ref[ch][vis] - got[ch][vis])
        assert (d <= np.maximum(np.abs(got[ch][vis]), 2.0 ** -14) * 2.0 ** -9).all(), ch
    assert np.array_equal(ref["a"][vis], got["a"][vis])
    cr, cg = _cov(ref)[vis], _cov(got)[vis]
    rel = np.abs(cr - cg).reshape(-1, 4).max(1) / (cg[:, 0, 0] + cg[:, 1, 1])
    assert np.percentile(rel, 50) < 1e-6 and np.percentile(rel, 99) < 2e-5 and rel.max() < 2e-3
    # sorted keys: the reference's keys for the same (sorted) order are the same depths to a couple of ulp, and ascending
    kr = O.ref_calc_distances(asset, fp, order)
    assert np.abs(_unsortable(kr) - _unsortable(keys)).max() <= 4e-6
    assert (np.diff(_unsortable(kr).astype(np.float64)) >= -8e-6).all()
    r.Dispose()
