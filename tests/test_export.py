"""PLY export (SURVEY 8f N3, E/GaussianSplatRendererEditor.cs:394-445 + CSExportData): record layout, alive filtering and
the export -> import round trip.  CPU part uses the oracle's CSExportData; the GPU part checks gs_export_splats against it."""
import numpy as np
import pytest


def _quat_from_record(rec):
    """raw rot_0..3 = (w, x, y, z) -> unit xyzw"""
    q = np.stack([rec[:, 59], rec[:, 60], rec[:, 61], rec[:, 58]], 1).astype(np.float64)
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def test_export_record_layout_and_round_trip(g, O, tmp_path):
    n = 3000
    src = g.generate_input_splats(g.SCENE_CLUSTERED, n, 0x5EED0041)       # linearised importer records
    asset = g.create_asset(src.copy(), "VeryHigh")                          # lossless formats: only rot is quantised (10.10.10.2)
    rec = O.export_data(asset)
    assert rec.shape == (n, 62) and not rec[:, 3:6].any()
    path = tmp_path / "export.ply"
    assert g.write_ply(str(path), rec) == n
    head = path.read_bytes()[:2000].split(b"end_header\n")[0].decode()
    lines = head.split("\n")
    assert lines[0] == "ply" and lines[1] == "format binary_little_endian 1.0" and lines[2] == "element vertex %d" % n
    assert lines[3] == "property float x" and lines[9] == "property float f_dc_0" and lines[-2] == "property float rot_3" and len(lines) == 66
    back = g.read_ply(str(path))                                            # re-import: LinearizeData applied again
    again = g.create_asset(back.copy(), "VeryHigh")
    # the re-imported asset is the same asset up to log/exp, logit/sigmoid and the SH0 <-> colour affine map
    assert np.array_equal(again.posData, asset.posData)
    assert np.array_equal(again.shData, asset.shData)
    sc0 = asset.otherData[:n * 16].view(np.float32).reshape(n, 4)[:, 1:]
    sc1 = again.otherData[:n * 16].view(np.float32).reshape(n, 4)[:, 1:]
    assert np.allclose(sc0, sc1, rtol=2e-6)
    rq0, rq1 = asset.otherData[:n * 16].view(np.uint32).reshape(n, 4)[:, 0], again.otherData[:n * 16].view(np.uint32).reshape(n, 4)[:, 0]
    same = (rq0 == rq1).mean()
    assert same > 0.9, "10-bit quaternions re-quantise to the same code except where rounding sits on a boundary (%.3f)" % same
    c0, c1 = asset.colorData.view(np.float32), again.colorData.view(np.float32)
    assert np.allclose(c0, c1, atol=3e-6)


def test_export_drops_deleted_and_cut_splats(g, O, tmp_path):
    n = 2000
    asset = g.synthetic_asset(g.SCENE_CLUSTERED, n, 0x5EED0042, "Medium")
    from util import camera
    # a box cutout around the origin, not inverted: everything OUTSIDE it is cut (S/SplatUtilities.compute:164-187)
    box = np.diag([1 / 4.0, 1 / 4.0, 1 / 4.0, 1.0]).astype(np.float32)
    fp, _keep = g.make_frame_params(camera(g, 64, 64), cutouts=[(box, 1)])
    rec = O.export_data(asset, fp)
    inside = (np.abs(rec[:, 0:3]) <= 4.0).all(axis=1)
    assert 0 < inside.sum() < n
    assert np.array_equal(rec[:, 3] != 0, ~inside) and np.array_equal(rec[:, 3], rec[:, 5])
    deleted = np.zeros((n + 31) // 32, np.uint32)
    victims = np.nonzero(inside)[0][::3]
    for v in victims:
        deleted[v >> 5] |= np.uint32(1 << (v & 31))
    path = tmp_path / "edit.ply"
    alive = g.write_ply(str(path), rec, deleted)
    assert alive == inside.sum() - len(victims)
    back = g.read_ply(str(path))
    keep = inside.copy()
    keep[victims] = False
    assert np.array_equal(back[:, 0:3], rec[keep, 0:3])          # order preserved, positions untouched by LinearizeData
    with pytest.raises(ValueError):
        g.write_ply(str(path), rec[:, :10])


@pytest.mark.gpu
@pytest.mark.parametrize("quality", ["VeryLow", "Medium", "High", "VeryHigh"])
def test_gpu_export_matches_oracle(g, O, ctx, quality, tmp_path):
    n = 20000
    asset = g.synthetic_asset(g.SCENE_CLUSTERED, n, 0x5EED0043, quality)
    r = g.GaussianSplatRenderer(asset, ctx)
    box = np.diag([1 / 6.0, 1 / 6.0, 1 / 6.0, 1.0]).astype(np.float32)
    r.m_Cutouts = [(box, 1)]
    got = r.EditExportData()
    from util import camera
    fp, _keep = g.make_frame_params(camera(g, 64, 64), cutouts=r.m_Cutouts)
    ref = O.export_data(asset, fp, threads=O.max_threads())
    exact = list(range(0, 6)) + list(range(9, 54)) + list(range(58, 62))     # pos, nor, SH, rot: pure decode -> bit-exact
    assert np.array_equal(got[:, exact].view(np.uint32), ref[:, exact].view(np.uint32))
    assert np.array_equal(got[:, 6:9].view(np.uint32), ref[:, 6:9].view(np.uint32))          # ColorToSH0: IEEE sub + div
    # log(): libm on the CPU, CUDA's logf on the GPU (both <= 1-2 ulp): the only toleranced values on this path
    assert np.allclose(got[:, 54:58], ref[:, 54:58], rtol=2e-6, atol=2e-6)
    assert 0 < (got[:, 3] != 0).sum() < n
    assert r.ExportPlyFile(str(tmp_path / "gpu.ply")) == int((got[:, 3] == 0).sum())
    with pytest.raises(g.GsError) as e:
        r.EditExportData(bakeTransform=True)
    assert e.value.code == -4
    r.Dispose()
