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


def _trs(g, t, axis, angle_deg, s):
    axis = np.asarray(axis, np.float64) / np.linalg.norm(axis)
    h = np.radians(angle_deg) / 2
    q = np.array([*(axis * np.sin(h)), np.cos(h)])
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    m = np.eye(4)
    m[:3, :3] = R * s
    m[:3, 3] = t
    return m.astype(np.float32), q.astype(np.float32)


def test_baked_export_renders_like_the_transformed_asset(g, O, tmp_path):
    """The property that defines "bake transform" (S/SplatUtilities.compute:626-643): exporting with the transform baked
    in and re-importing gives an asset that, untransformed, is seen exactly like the original under the transform --
    positions, projected axes and, through the rotated SH bands, the view-dependent colour of every splat."""
    from util import camera, view_fields
    n = 400
    src = g.generate_input_splats(g.SCENE_CLUSTERED, n, 0x5EED0071)
    src[:, 0:3] *= 0.05                                             # keep the cloud in front of the camera
    src[:, 9:54] *= 3.0                                             # stronger view dependence than the generator's default
    a = g.create_asset(src.copy(), "VeryHigh")
    T, q = _trs(g, (0.4, -0.2, 0.3), (0.3, 1.0, -0.5), 67.0, 1.3)
    qd, sd = __import__("unitygaussiansplatting_b200.renderer", fromlist=["decompose_trs"]).decompose_trs(T)
    assert np.allclose(sd, 1.3, atol=1e-5) and min(np.abs(qd - q).max(), np.abs(qd + q).max()) < 1e-5
    rec = O.export_data(a)
    from unitygaussiansplatting_b200.renderer import bake_transform
    bake_transform(rec, T)
    path = tmp_path / "baked.ply"
    assert g.write_ply(str(path), rec) == n
    b = g.create_asset(g.read_ply(str(path)), "VeryHigh")
    cam = camera(g, 320, 240, fov=50.0, pos=(0.1, 0.2, -3.0))
    fa, _k1 = g.make_frame_params(cam, localToWorld=T, sh_order=3)
    fb, _k2 = g.make_frame_params(cam, sh_order=3)
    va, vb = view_fields(O.calc_view(a, fa)), view_fields(O.calc_view(b, fb))
    # match splats across the two assets (the importer re-sorts by Morton code of the new bounds) by clip-space position
    ka = np.lexsort(np.round(va["pos"][:, :3], 4).T)
    kb = np.lexsort(np.round(vb["pos"][:, :3], 4).T)
    assert np.allclose(va["pos"][ka], vb["pos"][kb], atol=2e-5)
    for ch in ("r", "g", "b", "a"):
        assert np.allclose(va[ch][ka], vb[ch][kb], rtol=2e-3, atol=2e-3), ch
    vis = va["pos"][ka][:, 3] > 0
    # projected ellipse axes: equal up to the 10-bit re-quantisation of the composed rotation, and up to sign
    for ax in ("axis1", "axis2"):
        A, B = va[ax][ka][vis], vb[ax][kb][vis]
        err = np.minimum(np.abs(A - B).max(1), np.abs(A + B).max(1)) / (np.linalg.norm(A, axis=1) + 1e-6)
        assert np.median(err) < 5e-3 and (err < 0.08).mean() > 0.97, (ax, np.median(err))
    # colours really are view dependent here: without rotating the SH the check above fails
    rec2 = O.export_data(a)
    unrot = rec2.copy()
    bake_transform(rec2, T)
    assert np.abs(rec2[:, 9:54] - unrot[:, 9:54]).max() > 0.05
    assert np.array_equal(rec2[:, 6:9], unrot[:, 6:9])              # band 0 is rotation invariant


def test_bake_negative_scale_flips_and_identity(g, O):
    from unitygaussiansplatting_b200.renderer import bake_transform
    a = g.synthetic_asset(g.SCENE_CLUSTERED, 300, 0x5EED0072, "VeryHigh")
    rec = O.export_data(a)
    same = bake_transform(rec.copy(), np.eye(4, dtype=np.float32))
    assert np.allclose(same, rec, atol=1e-6)                        # identity transform: nothing moves (SH matrices = I)
    m = np.diag([-1.0, 1.0, 1.0, 1.0]).astype(np.float32)           # mirror in x: rot.yz negated (:631-632), |scale| = 1
    mir = bake_transform(rec.copy(), m, rotation=np.array([0, 0, 0, 1], np.float32), scale=np.array([-1, 1, 1], np.float32))
    assert np.allclose(mir[:, 0], -rec[:, 0]) and np.allclose(mir[:, 1:3], rec[:, 1:3])
    assert np.allclose(mir[:, 58], rec[:, 58]) and np.allclose(mir[:, 59], rec[:, 59])           # w, x kept
    assert np.allclose(mir[:, 60:62], -rec[:, 60:62]) and np.allclose(mir[:, 55:58], rec[:, 55:58], atol=1e-6)
    with pytest.raises(ValueError):
        bake_transform(rec.copy(), np.zeros((4, 4), np.float32))
