"""The oracle against THE REFERENCE'S OWN SHADER CODE.

oracle/_ref/libref_hlsl.so is package/Shaders/{GaussianSplatting.hlsl, SplatUtilities.compute, SphericalHarmonics.hlsl,
RenderGaussianSplats.shader} themselves -- read from /root/reference, syntactically rewritten to C++ spelling and compiled with
g++ against a shim of HLSL's types and intrinsics (oracle/refhlsl/).  Expressions, constants, operation order and branches
are the reference's; scalar arithmetic is IEEE float32 without contraction.  The oracle (and the CUDA path, bit-identical to
it) evaluates the same formulas under its own arithmetic contract (explicit fmaf chains, reciprocal constants), so the two
agree to float rounding, not to the bit: tolerances below are a few ulp, scaled by conditioning where a formula cancels."""
import numpy as np
import pytest

from util import camera, one_splat, view_fields


@pytest.fixture(scope="module")
def R(O):
    if O.ref_hlsl() is None:
        pytest.skip("oracle/_ref/libref_hlsl.so not built and /root/reference not present")
    return O


def _unsortable(keys):
    k = keys.astype(np.uint32)
    u = np.where(k & 0x80000000, k ^ np.uint32(0x80000000), ~k).astype(np.uint32)
    return u.view(np.float32)


def _cov(v):
    a1, a2 = v["axis1"].astype(np.float64), v["axis2"].astype(np.float64)
    return 0.5 * (a1[:, :, None] * a1[:, None, :] + a2[:, :, None] * a2[:, None, :])


@pytest.mark.parametrize("quality", ["Medium", "VeryHigh", "High", "Low", "custom-f16sh", "custom-norm6pos"])
def test_view_data_matches_the_reference_shader_code(g, R, quality):
    n = 20000
    if quality.startswith("custom"):     # format combinations no preset uses: Float16 SH, Norm6 positions, Norm16 scale
        fmts = {"custom-f16sh": (g.VectorFormat.Norm16, g.VectorFormat.Norm6, g.ColorFormat.Float16x4, g.SHFormat.Float16),
                "custom-norm6pos": (g.VectorFormat.Norm6, g.VectorFormat.Norm16, g.ColorFormat.Float32x4, g.SHFormat.Norm11)}[quality]
        asset = g.create_asset(g.generate_input_splats(g.SCENE_CLUSTERED, n, 0x5EED0091), formats=fmts)
    else:
        asset = g.synthetic_asset(g.SCENE_CLUSTERED, n, 0x5EED0091, quality)
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = np.array([[0.8, -0.6, 0.0], [0.6, 0.8, 0.0], [0.0, 0.0, 1.0]], np.float32) * 1.1
    T[:3, 3] = (0.3, -0.1, 0.2)
    box = np.diag([1 / 9.0, 1 / 9.0, 1 / 9.0, 1.0]).astype(np.float32)
    deleted = np.zeros((n + 31) // 32, np.uint32)
    deleted[3] = 0xF0F0F0F0
    fp, _keep = g.make_frame_params(camera(g, 320, 240), localToWorld=T, splat_scale=0.9, opacity_scale=1.3, sh_order=3,
                                    cutouts=[(box, 1)], deleted_bits=deleted, splat_count=n)
    ref, ora = view_fields(R.ref_calc_view(asset, fp)), view_fields(R.calc_view(asset, fp))
    # which splats are culled (w = 0: deleted / cut) or behind the camera is decided identically
    assert np.array_equal(ref["pos"][:, 3] <= 0, ora["pos"][:, 3] <= 0)
    assert ((ref["pos"][:, 3] == 0) & (ora["pos"][:, 3] == 0)).sum() > 100
    assert (np.abs(ref["pos"] - ora["pos"]).max(1) <= 2e-6 * (1 + np.abs(ora["pos"]).max(1))).all()
    vis = ora["pos"][:, 3] > 0
    for ch in "rgb":                                         # half-precision results: at most one unit in the last place apart
        d = np.abs(ref[ch][vis] - ora[ch][vis])
        assert (d <= np.maximum(np.abs(ora[ch][vis]), 2.0 ** -14) * 2.0 ** -9).all(), ch     # <= 2 half ulps
        assert (d == 0).mean() > 0.9
    assert np.array_equal(ref["a"][vis], ora["a"][vis])
    # the two screen axes, through the covariance they encode (eigenvectors of near-circular footprints are ill-defined)
    cr, co = _cov(ref)[vis], _cov(ora)[vis]
    rel = np.abs(cr - co).reshape(-1, 4).max(1) / (co[:, 0, 0] + co[:, 1, 1])
    assert np.percentile(rel, 50) < 1e-6 and np.percentile(rel, 99) < 2e-5 and rel.max() < 2e-3


@pytest.mark.parametrize("quality", ["Medium", "VeryHigh"])
def test_sort_keys_match_the_reference_shader_code(g, R, quality):
    n = 30000
    asset = g.synthetic_asset(g.SCENE_CLUSTERED, n, 0x5EED0092, quality)
    fp, _keep = g.make_frame_params(camera(g, 320, 240))
    order = np.random.default_rng(1).permutation(n).astype(np.uint32)
    kr, ko = R.ref_calc_distances(asset, fp, order), R.calc_distances(asset, fp, order)
    zr, zo = _unsortable(kr), _unsortable(ko)
    assert np.abs(zr - zo).max() <= 4e-6 and (kr == ko).mean() > 0.5     # view-space depth of +-36: a couple of ulp
    assert np.array_equal(np.argsort(kr, kind="stable")[:100], np.argsort(ko, kind="stable")[:100]) or (kr == ko).mean() < 1.0


def test_export_and_baked_transform_match_the_reference_shader_code(g, R):
    """CSExportData incl. the _ExportTransformFlags branch: QuatMul, scale, and RotateSH (the closed form after sh-lib,
    S/SphericalHarmonics.hlsl) against the oracle's export and the product's gsa_bake_transform (least-squares band matrices)."""
    from unitygaussiansplatting_b200.renderer import bake_transform, decompose_trs
    n = 3000
    asset = g.synthetic_asset(g.SCENE_CLUSTERED, n, 0x5EED0093, "VeryHigh")
    fp, _keep = g.make_frame_params(camera(g, 64, 64))
    plain_ref, plain_ora = R.ref_export(asset, fp), R.export_data(asset, fp)
    assert np.allclose(plain_ref, plain_ora, rtol=2e-6, atol=2e-6)
    ang = np.radians(50.0)
    Rm = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]]) @ \
        np.array([[1, 0, 0], [0, np.cos(0.4), -np.sin(0.4)], [0, np.sin(0.4), np.cos(0.4)]])
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = (Rm * 1.25).astype(np.float32)
    T[:3, 3] = (0.5, 0.25, -0.75)
    q, s = decompose_trs(T)
    fpT, _k2 = g.make_frame_params(camera(g, 64, 64), localToWorld=T)
    baked_ref = R.ref_export(asset, fpT, bake=True, rotation=q, scale=s)
    baked_ours = bake_transform(R.export_data(asset, fp), T, rotation=q, scale=s)
    assert np.allclose(baked_ref[:, 0:3], baked_ours[:, 0:3], atol=1e-5)                  # positions
    assert np.allclose(baked_ref[:, 55:58], baked_ours[:, 55:58], atol=1e-5)              # log scale
    same = np.abs(baked_ref[:, 58:62] - baked_ours[:, 58:62]).max(1)
    flip = np.abs(baked_ref[:, 58:62] + baked_ours[:, 58:62]).max(1)
    assert np.minimum(same, flip).max() < 1e-5                                            # rotation (q and -q are the same)
    assert np.array_equal(baked_ref[:, 6:9], baked_ours[:, 6:9])                          # band 0
    assert np.abs(baked_ref[:, 9:54] - plain_ref[:, 9:54]).max() > 0.05                   # the rotation does something ...
    assert np.abs(baked_ref[:, 9:54] - baked_ours[:, 9:54]).max() < 2e-5                  # ... and ours is the same rotation


def test_draw_stages_match_the_reference_shader_code(g, R):
    """vert + frag of RenderGaussianSplats.shader on one rotated, anisotropic splat: the quad the reference emits maps pixel
    centres to quad coordinates (the rasteriser's linear interpolation), frag gives the fragment; the oracle's render of that
    splat (fp32 blend, so the image IS the fragment) must be the same picture."""
    W, H = 200, 150
    cam = camera(g, W, H, fov=50.0, pos=(0.1, 0.0, -3.0))
    asset = one_splat(g, pos=(0.2, -0.1, 0.3), scale=(0.25, 0.08, 0.05), quat=(0.3, 0.5, -0.2, 0.78), opacity=0.65, dc0=(0.7, 0.5, 0.9))
    fp, _keep = g.make_frame_params(cam, sh_order=0)
    view = R.calc_view(asset, fp)
    order = np.arange(1, dtype=np.uint32)
    rt = R.render(view, order, W, H, blend_mode=1)
    clip, qpos, col = R.ref_vert(view, order, 0, W, H)
    assert np.array_equal(np.abs(qpos), np.full((4, 2), 2.0, np.float32))                 # quad corners at +-2 (:54-55)
    # pixel position of each corner; interpolation inside a parallelogram is affine: solve quad coords from three corners
    px = np.stack([(clip[:, 0] / clip[:, 3] * 0.5 + 0.5) * W, (0.5 - 0.5 * clip[:, 1] / clip[:, 3]) * H], 1).astype(np.float64)
    A = np.linalg.solve(np.column_stack([px[:3], np.ones(3)]), qpos[:3].astype(np.float64))   # [x y 1] @ A = q
    assert np.allclose(np.array([*px[3], 1.0]) @ A, qpos[3], atol=1e-4)
    checked = drawn = 0
    for y in range(0, H):
        for x in range(0, W):
            qx, qy = np.array([x + 0.5, y + 0.5, 1.0]) @ A
            if abs(qx) > 2.3 or abs(qy) > 2.3:
                assert rt[y, x, 3] == 0
                continue
            edge = min(abs(abs(qx) - 2), abs(abs(qy) - 2)) < 2e-3
            out, discarded = R.ref_frag(col, float(qx), float(qy))
            inside = abs(qx) <= 2 and abs(qy) <= 2
            want = np.zeros(4, np.float32) if (discarded or not inside) else out
            if edge or abs(float(np.exp(-(qx * qx + qy * qy))) * col[3] - 1 / 255) < 2e-5:
                continue                                                                   # on the quad edge / discard threshold
            assert np.abs(rt[y, x] - want).max() < 3e-6, (x, y, rt[y, x], want)
            checked += 1
            drawn += int(want[3] > 0)
    assert checked > 700 and drawn > 400


def test_selected_splat_branch_matches_the_reference_shader_code(g, R):
    """With an edit selection bound the reference's vertex shader hands a selected splat to the pixel shader with col.a = -1
    (S/RenderGaussianSplats.shader:63-73) and the pixel shader takes its "selected" branch (:87-101: opacity from the gaussian
    alone, +0.3, a solid magenta ring where exp(power) is in (7/255, 10/255), magenta tint).  The oracle's draw with
    selected_bits must give the picture the compiled reference vert + frag give."""
    W, H = 200, 150
    cam = camera(g, W, H, fov=50.0, pos=(0.1, 0.0, -3.0))
    asset = one_splat(g, pos=(0.2, -0.1, 0.3), scale=(0.25, 0.08, 0.05), quat=(0.3, 0.5, -0.2, 0.78), opacity=0.02, dc0=(0.7, 0.5, 0.9), n_pad=40)
    fp, _keep = g.make_frame_params(cam, sh_order=0)
    view = R.calc_view(asset, fp)
    order = np.arange(asset.splatCount, dtype=np.uint32)
    bits = np.zeros(2, np.uint32)
    bits[0] = 1                                                                            # splat 0 selected, the padding splats not
    col = R.ref_vert_selected(view, order, 0, W, H, bits)
    assert col[3] == -1.0
    assert R.ref_vert_selected(view, order, 1, W, H, bits)[3] >= 0.0
    plain = R.render(view, order[:1], W, H, blend_mode=1)
    rt = R.render(view, order[:1], W, H, blend_mode=1, selected_bits=bits)
    assert plain[..., 3].max() < 0.03 and rt[..., 3].max() == 1.0                          # opacity 0.02 alone is almost nothing
    clip, qpos, _col = R.ref_vert(view, order, 0, W, H)
    px = np.stack([(clip[:, 0] / clip[:, 3] * 0.5 + 0.5) * W, (0.5 - 0.5 * clip[:, 1] / clip[:, 3]) * H], 1).astype(np.float64)
    A = np.linalg.solve(np.column_stack([px[:3], np.ones(3)]), qpos[:3].astype(np.float64))
    checked = ring = 0
    for y in range(H):
        for x in range(W):
            qx, qy = np.array([x + 0.5, y + 0.5, 1.0]) @ A
            if abs(qx) > 2.3 or abs(qy) > 2.3:
                assert rt[y, x, 3] == 0
                continue
            e = float(np.exp(-(qx * qx + qy * qy)))
            if min(abs(abs(qx) - 2), abs(abs(qy) - 2)) < 2e-3 or min(abs(e - 1 / 255), abs(e - 7 / 255), abs(e - 10 / 255)) < 3e-5:
                continue                                                                   # on the quad edge / one of the three thresholds
            out, discarded = R.ref_frag(col, float(qx), float(qy))
            inside = abs(qx) <= 2 and abs(qy) <= 2
            want = np.zeros(4, np.float32) if (discarded or not inside) else out
            assert np.abs(rt[y, x] - want).max() < 3e-6, (x, y, rt[y, x], want)
            checked += 1
            ring += int(want[3] == 1.0 and want[0] == 1.0 and want[1] == 0.0)
    assert checked > 700 and ring > 20


def test_rotation_packing_matches_the_reference_shader_code(g, R):
    """The importer packs rotations with C# twins of these HLSL functions; the packer (gsa_pack_smallest3 + its 10.10.10.2
    encoder, checked through a one-splat asset) must produce the same code words, and decoding must agree."""
    import ctypes as C
    from unitygaussiansplatting_b200 import _native as N
    L, lib = R.ref_hlsl(), N.asset_lib()
    rng = np.random.default_rng(3)
    for _ in range(300):
        q = rng.standard_normal(4).astype(np.float32)
        q /= np.linalg.norm(q)
        ours, ref = np.zeros(4, np.float32), np.zeros(4, np.float32)
        lib.gsa_pack_smallest3(q.ctypes.data, ours.ctypes.data)
        enc_ref = L.refhlsl_pack_rotation(q.ctypes.data, ref.ctypes.data)
        assert np.allclose(ours, ref, atol=1e-7)
        enc_ours = (int(ours[0] * 1023.5) | (int(ours[1] * 1023.5) << 10) | (int(ours[2] * 1023.5) << 20) | (int(ours[3] * 3.5) << 30)) & 0xFFFFFFFF
        assert enc_ours == enc_ref or np.abs(ours - ref).max() > 0      # same code word whenever the packed floats are identical
        back = np.zeros(4, np.float32)
        L.refhlsl_decode_rotation(enc_ref, back.ctypes.data)
        assert min(np.abs(back - q).max(), np.abs(back + q).max()) < 2.5e-3                # 10-bit components
