"""Analytic sanity checks of the CPU oracle (it is the parity pin, so it gets its own checks;
SURVEY.md 8c: the reference's own tests pin nothing here; tests/test_reference_hlsl.py adds the compiled shader source)."""
import numpy as np
import pytest

from util import camera, one_splat, view_fields


def _view(g, O, asset, cam, **kw):
    fp, _keep = g.make_frame_params(cam, **kw)
    return view_fields(O.calc_view(asset, fp)), fp


def test_isotropic_splat_near_the_optical_axis(g, O):
    cam = camera(g, 640, 480, fov=50.0, pos=(0, 0, -4.0))
    s = 0.07
    asset = one_splat(g, pos=(0.013, 0.021, 1.0), scale=(s, s, s), opacity=0.8, dc0=(0.9, 0.5, 0.2))
    v, fp = _view(g, O, asset, cam, sh_order=0)
    z = 5.0
    f = 640 * fp.mat_proj_gpu[0] / 2
    want = np.sqrt(2 * (f * f * s * s / (z * z) + 0.3))
    a1, a2 = v["axis1"][0], v["axis2"][0]
    assert abs(a1 @ a2) < 1e-3 * want * want                       # axes orthogonal
    assert abs(np.linalg.norm(a1) - want) < 2e-3 * want and abs(np.linalg.norm(a2) - want) < 2e-3 * want
    assert abs(v["pos"][0, 3] - z) < 1e-5                           # clip.w = view depth
    assert np.allclose([v["r"][0], v["g"][0], v["b"][0]], [0.9, 0.5, 0.2], atol=1e-3)   # SH order 0: colour = dc0
    assert abs(v["a"][0] - 0.8) < 1e-3


def test_on_axis_isotropic_splat_is_finite(g, O):
    """DecomposeCovariance normalises (offDiag, lambda1 - diag1); that vector is (0,0) only when cov is an exact multiple of
    the identity, which the 10-bit quaternion never produces (identity decodes to x = -0.0007), so axes stay finite."""
    cam = camera(g, 640, 480, fov=50.0, pos=(0, 0, -4.0))
    asset = one_splat(g, pos=(0, 0, 1.0), scale=(0.07, 0.07, 0.07))
    v, fp = _view(g, O, asset, cam)
    assert np.isfinite(v["axis1"][0]).all() and np.isfinite(v["axis2"][0]).all()
    assert abs(np.linalg.norm(v["axis1"][0]) - np.linalg.norm(v["axis2"][0])) < 0.05


def test_behind_camera_deleted_and_cut_splats_get_w_zero_or_negative(g, O):
    cam = camera(g, 320, 240, pos=(0, 0, -4.0))
    asset = one_splat(g, pos=(0.1, 0.1, -9.0))
    v, _ = _view(g, O, asset, cam)
    assert v["pos"][0, 3] < 0 and not v["axis1"][0].any() and v["a"][0] == 0
    asset = one_splat(g, pos=(0.1, 0.1, 1.0), n_pad=40)
    bits = np.zeros(2, np.uint32); bits[0] = 1
    v, _ = _view(g, O, asset, cam, deleted_bits=bits, splat_count=asset.splatCount)
    i0 = int(np.argmin(np.abs(v["pos"][:, 3] - 5.0)))        # Morton order may move the splat
    vd, _ = _view(g, O, asset, cam)
    assert (v["pos"][:, 3] == 0).sum() == 1 and (vd["pos"][:, 3] == 0).sum() == 0
    cut = [(g.trs((0, 0, 0), None, (0.5, 0.5, 0.5)).astype(np.float32), 0)]   # ellipsoid radius 2 around the origin: keeps inside
    vc, _ = _view(g, O, asset, cam, cutouts=cut)
    inside = np.abs(vc["pos"][:, 3] - 5.0) < 1.0
    assert (vc["pos"][inside, 3] > 0).all() and (vc["pos"][~inside, 3] == 0).all()
    cut[0] = (cut[0][0], 0 | 0x100)                                             # inverted: cuts the inside
    vi, _ = _view(g, O, asset, cam, cutouts=cut)
    assert (vi["pos"][inside, 3] == 0).all() and (vi["pos"][~inside, 3] != 0).all()
    assert i0 >= 0


def test_sh_band1_known_answer(g, O):
    cam = camera(g, 320, 240, pos=(0, 0, -4.0))
    sh = np.zeros((15, 3), np.float32)
    sh[0] = (0.3, 0.0, 0.0); sh[1] = (0.0, 0.2, 0.0); sh[2] = (0.0, 0.0, 0.1)      # sh1, sh2, sh3
    asset = one_splat(g, pos=(0.5, 0.25, 1.0), dc0=(0.5, 0.5, 0.5), sh=sh)
    v, _ = _view(g, O, asset, cam, sh_order=1)
    d = np.array([0.0, 0.0, -4.0]) - np.array([0.5, 0.25, 1.0])
    d = d / np.linalg.norm(d)            # objViewDir; ShadeSH negates it, then res += C1*(-sh1*y + sh2*z - sh3*x)
    x, y, z = -d
    c1 = 0.4886025
    want = [0.5 + c1 * (-0.3 * y), 0.5 + c1 * (0.2 * z), 0.5 + c1 * (-0.1 * x)]
    assert np.allclose([v["r"][0], v["g"][0], v["b"][0]], want, atol=1e-3)
    v0, _ = _view(g, O, asset, cam, sh_order=1, sh_only=True)
    assert np.allclose([v0["r"][0], v0["g"][0], v0["b"][0]], want, atol=1e-3)      # dc0 is 0.5 here


def test_knobs_scale_and_opacity(g, O):
    cam = camera(g, 320, 240, pos=(0, 0, -4.0))
    asset = one_splat(g, pos=(0.3, 0.2, 1.0), scale=(0.2, 0.05, 0.1), quat=(0.2, 0.3, 0.1, 0.9), opacity=0.5)
    v1, _ = _view(g, O, asset, cam)
    v2, _ = _view(g, O, asset, cam, splat_scale=2.0, opacity_scale=1.5)
    l1, l2 = np.linalg.norm(v1["axis1"][0]), np.linalg.norm(v2["axis1"][0])
    assert abs((l2 * l2 / 2 - 0.3) / (l1 * l1 / 2 - 0.3) - 4.0) < 0.02             # cov scales with splatScale^2, low-pass does not
    assert abs(v2["a"][0] - 0.75) < 1e-3
    v3, _ = _view(g, O, asset, cam, opacity_scale=20.0)
    assert v3["a"][0] == 10.0                                                       # min(opacity*scale, 65000), no clamp to 1 here


def test_projected_orientation_and_centre(g, O):
    """A splat stretched along world +x+y must appear stretched along the image direction its endpoints project to."""
    cam = camera(g, 400, 300, fov=45.0, pos=(0, 0, -5.0))
    c45, s45 = np.cos(np.pi / 8), np.sin(np.pi / 8)     # rotate x axis by 45 deg about z
    asset = one_splat(g, pos=(0.4, -0.3, 0.0), scale=(0.5, 0.02, 0.02), quat=(0, 0, s45, c45), opacity=1.0, dc0=(1, 1, 1))
    fp, _ = g.make_frame_params(cam)
    view = O.calc_view(asset, fp)
    rt = O.render(view, np.arange(asset.splatCount, dtype=np.uint32), 400, 300)
    a = rt[..., 3].astype(np.float64)
    ys, xs = np.mgrid[0:300, 0:400]
    m = a.sum()
    cx, cy = (a * (xs + 0.5)).sum() / m, (a * (ys + 0.5)).sum() / m
    v = view_fields(view)
    ndc = v["pos"][0, :2] / v["pos"][0, 3]
    assert abs(cx - (ndc[0] * 0.5 + 0.5) * 400) < 0.05 and abs(cy - (0.5 - 0.5 * ndc[1]) * 300) < 0.05

    def project(p):
        vp = np.array(fp.mat_proj_gpu[:], np.float64).reshape(4, 4).T @ np.array(fp.mat_view[:], np.float64).reshape(4, 4).T
        c = vp @ np.array([p[0], p[1], p[2], 1.0])
        return np.array([(c[0] / c[3] * 0.5 + 0.5) * 400, (0.5 - 0.5 * c[1] / c[3]) * 300])

    e = project((0.4 + 0.35, -0.3 + 0.35, 0.0)) - project((0.4 - 0.35, -0.3 - 0.35, 0.0))
    e /= np.linalg.norm(e)
    mxx = (a * (xs + 0.5 - cx) ** 2).sum(); myy = (a * (ys + 0.5 - cy) ** 2).sum(); mxy = (a * (xs + 0.5 - cx) * (ys + 0.5 - cy)).sum()
    w, vec = np.linalg.eigh(np.array([[mxx, mxy], [mxy, myy]]))
    major = vec[:, 1]
    assert abs(abs(major @ e) - 1.0) < 1e-3 and w[1] > 20 * w[0]


def test_single_splat_pixel_values(g, O):
    cam = camera(g, 200, 150, pos=(0, 0, -4.0))
    asset = one_splat(g, pos=(0.02, 0.01, 0.0), scale=(0.08, 0.08, 0.08), opacity=0.6, dc0=(0.8, 0.4, 0.2))
    fp, _ = g.make_frame_params(cam, sh_order=0)
    view = O.calc_view(asset, fp)
    v = view_fields(view)
    rt = O.render(view, np.arange(asset.splatCount, dtype=np.uint32), 200, 150, blend_mode=1)
    ndc = v["pos"][0, :2] / v["pos"][0, 3]
    cx, cy = (ndc[0] * 0.5 + 0.5) * 200, (0.5 - 0.5 * ndc[1]) * 150
    L = np.linalg.norm(v["axis1"][0])
    ys, xs = np.mgrid[0:150, 0:200]
    r2 = ((xs + 0.5 - cx) ** 2 + (ys + 0.5 - cy) ** 2) / (L * L)      # isotropic: q = d / |axis|
    alpha = np.clip(np.exp(-r2) * v["a"][0], 0, 1)
    a1, a2 = v["axis1"][0] / L, v["axis2"][0] / np.linalg.norm(v["axis2"][0])
    dx, dy = xs + 0.5 - cx, cy - (ys + 0.5)
    inside = (np.abs(dx * a1[0] + dy * a1[1]) <= 2 * L) & (np.abs(dx * a2[0] + dy * a2[1]) <= 2 * L)
    alpha = np.where(inside & (alpha >= 1 / 255), alpha, 0)
    edge = np.abs(np.abs(dx * a1[0] + dy * a1[1]) - 2 * L) < 0.02      # pixels within rounding of the quad edge
    edge |= np.abs(np.abs(dx * a2[0] + dy * a2[1]) - 2 * L) < 0.02
    edge |= np.abs(np.exp(-r2) * v["a"][0] - 1 / 255) < 1e-5
    assert np.abs(rt[..., 3] - alpha)[~edge].max() < 3e-5      # (|axis1| and |axis2| differ in the 4th digit: 10-bit rotation)
    assert np.abs(rt[..., 0] - alpha * v["r"][0])[~edge].max() < 3e-5          # premultiplied colour
    assert rt[..., 3].max() > 0.55 and (rt[..., 3] > 0).sum() > 50


def test_front_to_back_blend_and_fp16_rounding(g, O):
    """Two coincident splats: dst = src*(1-dst.a) + dst in sorted (near first) order, rounded to half per splat in ROP mode."""
    cam = camera(g, 64, 64, pos=(0, 0, -4.0))
    raw = np.zeros((2, 62), np.float32)
    for i, (z, col, op) in enumerate([(0.0, (1.0, 0.0, 0.0), 0.5), (1.0, (0.0, 1.0, 0.0), 0.7)]):
        raw[i, 0:3] = (0.011, 0.007, z); raw[i, 6:9] = col; raw[i, 54] = op; raw[i, 55:58] = 0.3; raw[i, 58:62] = (0.5, 0.5, 0.5, 1.0)
    asset = g.create_asset(raw, "VeryHigh")
    fp, _ = g.make_frame_params(cam, sh_order=0)
    out = O.frame(asset, fp, blend_mode=1)
    v = view_fields(out["view"])
    near = int(np.argmin(v["pos"][:, 3]))
    assert out["order"][0] == near                      # ascending view-space z: nearest first
    px = out["rt"][32, 32]
    # at the centre pixel both gaussians are ~1: a0 ~ 0.5, a1 ~ 0.7
    assert px[0] > 0.45 and abs(px[1] - 0.7 * (1 - px[3] + 0.7 * (1 - 0.5)) / 1.0) < 0.2
    a0 = px[0]                                          # red only comes from the near splat: = alpha0
    g1 = px[1]                                          # green only from the far one: alpha1 * (1 - alpha0)
    assert abs(px[3] - (a0 + g1)) < 1e-6
    rop = O.frame(asset, fp, blend_mode=0)["rt"]
    assert np.array_equal(rop, rop.astype(np.float16).astype(np.float32))          # every value is a half
    assert np.abs(rop - out["rt"]).max() < 1e-3 and not np.array_equal(rop, out["rt"])


def test_threads_do_not_change_results(g, O):
    asset = g.synthetic_asset(g.SCENE_CLUSTERED, 30000, 21, "Medium")
    fp, _ = g.make_frame_params(camera(g, 200, 133))
    a, b = O.frame(asset, fp, threads=1), O.frame(asset, fp, threads=5)
    for k in a:
        assert np.array_equal(a[k], b[k]), k


def test_sort_is_a_stable_ascending_sort(O):
    rng = np.random.default_rng(0)
    for n in (0, 1, 2, 1000, 70001):
        keys = rng.integers(0, 50, n, dtype=np.uint32) * 0x01000100
        k, p = keys.copy(), np.arange(n, dtype=np.uint32)
        O.sort_pairs(k, p, threads=3)
        want = np.argsort(keys, kind="stable")
        assert np.array_equal(p, want.astype(np.uint32)) and np.array_equal(k, keys[want])


def test_distances_follow_the_previous_order_and_are_monotonic_in_depth(g, O):
    asset = g.synthetic_asset(g.SCENE_LATTICE, 1000, 0x5EED0001, "VeryHigh")
    cam = camera(g, 64, 64, pos=(0, 0, -4.0))
    fp, _ = g.make_frame_params(cam)
    ident = np.arange(1000, dtype=np.uint32)
    k0 = O.calc_distances(asset, fp, ident)
    perm = np.random.default_rng(1).permutation(1000).astype(np.uint32)
    assert np.array_equal(O.calc_distances(asset, fp, perm), k0[perm])
    z = asset.posData[:12000].view(np.float32).reshape(-1, 3)[:, 2]
    assert np.array_equal(np.argsort(k0, kind="stable"), np.argsort(z, kind="stable"))     # camera looks down +z
    assert (np.diff(np.sort(k0)) == 0).sum() > 10                                           # deliberate ties


def test_composite_formula(O):
    rng = np.random.default_rng(0)
    rt = np.zeros((4, 5, 4), np.float32)
    a = rng.random((4, 5)).astype(np.float32)
    c = rng.random((4, 5, 3)).astype(np.float32)
    rt[..., 3] = a; rt[..., :3] = c * a[..., None]
    rt[0, 0] = 0                                       # no splats here
    bg = rng.random((4, 5, 4)).astype(np.float32)
    out = O.composite(rt, bg)
    lin = c * (c * (c * 0.305306011 + 0.682171111) + 0.012522878)       # GammaToLinearSpace
    want = lin * a[..., None] + bg[..., :3] * (1 - a[..., None])
    m = np.ones((4, 5), bool); m[0, 0] = False
    assert np.abs(out[..., :3] - want)[m].max() < 2e-6
    assert np.abs(out[..., 3] - (a * a + bg[..., 3] * (1 - a)))[m].max() < 2e-6
    assert np.array_equal(out[0, 0], bg[0, 0])


def test_anisotropic_splats_match_the_textbook_conic_form(g, O):
    """Independent restatement, float64: the 3DGS paper's EWA projection (cov2d = J W Sigma W^T J^T + 0.3 I) and its pixel
    weight alpha = o * exp(-1/2 d^T cov2d^-1 d).  The reference draws the same Gaussian as an oriented quad with
    exp(-|q|^2) in eigen-coordinates (S/SplatUtilities.compute:107-162, S/RenderGaussianSplats.shader:54-86); for rotated,
    anisotropic splats well inside the screen (no clamp, no lambda2 floor, inside the +-2 quad) the two must agree."""
    rng = np.random.default_rng(5)
    W, H = 320, 240
    cam = camera(g, W, H, fov=50.0, pos=(0.2, -0.1, -3.0))
    for _ in range(6):
        q = rng.standard_normal(4)
        q /= np.linalg.norm(q)
        scale = np.exp(rng.uniform(np.log(0.03), np.log(0.25), 3))
        asset = one_splat(g, pos=tuple(rng.uniform(-0.3, 0.3, 3)), scale=tuple(scale), quat=tuple(q), opacity=0.7, dc0=(0.5, 0.6, 0.7))
        s = O.load_splat(asset, 0)                               # rotation as stored (10.10.10.2), not as requested
        fp, _ = g.make_frame_params(cam, sh_order=0)
        view = O.calc_view(asset, fp)
        v = view_fields(view)
        rt = O.render(view, np.arange(1, dtype=np.uint32), W, H, blend_mode=1)
        x, y, z, w = (float(t) for t in s["rot"])
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                      [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
        S2 = np.diag(np.asarray(s["scale"], np.float64) ** 2)
        V = np.array(fp.mat_view[:], np.float64).reshape(4, 4).T
        P = np.array(fp.mat_proj_gpu[:], np.float64).reshape(4, 4).T
        t = V @ np.array(list(s["pos"]) + [1.0], np.float64)
        f = W * P[0, 0] / 2
        J = np.array([[f / t[2], 0, -f * t[0] / t[2] ** 2], [0, f / t[2], -f * t[1] / t[2] ** 2]])
        cov = J @ V[:3, :3] @ R @ S2 @ R.T @ V[:3, :3].T @ J.T + 0.3 * np.eye(2)
        assert np.linalg.eigvalsh(cov)[0] > 0.2                  # the lambda2 >= 0.1 floor is not in play
        clip = P @ t
        cx, cy = (clip[0] / clip[3] * 0.5 + 0.5) * W, (0.5 - 0.5 * clip[1] / clip[3]) * H
        ys, xs = np.mgrid[0:H, 0:W]
        # cov2d lives in view space (y up); the render-into-texture projection flips y, hence +(py - cy) here
        d = np.stack([xs + 0.5 - cx, (ys + 0.5) - cy], -1)
        e = np.einsum("...i,ij,...j->...", d, np.linalg.inv(cov), d)
        alpha = np.clip(float(v["a"][0]) * np.exp(-0.5 * e), 0, 1)
        drawn = rt[..., 3] > 0
        assert drawn.sum() > 300
        assert np.abs(rt[..., 3] - alpha)[drawn].max() < 2e-5
        assert np.abs(rt[..., 1] - alpha * float(v["g"][0]))[drawn].max() < 2e-5
        # and nothing else was dropped: an undrawn pixel is below the 1/255 discard or outside the +-2 quad of the eigenbasis
        lam, vec = np.linalg.eigh(cov)
        qe = np.abs(d @ vec) / np.sqrt(2 * lam)
        undrawn_ok = (alpha < 1 / 255 + 1e-5) | (qe.max(-1) > 2 - 1e-3)
        assert undrawn_ok[~drawn].all()


def test_sh_order3_matches_the_published_real_sh_basis(g, O):
    """Independent restatement, float64: colour = dc + sum_lm c_lm Y_lm(dir) with the real spherical-harmonic basis of the
    3DGS paper's eval_sh (constants 0.4886, 1.0925, 0.3154, 0.5463, 0.5900, 2.8906, 0.4570, 0.3732, 1.4453), dir = unit
    vector from the camera to the splat.  ShadeSH (S/GaussianSplatting.hlsl:139-179) is that sum written per band."""
    rng = np.random.default_rng(9)
    cam_pos = np.array([0.3, 0.2, -4.0])
    cam = camera(g, 320, 240, pos=tuple(cam_pos))
    for _ in range(5):
        sh = (0.2 * rng.standard_normal((15, 3))).astype(np.float32)
        pos = rng.uniform(-0.8, 0.8, 3)
        asset = one_splat(g, pos=tuple(pos), dc0=(0.9, 1.0, 1.1), sh=sh)
        s = O.load_splat(asset, 0)
        d = np.asarray(s["pos"], np.float64) - cam_pos
        x, y, z = d / np.linalg.norm(d)
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        Y = [-0.4886025119029199 * y, 0.4886025119029199 * z, -0.4886025119029199 * x,
             1.0925484305920792 * xy, -1.0925484305920792 * yz, 0.31539156525252005 * (2 * zz - xx - yy), -1.0925484305920792 * xz,
             0.5462742152960396 * (xx - yy),
             -0.5900435899266435 * y * (3 * xx - yy), 2.890611442640554 * xy * z, -0.4570457994644658 * y * (4 * zz - xx - yy),
             0.3731763325901154 * z * (2 * zz - 3 * xx - 3 * yy), -0.4570457994644658 * x * (4 * zz - xx - yy),
             1.445305721320277 * z * (xx - yy), -0.5900435899266435 * x * (xx - 3 * yy)]
        coeff = np.asarray(s["sh"], np.float64).reshape(15, 3)
        for order, nterms in ((1, 3), (2, 8), (3, 15)):
            want = np.maximum(np.asarray(s["col"], np.float64) + sum(Y[k] * coeff[k] for k in range(nterms)), 0.0)
            v, _ = _view(g, O, asset, cam, sh_order=order)
            got = np.array([v["r"][0], v["g"][0], v["b"][0]], np.float64)
            assert np.allclose(got, want, rtol=1.5e-3, atol=2e-4), (order, got, want)     # the result is stored as half


def test_scene_depth_buffer_hides_what_lies_behind_it(g, O):
    """The splat pass keeps ShaderLab's default ZTest LEqual against the camera's depth buffer (S/RenderGaussianSplats.shader:8-12,
    depth target bound at R/GaussianSplatRenderer.cs:195).  Oracle semantics: every fragment of a splat has the depth
    clip.z / clip.w of the centre; under the GPU projection's reversed Z it passes iff that depth >= the stored one."""
    W, H = 96, 64
    cam = camera(g, W, H, fov=50.0, pos=(0.0, 0.0, -3.0))
    near = one_splat(g, pos=(-0.4, 0.0, 0.0), scale=(0.2, 0.2, 0.2), opacity=0.9, dc0=(0.9, 0.1, 0.1))
    fp, _keep = g.make_frame_params(cam, sh_order=0)
    view = O.calc_view(near, fp)
    z = view.view(np.float32)[0, 2] / view.view(np.float32)[0, 3]        # the quad's depth, reversed Z: nearer = larger
    assert 0.0 < z < 1.0
    order = np.arange(1, dtype=np.uint32)
    free = O.render(view, order, W, H, 1)
    assert free[..., 3].max() > 0.5
    for stored, visible in ((np.float32(z), True), (np.nextafter(np.float32(z), np.float32(2)), False), (np.float32(0.0), True), (np.float32(1.0), False)):
        depth = np.full((H, W), stored, np.float32)
        img = O.render(view, order, W, H, 1, scene_depth=depth)
        assert np.array_equal(img, free) if visible else not img.any(), (stored, visible)
    half = np.zeros((H, W), np.float32)
    half[:, W // 2:] = 1.0                                                # an occluder over the right half of the screen only
    img = O.render(view, order, W, H, 1, scene_depth=half)
    assert np.array_equal(img[:, :W // 2], free[:, :W // 2]) and not img[:, W // 2:].any()
