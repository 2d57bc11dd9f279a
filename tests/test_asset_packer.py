"""The host-side packer (csrc/asset_creator.cpp) against the reference's documented layout, and
encode -> oracle-decode round trips for every supported format."""
import numpy as np
import pytest

from util import one_splat


def _raw(g, kind, n, seed):
    return g.generate_input_splats(kind, n, seed)


def test_sizes_match_the_reference_formulas(g):
    # R/GaussianSplatAsset.cs:152-203; readme.md:80 quotes 282 MB for bicycle at Medium
    from unitygaussiansplatting_b200 import _native as N
    import ctypes as C
    sz = N.GsaSizes()
    n = 6_131_954
    assert N.asset_lib().gsa_calc_sizes(n, 2, 2, 2, 3, C.byref(sz)) == 0
    assert (sz.tex_width, sz.tex_height) == (2048, 3008)
    assert sz.pos_bytes == (n * 4 + 7) // 8 * 8 and sz.other_bytes == n * 8 and sz.sh_bytes == n * 32
    assert sz.color_bytes == 2048 * 3008 * 4 and sz.chunk_bytes == ((n + 255) // 256) * 64
    total = sz.pos_bytes + sz.other_bytes + sz.sh_bytes + sz.color_bytes + sz.chunk_bytes
    assert abs(total / 2**20 - 282.3) < 0.1
    assert N.asset_lib().gsa_calc_sizes(n, 0, 0, 0, 0, C.byref(sz)) == 0 and sz.chunk_bytes == 0   # lossless: no chunks
    # VeryLow preset (E/GaussianSplatAssetCreator.cs:195-200): BC7 = 1 byte per texel, palette of 4096 x 96 B, u16 index per splat
    assert N.asset_lib().gsa_calc_sizes(n, 2, 3, 3, 8, C.byref(sz)) == 0
    assert sz.color_bytes == 2048 * 3008 and sz.sh_bytes == 4096 * 96 and sz.other_bytes == (n * 8 + 7) // 8 * 8
    assert N.asset_lib().gsa_calc_sizes(n, 2, 2, 2, 4, C.byref(sz)) == 0 and sz.sh_bytes == 65536 * 96 and sz.other_bytes == (n * 10 + 7) // 8 * 8
    assert N.asset_lib().gsa_calc_sizes(4096, 2, 3, 3, 8, C.byref(sz)) != 0   # palette not smaller than the data
    assert N.asset_lib().gsa_calc_sizes(n, 2, 2, 4, 3, C.byref(sz)) != 0 and N.asset_lib().gsa_calc_sizes(n, 2, 2, 2, 9, C.byref(sz)) != 0


def test_generation_is_deterministic(g):
    a, b = _raw(g, g.SCENE_CLUSTERED, 5000, 7), _raw(g, g.SCENE_CLUSTERED, 5000, 7)
    assert np.array_equal(a, b) and not np.array_equal(a, _raw(g, g.SCENE_CLUSTERED, 5000, 8))
    assert np.isfinite(a).all()


def test_morton_reorder(g):
    from unitygaussiansplatting_b200 import _native as N
    raw = _raw(g, g.SCENE_UNIFORM, 20000, 3)
    before = raw.copy()
    asset = g.create_asset(raw, "VeryHigh")     # lossless: `raw` is now just the reordered input
    assert sorted(map(bytes, raw)) == sorted(map(bytes, before))
    bmin, bmax = before[:, :3].min(0), before[:, :3].max(0)
    assert np.array_equal(asset.boundsMin, bmin) and np.array_equal(asset.boundsMax, bmax)
    inv = np.float32(1.0) / (bmax - bmin)
    ip = ((raw[:, :3] - bmin) * inv * np.float32((1 << 21) - 1)).astype(np.uint32)
    codes = np.array([N.asset_lib().gsa_morton_encode3(int(x), int(y), int(z)) for x, y, z in ip], np.uint64)
    assert np.all(codes[1:] >= codes[:-1])
    assert N.asset_lib().gsa_morton_encode3(1, 0, 0) == 1 and N.asset_lib().gsa_morton_encode3(0, 1, 0) == 2 and \
        N.asset_lib().gsa_morton_encode3(0, 0, 1) == 4 and N.asset_lib().gsa_morton_encode3(0x1FFFFF, 0x1FFFFF, 0x1FFFFF) == (1 << 63) - 1


def _rotmat(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def _unpack_rot(packed):
    # inverse of PackSmallest3Rotation (R/GaussianUtils.cs:46-76), float64
    idx = int(round(packed[3] * 3))
    xyz = packed[:3] * np.sqrt(2.0) - 1 / np.sqrt(2.0)
    w = np.sqrt(max(0.0, 1 - float(xyz @ xyz)))
    q = [xyz[0], xyz[1], xyz[2], w]
    return {0: [w, xyz[0], xyz[1], xyz[2]], 1: [xyz[0], w, xyz[1], xyz[2]], 2: [xyz[0], xyz[1], w, xyz[2]], 3: q}[idx]


@pytest.mark.parametrize("quality", ["VeryHigh", "High", "Medium"])
def test_encode_decode_round_trip(g, O, quality):
    raw = _raw(g, g.SCENE_CLUSTERED, 3000, 11)
    src = raw.copy()
    # reproduce the Morton permutation so decoded splat i can be compared with its source record
    ref = g.create_asset(src.copy(), "VeryHigh")
    order_src = src.copy()
    g.create_asset(order_src, "VeryHigh")          # order_src is now the Morton-ordered, unmodified input
    asset = g.create_asset(raw, quality)
    tol = {"VeryHigh": dict(pos=0, scale=0, col=0, sh=0, op=0),
           "High": dict(pos=2e-5, scale=6e-3, col=2e-3, sh=2e-3, op=1.5e-3),
           "Medium": dict(pos=1.2e-3, scale=0.02, col=6e-3, sh=0.04, op=6e-3)}[quality]
    for i in range(0, 3000, 7):
        s, d = order_src[i], O.load_splat(asset, i)
        chunk = order_src[(i // 256) * 256:(i // 256) * 256 + 256]
        prange = (chunk[:, :3].max(0) - chunk[:, :3].min(0)).max() + 1e-5
        assert np.abs(d["pos"] - s[0:3]).max() <= tol["pos"] * prange + 1e-6 * (quality != "VeryHigh")
        assert np.abs(d["scale"] / s[55:58] - 1).max() <= tol["scale"]
        assert np.abs(d["col"] - s[6:9]).max() <= tol["col"]
        # opacity is stored as SquareCentered01(opacity), whose inverse has infinite slope at 0.5: compare in the stored domain
        sq = lambda x: (x - 0.5) * abs(x - 0.5) * 2 + 0.5
        assert abs(sq(float(d["opacity"])) - sq(float(s[54]))) <= tol["op"]
        shrange = float(chunk[:, 9:54].max() - chunk[:, 9:54].min()) + 1e-5
        assert np.abs(d["sh"].reshape(45) - s[9:54]).max() <= tol["sh"] * shrange
        R0, R1 = _rotmat(_unpack_rot(s[58:62].astype(np.float64))), _rotmat(d["rot"].astype(np.float64))
        assert np.abs(R0 - R1).max() < 6e-3       # 10-bit smallest-three
    if quality == "VeryHigh":
        assert asset.chunkData is None
        assert np.array_equal(asset.posData[:3000 * 12].view(np.float32).reshape(-1, 3), order_src[:, :3])
    else:
        assert asset.chunkData is not None and asset.chunkData.nbytes == ((3000 + 255) // 256) * 64
    assert ref.splatCount == 3000


def test_pack_smallest3_and_rotation_decode_agree(g, O):
    rng = np.random.default_rng(2)
    for _ in range(50):
        q = rng.standard_normal(4)
        q /= np.linalg.norm(q)
        asset = one_splat(g, quat=q)
        d = O.load_splat(asset, 0)
        idx0 = 0 if np.allclose(d["pos"], 0, atol=1e-6) else None
        assert idx0 == 0
        assert np.abs(_rotmat(q) - _rotmat(d["rot"].astype(np.float64))).max() < 6e-3
        assert abs(np.linalg.norm(d["rot"]) - 1) < 3e-3   # DecodeRotation does not renormalise


def test_unity_f32tof16_rounds_half_up_on_bit_12(g):
    from unitygaussiansplatting_b200 import _native as N
    L = N.asset_lib()
    rng = np.random.default_rng(4)
    v = (rng.random(5000).astype(np.float32) * 8 - 4)
    got = np.array([L.gsa_f32tof16(float(x)) for x in v], np.uint32)
    want = v.astype(np.float16).view(np.uint16).astype(np.uint32)
    assert (got != want).mean() < 0.01 and np.abs(got.astype(np.int64) - want.astype(np.int64)).max() <= 1
    assert L.gsa_f32tof16(1.0) == 0x3C00 and L.gsa_f32tof16(-2.0) == 0xC000 and L.gsa_f32tof16(65536.0) == 0x7C00
    assert L.gsa_f32tof16(1.00048828125) == 0x3C01     # exact tie rounds up (RNE would give 0x3C00)


def test_chunk_bounds_enclose_and_decode_is_inside_them(g, O):
    raw = _raw(g, g.SCENE_CLUSTERED, 1000, 5)
    asset = g.create_asset(raw, "Medium")
    ch = asset.chunkData.view(np.uint32).reshape(-1, 16)
    pos_minmax = ch[:, 4:10].view(np.float32).reshape(-1, 3, 2)
    for i in range(0, 1000, 13):
        d = O.load_splat(asset, i)
        lo, hi = pos_minmax[i // 256, :, 0], pos_minmax[i // 256, :, 1]
        assert np.all(d["pos"] >= lo - 1e-6) and np.all(d["pos"] <= hi + 1e-6)
        assert 0.0 <= float(d["opacity"]) <= 1.0 and np.all(d["scale"] > 0)


def _write_ply(path, cols, names, crlf=False, extra_uchar=False):
    nl = "\r\n" if crlf else "\n"
    n = cols.shape[0]
    hdr = "ply" + nl + "format binary_little_endian 1.0" + nl + "element vertex %d" % n + nl
    for nm in names:
        hdr += "property float %s" % nm + nl
    if extra_uchar:
        hdr += "property uchar flag" + nl
    hdr += "end_header" + nl
    rec = np.zeros(n, dtype=[("f", "<f4", (len(names),))] + ([("u", "u1")] if extra_uchar else []))
    rec["f"] = cols
    with open(path, "wb") as f:
        f.write(hdr.encode())
        f.write(rec.tobytes())


def test_ply_reader_follows_the_importer(g, tmp_path):
    """E/Utils/GaussianFileReader.cs:45-232: attribute mapping by name, SH re-interleave, LinearizeData."""
    rng = np.random.default_rng(9)
    names = ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"] + ["f_rest_%d" % i for i in range(45)] + \
            ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    n = 300
    cols = rng.standard_normal((n, 62)).astype(np.float32)
    perm = rng.permutation(62)                      # attribute order in the file is arbitrary: mapping is by name
    _write_ply(tmp_path / "a.ply", cols[:, perm], [names[i] for i in perm], crlf=True, extra_uchar=True)
    got = g.read_ply(tmp_path / "a.ply")
    assert got.shape == (n, 62)
    assert np.array_equal(got[:, 0:6], cols[:, 0:6])                                   # pos, normal untouched
    assert np.allclose(got[:, 6:9], cols[:, 6:9] * 0.2820948 + 0.5, atol=1e-6)         # SH0ToColor
    want_sh = cols[:, 9:54].reshape(n, 3, 15).transpose(0, 2, 1).reshape(n, 45)        # channel-major -> 15 x RGB
    assert np.array_equal(got[:, 9:54], want_sh)
    assert np.allclose(got[:, 54], 1 / (1 + np.exp(-cols[:, 54].astype(np.float64))), atol=1e-6)   # sigmoid
    assert np.allclose(got[:, 55:58], np.exp(cols[:, 55:58].astype(np.float64)), rtol=1e-6)         # exp(log scale)
    wxyz = cols[:, 58:62].astype(np.float64)
    q = wxyz / np.linalg.norm(wxyz, axis=1, keepdims=True)
    for i in range(0, n, 17):
        R0 = _rotmat([q[i, 1], q[i, 2], q[i, 3], q[i, 0]])
        R1 = _rotmat(_unpack_rot(got[i, 58:62].astype(np.float64)))
        assert np.abs(R0 - R1).max() < 1e-5
    # the records go straight into the packer
    asset = g.create_asset(got.copy(), "Medium")
    assert asset.splatCount == n
    # missing required attribute / ascii ply -> rejected
    _write_ply(tmp_path / "b.ply", cols[:, :61], names[:61])
    with pytest.raises(ValueError):
        g.read_ply(tmp_path / "b.ply")
    (tmp_path / "c.ply").write_text("ply\nformat ascii 1.0\nelement vertex 1\nend_header\n")
    with pytest.raises(ValueError):
        g.read_ply(tmp_path / "c.ply")
    # a property type the importer does not know (its TypeToSize throws), and a header that promises more vertices than
    # the file holds: rejected before anything is allocated from the untrusted count
    good = (tmp_path / "a.ply").read_bytes()
    (tmp_path / "d.ply").write_bytes(good.replace(b"property uchar", b"property int", 1))
    with pytest.raises(ValueError):
        g.read_ply(tmp_path / "d.ply")
    (tmp_path / "e.ply").write_bytes(good.replace(b"element vertex %d" % n, b"element vertex %d" % (n * 1000000), 1))
    with pytest.raises(ValueError):
        g.read_ply(tmp_path / "e.ply")
    (tmp_path / "f.ply").write_bytes(good[:-7])
    with pytest.raises(ValueError):
        g.read_ply(tmp_path / "f.ply")


def test_spz_reader_follows_the_importer(g, tmp_path):
    """E/Utils/SPZFileReader.cs:66-195: gzip stream, 24-bit fixed-point positions, byte-quantised everything else."""
    import gzip
    import struct
    rng = np.random.default_rng(10)
    n, sh_level, fract = 257, 2, 12
    shc = 8
    pos = rng.integers(-2**23, 2**23, (n, 3), dtype=np.int64)
    alpha = rng.integers(0, 256, n, dtype=np.uint8)
    col = rng.integers(0, 256, (n, 3), dtype=np.uint8)
    scale = rng.integers(0, 256, (n, 3), dtype=np.uint8)
    rot = rng.integers(60, 196, (n, 3), dtype=np.uint8)
    sh = rng.integers(0, 256, (n, shc * 3), dtype=np.uint8)
    p24 = np.zeros((n, 3, 3), np.uint8)
    u = pos & 0xFFFFFF
    p24[..., 0], p24[..., 1], p24[..., 2] = u & 255, (u >> 8) & 255, (u >> 16) & 255
    blob = struct.pack("<IIII", 0x5053474E, 2, n, sh_level | (fract << 8)) + p24.tobytes() + alpha.tobytes() + col.tobytes() + \
        scale.tobytes() + rot.tobytes() + sh.tobytes()
    with gzip.open(tmp_path / "a.spz", "wb") as f:
        f.write(blob)
    got = g.read_spz(tmp_path / "a.spz")
    assert got.shape == (n, 62)
    assert np.array_equal(got[:, 0:3], (pos.astype(np.float32) * np.float32(1.0 / (1 << fract))))
    assert np.allclose(got[:, 55:58], np.exp(scale.astype(np.float64) / 16.0 - 10.0), rtol=2e-6)
    assert np.array_equal(got[:, 54], alpha.astype(np.float32) / np.float32(255.0))
    assert np.allclose(got[:, 6:9], ((col / 255.0 - 0.5) / 0.15) * 0.2820948 + 0.5, atol=2e-6)
    assert np.array_equal(got[:, 9:9 + shc * 3], (sh.astype(np.float32) - 128.0) / 128.0) and not got[:, 9 + shc * 3:54].any()
    xyz = rot.astype(np.float64) / 127.5 - 1.0
    w = np.sqrt(np.maximum(0.0, 1.0 - (xyz ** 2).sum(1)))
    for i in range(0, n, 13):
        q = np.array([xyz[i, 0], xyz[i, 1], xyz[i, 2], w[i]]); q /= np.linalg.norm(q)
        assert np.abs(_rotmat(q) - _rotmat(_unpack_rot(got[i, 58:62].astype(np.float64)))).max() < 1e-5
    g.create_asset(got.copy(), "Medium")
    with gzip.open(tmp_path / "b.spz", "wb") as f:
        f.write(struct.pack("<IIII", 0x5053474E, 3, n, 0))        # unsupported version
    with pytest.raises(ValueError):
        g.read_spz(tmp_path / "b.spz")


@pytest.mark.parametrize("quality", ["VeryLow", "Medium", "VeryHigh"])
def test_asset_files_round_trip(g, tmp_path, quality):
    """save_asset writes the importer's file set (E/GaussianSplatAssetCreator.cs:300-305 + the serialized fields of
    R/GaussianSplatAsset.cs:18-22,205-216); load_asset reads it back identically and rejects damaged sets."""
    n = 9000
    a = g.synthetic_asset(g.SCENE_CLUSTERED, n, 0x5EED0061, quality)
    g.save_asset(a, tmp_path, "scene")
    names = sorted(p.name for p in tmp_path.iterdir())
    expect = ["scene.asset", "scene_col.bytes", "scene_oth.bytes", "scene_pos.bytes", "scene_shs.bytes"] + (["scene_chk.bytes"] if quality != "VeryHigh" else [])
    assert names == sorted(expect)
    b = g.load_asset(tmp_path, "scene")
    assert (b.splatCount, b.posFormat, b.scaleFormat, b.colorFormat, b.shFormat) == (a.splatCount, a.posFormat, a.scaleFormat, a.colorFormat, a.shFormat)
    for attr in ("posData", "otherData", "colorData", "shData"):
        assert np.array_equal(getattr(a, attr), getattr(b, attr))
    assert (a.chunkData is None) == (b.chunkData is None) and (a.chunkData is None or np.array_equal(a.chunkData, b.chunkData))
    assert np.allclose(a.boundsMin, b.boundsMin) and np.allclose(a.boundsMax, b.boundsMax)
    # a Unity-written .asset has many more lines (script reference, GUID'd TextAsset references, cameras): only the scalars matter
    text = (tmp_path / "scene.asset").read_text()
    (tmp_path / "scene.asset").write_text(text + "  m_PosData: {fileID: 4900000, guid: 0123456789abcdef0123456789abcdef, type: 3}\n  m_Cameras: []\n")
    assert g.load_asset(tmp_path, "scene").splatCount == n
    (tmp_path / "scene_pos.bytes").write_bytes(b"\0" * 16)
    with pytest.raises(ValueError):
        g.load_asset(tmp_path, "scene")
    (tmp_path / "scene.asset").write_text(text.replace("m_FormatVersion: 20231020", "m_FormatVersion: 20230101"))
    with pytest.raises(ValueError):
        g.load_asset(tmp_path, "scene")
