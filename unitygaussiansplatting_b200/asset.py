"""Host-side mirror of GaussianSplatAsset (package/Runtime/GaussianSplatAsset.cs) and of the
importer's quality presets (package/Editor/GaussianSplatAssetCreator.cs:189-228).

The byte blobs are produced by libgsplat_asset.so (csrc/asset_creator.cpp), which follows the
reference importer's packing bit for bit; this module only owns the numpy buffers.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from enum import IntEnum
from typing import Optional

import numpy as np

from . import _native as N


class VectorFormat(IntEnum):  # R/GaussianSplatAsset.cs:31-37
    Float32 = 0
    Norm16 = 1
    Norm11 = 2
    Norm6 = 3


class ColorFormat(IntEnum):  # R/GaussianSplatAsset.cs:51-57
    Float32x4 = 0
    Float16x4 = 1
    Norm8x4 = 2
    BC7 = 3


class SHFormat(IntEnum):  # R/GaussianSplatAsset.cs:70-81
    Float32 = 0
    Float16 = 1
    Norm11 = 2
    Norm6 = 3
    Cluster64k = 4
    Cluster32k = 5
    Cluster16k = 6
    Cluster8k = 7
    Cluster4k = 8


# E/GaussianSplatAssetCreator.cs:195-224 -- the importer's five presets
QUALITY = {
    "VeryLow": (VectorFormat.Norm11, VectorFormat.Norm6, ColorFormat.BC7, SHFormat.Cluster4k),
    "Low": (VectorFormat.Norm11, VectorFormat.Norm6, ColorFormat.Norm8x4, SHFormat.Cluster16k),
    "Medium": (VectorFormat.Norm11, VectorFormat.Norm11, ColorFormat.Norm8x4, SHFormat.Norm6),
    "High": (VectorFormat.Norm16, VectorFormat.Norm16, ColorFormat.Float16x4, SHFormat.Norm11),
    "VeryHigh": (VectorFormat.Float32, VectorFormat.Float32, ColorFormat.Float32x4, SHFormat.Float32),
}

SCENE_LATTICE, SCENE_CLUSTERED, SCENE_UNIFORM = 0, 1, 2
INPUT_SPLAT_FLOATS = 62  # InputSplatData, E/Utils/GaussianFileReader.cs:17-26


@dataclass
class GaussianSplatAsset:
    splatCount: int
    posFormat: VectorFormat
    scaleFormat: VectorFormat
    colorFormat: ColorFormat
    shFormat: SHFormat
    posData: np.ndarray
    otherData: np.ndarray
    colorData: np.ndarray
    shData: np.ndarray
    chunkData: Optional[np.ndarray]
    boundsMin: np.ndarray = field(default_factory=lambda: np.zeros(3, np.float32))
    boundsMax: np.ndarray = field(default_factory=lambda: np.zeros(3, np.float32))

    @property
    def total_bytes(self) -> int:
        return sum(a.nbytes for a in (self.posData, self.otherData, self.colorData, self.shData) if a is not None) + (
            self.chunkData.nbytes if self.chunkData is not None else 0)

    def desc(self) -> N.GsAssetDesc:
        d = N.GsAssetDesc()
        d.splat_count = self.splatCount
        d.pos_format, d.scale_format = int(self.posFormat), int(self.scaleFormat)
        d.sh_format, d.color_format = int(self.shFormat), int(self.colorFormat)
        d.pos, d.pos_bytes = self.posData.ctypes.data, self.posData.nbytes
        d.other, d.other_bytes = self.otherData.ctypes.data, self.otherData.nbytes
        d.sh, d.sh_bytes = self.shData.ctypes.data, self.shData.nbytes
        d.color, d.color_bytes = self.colorData.ctypes.data, self.colorData.nbytes
        if self.chunkData is not None and self.chunkData.nbytes:
            d.chunks, d.chunk_bytes = self.chunkData.ctypes.data, self.chunkData.nbytes
        else:
            d.chunks, d.chunk_bytes = None, 0
        return d


def generate_input_splats(kind: int, n: int, seed: int) -> np.ndarray:
    """Deterministic synthetic InputSplatData records (n x 62 float32), SURVEY.md 8d."""
    out = np.empty((n, INPUT_SPLAT_FLOATS), np.float32)
    rc = N.asset_lib().gsa_generate(kind, n, seed & 0xFFFFFFFF, out.ctypes.data)
    if rc != 0:
        raise ValueError("gsa_generate failed (%d)" % rc)
    return out


def read_ply(path: str) -> np.ndarray:
    """INRIA gaussian-splat .ply -> linearised InputSplatData records (n x 62 float32), as the reference importer reads
    them (E/Utils/GaussianFileReader.cs:45-232)."""
    lib = N.asset_lib()
    n = lib.gsa_ply_vertex_count(str(path).encode())
    if n < 0:
        raise ValueError("%s is not a binary little-endian gaussian splat PLY (code %d)" % (path, n))
    out = np.empty((n, INPUT_SPLAT_FLOATS), np.float32)
    rc = lib.gsa_ply_read(str(path).encode(), out.ctypes.data, n)
    if rc != 0:
        raise ValueError("gsa_ply_read failed (%d)" % rc)
    return out


def write_ply(path: str, records: np.ndarray, deleted_bits=None) -> int:
    """ExportPlyFile (E/GaussianSplatRendererEditor.cs:394-445): raw attribute records (n x 62, e.g. from
    GaussianSplatRenderer.EditExportData) -> binary little-endian .ply; deleted / cut-marked records are dropped."""
    records = np.ascontiguousarray(records, np.float32)
    if records.ndim != 2 or records.shape[1] != INPUT_SPLAT_FLOATS:
        raise ValueError("records must be (n, 62) float32")
    bits = None
    if deleted_bits is not None:
        bits = np.ascontiguousarray(deleted_bits, np.uint32)
        if bits.size < (records.shape[0] + 31) // 32:
            raise ValueError("deleted_bits is shorter than splat_count / 32 words")
    n = N.asset_lib().gsa_ply_write(str(path).encode(), records.ctypes.data, records.shape[0], bits.ctypes.data if bits is not None else None)
    if n < 0:
        raise OSError("could not write %s" % path)
    return int(n)


def read_spz(path: str) -> np.ndarray:
    """Niantic/Scaniverse .spz -> InputSplatData records, as E/Utils/SPZFileReader.cs unpacks them."""
    lib = N.asset_lib()
    n = lib.gsa_spz_vertex_count(str(path).encode())
    if n < 0:
        raise ValueError("%s is not a version-2 SPZ file (code %d)" % (path, n))
    out = np.empty((n, INPUT_SPLAT_FLOATS), np.float32)
    rc = lib.gsa_spz_read(str(path).encode(), out.ctypes.data, n)
    if rc != 0:
        raise ValueError("gsa_spz_read failed (%d)" % rc)
    return out


def create_asset(splats: np.ndarray, quality: str = "Medium", formats=None) -> GaussianSplatAsset:
    """CreateAsset: Morton reorder, chunking, packing.  `splats` (n x 62 float32) is consumed."""
    pf, sf, cf, shf = formats if formats is not None else QUALITY[quality]
    if not (splats.dtype == np.float32 and splats.ndim == 2 and splats.shape[1] == INPUT_SPLAT_FLOATS and splats.flags.c_contiguous):
        raise ValueError("splats must be a C-contiguous (n, 62) float32 array")
    n = splats.shape[0]
    lib = N.asset_lib()
    sz = N.GsaSizes()
    if lib.gsa_calc_sizes(n, int(pf), int(sf), int(cf), int(shf), C.byref(sz)) != 0:
        raise ValueError("unsupported format combination (clustered SH needs more splats than palette entries)")
    pos = np.zeros(sz.pos_bytes, np.uint8)
    other = np.zeros(sz.other_bytes, np.uint8)
    color = np.zeros(sz.color_bytes, np.uint8)
    sh = np.zeros(sz.sh_bytes, np.uint8)
    chunks = np.zeros(sz.chunk_bytes, np.uint8) if sz.chunk_bytes else None
    bounds = np.zeros(6, np.float32)
    rc = lib.gsa_create_asset(splats.ctypes.data, n, int(pf), int(sf), int(cf), int(shf), pos.ctypes.data, other.ctypes.data,
                              color.ctypes.data, sh.ctypes.data, chunks.ctypes.data if chunks is not None else None,
                              bounds.ctypes.data)
    if rc != 0:
        raise ValueError("gsa_create_asset failed (%d)" % rc)
    return GaussianSplatAsset(n, VectorFormat(pf), VectorFormat(sf), ColorFormat(cf), SHFormat(shf), pos, other, color, sh, chunks,
                              bounds[:3].copy(), bounds[3:].copy())


# ---- on-disk form: the five .bytes blobs + the serialized ScriptableObject (E/GaussianSplatAssetCreator.cs:296-330) ----
FORMAT_VERSION = 20231020  # GaussianSplatAsset.kCurrentVersion, R/GaussianSplatAsset.cs:13
_BLOBS = (("chunkData", "chk"), ("posData", "pos"), ("otherData", "oth"), ("colorData", "col"), ("shData", "shs"))


def save_asset(asset: GaussianSplatAsset, folder, name: str) -> None:
    """Writes `<name>_chk|pos|oth|col|shs.bytes` (the importer's file names, E/...:300-305) and `<name>.asset`: the fields
    Unity serializes for a GaussianSplatAsset (R/GaussianSplatAsset.cs:18-22,205-216) in its text-YAML form.  The object
    references to the TextAssets are by GUID in a Unity project; here they are by file name, which load_asset resolves."""
    from pathlib import Path
    folder = Path(folder)
    folder.mkdir(parents=True, exist_ok=True)
    for attr, tag in _BLOBS:
        blob = getattr(asset, attr)
        if blob is not None and blob.nbytes:
            (folder / ("%s_%s.bytes" % (name, tag))).write_bytes(np.ascontiguousarray(blob).tobytes())
    v = lambda a: "{x: %r, y: %r, z: %r}" % tuple(float(t) for t in a)
    lines = ["%YAML 1.1", "%TAG !u! tag:unity3d.com,2011:", "--- !u!114 &11400000", "MonoBehaviour:", "  m_Name: %s" % name,
             "  m_FormatVersion: %d" % FORMAT_VERSION, "  m_SplatCount: %d" % asset.splatCount,
             "  m_BoundsMin: %s" % v(asset.boundsMin), "  m_BoundsMax: %s" % v(asset.boundsMax),
             "  m_PosFormat: %d" % int(asset.posFormat), "  m_ScaleFormat: %d" % int(asset.scaleFormat),
             "  m_SHFormat: %d" % int(asset.shFormat), "  m_ColorFormat: %d" % int(asset.colorFormat)]
    (folder / (name + ".asset")).write_text("\n".join(lines) + "\n")


def load_asset(folder, name: str) -> GaussianSplatAsset:
    """Reads what save_asset -- or the Unity importer -- wrote: `<name>.asset` (YAML; only the m_* scalars are needed) and the
    `.bytes` blobs next to it.  Sizes are checked against the formats (R/GaussianSplatAsset.cs:174-203)."""
    import re
    from pathlib import Path
    folder = Path(folder)
    text = (folder / (name + ".asset")).read_text()

    def field(key, cast=int):
        m = re.search(r"^\s*%s:\s*(\S+)\s*$" % re.escape(key), text, re.M)
        if not m:
            raise ValueError("%s.asset has no %s" % (name, key))
        return cast(m.group(1))

    def vec(key):
        m = re.search(r"^\s*%s:\s*\{x:\s*([^,]+),\s*y:\s*([^,]+),\s*z:\s*([^}]+)\}" % re.escape(key), text, re.M)
        return np.array([float(t) for t in m.groups()], np.float32) if m else np.zeros(3, np.float32)

    version = field("m_FormatVersion")
    if version != FORMAT_VERSION:
        raise ValueError("asset format version %d, expected %d (HasValidAsset rejects it too, R/GaussianSplatRenderer.cs:361-368)" % (version, FORMAT_VERSION))
    n = field("m_SplatCount")
    pf, sf = VectorFormat(field("m_PosFormat")), VectorFormat(field("m_ScaleFormat"))
    shf, cf = SHFormat(field("m_SHFormat")), ColorFormat(field("m_ColorFormat"))
    blobs = {}
    for attr, tag in _BLOBS:
        path = folder / ("%s_%s.bytes" % (name, tag))
        blobs[attr] = np.fromfile(path, np.uint8) if path.exists() else None
    for attr in ("posData", "otherData", "colorData", "shData"):
        if blobs[attr] is None:
            raise ValueError("missing %s blob of %s" % (attr, name))
    sz = N.GsaSizes()
    if N.asset_lib().gsa_calc_sizes(n, int(pf), int(sf), int(cf), int(shf), C.byref(sz)) != 0:
        raise ValueError("unsupported format combination in %s.asset" % name)
    # pos / other are padded to 8 bytes by the importer; a clustered-SH file may also carry CalcSHDataSize's n*2 slack
    need = {"posData": n * (12, 6, 4, 2)[int(pf)], "otherData": sz.other_bytes - 7, "colorData": sz.color_bytes, "shData": sz.sh_bytes}
    for attr, least in need.items():
        if blobs[attr].nbytes < least:
            raise ValueError("%s of %s is %d bytes, its format needs %d" % (attr, name, blobs[attr].nbytes, least))
    uses_chunks = sz.chunk_bytes != 0
    if uses_chunks and (blobs["chunkData"] is None or blobs["chunkData"].nbytes < sz.chunk_bytes):
        raise ValueError("%s needs chunk data (lossy formats) but has none / too little" % name)
    return GaussianSplatAsset(n, pf, sf, cf, shf, blobs["posData"], blobs["otherData"], blobs["colorData"], blobs["shData"],
                              blobs["chunkData"] if uses_chunks else None, vec("m_BoundsMin"), vec("m_BoundsMax"))


def synthetic_asset(kind: int, n: int, seed: int, quality: str = "Medium") -> GaussianSplatAsset:
    return create_asset(generate_input_splats(kind, n, seed), quality)
