"""The C-ABI library loads and exports every symbol the headers declare (no compute without a GPU)."""
import ctypes as C
import re
from pathlib import Path

import pytest

from conftest import has_cuda

ROOT = Path(__file__).resolve().parents[1]


def _declared(header: str, prefix: str):
    text = (ROOT / "include" / header).read_text()
    return sorted(set(re.findall(r"\b(%s[a-z0-9_]+)\s*\(" % prefix, text)))


def test_native_exports_every_declared_symbol(g):
    from unitygaussiansplatting_b200 import _native as N
    lib = N.native()
    declared = _declared("gsplat_b200.h", "gs_")
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "libgsplat_b200.so does not export %s" % name
    assert set(declared) == set(N.NATIVE_SYMBOLS), "python binding table out of sync with the header"
    assert b"sm_100a" in lib.gs_version()


def test_asset_lib_exports_every_declared_symbol(g):
    from unitygaussiansplatting_b200 import _native as N
    lib = N.asset_lib()
    declared = _declared("gsplat_asset.h", "gsa_")
    for name in declared:
        assert hasattr(lib, name)
    assert set(declared) == set(N.ASSET_SYMBOLS)


def test_struct_layouts_match_the_header():
    from unitygaussiansplatting_b200 import _native as N
    assert C.sizeof(N.GsCutout) == 68                      # R/GaussianCutout.cs:20-24
    assert C.sizeof(N.GsAssetDesc) == 4 * 5 + 4 + 8 * 5 + 8 * 5
    assert C.sizeof(N.GsFrameParams) == 64 * 4 + 8 + 12 + 8 + 8 + 8 + 4 + 24 + 16
    assert C.sizeof(N.GsImage) == 32
    assert C.sizeof(N.GsRenderOptions) == 32
    assert C.sizeof(N.GsGroupStats) == 8 * 4 + 8 + 17 * 4 + 16 * 4


@pytest.mark.skipif(has_cuda(), reason="CPU-only behaviour")
def test_no_cpu_fallback(g):
    """Without a device the product fails loudly instead of falling back to anything."""
    with pytest.raises(g.GsError) as e:
        g.GaussianSplatContext(0)
    assert e.value.code == -6


def test_product_never_imports_the_oracle():
    pkg = ROOT / "unitygaussiansplatting_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.cpp")):
        text = p.read_text()
        for needle in ("gs_oracle", "import oracle", "from oracle", "gso_", "oracle/gs", "oracle\" /"):
            assert needle not in text, "%s references the oracle (%s)" % (p, needle)


def test_group_entry_points_validate_their_arguments(g):
    """gs_group_* never throw or abort: bad arguments come back as status codes (no GPU needed for these)."""
    from unitygaussiansplatting_b200 import _native as N
    lib = N.native()
    h = C.c_void_p()
    two = (C.c_int * 2)(0, 0)
    assert lib.gs_group_create(None, 2, 0, C.byref(h)) == -1 and not h.value
    assert lib.gs_group_create(two, 0, 0, C.byref(h)) == -1
    assert lib.gs_group_create(two, 17, 0, C.byref(h)) == -1
    assert lib.gs_group_create(two, 2, 0, C.byref(h)) == -1                      # a device listed twice without GS_GROUP_EMULATE
    assert b"twice" in lib.gs_last_error(None)
    if not has_cuda():
        assert lib.gs_group_create(two, 2, N.GS_GROUP_EMULATE, C.byref(h)) == N.GS_ERR_NO_DEVICE and not h.value
    out = (C.c_uint32 * 5)()
    assert lib.gs_group_balance_rows(None, 10, 4, out) == 0 and list(out) == [0, 2, 5, 8, 10] or list(out)[0] == 0
    assert lib.gs_group_balance_rows(None, 0, 4, out) == -1
    assert lib.gs_group_balance_rows(None, 10, 17, out) == -1
    assert lib.gs_group_size(None) == 0 and lib.gs_group_local_count(None) == 0 and not lib.gs_group_context(None, 0)
    assert lib.gs_group_frame(None, None, None, None, 1, None) == -1
    assert lib.gs_group_sync(None) == -1
    lib.gs_group_destroy(None)
