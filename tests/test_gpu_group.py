"""The group path (gs_group_*: key-range-sharded sort + row-range-sharded view-calc / binning / compositing) against the
single-GPU frame: draw order and render target must be BIT-identical on every member, frame after frame, while the camera
moves (so the order carried from frame to frame, the splitters and the re-cut row ranges are all exercised).

On a one-GPU box the group is emulated (GS_GROUP_EMULATE: G contexts on the same device, exchanges by device copies), which
runs every kernel and every piece of host logic of the real thing except the NCCL calls themselves; test_group_nccl_*
needs >= 2 devices and is skipped otherwise (bench.py --gpus N makes the same assertion on the NCCL path)."""
import numpy as np
import pytest

from util import camera

pytestmark = pytest.mark.gpu


def _cams(g, w, h, k):
    # a short fly-through: position and heading change every frame
    out = []
    for i in range(k):
        a = 0.12 * i
        out.append(camera(g, w, h, pos=(0.4 * i - 0.5, 0.5 + 0.05 * i, -6.0 + 0.35 * i), forward=(np.sin(a), -0.02 * i, np.cos(a))))
    return out


def _single_gpu_sequence(g, ctx, asset, cams, sort_nth=1, **knobs):
    r = g.GaussianSplatRenderer(asset, ctx)
    r.m_SortNthFrame = sort_nth
    for k, v in knobs.items():
        setattr(r, k, v)
    frames = []
    for cam in cams:
        rt = np.zeros((cam.pixelHeight, cam.pixelWidth, 4), np.float16)
        r.SortAndRenderSplats(cam, rt=rt)
        frames.append((r.readback_order(), rt))
    r.Dispose()
    return frames


def _check_group(g, grp, cams, want):
    for k, cam in enumerate(cams):
        rts = [np.zeros((cam.pixelHeight, cam.pixelWidth, 4), np.float16) for _ in range(grp.local_count)]
        grp.SortAndRenderSplats(cam, rts=rts)
        st = grp.stats()
        bounds = list(st.row_bounds[: grp.size + 1])
        assert bounds[0] == 0 and bounds[-1] == (cam.pixelHeight + 15) // 16 and all(b1 >= b0 for b0, b1 in zip(bounds, bounds[1:]))
        assert sum(st.slab_counts[: grp.size]) == grp.splatCount
        for i in range(grp.local_count):
            assert np.array_equal(grp.readback_order(i), want[k][0]), "frame %d member %d: draw order differs from the single-GPU sort" % (k, i)
            assert rts[i].any()
            assert np.array_equal(rts[i].view(np.uint16), want[k][1].view(np.uint16)), "frame %d member %d: render target differs" % (k, i)


@pytest.mark.parametrize("gpus", [1, 2, 3, 4, 8])
def test_group_emulated_equals_single_gpu(g, ctx, gpus):
    asset = g.synthetic_asset(g.SCENE_CLUSTERED, 150000, 0x5EED0071, "Medium")
    cams = _cams(g, 640, 400, 5)
    want = _single_gpu_sequence(g, ctx, asset, cams)
    from unitygaussiansplatting_b200.multigpu import GaussianSplatGroup
    grp = GaussianSplatGroup.create(asset, [0] * gpus, emulate=True)
    _check_group(g, grp, cams, want)
    grp.close()


def test_group_more_gpus_than_rows_and_ties(g, ctx):
    """A 40-pixel-high screen has 3 rows of 16 pixels for 8 members; the lattice asset is full of depth ties."""
    asset = g.synthetic_asset(g.SCENE_LATTICE, 1000, 0x5EED0001, "Medium")
    cams = [camera(g, 96, 40, pos=(0.1 * i, 0.0, -4.0)) for i in range(4)]
    want = _single_gpu_sequence(g, ctx, asset, cams)
    from unitygaussiansplatting_b200.multigpu import GaussianSplatGroup
    grp = GaussianSplatGroup.create(asset, [0] * 8, emulate=True)
    _check_group(g, grp, cams, want)
    grp.close()


def test_group_sort_every_other_frame_and_knobs(g, ctx):
    asset = g.synthetic_asset(g.SCENE_CLUSTERED, 60000, 0x5EED0072, "VeryHigh")
    cams = _cams(g, 333, 211, 5)
    knobs = dict(m_SplatScale=1.3, m_OpacityScale=0.8, m_SHOrder=2)
    want = _single_gpu_sequence(g, ctx, asset, cams, sort_nth=2, **knobs)
    from unitygaussiansplatting_b200.multigpu import GaussianSplatGroup
    grp = GaussianSplatGroup.create(asset, [0, 0, 0], emulate=True)
    grp.m_SortNthFrame = 2
    for k, v in knobs.items():
        setattr(grp, k, v)
    _check_group(g, grp, cams, want)
    grp.close()


def test_group_device_images(g, ctx):
    """Members may hand in device images (the zero-copy host): the exchange happens in place in them."""
    import torch
    asset = g.synthetic_asset(g.SCENE_CLUSTERED, 80000, 0x5EED0073, "Medium")
    cams = _cams(g, 512, 320, 3)
    want = _single_gpu_sequence(g, ctx, asset, cams)
    from unitygaussiansplatting_b200.multigpu import GaussianSplatGroup
    grp = GaussianSplatGroup.create(asset, [0, 0, 0, 0], emulate=True)
    rts = [torch.zeros((320, 512, 4), dtype=torch.float16, device="cuda") for _ in range(4)]
    for k, cam in enumerate(cams):
        grp.SortAndRenderSplats(cam, rts=rts)
        grp.sync()
        for i in range(4):
            assert np.array_equal(rts[i].cpu().numpy().view(np.uint16), want[k][1].view(np.uint16)), "frame %d member %d" % (k, i)
    grp.close()


def test_group_full_size_cfg2(g, ctx):
    """BASELINE configs[1]-sized: 6,131,954 Medium @1200x797 on an emulated group of 4."""
    import bench
    _g, asset, cam0 = bench.make_scene()
    cams = [bench.orbit_camera(k) for k in range(3)]
    want = _single_gpu_sequence(g, ctx, asset, cams)
    from unitygaussiansplatting_b200.multigpu import GaussianSplatGroup
    grp = GaussianSplatGroup.create(asset, [0, 0, 0, 0], emulate=True)
    _check_group(g, grp, cams, want)
    grp.close()


def test_group_nccl_single_process(g, ctx):
    """gs_group_create over real devices: NCCL (ncclCommInitAll) inside the library."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 CUDA devices")
    n = min(torch.cuda.device_count(), 4)
    asset = g.synthetic_asset(g.SCENE_CLUSTERED, 150000, 0x5EED0071, "Medium")
    cams = _cams(g, 640, 400, 4)
    want = _single_gpu_sequence(g, ctx, asset, cams)
    from unitygaussiansplatting_b200.multigpu import GaussianSplatGroup
    grp = GaussianSplatGroup.create(asset, list(range(n)))
    _check_group(g, grp, cams, want)
    grp.close()


def test_large_asset_bitmap_path(g, ctx, tmp_path):
    """Above ~33 M splats the walkers' flag bitmaps no longer fit beside their static shared memory and are read through L1
    instead (found by tools/big_scene.py at 50 M).  GS_WALK_BITS_GLOBAL=1 forces that path on a small scene: single GPU and an
    emulated group of 3 must give the same order and pixels as the default path."""
    import os
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    script = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import unitygaussiansplatting_b200 as g\n"
        "from unitygaussiansplatting_b200.multigpu import GaussianSplatGroup\n"
        "from util import camera\n"
        "asset = g.synthetic_asset(g.SCENE_CLUSTERED, 120000, 0x5EED0075, 'Medium')\n"
        "r = g.GaussianSplatRenderer(asset, g.GaussianSplatContext(0))\n"
        "grp = GaussianSplatGroup.create(asset, [0, 0, 0], emulate=True)\n"
        "out = {}\n"
        "for k in range(3):\n"
        "    cam = camera(g, 480, 300, pos=(0.3 * k, 0.5, -6.0 + 0.2 * k))\n"
        "    rt = np.zeros((300, 480, 4), np.float16)\n"
        "    r.SortAndRenderSplats(cam, rt=rt)\n"
        "    rts = [np.zeros_like(rt) for _ in range(3)]\n"
        "    grp.SortAndRenderSplats(cam, rts=rts)\n"
        "    assert all(np.array_equal(x, rt) for x in rts) and np.array_equal(grp.readback_order(1), r.readback_order())\n"
        "    out['rt%%d' %% k], out['order%%d' %% k] = rt, r.readback_order()\n"
        "np.savez(sys.argv[1], **out)\n" % (str(root), str(root / "tests")))
    res = {}
    for flag in ("0", "1"):
        path = tmp_path / ("walk%s.npz" % flag)
        subprocess.run([sys.executable, "-c", script, str(path)], check=True, env=dict(os.environ, GS_WALK_BITS_GLOBAL=flag), timeout=300)
        res[flag] = np.load(path)
    for k in res["0"].files:
        assert res["0"][k].any()
        assert np.array_equal(res["0"][k], res["1"][k]), k


def test_group_async_host_readback(g, ctx):
    """GS_FLAG_ASYNC_READBACK on the group path: host images are filled by the transfer stream after the call has returned;
    gs_group_sync completes them.  Two pinned images per member in rotation, three frames in flight before the first sync."""
    import torch
    asset = g.synthetic_asset(g.SCENE_CLUSTERED, 90000, 0x5EED0076, "Medium")
    cams = _cams(g, 512, 320, 4)
    want = _single_gpu_sequence(g, ctx, asset, cams)
    from unitygaussiansplatting_b200.multigpu import GaussianSplatGroup
    grp = GaussianSplatGroup.create(asset, [0, 0, 0], emulate=True)
    grp.async_readback = True
    pins = [[torch.zeros((320, 512, 4), dtype=torch.float16).pin_memory() for _ in range(2)] for _ in range(3)]
    for k, cam in enumerate(cams):
        if k >= 2:                       # the images about to be reused hold frame k-2: complete and check them first
            grp.sync()
            for i in range(3):
                assert np.array_equal(pins[i][k & 1].numpy().view(np.uint16), want[k - 2][1].view(np.uint16)), "frame %d member %d" % (k - 2, i)
        grp.SortAndRenderSplats(cam, rts=[pins[i][k & 1].numpy() for i in range(3)])
    grp.sync()
    for k in (2, 3):
        for i in range(3):
            assert np.array_equal(pins[i][k & 1].numpy().view(np.uint16), want[k][1].view(np.uint16)), "frame %d member %d" % (k, i)
    assert np.array_equal(grp.readback_order(2), want[3][0])
    grp.close()
