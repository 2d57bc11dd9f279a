#!/usr/bin/env python
"""A small emulated group frame (3 contexts on one GPU) for compute-sanitizer: slab compaction + slab sort + exchange copies +
range-partitioned view-calc / binning / compositing, checked against the single-GPU frame."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import unitygaussiansplatting_b200 as g
from unitygaussiansplatting_b200.multigpu import GaussianSplatGroup

asset = g.synthetic_asset(g.SCENE_CLUSTERED, 30000, 0x5EED0071, "Medium")
cams = [g.Camera(position=np.array([0.3 * k, 0.5, -6.0 + 0.3 * k]), rotation=g.look_rotation([0, 0, 1]), fieldOfView=39.1, pixelWidth=320, pixelHeight=200)
        for k in range(3)]
ctx = g.GaussianSplatContext(0)
r = g.GaussianSplatRenderer(asset, ctx)
grp = GaussianSplatGroup.create(asset, [0, 0, 0], emulate=True)
for cam in cams:
    want = np.zeros((200, 320, 4), np.float16)
    r.SortAndRenderSplats(cam, rt=want)
    rts = [np.zeros_like(want) for _ in range(3)]
    grp.SortAndRenderSplats(cam, rts=rts)
    assert np.array_equal(grp.readback_order(0), r.readback_order())
    assert all(np.array_equal(x, want) for x in rts)
grp.close()
r.Dispose()
print("group frames ok")
