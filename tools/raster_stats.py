#!/usr/bin/env python
"""Work counters of the compositor for the cfg2 frame (GS_RASTER_STATS=1)."""
import os, sys
from pathlib import Path
os.environ["GS_RASTER_STATS"] = "1"
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch, bench
from unitygaussiansplatting_b200 import _native as N
if len(sys.argv) > 2:   # raster_stats.py WIDTH HEIGHT
    bench.WIDTH, bench.HEIGHT = int(sys.argv[1]), int(sys.argv[2])
g, asset, cam = bench.make_scene()
ctx = g.GaussianSplatContext(0)
r = g.GaussianSplatRenderer(asset, ctx)
rt = torch.zeros((cam.pixelHeight, cam.pixelWidth, 4), dtype=torch.float16, device="cuda")
for _ in range(2):
    r.SortAndRenderSplats(cam, rt=rt)
out = np.zeros(8, np.uint64)
N.check(ctx.handle, N.native().gs_debug_raster_stats(ctx.handle, out.ctypes.data))
names = ["warp-batches", "warp cull ballots (32 entries each)", "warp candidates", "warp evaluations", "pixel blends", "list entries x warps"]
for n_, v in zip(names, out):
    print("%-40s %12d" % (n_, int(v)))
ntiles = ((cam.pixelWidth + 15) // 16) * ((cam.pixelHeight + 15) // 16)
print("warp time: mean %.1f us, max %.1f us" % (float(out[6]) / (ntiles * 8) / 1e3, float(out[7]) / 1e3))
print("entries scanned per warp / listed: %.3f" % (float(out[1]) * 32 / max(1.0, float(out[5]))))
print("blends per evaluation (of 32 lanes): %.2f" % (float(out[4]) / max(1.0, float(out[3]))))
