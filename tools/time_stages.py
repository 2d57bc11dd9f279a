#!/usr/bin/env python
"""Per-stage device times of the cfg2 frame (library CUDA events), for quick A/B on the GPU box."""
import statistics
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import bench

# usage: time_stages.py [n] [quality] [width height fov seed]   (defaults: BASELINE cfg2)
n = int(sys.argv[1]) if len(sys.argv) > 1 else bench.N_SPLATS
quality = sys.argv[2] if len(sys.argv) > 2 else "Medium"
if len(sys.argv) > 6:
    bench.WIDTH, bench.HEIGHT, bench.FOV, bench.SEED = int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5]), int(sys.argv[6], 0)
g, asset, cam = bench.make_scene(n, quality)
ctx = g.GaussianSplatContext(0)
r = g.GaussianSplatRenderer(asset, ctx)
import torch
rt = torch.zeros((cam.pixelHeight, cam.pixelWidth, 4), dtype=torch.float16, device="cuda")
for _ in range(5):
    r.SortAndRenderSplats(cam, rt=rt)
ctx.set_timing(True)
acc = {}
for _ in range(30):
    r.SortAndRenderSplats(cam, rt=rt)
    st = ctx.stage_times()
    for k in ("distances_ms", "sort_ms", "view_ms", "bin_ms", "raster_ms", "total_ms"):
        acc.setdefault(k, []).append(getattr(st, k))
    acc.setdefault("p", []).append(list(st.sort_pass_ms))
print({k: round(statistics.median(v) * 1e3, 1) for k, v in acc.items() if k != "p"}, "us; passes",
      [round(statistics.median(p[i] for p in acc["p"]) * 1e3, 1) for i in range(4)], "entries", int(st.tile_entries), "| %s n=%d %dx%d" % (quality, n, cam.pixelWidth, cam.pixelHeight))
