#!/usr/bin/env python
"""k_calc_view time under different knobs (which part of the kernel costs what)."""
import statistics, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch, bench
g, asset, cam = bench.make_scene()
ctx = g.GaussianSplatContext(0)
r = g.GaussianSplatRenderer(asset, ctx)
rt = torch.zeros((cam.pixelHeight, cam.pixelWidth, 4), dtype=torch.float16, device="cuda")
ctx.set_timing(True)
def t_fused(**kw):
    for k, v in kw.items(): setattr(r, k, v)
    xs = []
    for _ in range(12):
        r.SortAndRenderSplats(cam, rt=rt); xs.append(ctx.stage_times().view_ms * 1e3)
    return statistics.median(xs[2:])
def t_full(**kw):
    for k, v in kw.items(): setattr(r, k, v)
    xs = []
    for _ in range(12):
        r.CalcViewData(cam); xs.append(ctx.stage_times().view_ms * 1e3)
    return statistics.median(xs[2:])
print("fused sh3", t_fused(m_SHOrder=3), "fused sh0", t_fused(m_SHOrder=0))
print("full  sh3", t_full(m_SHOrder=3), "full  sh0", t_full(m_SHOrder=0))
import copy
cam2 = copy.deepcopy(cam); cam2.position = np.array([0.0, 0.5, -60.0])   # everything in view, tiny
print("far camera (all in frustum): fused", end=" ")
xs = []
for _ in range(8):
    r.m_SHOrder = 3; r.SortAndRenderSplats(cam2, rt=rt); xs.append(ctx.stage_times().view_ms * 1e3)
print(statistics.median(xs[2:]))
cam3 = copy.deepcopy(cam); cam3.rotation = g.look_rotation([0, 0, -1]); cam3.position = np.array([0.0, 0.5, -40.0])  # everything behind
xs = []
for _ in range(8):
    r.SortAndRenderSplats(cam3, rt=rt); xs.append(ctx.stage_times().view_ms * 1e3)
print("looking away (all behind): fused", statistics.median(xs[2:]))
