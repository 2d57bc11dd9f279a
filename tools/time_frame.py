#!/usr/bin/env python
"""Whole-frame device time of cfg2 (CUDA events around 50 fused frames), with and without stage overlap."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch, bench
g, asset, cam = bench.make_scene()
s = torch.cuda.Stream()
ctx = g.GaussianSplatContext(0, s.cuda_stream)
r = g.GaussianSplatRenderer(asset, ctx)
with torch.cuda.stream(s):
    rt = torch.zeros((cam.pixelHeight, cam.pixelWidth, 4), dtype=torch.float16, device="cuda")
    for _ in range(5):
        r.SortAndRenderSplats(cam, rt=rt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(50):
        r.SortAndRenderSplats(cam, rt=rt)
    e1.record(s)
    torch.cuda.synchronize()
print("ms per frame: %.4f" % (e0.elapsed_time(e1) / 50))
