#!/usr/bin/env python
"""A handful of cfg2 frames (orbiting camera) and nothing else: the subject of ncu launch lists / --set full captures.
usage: few_frames.py [frames] [gpus-emulated] [workload]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 3
emu = int(sys.argv[2]) if len(sys.argv) > 2 else 0
name = sys.argv[3] if len(sys.argv) > 3 else "cfg2"
n, quality, w, h, fov, seed, _ = bench.WORKLOADS[name]
import unitygaussiansplatting_b200 as g
asset = bench.get_asset(n, quality, seed)
rt = torch.zeros((h, w, 4), dtype=torch.float16, device="cuda")
if emu:
    from unitygaussiansplatting_b200.multigpu import GaussianSplatGroup
    grp = GaussianSplatGroup.create(asset, [0] * emu, emulate=True)
    for k in range(frames):
        grp.SortAndRenderSplats(bench.orbit_camera(k, w, h, fov), rts=[rt] + [None] * (emu - 1))
    grp.sync()
else:
    ctx = g.GaussianSplatContext(0)
    r = g.GaussianSplatRenderer(asset, ctx)
    for k in range(frames):
        r.SortAndRenderSplats(bench.orbit_camera(k, w, h, fov), rt=rt)
    ctx.sync()
print("done")
