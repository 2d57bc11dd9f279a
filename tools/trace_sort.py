#!/usr/bin/env python
"""Run the 6.1M sort with GS_SORT_TRACE and print the per-phase time breakdown of pass 1 tiles."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
os.environ["GS_SORT_TRACE"] = "/tmp/sort_trace.bin"
import unitygaussiansplatting_b200 as g
import torch
n = 6_131_954
rng = np.random.default_rng(11)
z = rng.normal(8.0, 6.0, n).astype(np.float32)
u = z.view(np.uint32)
keys = (u ^ np.where(u >> 31, 0xFFFFFFFF, 0x80000000).astype(np.uint32)).astype(np.uint32)
ctx = g.GaussianSplatContext(0)
dk = torch.from_numpy(keys.view(np.int32)).cuda(); dv = torch.arange(n, dtype=torch.int32, device="cuda")
for _ in range(3):
    d1, d2 = dk.clone(), dv.clone()
    ctx.sort_pairs_device(d1.data_ptr(), d2.data_ptr(), n)
    ctx.sync()
t = np.fromfile("/tmp/sort_trace.bin", dtype=np.uint64).reshape(-1, 8).astype(np.int64)
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
names = ["load+hist", "publish+lookback", "rank", "prefix", "scatter", "writeout"]
print("threads", os.environ.get("GS_SORT_THREADS","256"), "tiles", len(t), "kernel span us", (t[:, 6].max() - t0) / 1e3)
for i, nm in enumerate(names):
    d = (t[:, i + 1] - t[:, i]) / 1e3
    print("%-26s mean %7.2f us  p50 %7.2f  p90 %7.2f  max %7.2f" % (nm, d.mean(), np.median(d), np.percentile(d, 90), d.max()))
tot = (t[:, 6] - t[:, 0]) / 1e3
print("tile total mean %.2f p50 %.2f max %.2f" % (tot.mean(), np.median(tot), tot.max()))
start = (t[:, 0] - t0) / 1e3
for k in (0, 100, 300, 591, 592, 700, 1000, 1400):
    if k < len(t):
        print("tile %4d start %7.2f lookback %6.2f total %6.2f" % (k, start[k], (t[k, 2] - t[k, 1]) / 1e3, tot[k]))
