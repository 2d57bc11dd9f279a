#!/usr/bin/env python
"""Regenerates tests/golden/*.npz: outputs of the CPU oracle on two small seeded scenes.

The reference ships no golden vectors for this path and cannot run
here, so these fixtures pin the ORACLE (and the packer) against silent drift, and give the GPU
tests a file to compare against that does not depend on the oracle being rebuilt on the GPU box.
Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

CASES = {
    # name: (scene kind, n, seed, quality, width, height, camera pos, sh_order)
    "lattice_1k_medium": (0, 1000, 0x5EED0001, "Medium", 128, 96, (0.0, 0.0, -4.0), 3),
    "clustered_4k_veryhigh": (1, 4000, 0x5EED0003, "VeryHigh", 160, 100, (0.0, 0.5, -6.0), 3),
    "clustered_4k_medium_sh1": (1, 4000, 0x5EED0002, "Medium", 160, 100, (1.0, 0.2, -5.0), 1),
}


def build(name):
    import unitygaussiansplatting_b200 as g
    from oracle import gs_oracle_py as O
    from util import camera
    kind, n, seed, quality, w, h, pos, sh = CASES[name]
    asset = g.synthetic_asset(kind, n, seed, quality)
    cam = camera(g, w, h, pos=pos)
    fp, _keep = g.make_frame_params(cam, sh_order=sh)
    out = O.frame(asset, fp, threads=1)
    blobs = [asset.posData, asset.otherData, asset.colorData, asset.shData] + ([asset.chunkData] if asset.chunkData is not None else [])
    digest = hashlib.sha256(b"".join(b.tobytes() for b in blobs)).hexdigest()
    tgt = O.composite(out["rt"], np.full((h, w, 4), 0.25, np.float32))
    return asset, cam, sh, {"keys": out["keys"], "order": out["order"], "view": out["view"], "rt": out["rt"].astype(np.float16),
                            "composite": tgt, "asset_sha256": np.frombuffer(bytes.fromhex(digest), np.uint8)}


if __name__ == "__main__":
    for name in CASES:
        _a, _c, _s, data = build(name)
        np.savez_compressed(Path(__file__).parent / (name + ".npz"), **data)
        print(name, {k: v.shape for k, v in data.items()})
