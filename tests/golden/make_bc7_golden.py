#!/usr/bin/env python
"""Regenerates tests/golden/bc7_blocks.npz: random BC7 blocks of every mode and their pixels as decoded by an
INDEPENDENT implementation (Pillow's DDS/BC7 reader).  ColorFormat.BC7 is decoded by the GPU texture unit in the
reference (R/GaussianSplatAsset.cs:169), i.e. by the published block format, so a third-party conformant decoder is the
golden here -- this pins the oracle's and the CUDA path's BC7 decode.  Run: python tests/golden/make_bc7_golden.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "tools"))
from gen_bc7_tables import pillow_decode  # noqa: E402

PER_MODE = 512


def main():
    rng = np.random.default_rng(0xBC7)
    blocks = rng.integers(0, 256, (8 * PER_MODE, 16), dtype=np.uint8)
    for mode in range(8):  # force the unary mode prefix: `mode` zero bits, then a one
        sl = slice(mode * PER_MODE, (mode + 1) * PER_MODE)
        keep = np.uint8((~((1 << (mode + 1)) - 1)) & 0xFF)
        blocks[sl, 0] = (blocks[sl, 0] & keep) | np.uint8(1 << mode)
    # a few structured blocks: all-ones indices, extreme endpoints
    blocks[::97, 8:] = 0xFF
    blocks[5::131, 1:8] = 0x00
    nb = blocks.shape[0]
    img = pillow_decode(blocks.tobytes(), 4 * 64, 4 * (nb // 64))
    pixels = img.reshape(nb // 64, 4, 64, 4, 4).transpose(0, 2, 1, 3, 4).reshape(nb, 16, 4)
    np.savez_compressed(Path(__file__).with_name("bc7_blocks.npz"), blocks=blocks, pixels=pixels.astype(np.uint8))
    print("wrote", nb, "blocks")


if __name__ == "__main__":
    main()
