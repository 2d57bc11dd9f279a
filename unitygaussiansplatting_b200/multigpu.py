"""Screen-tile partition across the GPUs of one box (the reference is single-GPU).

Every rank holds the whole asset, sorts and view-calcs it redundantly, and composites only its own
bands of 64-pixel bin rows (interleaved round-robin for load balance).  Each rank renders straight
into its slice of the all-gather buffer ("band-packed": own bin row k -> pixel rows [64k,64k+64)),
ONE all-gather moves the bands, and gs_unshuffle_bands assembles the image.  torch.distributed is
plumbing only (process group + the collective); the partition arithmetic below is mirrored by the
device code in csrc/gs_raster.cu (struct Partition) and tested against it.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

from . import _native as N

TILE = 64   # band granularity in pixels: the library bins (and partitions) in 64-pixel rows (csrc/gs_common.cuh kBin)


@dataclass
class BandPartition:
    height: int
    count: int
    index: int
    band_rows: int = 1   # tile rows per band

    @property
    def tiles_y(self) -> int:
        return (self.height + TILE - 1) // TILE

    def owner(self, tile_row: int) -> int:
        return 0 if self.count <= 1 else (tile_row // self.band_rows) % self.count

    def own_rows_below(self, y: int, index=None) -> int:
        """number of tile rows in [0, y) owned by partition `index`"""
        index = self.index if index is None else index
        if self.count <= 1:
            return y
        cyc = self.band_rows * self.count
        q, r = divmod(y, cyc)
        lo = index * self.band_rows
        return q * self.band_rows + min(max(r - lo, 0), self.band_rows)

    def own_tile_rows(self, index=None) -> int:
        return self.own_rows_below(self.tiles_y, index)

    def kth_own_row(self, k: int, index=None) -> int:
        index = self.index if index is None else index
        if self.count <= 1:
            return k
        return ((k // self.band_rows) * self.count + index) * self.band_rows + (k % self.band_rows)

    @property
    def rows_per_partition(self) -> int:
        """pixel rows of every rank's slice in the gather buffer (max over ranks, so slices are equal-sized)"""
        return TILE * max(self.own_tile_rows(i) for i in range(max(1, self.count)))

    def options(self):
        return (self.index, self.count if self.count > 1 else 0, self.band_rows)

    def source_row(self, y: int):
        """image row y -> (partition, row inside that partition's band-packed buffer)"""
        ty = y // TILE
        o = self.owner(ty)
        return o, self.own_rows_below(ty, o) * TILE + (y - ty * TILE)


def alloc_gather(part: BandPartition, width: int, device, dtype=None):
    import torch
    return torch.zeros((max(1, part.count), part.rows_per_partition, width, 4), dtype=dtype or torch.float16, device=device)


def render_partitioned(renderer, cam, part: BandPartition, gathered, out_image, stream=None):
    """One multi-GPU frame.  `gathered`: alloc_gather() tensor; `out_image`: (H, W, 4) CUDA tensor."""
    import torch.distributed as dist
    own_px = part.own_tile_rows() * TILE
    mine = gathered[part.index]
    renderer.partition = part.options()
    renderer.band_packed = True
    if own_px:
        renderer.SortAndRenderSplats(cam, rt=mine[:own_px])
    else:   # more partitions than 64-pixel rows: nothing to composite here, but keep the draw order current
        renderer.SortPoints(cam)
    if part.count > 1:
        dist.all_gather_into_tensor(gathered.view(-1), mine.reshape(-1))   # in place: input is our slice of the output
    unshuffle(renderer.context, gathered, part, out_image)


def unshuffle(context, gathered, part: BandPartition, out_image):
    from .renderer import _image
    h, w = out_image.shape[0], out_image.shape[1]
    im = _image(out_image, w, h)
    fmt = N.GS_PIX_RGBA16F if gathered.element_size() == 2 else N.GS_PIX_RGBA32F
    N.check(context.handle, N.native().gs_unshuffle_bands(context.handle, C.c_void_p(gathered.data_ptr()), max(1, part.count), part.band_rows,
                                                          part.rows_per_partition, fmt, C.byref(im)))
