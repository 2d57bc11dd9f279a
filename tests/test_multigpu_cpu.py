"""Host-side logic of the multi-GPU path (screen-tile bands + one all-gather), on CPU:
partition arithmetic, and a world_size-2 gloo run that gathers band-packed buffers and reassembles the frame."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_band_partition_arithmetic():
    from unitygaussiansplatting_b200.multigpu import BandPartition, TILE
    for height in (16, 17, 797, 1080, 2160):
        for count in (1, 2, 3, 4, 8):
            for band in (1, 2, 8):
                parts = [BandPartition(height, count, i, band) for i in range(count)]
                ty = parts[0].tiles_y
                owners = [parts[0].owner(t) for t in range(ty)]
                for i, p in enumerate(parts):
                    own = [t for t in range(ty) if owners[t] == i]
                    assert p.own_tile_rows() == len(own)
                    assert [p.kth_own_row(k) for k in range(len(own))] == own
                    for y in range(ty + 1):
                        assert p.own_rows_below(y) == sum(1 for t in own if t < y)
                assert sum(p.own_tile_rows() for p in parts) == ty
                rpp = parts[0].rows_per_partition
                assert all(p.rows_per_partition == rpp for p in parts) and rpp >= TILE * max(p.own_tile_rows() for p in parts)
                seen = set()
                for y in range(height):
                    o, row = parts[0].source_row(y)
                    assert 0 <= row < rpp and (o, row) not in seen
                    seen.add((o, row))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch
    import torch.distributed as dist
    import unitygaussiansplatting_b200 as g
    from oracle import gs_oracle_py as O
    from unitygaussiansplatting_b200.multigpu import BandPartition, TILE
    from util import camera
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W, H = 200, 150
    asset = g.synthetic_asset(g.SCENE_CLUSTERED, 5000, 33, "Medium")      # replicated on every rank
    fp, _keep = g.make_frame_params(camera(g, W, H))
    full = O.frame(asset, fp)["rt"].astype(np.float16)                     # stands in for the rank's renderer
    part = BandPartition(H, world, rank, band_rows=2)
    mine = np.zeros((part.rows_per_partition, W, 4), np.float16)
    for k in range(part.own_tile_rows()):                                  # band-packed send buffer: own 64-pixel row k -> rows [64k, 64k+64)
        ty = part.kth_own_row(k)
        rows = full[ty * TILE:min((ty + 1) * TILE, H)]
        mine[k * TILE:k * TILE + rows.shape[0]] = rows
    send = torch.from_numpy(mine.view(np.uint8))
    recv = [torch.zeros_like(send) for _ in range(world)]
    dist.all_gather(recv, send)                                            # the ONE collective of the frame
    gathered = np.stack([t.numpy().view(np.float16) for t in recv])
    out = np.zeros_like(full)
    for y in range(H):
        o, row = part.source_row(y)
        out[y] = gathered[o, row]
    q.put((rank, bool(np.array_equal(out, full)), float(np.abs(mine.astype(np.float32)).sum())))
    dist.destroy_process_group()


def test_two_rank_gloo_gather_reassembles_the_frame():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _r, ok, _s in res)
    assert all(s > 0 for _r, _ok, s in res)       # both ranks actually contributed pixels
