"""Host-side logic of the group path (gs_group_*), on CPU:
  * the row balancer of the library itself (gs_group_balance_rows is host arithmetic: callable without a GPU);
  * the key-range slab scheme restated in numpy -- splitters from last frame's quantile positions, slab = number of
    splitters <= key, offsets from the ">= splitter" counts, each slab compacted out of last frame's order and stably
    sorted -- must reproduce the reference's stable sort (the oracle's) exactly, frame after frame, with heavy ties;
  * a world_size-2 gloo run of that scheme: every rank sorts only its slab and composites only its rows, two exchanges,
    and both ranks end up with the oracle's order and frame."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def slab_table(keys_nat, prev_order, G):
    """What k_calc_distances derives on every GPU: ascending splitters, slab of every splat, slab offsets."""
    n = keys_nat.size
    qpos = [(n * (j + 1)) // G for j in range(G - 1)]
    thr = np.sort(keys_nat[prev_order[qpos]]) if G > 1 else np.zeros(0, np.uint32)
    slab = np.searchsorted(thr, keys_nat, side="right")          # #{splitters <= key}
    ge = [(keys_nat >= t).sum() for t in thr]
    off = [0] + [n - int(c) for c in ge] + [n]
    return thr, slab, off


def slab_sort(keys_nat, prev_order, slab, g):
    """GPU g's share: compact its slab out of last frame's order, stable sort by key."""
    mine = prev_order[slab[prev_order] == g]
    return mine[np.argsort(keys_nat[mine], kind="stable")]


@pytest.mark.parametrize("G", [1, 2, 3, 4, 8, 16])
def test_slab_scheme_equals_stable_sort(G):
    rng = np.random.default_rng(1234 + G)
    n = 20000
    order = np.arange(n, dtype=np.uint32)
    for frame in range(5):
        # few distinct values + clustered values: ties inside slabs, ties AT splitters, empty slabs
        keys_nat = rng.integers(0, 200 if frame % 2 else 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
        if frame == 3:
            keys_nat[:] = 7                                       # every key equal: one slab holds everything
        want = order[np.argsort(keys_nat[order], kind="stable")]  # the reference's contract (R/GpuSorting.cs:142-198)
        thr, slab, off = slab_table(keys_nat, order, G)
        assert all(b >= a for a, b in zip(off, off[1:]))
        new = np.empty(n, np.uint32)
        for g in range(G):
            part = slab_sort(keys_nat, order, slab, g)
            assert part.size == off[g + 1] - off[g], "slab size must follow from the >= counts alone"
            new[off[g]:off[g + 1]] = part
        assert np.array_equal(new, want)
        order = new


def test_balance_rows_properties():
    from unitygaussiansplatting_b200.multigpu import balance_rows
    rng = np.random.default_rng(7)
    for rows in (1, 3, 50, 68, 135):
        for parts in (1, 2, 3, 4, 8, 16):
            for kind in range(4):
                cost = [np.zeros(rows), rng.integers(0, 1000, rows), np.r_[np.zeros(rows // 2), rng.integers(1000, 2000, rows - rows // 2)],
                        (np.arange(rows) == rows // 3) * 100000][kind].astype(np.uint32)
                b = balance_rows(cost, parts)
                assert b[0] == 0 and b[-1] == rows and b.size == parts + 1
                assert (np.diff(b.astype(np.int64)) >= 0).all()
                if kind == 0:                                      # no history: even split
                    assert np.abs(np.diff(b.astype(np.int64)) - rows / parts).max() <= 1.0
                if kind == 1 and rows >= 8 * parts:                # no part much above the mean + one row
                    w = cost.astype(np.float64) + (cost.sum() // (rows * 8) + 1)
                    share = np.array([w[b[i]:b[i + 1]].sum() for i in range(parts)])
                    assert share.max() <= w.sum() / parts + w.max()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        _worker_body(rank, world, port, q)
    except Exception as e:   # report instead of leaving the parent waiting for its time-out
        q.put((rank, False))
        raise


def _worker_body(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch
    import torch.distributed as dist
    import unitygaussiansplatting_b200 as g
    from oracle import gs_oracle_py as O
    from unitygaussiansplatting_b200.multigpu import balance_rows
    from util import camera
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W, H = 200, 150
    rows = (H + 15) // 16
    asset = g.synthetic_asset(g.SCENE_CLUSTERED, 6000, 35, "Medium")       # replicated on every rank
    n = asset.splatCount
    order = np.arange(n, dtype=np.uint32)
    ref_order = order.copy()
    ok = True
    cost_hist = [None, None]
    for frame in range(4):
        fp, _keep = g.make_frame_params(camera(g, W, H, pos=(0.3 * frame, 0.5, -6.0 + 0.4 * frame)))
        # the reference contract, single "GPU": oracle sort seeded by last frame's order, oracle frame
        ref = O.frame(asset, fp, prev_order=ref_order)
        ref_order = ref["order"]
        # ---- this rank: replicated key table, own slab only ----
        keys_nat = O.calc_distances(asset, fp, np.arange(n, dtype=np.uint32))
        thr, slab, off = slab_table(keys_nat, order, world)
        mine = slab_sort(keys_nat, order, slab, rank)
        new = np.zeros(n, np.uint32)
        new[off[rank]:off[rank + 1]] = mine
        # exchange 1: the order slabs (sizes known everywhere from `off`): one broadcast per slab, in place
        t = torch.from_numpy(new.view(np.int32))                             # (gloo has no uint32)
        for c in range(world):
            if off[c + 1] > off[c]:
                dist.broadcast(t[off[c]:off[c + 1]], src=c)
        order = new
        ok &= bool(np.array_equal(order, ref_order))
        # ---- own rows only: the range two frames ago's costs give (identical on both ranks) ----
        b = balance_rows(cost_hist[frame & 1] if cost_hist[frame & 1] is not None else np.zeros(rows, np.uint32), world)
        full = ref["rt"].astype(np.float16)                                 # stands in for the rank's compositor
        img = np.zeros_like(full)
        y0, y1 = min(int(b[rank]) * 16, H), min(int(b[rank + 1]) * 16, H)
        img[y0:y1] = full[y0:y1]
        cost = np.zeros(rows, np.uint32)
        lit = (full.astype(np.float32) != 0).any(axis=2).sum(axis=1)        # lit pixels per pixel row ...
        per_row = np.add.reduceat(lit, np.arange(0, H, 16)).astype(np.uint32)   # ... per 16-pixel row: the stand-in for the measured cost
        cost[b[rank]:b[rank + 1]] = per_row[b[rank]:b[rank + 1]]
        # exchange 2: composited rows + their costs
        ti, tc = torch.from_numpy(img.view(np.uint8).reshape(H, -1)), torch.from_numpy(cost.view(np.int32))
        for c in range(world):
            a0, a1 = min(int(b[c]) * 16, H), min(int(b[c + 1]) * 16, H)
            if a1 > a0:
                dist.broadcast(ti[a0:a1], src=c)
            if b[c + 1] > b[c]:
                dist.broadcast(tc[int(b[c]):int(b[c + 1])], src=c)
        ok &= bool(np.array_equal(img, full))
        cost_hist[frame & 1] = cost
    q.put((rank, ok))
    dist.destroy_process_group()


def test_two_rank_gloo_slab_sort_and_row_exchange():
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
    assert all(ok for _r, ok in res)
