#!/usr/bin/env python
"""What NCCL can do for the two exchanges of the group path, by collective and message size (run under torchrun).
Prints, per variant, the device time of one exchange (CUDA events, median of 50)."""
import os
import statistics
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)


def timed(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = torch.tensor([statistics.median(ts)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for name, total in (("order 24.5 MB", 6_131_954 * 4), ("rows cfg2 7.65 MB", 1200 * 797 * 8), ("rows 1080p 16.6 MB", 1920 * 1080 * 8), ("rows 4K 66 MB", 3840 * 2160 * 8)):
    per = (total // world + 15) // 16 * 16
    buf = torch.zeros(per * world, dtype=torch.uint8, device=dev)
    mine = buf[rank * per:(rank + 1) * per]
    res = {}
    res["all_gather(in place)"] = timed(lambda: dist.all_gather_into_tensor(buf, mine))

    def bcasts():
        with dist._coalescing_manager(device=dev, async_ops=False):
            for c in range(world):
                dist.broadcast(buf[c * per:(c + 1) * per], src=c)
    try:
        res["grouped broadcasts"] = timed(bcasts)
    except Exception as e:
        res["grouped broadcasts"] = "n/a (%s)" % type(e).__name__

    def sendrecv():
        ops = []
        for p in range(world):
            if p != rank:
                ops.append(dist.P2POp(dist.isend, mine, p))
                ops.append(dist.P2POp(dist.irecv, buf[p * per:(p + 1) * per], p))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    res["grouped send/recv"] = timed(sendrecv)
    if rank == 0:
        print(name, "x%d:" % world, {k: (round(v, 1) if isinstance(v, float) else v) for k, v in res.items()}, "us", flush=True)
dist.destroy_process_group()
