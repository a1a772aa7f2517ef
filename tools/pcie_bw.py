"""Host<->device copy bandwidth with 1..N GPUs copying at the same time (run under torchrun, one rank per GPU).
Tells whether the end-to-end (host-buffer) numbers of bench.py are bounded by the platform's aggregate PCIe /
host-memory bandwidth rather than by anything in this repo."""
import os
import sys

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nbytes = 1 << 30
h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
h.fill_(1)


def timed(fn, active):
    res = []
    for _ in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if active:
            for _ in range(4):
                fn()
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1))
    return min(res)


for nact in sorted({1, 2, 4, 8, world} & set(range(1, world + 1))):
    for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
        ms = timed(fn, rank < nact)
        gbs = torch.tensor([4 * nbytes / (ms * 1e-3) / 1e9 if rank < nact else 0.0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(gbs)
        if rank == 0:
            print("%s with %d GPU(s) copying: aggregate %.1f GB/s (%.1f per GPU)" % (name, nact, gbs.item(), gbs.item() / nact), flush=True)
if world > 1:
    dist.destroy_process_group()
