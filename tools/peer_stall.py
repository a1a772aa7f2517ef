"""Stall hunt for the fused peer exchange: many back-to-back sharded commitments with per-stage
CUDA-event timing on every rank; prints the distribution of step times and every outlier with its stage split.
Run under torchrun, one rank per GPU:  tools/peer_stall.py [exchange] [steps] [n_log] [ncols] [pinned_copies 0/1]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

exchange = sys.argv[1] if len(sys.argv) > 1 else "peer"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
n_log = int(sys.argv[3]) if len(sys.argv) > 3 else 20
ncols = int(sys.argv[4]) if len(sys.argv) > 4 else 256
with_copies = bool(int(sys.argv[5])) if len(sys.argv) > 5 else False
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ["MP2_SHARDED_TIMING"] = "1"
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import mapreduce_plonky2_b200 as G
from mapreduce_plonky2_b200 import sharded as S

G.init(local)
n = 1 << n_log
c_loc = ncols // world
cols = torch.randint(0, 1 << 62, (c_loc, n), dtype=torch.int64, device="cuda")
engine = S.CudaEngine()
scratch = {}
host_out = None
if with_copies:
    N = n << 3
    host_out = S.HostOutputs(torch.empty((c_loc, n), dtype=torch.int64, pin_memory=True),
                             torch.empty((N // world, ncols), dtype=torch.int64, pin_memory=True),
                             torch.empty((2 * (N - 16) // world, 4), dtype=torch.int64, pin_memory=True),
                             torch.empty((16, 4), dtype=torch.int64, pin_memory=True), torch.cuda.Stream())
for _ in range(3):
    S.commit_sharded(cols, ncols, 3, 4, 0, engine, scratch=scratch, exchange=exchange, host_out=host_out)
torch.cuda.synchronize()
dist.barrier()
scratch["timing"] = []
evs = []
for i in range(steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    S.commit_sharded(cols, ncols, 3, 4, 0, engine, scratch=scratch, exchange=exchange, host_out=host_out)
    e1.record()
    evs.append((e0, e1))
    if with_copies:
        torch.cuda.synchronize()
torch.cuda.synchronize()
ms = [a.elapsed_time(b) for a, b in evs]
rep = S.timing_report(scratch)
med = sorted(ms)[len(ms) // 2]
out = ["[rank %d] %s exchange, %d steps: median %.2f ms, min %.2f, max %.2f" % (rank, exchange, steps, med, min(ms), max(ms))]
for i, t in enumerate(ms):
    if t > 1.05 * med:
        out.append("[rank %d]   outlier step %d: %.2f ms  %s" % (rank, i, t, " ".join("%s=%.2f" % kv for kv in rep[i].items())))
out.append("[rank %d]   typical step: %s" % (rank, " ".join("%s=%.2f" % kv for kv in rep[len(rep) // 2].items())))
for r in range(world):
    if r == rank:
        print("\n".join(out), flush=True)
    dist.barrier()
dist.destroy_process_group()
