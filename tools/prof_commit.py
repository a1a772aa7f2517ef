"""Short single-GPU driver for ncu captures: a few device-resident commitments of one shape.

    ncu --set full --clock-control none --import-source on -k regex:k_leaf_hash -c 2 \
        -o gpurun_out/prof python tools/prof_commit.py --ncols 135 --n-log 14 --hash 0 --iters 2
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mapreduce_plonky2_b200 import device as D

ap = argparse.ArgumentParser()
ap.add_argument("--ncols", type=int, default=135)
ap.add_argument("--n-log", type=int, default=14)
ap.add_argument("--rate-bits", type=int, default=3)
ap.add_argument("--cap", type=int, default=4)
ap.add_argument("--hash", type=int, default=0)
ap.add_argument("--iters", type=int, default=2)
a = ap.parse_args()
torch.cuda.set_device(0); D.bind_current_device()
cols = torch.randint(0, 2**62, (a.ncols, 1 << a.n_log), dtype=torch.int64, device="cuda")
bufs = D.CommitBuffers(a.ncols, a.n_log, a.rate_bits, a.cap, True)
for _ in range(a.iters):
    D.commit_resident(cols, bufs, a.hash)
torch.cuda.synchronize()
print("done")
