"""Merkle levels only (digests of the leaves already in place): fused top-of-tree vs one launch per level."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mapreduce_plonky2_b200 import device as D
torch.cuda.set_device(0); D.bind_current_device()
for kind in (0, 1):
    for n_log, cap in ((17, 4), (13, 4), (9, 4), (6, 0)):
        N = 1 << n_log
        dig = torch.randint(0, 2**62, (max(2 * (N - (1 << cap)), 1), 4), dtype=torch.int64, device="cuda")
        capb = torch.empty((1 << cap, 4), dtype=torch.int64, device="cuda")
        for _ in range(5): D.merkle_levels(N, cap, kind, dig, capb)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(10):
            e0.record(); D.merkle_levels(N, cap, kind, dig, capb); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print("levels of 2^%d leaves cap %d kind %d: %.1f us" % (n_log, cap, kind, best * 1e3), flush=True)
