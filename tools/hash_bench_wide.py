import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mapreduce_plonky2_b200 import device as D
torch.cuda.set_device(0); D.bind_current_device()
ncols, N = 256, 1 << 21   # quarter of the wide batch's leaves: same per-SM behaviour, shorter runs
lde = torch.randint(0, 2**62, (ncols, N), dtype=torch.int64, device="cuda")
leaves = torch.empty((N, ncols), dtype=torch.int64, device="cuda")
dig = torch.empty((2 * (N - 16), 4), dtype=torch.int64, device="cuda")
cap = torch.empty((16, 4), dtype=torch.int64, device="cuda")
for kind in (0, 1):
    for _ in range(2): D.merkle_colmajor(lde, 4, kind, leaves, dig, cap)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); D.merkle_colmajor(lde, 4, kind, leaves, dig, cap); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1); perms = N * 32 + N - 16
    print("wide/4 kind=%d: %.2f ms %.1f Mperm/s %.0f clk/perm/SM" % (kind, ms, perms / ms / 1e3, ms * 1e-3 * 1.94e9 * 148 / perms), flush=True)
