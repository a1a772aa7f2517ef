"""Scratch: correctness spot-check + timing of the Merkle stage alone (device-resident)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mapreduce_plonky2_b200 as G
from mapreduce_plonky2_b200 import device as D
import oracle as O

torch.cuda.set_device(0); D.bind_current_device()
# correctness
rng = np.random.default_rng(1)
st = rng.integers(0, 2**64, size=(512, 12), dtype=np.uint64)
for kind in (0, 1):
    got = G.permute(st, kind)
    ok = all(np.array_equal(got[i], O.permute(st[i], kind)) for i in range(0, 512, 37))
    print("permute kind", kind, "OK" if ok else "MISMATCH", flush=True)

def run(ncols, n_log, kind, iters=5):
    N = (1 << n_log) << 3
    lde = torch.randint(0, 2**62, (ncols, N), dtype=torch.int64, device="cuda")
    leaves = torch.empty((N, ncols), dtype=torch.int64, device="cuda")
    dig = torch.empty((2 * (N - 16), 4), dtype=torch.int64, device="cuda")
    cap = torch.empty((16, 4), dtype=torch.int64, device="cuda")
    t0 = time.time()
    while time.time() - t0 < 0.3:
        D.merkle_colmajor(lde, 4, kind, leaves, dig, cap); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(iters):
        e0.record(); D.merkle_colmajor(lde, 4, kind, leaves, dig, cap); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    perms = N * ((ncols + 7) // 8) + N - 16
    print("merkle c=%d n=2^%d kind=%d: %.3f ms  %.1f Mperm/s  %.0f clk/perm/SM@1.94GHz" % (
        ncols, n_log, kind, best, perms / best / 1e3, best * 1e-3 * 1.94e9 * 148 / perms), flush=True)

for kind in (0, 1):
    run(135, 14, kind)
    run(20, 14, kind)
