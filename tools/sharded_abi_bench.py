"""End-to-end timing of the multi-GPU C ABI (mp2gpu_comm_init + mp2gpu_commit_from_values_sharded): ONE process, G devices,
pinned host buffers in and out -- the call a patched plonky2 makes for a wide batch.  usage: sharded_abi_bench.py G [leaves 0/1]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mapreduce_plonky2_b200 as G_
from mapreduce_plonky2_b200 import _lib, plonky2 as P2

ndev = int(sys.argv[1]) if len(sys.argv) > 1 else 2
want_leaves = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
n_log, ncols, rate_bits, cap_h, kind = 20, 256, 3, 4, 0
n, N, ncap = 1 << n_log, (1 << n_log) << 3, 16
G_.init(0)
comm = P2.Communicator(list(range(ndev)))
cols = P2.pinned_empty((ncols, n))
rng = np.random.default_rng(1)
for c in range(ncols):
    cols[c] = rng.integers(0, 1 << 62, n, dtype=np.uint64)
coeffs = P2.pinned_empty((ncols, n))
leaves = P2.pinned_empty((N, ncols)) if want_leaves else None
dig = P2.pinned_empty((2 * (N - ncap), 4))
cap = P2.pinned_empty((ncap, 4))
def call():
    _lib.call("mp2gpu_commit_from_values_sharded", comm._handle, P2._col_ptrs(cols), ncols, n_log, rate_bits, cap_h, kind, 0,
              P2._col_ptrs(coeffs), P2._ptr(leaves), P2._ptr(dig), P2._ptr(cap))
call(); call()
ts = []
for _ in range(4):
    t0 = time.perf_counter(); call(); ts.append(time.perf_counter() - t0)
best, med = min(ts), sorted(ts)[len(ts) // 2]
print("C ABI sharded, %d GPUs, leaves_out %s: median %.1f ms (best %.1f) -> %.2f Gelem/s; cap_xor %016x" % (
    ndev, "set" if want_leaves else "NULL", med * 1e3, best * 1e3, ncols * N / med / 1e9, int(np.bitwise_xor.reduce(cap.reshape(-1)))))
comm.free()
