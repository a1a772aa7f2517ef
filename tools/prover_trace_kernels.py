"""Per-kernel device time of the whole-prover replay (library profiler: CUDA events around every launch)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mapreduce_plonky2_b200 as G
from mapreduce_plonky2_b200 import trace as T, device as D
torch.cuda.set_device(0); G.init(0)
nthreads = int(sys.argv[1]) if len(sys.argv) > 1 else 1
tr = T.ProverTrace(T.LEAF_PROOF_DEGREES, 1, nthreads)
tr.run(2 * nthreads); tr.run(2 * nthreads)
D.profile_enable(True); D.profile_report()
n = 8 * nthreads
t0 = time.perf_counter(); tr.run(n); dt = time.perf_counter() - t0
D.profile_enable(False)
prof = D.profile_report()
tot = sum(v[1] for v in prof.values())
print("%d threads, %d proofs in %.1f ms (profiling on); kernel time %.1f ms = %.2f ms/proof" % (nthreads, n, dt * 1e3, tot, tot / n))
for k, (cnt, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print("  %-24s %6d launches %9.2f ms  %5.1f %%  (%.3f ms/proof)" % (k, cnt, ms, 100 * ms / tot, ms / n))
tr.free()
