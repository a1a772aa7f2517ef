"""Whole device-side prover replay (trace.ProverTrace): proofs/s on one GPU for a few thread counts."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mapreduce_plonky2_b200 as G
from mapreduce_plonky2_b200 import trace as T

torch.cuda.set_device(0)
G.init(0)
for nthreads in [int(x) for x in (sys.argv[1:] or ["1", "4", "8", "16"])]:
    tr = T.ProverTrace(T.LEAF_PROOF_DEGREES, 1, nthreads)
    tr.run(nthreads)  # warm-up: tables, pools
    n = max(8, 2 * nthreads)
    tr.stage_log = []
    t0 = time.perf_counter()
    tr.run(n)
    dt = time.perf_counter() - t0
    worst = sorted(tr.stage_log, key=lambda r: -sum(ms for _, ms in r[1]))[:2]
    typical = sorted(tr.stage_log, key=lambda r: sum(ms for _, ms in r[1]))[len(tr.stage_log) // 2]
    for tag, rec in [("slowest", w) for w in worst] + [("median", typical)]:
        print("   %s prove(2^%d): %s" % (tag, rec[0], " ".join("%s=%.2f" % kv for kv in rec[1])))
    print("threads %2d: %d leaf proofs (3 prove() each) in %.1f ms -> %.1f proofs/s, %.2f ms/proof" % (
        nthreads, n, dt * 1e3, n / dt, dt / n * 1e3), flush=True)
    tr.free()
