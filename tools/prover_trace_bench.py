"""Whole device-side prover replay (trace.ProverTrace): proofs/s on one GPU for a few thread counts, through the native
mp2gpu_prove call and through the Python mirror's call sequence (which also logs per-stage times)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mapreduce_plonky2_b200 as G
from mapreduce_plonky2_b200 import trace as T

torch.cuda.set_device(0)
G.init(0)
for nthreads in [int(x) for x in (sys.argv[1:] or ["1", "4", "8", "16"])]:
    for native in (True, False):
        tr = T.ProverTrace(T.LEAF_PROOF_DEGREES, 1, nthreads, native=native)
        for _ in range(3):
            tr.run(2 * nthreads)  # warm-up: tables, pools
        n = max(16, 4 * nthreads)
        if not native:
            tr.stage_log = []
        rates = []
        for _ in range(3):
            t0 = time.perf_counter()
            tr.run(n)
            rates.append(n / (time.perf_counter() - t0))
        if not native:
            typical = sorted(tr.stage_log, key=lambda r: sum(ms for _, ms in r[1]))[len(tr.stage_log) // 2]
            print("   median prove(2^%d): %s" % (typical[0], " ".join("%s=%.2f" % kv for kv in typical[1])))
        print("threads %2d %-6s: %d leaf proofs (3 prove() each) per run -> %s proofs/s" % (
            nthreads, "native" if native else "python", n, " ".join("%.1f" % r for r in rates)), flush=True)
        tr.free()
