"""Run-to-run variance of trace.ProverTrace (8 threads, 32 proofs), alone and after the stream-based TraceRunner."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mapreduce_plonky2_b200 as G
from mapreduce_plonky2_b200 import trace as T
torch.cuda.set_device(0); G.init(0)
def go(tag, reps=5):
    tr = T.ProverTrace(T.LEAF_PROOF_DEGREES, 1, 8)
    tr.run(8); tr.run(8)
    for r in range(reps):
        tr.stage_log = []
        t0 = time.perf_counter(); tr.run(32); dt = time.perf_counter() - t0
        tot = {}
        for _, rec in tr.stage_log:
            for k, v in rec: tot[k] = tot.get(k, 0) + v
        print("%s run %d: %.1f proofs/s; stage sums (ms over all threads): %s" % (tag, r, 32 / dt, " ".join("%s=%.0f" % kv for kv in tot.items())), flush=True)
    tr.free()
go("alone")
runner = T.TraceRunner(T.LEAF_PROOF_DEGREES, 1, 8)
for _ in range(3): runner.run(32)
torch.cuda.synchronize()
del runner; torch.cuda.empty_cache()
go("after TraceRunner")
