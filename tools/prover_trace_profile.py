"""cProfile of one prover thread replaying whole proofs (where the host time of trace.ProverTrace goes)."""
import cProfile, os, pstats, sys, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mapreduce_plonky2_b200 as G
from mapreduce_plonky2_b200 import trace as T
torch.cuda.set_device(0); G.init(0)
tr = T.ProverTrace(T.LEAF_PROOF_DEGREES, 1, 1)
for d in tr.degrees: tr.prove(d, *tr.inputs[0][d])
pr = cProfile.Profile(); pr.enable()
for _ in range(4):
    for d in tr.degrees: tr.prove(d, *tr.inputs[0][d])
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
