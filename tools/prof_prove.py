"""Two whole native proofs at 2^14 over the recursion gate set (for ncu: the second one's kernels are the capture)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mapreduce_plonky2_b200 as G
from mapreduce_plonky2_b200 import trace as T
torch.cuda.set_device(0); G.init(0)
tr = T.ProverTrace((14,), 1, 1)
for _ in range(2):
    tr.prove(14, *tr.inputs[0][14])
print("done", G.launch_count())
