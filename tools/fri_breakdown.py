"""Where the wall-clock of one fri_proof goes (degree 2^14, standard_recursion_config): per-step host timers around
the calls fri.fri_proof makes."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mapreduce_plonky2_b200 as G
from mapreduce_plonky2_b200 import fri as GF
P = 0xFFFFFFFF00000001
kind = 1
G.init(0)
rng = np.random.default_rng(7)
degree_bits = 14
n = 1 << degree_bits
widths = (85, 135, 20, 16)
oracles = [G.PolynomialBatch.from_coeffs(list(rng.integers(0, P, (w, n), dtype=np.uint64)), 3, False, 4, hash_kind=kind,
                                         keep_on_device=True, fetch_leaves=False) for w in widths]
zeta, gzeta = rng.integers(0, P, 2, dtype=np.uint64), rng.integers(0, P, 2, dtype=np.uint64)
batches = [GF.FriBatchInfo(zeta, [(o, p) for o, w in enumerate(widths) for p in range(w)]), GF.FriBatchInfo(gzeta, [(2, 0), (2, 1)])]
params = GF.FriConfig().fri_params(degree_bits)
for it in range(3):
    ch = GF.Challenger(kind)
    alpha = ch.get_extension_challenge()
    phase = G.FriCommitPhase.from_openings(oracles, [(b.point, b.polynomials) for b in batches], alpha, 4, kind)
    T = {}
    def tick(name, t0):
        T[name] = T.get(name, 0) + (time.perf_counter() - t0) * 1e3
    for arity_bits in params.reduction_arity_bits:
        t0 = time.perf_counter(); cap = phase.commit_layer(arity_bits); tick("commit_layer", t0)
        t0 = time.perf_counter(); ch.observe_cap(cap); beta = ch.get_extension_challenge(); tick("transcript", t0)
        t0 = time.perf_counter(); phase.fold(beta); tick("fold", t0)
    t0 = time.perf_counter(); final_poly = phase.finish(); tick("finish", t0)
    t0 = time.perf_counter(); ch.observe_extension_elements(final_poly); tick("transcript", t0)
    t0 = time.perf_counter(); w = GF.fri_proof_of_work(ch, params.config); tick("pow", t0)
    t0 = time.perf_counter(); x_indices = [c % (1 << params.lde_bits) for c in ch.get_n_challenges(28)]; tick("transcript", t0)
    t0 = time.perf_counter(); initial = [o.open(x_indices) for o in oracles]; tick("open_initial", t0)
    idx = list(x_indices)
    t0 = time.perf_counter()
    for i, ab in enumerate(params.reduction_arity_bits):
        idx = [x >> ab for x in idx]
        phase.open_layer(i, idx)
    tick("open_layers", t0)
    phase.free()
print(" ".join("%s=%.3f" % kv for kv in T.items()), "total=%.3f ms" % sum(T.values()))
