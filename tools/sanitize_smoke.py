"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): both hashers, single-pass and four-step transforms,
quotient, FRI proof, the native prove() -- sizes chosen so the run finishes in a minute under the sanitizer."""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mapreduce_plonky2_b200 as G
from mapreduce_plonky2_b200 import fri as GF, quotient as Q
import plonk_ref as PR
G.init(0)
rng = np.random.default_rng(3)
P = 0xFFFFFFFF00000001
for kind in (0, 1):
    for n_log, ncols in ((5, 3), (10, 20), (15, 9)):   # hash_or_noop leaves, single pass, four-step
        cols = rng.integers(0, P, (ncols, 1 << n_log), dtype=np.uint64)
        pb = G.PolynomialBatch.from_values(cols, 3, False, min(4, n_log + 3), hash_kind=kind, keep_on_device=True)
        pb.open([0, 1, (1 << n_log) - 1]); pb.eval(np.array([[5, 7]], dtype=np.uint64)); pb.free()
    inst = PR.synthetic_instance(3 + kind, degree_bits=5, two_groups=True, with_poseidon=True, extra_gates=True)
    c = inst.circuit
    r = random.Random(9)
    betas, gammas, alphas = ([r.randrange(P) for _ in range(2)] for _ in range(3))
    zs = PR.zs_partial_products(inst, betas, gammas)
    mk = lambda v: G.PolynomialBatch.from_values(np.array(v, dtype=np.uint64), 3, False, 4, hash_kind=kind, keep_on_device=True, fetch_leaves=False)
    b = [mk(inst.constants + inst.sigmas), mk(inst.wires), mk(zs)]
    q = Q.compute_quotient_polys(Q.CircuitDesc.from_circuit(c), b[0], b[1], b[2], betas, gammas, alphas, inst.public_inputs_hash, 3, 4, hash_kind=kind)
    oracles = b + [q]
    ch = GF.Challenger(kind)
    zeta = ch.get_extension_challenge()
    batches = [GF.FriBatchInfo(zeta, [(o, p) for o, bb in enumerate(oracles) for p in range(bb.num_polys)])]
    GF.open_batches(batches, oracles)
    proof = GF.prove_openings(batches, oracles, ch, GF.FriConfig().fri_params(5))
    for x in oracles: x.free()
    # the permutation argument's kernels and the native prove() (all 15 gate kinds incl. CosetInterpolationGate rows)
    from mapreduce_plonky2_b200 import prover as GP
    cfg = GF.FriConfig(proof_of_work_bits=8, num_query_rounds=3)
    cs = mk(inst.constants + inst.sigmas)
    data = GP.prove_native(Q.CircuitDesc.from_circuit(c), cs, [1, 2, 3, 4], np.array(inst.wires, dtype=np.uint64), [7],
                           inst.public_inputs_hash, cfg, hash_kind=kind)
    assert len(data) > 1000
    cs.free()
print("sanitize smoke done; launches:", G.launch_count())
