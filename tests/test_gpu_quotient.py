"""GPU parity of mp2gpu_quotient_polys (plonky2 plonk/prover.rs compute_quotient_polys + the quotient commitment;
SURVEY.md 8(f) row 3) through the C ABI: bit-exact against the oracle's restatement (oracle/quotient.py), accepted by
the by-definition verifier identity (tests/plonk_ref.py) -- prove then verify, as every `run_circuit` test of the
reference does -- and the commitment of the chunks equals from_coeffs of the same chunks."""
import os
import random

import numpy as np
import pytest

import plonk_ref as PR

pytestmark = pytest.mark.gpu
P = PR.P


def _coeffs(cols):
    import pyref as R
    return np.array([R.ifft(list(c)) for c in cols], dtype=np.uint64)


def _commit3(G, inst, zs_pp, rate_bits, cap, kind):
    vals = lambda cols: np.array(cols, dtype=np.uint64)
    mk = lambda cols: G.PolynomialBatch.from_values(vals(cols), rate_bits, False, cap, hash_kind=kind, keep_on_device=True,
                                                    fetch_leaves=False)
    return mk(inst.constants + inst.sigmas), mk(inst.wires), mk(zs_pp)


@pytest.mark.parametrize("seed,degree_bits,two_groups,kind,rate_bits,qbits,with_poseidon", [
    (1, 3, False, 0, 3, 3, False), (2, 4, False, 1, 3, 3, False), (3, 4, True, 0, 3, 3, False), (4, 5, True, 1, 3, 3, False),
    (5, 6, True, 0, 3, 2, False), (6, 12, True, 1, 3, 3, False),
    (7, 3, False, 0, 3, 3, True), (8, 5, True, 1, 3, 3, True), (9, 10, True, 1, 3, 3, True),
    (10, 4, False, 0, 3, 3, "extra"), (11, 5, True, 1, 3, 3, "both"), (12, 9, True, 0, 3, 3, "both")])
def test_quotient_matches_oracle_and_passes_the_verifier_identity(oracle, seed, degree_bits, two_groups, kind, rate_bits, qbits,
                                                                  with_poseidon):
    import mapreduce_plonky2_b200 as G
    from mapreduce_plonky2_b200 import quotient as Q
    from oracle import quotient as OQ

    G.init(0)
    rng = random.Random(0x7171 + seed)
    inst = PR.synthetic_instance(seed, degree_bits=degree_bits, two_groups=two_groups,
                                 with_poseidon=with_poseidon in (True, "both"), extra_gates=with_poseidon in ("extra", "both"))
    c = inst.circuit
    c.quotient_degree_bits = qbits
    betas, gammas, alphas = ([rng.randrange(P) for _ in range(c.num_challenges)] for _ in range(3))
    zs_pp = PR.zs_partial_products(inst, betas, gammas)
    cap = min(4, degree_bits + rate_bits)
    b_cs, b_w, b_z = _commit3(G, inst, zs_pp, rate_bits, cap, kind)
    desc = Q.CircuitDesc.from_circuit(c)
    # odd seeds read the batches' row-major leaves (what big batches keep), even seeds the column-major LDE
    os.environ["MP2_QUOTIENT_ROWMAJOR"] = "1" if seed % 2 else "0"
    try:
        qb = Q.compute_quotient_polys(desc, b_cs, b_w, b_z, betas, gammas, alphas, inst.public_inputs_hash, rate_bits, cap,
                                      hash_kind=kind, fetch_leaves=True)
    finally:
        os.environ.pop("MP2_QUOTIENT_ROWMAJOR", None)
    chunks = qb.polynomials
    assert chunks.shape == (c.num_challenges * c.max_degree, c.n)
    if degree_bits <= 6:  # the pure-Python restatement is O(N * terms) big-int work
        ref = OQ.compute_quotient_polys(c, _coeffs(inst.constants + inst.sigmas), _coeffs(inst.wires), _coeffs(zs_pp),
                                        betas, gammas, alphas, inst.public_inputs_hash)
        assert np.array_equal(chunks, ref)
        for _ in range(2):
            zeta = rng.randrange(2, P)
            assert PR.check_quotient_identity(inst, zs_pp, [list(map(int, ch)) for ch in chunks], betas, gammas, alphas, zeta)
    else:
        # size-independent property: the chunks must open consistently with the verifier's equation at a random
        # point, with every opening taken from the DEVICE batches (mp2gpu_batch_eval)
        zeta = rng.randrange(2, P)
        g = PR.R.root_of_unity(c.degree_bits)
        pts = np.array([[zeta, 0], [g * zeta % P, 0]], dtype=np.uint64)
        o_cs, o_w, o_z, o_q = (b.eval(pts)[..., 0] for b in (b_cs, b_w, b_z, qb))   # base-field points: second component 0
        nch, npp = c.num_challenges, c.num_partial_products
        van = PR.eval_vanishing_poly(c, zeta, [int(v) for v in o_cs[0][:c.num_constants]], [int(v) for v in o_w[0]],
                                     inst.public_inputs_hash, [int(v) for v in o_z[0][:nch]], [int(v) for v in o_z[1][:nch]],
                                     [int(v) for v in o_z[0][nch:]], [int(v) for v in o_cs[0][c.num_constants:]],
                                     betas, gammas, alphas)
        z_h, zeta_n = (pow(zeta, c.n, P) - 1) % P, pow(zeta, c.n, P)
        for i in range(nch):
            t = 0
            for k in reversed(range(c.max_degree)):
                t = (t * zeta_n + int(o_q[0][i * c.max_degree + k])) % P
            assert van[i] == z_h * t % P
    # the quotient commitment is from_coeffs of the chunks
    want = oracle.commit(chunks, rate_bits, cap, kind, True, want_leaves=True)
    assert np.array_equal(qb.merkle_tree.cap.hashes, want["cap"])
    assert np.array_equal(qb.merkle_tree.digests, want["digests"])
    assert np.array_equal(qb.merkle_tree.leaves, want["leaves"])
    for b in (b_cs, b_w, b_z, qb):
        b.free()


def test_unsupported_gate_and_shape_errors(oracle):
    import mapreduce_plonky2_b200 as G
    from mapreduce_plonky2_b200 import quotient as Q

    G.init(0)
    inst = PR.synthetic_instance(11, degree_bits=3)
    c = inst.circuit
    betas = gammas = alphas = [3, 5]
    zs_pp = PR.zs_partial_products(inst, betas, gammas)
    b_cs, b_w, b_z = _commit3(G, inst, zs_pp, 3, 2, 0)
    desc = Q.CircuitDesc.from_circuit(c)
    desc.gates[2] = Q.GateDesc("lookup")
    with pytest.raises(G.Mp2GpuError, match="outside the supported subset"):
        Q.compute_quotient_polys(desc, b_cs, b_w, b_z, betas, gammas, alphas, inst.public_inputs_hash, 3, 2, hash_kind=0)
    for bad, msg in ((Q.GateDesc("coset_interpolation", 7, 4), "subgroup_bits must be"), (Q.GateDesc("coset_interpolation", 2, 1), "degree must be"),
                     (Q.GateDesc("coset_interpolation", 4, 6), "exceeds the wires")):
        desc = Q.CircuitDesc.from_circuit(c)
        desc.gates[2] = bad
        with pytest.raises(G.Mp2GpuError, match=msg):
            Q.compute_quotient_polys(desc, b_cs, b_w, b_z, betas, gammas, alphas, inst.public_inputs_hash, 3, 2, hash_kind=0)
    desc = Q.CircuitDesc.from_circuit(c)
    with pytest.raises(G.Mp2GpuError, match="wires batch must hold"):
        Q.compute_quotient_polys(desc, b_cs, b_z, b_z, betas, gammas, alphas, inst.public_inputs_hash, 3, 2, hash_kind=0)
    desc.quotient_degree_bits = 4
    with pytest.raises(G.Mp2GpuError, match="exceeds a batch's rate_bits"):
        Q.compute_quotient_polys(desc, b_cs, b_w, b_z, betas, gammas, alphas, inst.public_inputs_hash, 3, 2, hash_kind=0)


@pytest.mark.parametrize("nch,routed,qbits,degree_bits", [(1, 8, 3, 4), (3, 12, 2, 5), (4, 20, 3, 4), (2, 9, 1, 6)])
def test_quotient_other_shapes(oracle, nch, routed, qbits, degree_bits):
    """1 / 3 / 4 challenges, routed-wire counts that are not multiples of the chunk size, quotient degree factors 2 / 4 / 8."""
    import mapreduce_plonky2_b200 as G
    from mapreduce_plonky2_b200 import quotient as Q
    from oracle import quotient as OQ

    G.init(0)
    rng = random.Random(0xA5 + nch * 7 + routed)
    inst = PR.synthetic_instance(100 + nch, degree_bits=degree_bits, num_wires=routed + 3, num_routed_wires=routed,
                                 two_groups=(nch % 2 == 0))
    c = inst.circuit
    c.num_challenges, c.quotient_degree_bits = nch, qbits
    betas, gammas, alphas = ([rng.randrange(P) for _ in range(nch)] for _ in range(3))
    zs_pp = PR.zs_partial_products(inst, betas, gammas)
    b_cs, b_w, b_z = _commit3(G, inst, zs_pp, 3, 4, 0)
    qb = Q.compute_quotient_polys(Q.CircuitDesc.from_circuit(c), b_cs, b_w, b_z, betas, gammas, alphas, inst.public_inputs_hash,
                                  3, 4, hash_kind=0)
    ref = OQ.compute_quotient_polys(c, _coeffs(inst.constants + inst.sigmas), _coeffs(inst.wires), _coeffs(zs_pp), betas, gammas,
                                    alphas, inst.public_inputs_hash)
    # bit-exact against the oracle whatever the degrees (for qbits < 3 the pointwise "quotient" is not low-degree -- the
    # circuit's constraints have degree up to 7 -- but both sides compute the same function on the same coset)
    assert np.array_equal(qb.polynomials, ref)
    if qbits == 3:
        zeta = rng.randrange(2, P)
        assert PR.check_quotient_identity(inst, zs_pp, [list(map(int, ch)) for ch in qb.polynomials], betas, gammas, alphas, zeta)
    for b in (b_cs, b_w, b_z, qb):
        b.free()


@pytest.mark.parametrize("seed,degree_bits,kind", [(31, 5, 1), (32, 7, 0)])
def test_recursion_gate_set_with_both_coset_interpolation_shapes(oracle, seed, degree_bits, kind):
    """standard_recursion_config's shape (135 wires, 80 routed) with every supported gate kind, including the
    CosetInterpolationGate the recursive FRI verifier uses (16 points, degree 6: runs of 6 + 5 + 5) next to an 8-point one."""
    import mapreduce_plonky2_b200 as G
    from mapreduce_plonky2_b200 import quotient as Q
    from oracle import quotient as OQ

    G.init(0)
    rng = random.Random(0xC0 + seed)
    inst = PR.synthetic_instance(seed, degree_bits=degree_bits, num_wires=135, num_routed_wires=80, two_groups=True,
                                 with_poseidon=True, extra_gates=True)
    c = inst.circuit
    kinds = [(g.kind, g.num_ops, g.param) for g in c.gates]
    assert ("coset_interpolation", 4, 6) in kinds and ("coset_interpolation", 3, 4) in kinds
    used = {c.gates[g].kind for g in inst.row_gate}
    betas, gammas, alphas = ([rng.randrange(P) for _ in range(c.num_challenges)] for _ in range(3))
    zs_pp = PR.zs_partial_products(inst, betas, gammas)
    b_cs, b_w, b_z = _commit3(G, inst, zs_pp, 3, 4, kind)
    qb = Q.compute_quotient_polys(Q.CircuitDesc.from_circuit(c), b_cs, b_w, b_z, betas, gammas, alphas, inst.public_inputs_hash,
                                  3, 4, hash_kind=kind)
    chunks = qb.polynomials
    if degree_bits <= 5:
        want = OQ.compute_quotient_polys(c, _coeffs(inst.constants + inst.sigmas), _coeffs(inst.wires), _coeffs(zs_pp), betas,
                                         gammas, alphas, inst.public_inputs_hash)
        assert np.array_equal(chunks, want)
    for _ in range(2):
        assert PR.check_quotient_identity(inst, zs_pp, [list(map(int, ch)) for ch in chunks], betas, gammas, alphas,
                                          rng.randrange(2, P))
    assert "coset_interpolation" in used, "the seed must place the gate on some row"
    for b in (b_cs, b_w, b_z, qb):
        b.free()
