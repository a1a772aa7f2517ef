"""The oracle's restatement of plonky2's `compute_quotient_polys` (oracle/quotient.py) against the VERIFIER's
equation evaluated by definition (tests/plonk_ref.py): prove, then verify -- the way every `run_circuit` test of
the reference pins its prover.  CPU only."""
import random

import numpy as np
import pytest

import plonk_ref as PR
from oracle import quotient as OQ

P = PR.P


def _coeffs(cols):
    import pyref as R
    return np.array([R.ifft(list(c)) for c in cols], dtype=np.uint64)


def test_poseidon_gate_witness_satisfies_its_constraints_and_breaks_when_tampered():
    import pyref as R
    rng = random.Random(0x90)
    for swap in (0, 1):
        ins = [rng.randrange(P) for _ in range(12)]
        w = PR.poseidon_gate_trace(ins, swap)
        assert PR.poseidon_gate_constraints(w) == [0] * 123
        perm_in = ins[4:8] + ins[0:4] + ins[8:] if swap else ins
        assert w[12:24] == R.poseidon(perm_in)          # the gate computes plonky2's permutation (with the swap)
        for j in (0, 24, 26, 40, 70, 100, 20):
            bad = list(w)
            bad[j] = (bad[j] + 1) % P
            assert any(PR.poseidon_gate_constraints(bad))


@pytest.mark.parametrize("seed,degree_bits,two_groups,with_poseidon,extra", [
    (1, 3, False, False, False), (2, 4, False, False, False), (3, 4, True, False, False), (4, 5, True, False, False),
    (5, 3, False, True, False), (6, 4, True, True, False), (7, 4, False, False, True), (8, 4, True, True, True)])
def test_quotient_passes_the_verifier_identity(oracle, seed, degree_bits, two_groups, with_poseidon, extra):
    rng = random.Random(0x5151 + seed)
    inst = PR.synthetic_instance(seed, degree_bits=degree_bits, two_groups=two_groups, with_poseidon=with_poseidon,
                                 extra_gates=extra)
    c = inst.circuit
    betas, gammas, alphas = ([rng.randrange(P) for _ in range(c.num_challenges)] for _ in range(3))
    zs_pp = PR.zs_partial_products(inst, betas, gammas)
    chunks = OQ.compute_quotient_polys(c, _coeffs(inst.constants + inst.sigmas), _coeffs(inst.wires), _coeffs(zs_pp),
                                       betas, gammas, alphas, inst.public_inputs_hash)
    assert chunks.shape == (c.num_challenges * c.max_degree, c.n)
    for _ in range(3):
        zeta = rng.randrange(2, P)
        assert PR.check_quotient_identity(inst, zs_pp, [list(map(int, ch)) for ch in chunks], betas, gammas, alphas, zeta)
    # a quotient of the right shape but wrong content is rejected
    bad = [list(map(int, ch)) for ch in chunks]
    bad[1][0] = (bad[1][0] + 1) % P
    assert not PR.check_quotient_identity(inst, zs_pp, bad, betas, gammas, alphas, rng.randrange(2, P))


def test_violated_gate_makes_the_vanishing_polynomial_indivisible(oracle):
    """With a broken witness Z_H does not divide the vanishing polynomial: the 'quotient' computed pointwise on the
    8n coset has degree >= 8n - ... and fails the identity (plonky2 panics in trim_to_len at this point)."""
    rng = random.Random(7)
    inst = PR.synthetic_instance(9, degree_bits=4)
    c = inst.circuit
    row = inst.row_gate.index(0)
    inst.wires[3][row] = (inst.wires[3][row] + 1) % P        # break one arithmetic output (its copy class may break too)
    betas, gammas, alphas = ([rng.randrange(P) for _ in range(2)] for _ in range(3))
    try:
        zs_pp = PR.zs_partial_products(inst, betas, gammas)
    except AssertionError:
        return  # the grand product already refuses the witness
    chunks = OQ.compute_quotient_polys(c, _coeffs(inst.constants + inst.sigmas), _coeffs(inst.wires), _coeffs(zs_pp),
                                       betas, gammas, alphas, inst.public_inputs_hash)
    assert not PR.check_quotient_identity(inst, zs_pp, [list(map(int, ch)) for ch in chunks], betas, gammas, alphas,
                                          rng.randrange(2, P))


def test_coset_interpolation_gate_interpolates_and_both_restatements_agree(oracle):
    """CosetInterpolationGate (plonky2 gates/coset_interpolation.rs): the generated witness satisfies the constraints, its
    evaluation_value IS the Lagrange interpolant of the values on shift * <w> at evaluation_point (by definition, in
    GF(p^2)), tampering breaks a constraint, and the oracle's restatement (closed-form weights w^k / m, one loop with
    cuts) agrees with the by-definition one (product weights, run by run) on random wires for several shapes."""
    rng = random.Random(0xC1)
    assert PR.coset_interpolation_degree(4, 8) == 6 and PR.coset_interpolation_degree(2, 8) == 4
    assert PR.coset_interpolation_layout(4, 6)[4:] == (37, 45, 47)      # routed wires, shifted point, total wires
    for bits, deg in [(2, 4), (3, 4), (3, 8), (4, 6), (4, 3), (5, 8), (1, 2)]:
        total = PR.coset_interpolation_layout(bits, deg)[-1]
        for _ in range(2):
            w = [rng.randrange(P) for _ in range(total)]
            a = PR.coset_interpolation_constraints(w, bits, deg)
            assert a == OQ._coset_interpolation_gate(w, bits, deg)
            assert len(a) == PR.Gate("coset_interpolation", bits, deg).num_constraints
    inst = PR.synthetic_instance(31, degree_bits=5, num_wires=135, num_routed_wires=80, two_groups=True, with_poseidon=True,
                                 extra_gates=True)
    c = inst.circuit
    rows = [r for r, g in enumerate(inst.row_gate) if c.gates[g].kind == "coset_interpolation"]
    assert {c.gates[inst.row_gate[r]].num_ops for r in rows} == {3, 4}
    E = PR.Ext
    for r in rows:
        g = c.gates[inst.row_gate[r]]
        w = [inst.wires[j][r] for j in range(c.num_wires)]
        assert PR.coset_interpolation_constraints(w, g.num_ops, g.param) == [0] * g.num_constraints
        npts, _, at_point, at_value, *_ = PR.coset_interpolation_layout(g.num_ops, g.param)
        dom, shift = PR.subgroup(g.num_ops), w[0]
        x = E(w[at_point], w[at_point + 1])
        total = E(0)
        for i in range(npts):
            num, den = E(1), 1
            for j in range(npts):
                if j != i:
                    num = num * (x - shift * dom[j] % P)
                    den = den * ((shift * dom[i] - shift * dom[j]) % P) % P
            total = total + E(w[1 + 2 * i], w[2 + 2 * i]) * num * pow(den, P - 2, P)
        assert total == E(w[at_value], w[at_value + 1])
        bad = list(w)
        bad[5] = (bad[5] + 1) % P
        assert any(PR.coset_interpolation_constraints(bad, g.num_ops, g.param))


def test_quotient_with_the_recursion_gate_set_passes_the_verifier_identity(oracle):
    """135 wires / 80 routed (standard_recursion_config) with all 14 supported gate kinds incl. both CosetInterpolationGate
    shapes: the oracle's quotient passes the verifier's identity."""
    rng = random.Random(0xC2)
    inst = PR.synthetic_instance(31, degree_bits=5, num_wires=135, num_routed_wires=80, two_groups=True, with_poseidon=True,
                                 extra_gates=True)
    c = inst.circuit
    betas, gammas, alphas = ([rng.randrange(P) for _ in range(c.num_challenges)] for _ in range(3))
    zs_pp = PR.zs_partial_products(inst, betas, gammas)
    chunks = OQ.compute_quotient_polys(c, _coeffs(inst.constants + inst.sigmas), _coeffs(inst.wires), _coeffs(zs_pp),
                                       betas, gammas, alphas, inst.public_inputs_hash)
    for _ in range(2):
        assert PR.check_quotient_identity(inst, zs_pp, [list(map(int, ch)) for ch in chunks], betas, gammas, alphas,
                                          rng.randrange(2, P))
