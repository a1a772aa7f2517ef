"""prove -> verify over the oracle's FRI pieces (CPU): a proof assembled from the oracle's commitments, its
prove_openings combination, commit phase, PoW and Merkle paths is accepted by the by-definition verifier in
pyref (plonky2 fri/verifier.rs restated), and tampering anywhere is rejected.  This is the structural pin the
reference's own tests use for this boundary (prove then verify: every run_circuit test; native tree + proof
accepted by the verifier: recursion-framework/src/universal_verifier_gadget/circuit_set.rs:296-371) -- it ties
leaf order, coset shifts, fold and quotient conventions to each other, not only to a second restatement."""
import copy

import numpy as np
import pytest

import fri_ref
import pyref
from util import field_elems


def _setup(oracle, degree_bits, widths, kind, rounds=5, pow_bits=6):
    n = 1 << degree_bits
    coeff_sets = [field_elems(0xF0 + 3 * k + degree_bits, (w, n)) for k, w in enumerate(widths)]
    zeta, gzeta = (tuple(int(v) for v in field_elems(0x5E7A + i, 2)) for i in range(2))
    batches = fri_ref.plonky2_instance(widths, zeta, gzeta)
    commits, openings, proof = fri_ref.oracle_fri_proof(oracle, coeff_sets, batches, degree_bits, kind,
                                                        pow_bits=pow_bits, num_query_rounds=rounds)
    return batches, commits, openings, proof


def _verify(batches, commits, openings, proof, degree_bits, kind, pow_bits=6):
    ch = pyref.Challenger(kind)
    caps = [c["cap"].tolist() for c in commits]
    fri_ref.transcript_head(ch.observe, caps, openings)
    p = {"caps": [c.tolist() for c in proof["caps"]], "final_poly": proof["final_poly"].tolist(),
         "pow_witness": proof["pow_witness"],
         "rounds": [{"initial": [(r.tolist(), s.tolist()) for r, s in rnd["initial"]],
                     "steps": [(e.tolist(), s.tolist()) for e, s in rnd["steps"]]} for rnd in proof["rounds"]]}
    pyref.verify_fri_proof(batches, openings, caps, p, ch, degree_bits, fri_ref.arity_schedule(degree_bits),
                           pow_bits=pow_bits, kind=kind)


@pytest.mark.parametrize("kind,degree_bits,widths", [(0, 6, (3, 5, 4, 2)), (1, 6, (2, 9, 3)), (1, 10, (3, 2))])
def test_oracle_fri_proof_verifies(oracle, kind, degree_bits, widths):
    batches, commits, openings, proof = _setup(oracle, degree_bits, widths, kind)
    assert len(proof["caps"]) == len(fri_ref.arity_schedule(degree_bits)) >= 1
    _verify(batches, commits, openings, proof, degree_bits, kind)


def test_tampered_fri_proofs_are_rejected(oracle):
    kind, degree_bits, widths = 1, 6, (3, 4, 2)
    batches, commits, openings, proof = _setup(oracle, degree_bits, widths, kind)
    _verify(batches, commits, openings, proof, degree_bits, kind)
    # a wrong claimed opening
    bad = copy.deepcopy(openings)
    bad[0][1] = ((bad[0][1][0] + 1) % pyref.P, bad[0][1][1])
    with pytest.raises(AssertionError):
        _verify(batches, commits, bad, proof, degree_bits, kind)
    # a wrong final polynomial coefficient (changes the transcript, hence PoW / indices / the final check)
    bad = copy.deepcopy(proof)
    bad["final_poly"][0, 0] ^= np.uint64(1)
    with pytest.raises(AssertionError):
        _verify(batches, commits, openings, bad, degree_bits, kind)
    # a wrong opened row
    bad = copy.deepcopy(proof)
    row, sib = bad["rounds"][0]["initial"][1]
    row = row.copy()
    row[0] ^= np.uint64(1)
    bad["rounds"][0]["initial"][1] = (row, sib)
    with pytest.raises(AssertionError, match="initial tree proof"):
        _verify(batches, commits, openings, bad, degree_bits, kind)
    # a wrong layer evaluation (not the one checked for consistency: caught by the layer's Merkle proof)
    bad = copy.deepcopy(proof)
    ev, sib = bad["rounds"][2]["steps"][0]
    ev = ev.copy()
    ev[:, 1] ^= np.uint64(2)
    bad["rounds"][2]["steps"][0] = (ev, sib)
    with pytest.raises(AssertionError):
        _verify(batches, commits, openings, bad, degree_bits, kind)
    # a wrong proof-of-work witness
    bad = copy.deepcopy(proof)
    bad["pow_witness"] += 1
    with pytest.raises(AssertionError):
        _verify(batches, commits, openings, bad, degree_bits, kind)


def test_oracle_equals_fri_golden(oracle):
    """tests/golden/fri_small.json (pyref, by definition): prove_openings' final polynomial, the layer caps of the
    commit phase and the remaining coefficients -- the oracle's FFT-based restatement must reproduce them."""
    cases = fri_ref.load_fri_golden()
    assert len(cases) == 2
    for c in cases:
        final = oracle.fri_combine([(z, [c["oracles"][o][p] for o, p in polys]) for z, polys in c["batches"]], c["alpha"])
        assert np.array_equal(final, c["final_poly"])
        n = 1 << c["degree_bits"]
        padded = np.zeros((n << c["rate_bits"], 2), dtype=np.uint64)
        padded[:n] = final
        trees, rest = oracle.fri_committed_trees(padded, oracle.coset_fft_ext(padded, 7), c["arity_bits"], c["betas"],
                                                 c["cap_height"], c["hash_kind"], c["rate_bits"])
        assert len(trees) == len(c["layer_caps"])
        for (leaves, digests, cap), want_cap, want_xor in zip(trees, c["layer_caps"], c["layer_digests_xor"]):
            assert np.array_equal(cap, want_cap)
            assert np.array_equal(np.bitwise_xor.reduce(digests, axis=0), want_xor)
        assert np.array_equal(rest, c["final_coeffs"])
