"""Whole proofs: mapreduce_plonky2_b200.prover.prove (plonky2 plonk/prover.rs from the witness on; every data-path step
on the device: three commitments, quotient polynomials, openings, FRI) accepted by a by-definition restatement of the
VERIFIER (tests/plonk_ref.verify_proof: transcript re-derived, vanishing identity evaluated in GF(p^2), pyref's FRI
verifier) -- the prove-then-verify pin every `run_circuit` test of the reference uses (e.g. mp2-test/src/circuit.rs:46,107),
here across all the components at once: a wrong transcript order, opening layout, quotient convention, leaf order or FRI
fold cannot pass.  Tampered proofs are rejected."""
import copy
import random

import numpy as np
import pytest

import plonk_ref as PR

pytestmark = pytest.mark.gpu
P = PR.P


def _as_dict(proof):
    o = proof.openings
    pairs = lambda a: [tuple(int(x) for x in v) for v in a]
    fri = proof.opening_proof
    return {"wires_cap": proof.wires_cap.hashes.tolist(), "zs_pp_cap": proof.plonk_zs_partial_products_cap.hashes.tolist(),
            "quotient_cap": proof.quotient_polys_cap.hashes.tolist(),
            "openings": {"constants": pairs(o.constants), "sigmas": pairs(o.plonk_sigmas), "wires": pairs(o.wires),
                         "zs": pairs(o.plonk_zs), "partial_products": pairs(o.partial_products),
                         "quotient": pairs(o.quotient_polys), "zs_next": pairs(o.plonk_zs_next)},
            "fri": {"caps": [c.hashes.tolist() for c in fri.commit_phase_merkle_caps], "final_poly": fri.final_poly.tolist(),
                    "pow_witness": fri.pow_witness,
                    "rounds": [{"initial": [(r.tolist(), m.siblings.tolist()) for r, m in rnd.initial_trees_proof],
                                "steps": [(s.evals.tolist(), s.merkle_proof.siblings.tolist()) for s in rnd.steps]}
                               for rnd in fri.query_round_proofs]}}


@pytest.mark.parametrize("seed,degree_bits,kind,with_poseidon,extra", [(1, 5, 0, False, False), (2, 6, 1, True, False),
                                                                       (3, 7, 1, True, True), (4, 9, 0, True, True)])
def test_gpu_proof_is_accepted_by_the_by_definition_verifier(seed, degree_bits, kind, with_poseidon, extra):
    import mapreduce_plonky2_b200 as G
    from mapreduce_plonky2_b200 import fri as GF, prover, quotient as Q

    G.init(0)
    inst = PR.synthetic_instance(seed, degree_bits=degree_bits, two_groups=True, with_poseidon=with_poseidon, extra_gates=extra)
    c = inst.circuit
    cfg = GF.FriConfig(proof_of_work_bits=10, num_query_rounds=6)
    cs = G.PolynomialBatch.from_values(np.array(inst.constants + inst.sigmas, dtype=np.uint64), cfg.rate_bits, False,
                                       cfg.cap_height, hash_kind=kind, keep_on_device=True, fetch_leaves=False)
    digest = G.circuit_digest(cs.merkle_tree.cap.hashes, degree_bits, kind)
    zs_fn = lambda betas, gammas: np.array(PR.zs_partial_products(inst, betas, gammas), dtype=np.uint64)
    proof = prover.prove(Q.CircuitDesc.from_circuit(c), cs, digest, np.array(inst.wires, dtype=np.uint64),
                         inst.public_inputs_hash, zs_fn, cfg, kind)
    d = _as_dict(proof)
    cs_cap = cs.merkle_tree.cap.hashes.tolist()
    PR.verify_proof(c, digest.tolist(), cs_cap, inst.public_inputs_hash, d, kind, pow_bits=10)
    # ... and as the bytes the reference moves between tasks: bincode(ProofWithVK) -> parse -> the same proof verifies
    from mapreduce_plonky2_b200 import wire as W

    pvk = W.ProofWithVK(W.ProofWithPublicInputs(proof, np.array(inst.public_inputs_hash, dtype=np.uint64)),
                        W.VerifierOnlyCircuitData(cs.merkle_tree.cap, digest))
    data = pvk.serialize()
    back = W.ProofWithVK.deserialize(data)
    assert back.serialize() == data
    PR.verify_proof(c, back.vk.circuit_digest.tolist(), back.vk.constants_sigmas_cap.hashes.tolist(), inst.public_inputs_hash,
                    _as_dict(back.proof.proof), kind, pow_bits=10)
    # tampering: an opening, a cap, the final polynomial, a public input
    rng = random.Random(seed)
    for what in ("opening", "quotient_opening", "cap", "final_poly", "public_input"):
        bad, pi = copy.deepcopy(d), list(inst.public_inputs_hash)
        if what == "opening":
            k = rng.randrange(len(bad["openings"]["wires"]))
            bad["openings"]["wires"][k] = ((bad["openings"]["wires"][k][0] + 1) % P, bad["openings"]["wires"][k][1])
        elif what == "quotient_opening":
            bad["openings"]["quotient"][0] = (bad["openings"]["quotient"][0][0], (bad["openings"]["quotient"][0][1] + 1) % P)
        elif what == "cap":
            bad["zs_pp_cap"][3][2] = (bad["zs_pp_cap"][3][2] + 1) % P
        elif what == "final_poly":
            bad["fri"]["final_poly"][0][0] = (bad["fri"]["final_poly"][0][0] + 1) % P
        else:
            pi[1] = (pi[1] + 1) % P
        with pytest.raises(AssertionError):
            PR.verify_proof(c, digest.tolist(), cs_cap, pi, bad, kind, pow_bits=10)
    cs.free()


def test_invalid_witness_does_not_yield_an_accepted_proof():
    """A broken gate output: Z_H no longer divides the vanishing polynomial, the 'quotient' the device computes pointwise is
    not the low-degree polynomial the identity needs, and the verifier refuses (plonky2's prover would panic earlier)."""
    import mapreduce_plonky2_b200 as G
    from mapreduce_plonky2_b200 import fri as GF, prover, quotient as Q

    G.init(0)
    inst = PR.synthetic_instance(21, degree_bits=5, two_groups=True)
    c = inst.circuit
    row = inst.row_gate.index(1)                      # a ConstantGate row: wire 0 must equal the gate constant
    inst.wires[0][row] = (inst.wires[0][row] + 1) % P
    cfg = GF.FriConfig(proof_of_work_bits=8, num_query_rounds=4)
    cs = G.PolynomialBatch.from_values(np.array(inst.constants + inst.sigmas, dtype=np.uint64), 3, False, 4, hash_kind=0,
                                       keep_on_device=True, fetch_leaves=False)
    digest = G.circuit_digest(cs.merkle_tree.cap.hashes, 5, 0)

    def zs_fn(betas, gammas):
        try:
            return np.array(PR.zs_partial_products(inst, betas, gammas), dtype=np.uint64)
        except AssertionError:      # the grand product does not close either: commit what the recurrence gives
            pytest.skip("the copy constraints already refuse this witness")

    proof = prover.prove(Q.CircuitDesc.from_circuit(c), cs, digest, np.array(inst.wires, dtype=np.uint64),
                         inst.public_inputs_hash, zs_fn, cfg, 0)
    with pytest.raises(AssertionError):
        PR.verify_proof(c, digest.tolist(), cs.merkle_tree.cap.hashes.tolist(), inst.public_inputs_hash, _as_dict(proof), 0, pow_bits=8)
    cs.free()
