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


@pytest.mark.parametrize("seed,degree_bits,two_groups,kind,with_poseidon,default_cfg", [
    (21, 5, False, 0, False, True), (22, 6, True, 1, False, False), (23, 8, True, 1, True, False),
    (24, 10, True, 0, "both", False), (25, 12, True, 1, True, False), (26, 15, True, 1, False, False)])
def test_native_prove_is_byte_identical_to_the_python_mirror_and_verifies(seed, degree_bits, two_groups, kind, with_poseidon,
                                                                          default_cfg):
    """mp2gpu_prove (csrc/prover.cpp: one native call per proof) against prover.py's sequence of ~40 calls, with the
    Z / partial-product values of the Python side coming from the by-definition restatement: same bytes."""
    import mapreduce_plonky2_b200 as G
    from mapreduce_plonky2_b200 import fri as GF
    from mapreduce_plonky2_b200 import prover as GP
    from mapreduce_plonky2_b200 import quotient as Q
    from mapreduce_plonky2_b200 import wire as W

    G.init(0)
    cfg = GF.FriConfig() if default_cfg else GF.FriConfig(proof_of_work_bits=10, num_query_rounds=6)
    inst = PR.synthetic_instance(seed, degree_bits=degree_bits, two_groups=two_groups,
                                 with_poseidon=with_poseidon in (True, "both"), extra_gates=with_poseidon in ("extra", "both"))
    desc = Q.CircuitDesc.from_circuit(inst.circuit)
    b_cs = G.PolynomialBatch.from_values(np.array(inst.constants + inst.sigmas, dtype=np.uint64), cfg.rate_bits, False,
                                         cfg.cap_height, hash_kind=kind, keep_on_device=True, fetch_leaves=False)
    digest = G.circuit_digest(b_cs.merkle_tree.cap.hashes, degree_bits, kind)
    wires = np.array(inst.wires, dtype=np.uint64)
    public_inputs = np.array([5, 6, 7], dtype=np.uint64)
    data = GP.prove_native(desc, b_cs, digest, wires, public_inputs, inst.public_inputs_hash, cfg, hash_kind=kind)
    if degree_bits <= 12:   # (the by-definition running products are pure-Python loops; 2^15 rows: two-pass transforms)
        ref = GP.prove(desc, b_cs, digest, wires, inst.public_inputs_hash,
                       lambda betas, gammas: np.array(PR.zs_partial_products(inst, betas, gammas), dtype=np.uint64), cfg, kind)
        assert data == W.write_proof_with_public_inputs(W.ProofWithPublicInputs(ref, public_inputs))
    # the Python mirror with Z / partial products from the device gives the same proof
    ref2 = GP.prove(desc, b_cs, digest, wires, inst.public_inputs_hash, None, cfg, kind)
    assert data == W.write_proof_with_public_inputs(W.ProofWithPublicInputs(ref2, public_inputs))
    # parsed back, the by-definition verifier accepts it; twice the same bytes (determinism pin, mp2-v1/src/api.rs:617-636)
    back = W.read_proof_with_public_inputs(data)
    assert np.array_equal(back.public_inputs, public_inputs)
    PR.verify_proof(inst.circuit, digest.tolist(), b_cs.merkle_tree.cap.hashes.tolist(), inst.public_inputs_hash,
                    _as_dict(back.proof), kind, pow_bits=cfg.proof_of_work_bits)
    assert GP.prove_native(desc, b_cs, digest, wires, public_inputs, inst.public_inputs_hash, cfg, hash_kind=kind) == data
    b_cs.free()


def test_native_prove_errors():
    import mapreduce_plonky2_b200 as G
    from mapreduce_plonky2_b200 import fri as GF
    from mapreduce_plonky2_b200 import prover as GP
    from mapreduce_plonky2_b200 import quotient as Q

    G.init(0)
    cfg = GF.FriConfig()
    inst = PR.synthetic_instance(3, degree_bits=5)
    desc = Q.CircuitDesc.from_circuit(inst.circuit)
    wires = np.array(inst.wires, dtype=np.uint64)
    cs_vals = np.array(inst.constants + inst.sigmas, dtype=np.uint64)
    b_cs = G.PolynomialBatch.from_values(cs_vals, 2, False, 4, hash_kind=1, keep_on_device=True, fetch_leaves=False)
    with pytest.raises(G.Mp2GpuError, match="another rate"):
        GP.prove_native(desc, b_cs, [1, 2, 3, 4], wires, [], inst.public_inputs_hash, cfg, hash_kind=1)
    b_cs.free()
    b_cs = G.PolynomialBatch.from_values(cs_vals, 3, False, 4, hash_kind=1, keep_on_device=True, fetch_leaves=False)
    with pytest.raises(G.Mp2GpuError, match="num_wires"):
        GP.prove_native(desc, b_cs, [1, 2, 3, 4], wires[:-1], [], inst.public_inputs_hash, cfg, hash_kind=1)
    desc.gates[0].kind = "lookup_table"
    with pytest.raises(G.Mp2GpuError, match="supported subset"):
        GP.prove_native(desc, b_cs, [1, 2, 3, 4], wires, [], inst.public_inputs_hash, cfg, hash_kind=1)
    b_cs.free()
