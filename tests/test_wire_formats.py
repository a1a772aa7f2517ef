"""bincode mirrors of FriProof / ProofWithPublicInputs / ProofWithVK (mapreduce_plonky2_b200/wire.py;
mp2-common/src/proof.rs:41-57): structural round trips, exact byte counts and malformed-input rejection.
Host logic only -- no GPU.  (Byte-level pinning against the Rust reference: tests/test_reference_goldens.py.)"""
import struct

import numpy as np
import pytest

from mapreduce_plonky2_b200 import wire as W
from mapreduce_plonky2_b200.fri import FriProof, FriQueryRound, FriQueryStep
from mapreduce_plonky2_b200.plonky2 import MerkleCap, MerkleProof, Mp2GpuError

P = 0xFFFFFFFF00000001


def _felts(rng, *shape):
    return rng.integers(0, P, size=shape, dtype=np.uint64)


def _fri_proof(rng, nlayers=2, nrounds=3, noracles=4, h=5):
    caps = [MerkleCap(_felts(rng, 16, 4)) for _ in range(nlayers)]
    rounds = []
    for _ in range(nrounds):
        init = [(_felts(rng, 3 + 5 * o), MerkleProof(_felts(rng, h, 4))) for o in range(noracles)]
        steps = [FriQueryStep(_felts(rng, 16, 2), MerkleProof(_felts(rng, h - 1 - i, 4))) for i in range(nlayers)]
        rounds.append(FriQueryRound(init, steps))
    return FriProof(caps, rounds, _felts(rng, 32, 2), int(_felts(rng, 1)[0]))


def _same_fri(a, b):
    assert len(a.commit_phase_merkle_caps) == len(b.commit_phase_merkle_caps)
    for x, y in zip(a.commit_phase_merkle_caps, b.commit_phase_merkle_caps):
        assert np.array_equal(x.hashes, y.hashes)
    assert len(a.query_round_proofs) == len(b.query_round_proofs)
    for qa, qb in zip(a.query_round_proofs, b.query_round_proofs):
        assert len(qa.initial_trees_proof) == len(qb.initial_trees_proof)
        for (ea, pa), (eb, pb) in zip(qa.initial_trees_proof, qb.initial_trees_proof):
            assert np.array_equal(ea, eb) and np.array_equal(pa.siblings, pb.siblings)
        assert len(qa.steps) == len(qb.steps)
        for sa, sb in zip(qa.steps, qb.steps):
            assert np.array_equal(sa.evals, sb.evals) and np.array_equal(sa.merkle_proof.siblings, sb.merkle_proof.siblings)
    assert np.array_equal(a.final_poly, b.final_poly) and a.pow_witness == b.pow_witness


def test_fri_proof_round_trip_and_size():
    rng = np.random.default_rng(1)
    p = _fri_proof(rng)
    data = W.write_fri_proof(p)
    # by hand: vec(2 caps: len + 16*32) | vec(3 rounds: vec(4 x (vec evals, vec siblings)) + vec(2 steps)) | poly | witness
    per_round = 8 + sum(8 + 8 * (3 + 5 * o) + 8 + 32 * 5 for o in range(4)) + 8 + sum(8 + 16 * 16 + 8 + 32 * (4 - i) for i in range(2))
    assert len(data) == 8 + 2 * (8 + 16 * 32) + 8 + 3 * per_round + 8 + 32 * 16 + 8
    _same_fri(p, W.read_fri_proof(data))
    assert W.write_fri_proof(W.read_fri_proof(data)) == data
    # the first field is the number of commit-phase caps, little endian u64; the last one the PoW witness
    assert struct.unpack("<Q", data[:8])[0] == 2
    assert struct.unpack("<Q", data[-8:])[0] == p.pow_witness


def test_proof_with_vk_round_trip():
    rng = np.random.default_rng(2)
    op = W.OpeningSet(constants=_felts(rng, 4, 2), plonk_sigmas=_felts(rng, 135, 2), wires=_felts(rng, 135, 2),
                      plonk_zs=_felts(rng, 2, 2), plonk_zs_next=_felts(rng, 2, 2), partial_products=_felts(rng, 18, 2),
                      quotient_polys=_felts(rng, 16, 2))
    proof = W.Proof(MerkleCap(_felts(rng, 16, 4)), MerkleCap(_felts(rng, 16, 4)), MerkleCap(_felts(rng, 16, 4)), op,
                    _fri_proof(rng, nlayers=3, nrounds=28))
    pwv = W.ProofWithVK(W.ProofWithPublicInputs(proof, _felts(rng, 9)),
                        W.VerifierOnlyCircuitData(MerkleCap(_felts(rng, 16, 4)), _felts(rng, 4)))
    data = pwv.serialize()
    back = W.ProofWithVK.deserialize(data)
    assert back.serialize() == data
    assert np.array_equal(back.proof.public_inputs, pwv.proof.public_inputs)
    for name in W.OPENING_FIELDS:
        assert np.array_equal(getattr(back.proof.proof.openings, name), getattr(op, name))
    assert back.proof.proof.openings.lookup_zs.shape == (0, 2)
    _same_fri(back.proof.proof.opening_proof, proof.opening_proof)
    assert np.array_equal(back.vk.circuit_digest, pwv.vk.circuit_digest)
    # the vk travels as serialize_bytes(to_bytes()): u64 length, then cap length + 16 hashes + the digest
    vk_bytes = pwv.vk.to_bytes()
    assert len(vk_bytes) == 8 + 16 * 32 + 32 and data.endswith(struct.pack("<Q", len(vk_bytes)) + vk_bytes)
    # ProofWithPublicInputs alone (serialize_proof) is a prefix of it
    assert data.startswith(W.write_proof_with_public_inputs(pwv.proof))


def test_non_canonical_elements_are_written_canonical_and_rejected_on_read():
    cap = MerkleCap(np.array([[P, P + 1, 2**64 - 1, 5]], dtype=np.uint64))
    vk = W.VerifierOnlyCircuitData(cap, np.zeros(4, dtype=np.uint64))
    b = vk.to_bytes()
    assert struct.unpack("<4Q", b[8:40]) == (0, 1, 2**64 - 1 - P, 5)
    bad = bytearray(b)
    bad[8:16] = struct.pack("<Q", P)
    with pytest.raises(Mp2GpuError):
        W.VerifierOnlyCircuitData.from_bytes(bytes(bad))


@pytest.mark.parametrize("mutate", ["truncate", "trailing", "huge_length"])
def test_malformed_inputs_raise(mutate):
    rng = np.random.default_rng(3)
    data = W.write_fri_proof(_fri_proof(rng, nlayers=1, nrounds=1, noracles=1))
    if mutate == "truncate":
        data = data[:-3]
    elif mutate == "trailing":
        data = data + b"\x00"
    else:
        data = struct.pack("<Q", 1 << 60) + data[8:]
    with pytest.raises(Mp2GpuError):
        W.read_fri_proof(data)
