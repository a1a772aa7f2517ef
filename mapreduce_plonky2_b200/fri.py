"""Host-side mirror of plonky2's FRI prover glue over the device primitives (plonky2 0.2.2).

What runs where:

* ``Challenger`` (plonky2 ``iop/challenger.rs``): the duplex sponge of the Fiat-Shamir transcript.  A few dozen
  strictly sequential permutations per proof over host data; they run on the host (``mp2gpu_transcript_permute``,
  as the reference's challenger does -- through the device each cost a ~60 us round trip, 6.7 ms per proof).
  ``Challenger(on_device=True)`` keeps the round-1 behaviour (``mp2gpu_permute_batch``); in the Rust integration the
  challenger stays plonky2's own host code, INTEGRATION.md section 4b.
* ``open_batches``: ``OpeningSet::new``'s polynomial evaluations, on the coefficients resident in HBM.
* ``prove_openings`` / ``fri_proof`` (``fri/oracle.rs``, ``fri/prover.rs``): alpha-batched quotient, commit phase,
  proof-of-work grind, query rounds -- leaves, digests and layer trees never leave the device; only caps, the
  final polynomial and the opened rows / Merkle paths come back.

Reached in the reference from every ``circuit_data.prove(pw)`` (recursion-framework/src/circuit_builder.rs:308);
the proof it assembles is what the universal verifier consumes at
recursion-framework/src/universal_verifier_gadget/verifier_gadget.rs:116-118.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import numpy as np

from . import plonky2 as P2
from .plonky2 import FriCommitPhase, MerkleCap, MerkleProof, Mp2GpuError, PolynomialBatch

P = 0xFFFFFFFF00000001
SPONGE_WIDTH, SPONGE_RATE = 12, 8


class Challenger:
    """``Challenger<F, H>``: overwrite-mode duplex sponge; challenges are popped from the END of the squeezed rate."""

    def __init__(self, hash_kind: int = P2.POSEIDON2, on_device: bool = False):
        self.hash_kind = hash_kind
        self.on_device = on_device
        self.sponge_state = np.zeros(SPONGE_WIDTH, dtype=np.uint64)
        self.input_buffer: List[int] = []
        self.output_buffer: List[int] = []

    def observe_element(self, x) -> None:
        self.output_buffer.clear()            # any buffered outputs are now invalid
        self.input_buffer.append(int(x) % P)
        if len(self.input_buffer) == SPONGE_RATE:
            self.duplexing()

    def observe_elements(self, xs) -> None:
        flat = np.ascontiguousarray(np.asarray(xs, dtype=np.uint64).reshape(-1))
        if self.on_device or flat.size < 4:
            for x in flat.tolist():
                self.observe_element(x)
            return
        # the same element-by-element semantics in one host call (mp2gpu_transcript_observe)
        import ctypes as C

        from . import _lib
        state = np.ascontiguousarray(self.sponge_state, dtype=np.uint64).copy()
        buf = np.zeros(SPONGE_RATE, dtype=np.uint64)
        buf[:len(self.input_buffer)] = self.input_buffer
        blen, last = C.c_uint32(len(self.input_buffer)), C.c_uint32(0)
        _lib.call("mp2gpu_transcript_observe", P2._ptr(state), P2._ptr(buf), C.byref(blen), P2._ptr(flat), flat.size,
                  self.hash_kind, C.byref(last))
        self.sponge_state = state
        self.input_buffer = [int(v) for v in buf[:blen.value]]
        self.output_buffer = [int(v) for v in state[:SPONGE_RATE]] if last.value else []

    def observe_extension_element(self, e) -> None:
        self.observe_elements(e)              # to_basefield_array()

    def observe_extension_elements(self, es) -> None:
        self.observe_elements(es)

    def observe_hash(self, h) -> None:
        self.observe_elements(h)

    def observe_cap(self, cap: MerkleCap) -> None:
        self.observe_elements(np.asarray(cap.hashes, dtype=np.uint64))   # = observe_hash of every cap entry, in order

    def get_challenge(self) -> int:
        if self.input_buffer or not self.output_buffer:
            self.duplexing()
        return self.output_buffer.pop()

    def get_n_challenges(self, n: int) -> List[int]:
        return [self.get_challenge() for _ in range(n)]

    def get_extension_challenge(self) -> np.ndarray:
        return np.array(self.get_n_challenges(2), dtype=np.uint64)

    def duplexing(self) -> None:
        assert len(self.input_buffer) <= SPONGE_RATE
        for i, x in enumerate(self.input_buffer):
            self.sponge_state[i] = x
        self.input_buffer.clear()
        if self.on_device:
            self.sponge_state = P2.permute(self.sponge_state.reshape(1, SPONGE_WIDTH), self.hash_kind).reshape(SPONGE_WIDTH)
        else:
            self.sponge_state = P2.transcript_permute(self.sponge_state, self.hash_kind)
        self.output_buffer = [int(v) for v in self.sponge_state[:SPONGE_RATE]]


@dataclass
class FriConfig:
    """``FriConfig`` of ``standard_recursion_config`` (mp2-common/src/lib.rs:45-47) by default."""
    rate_bits: int = 3
    cap_height: int = 4
    proof_of_work_bits: int = 16
    num_query_rounds: int = 28
    arity_bits: int = 4            # FriReductionStrategy::ConstantArityBits(arity_bits, final_poly_bits)
    final_poly_bits: int = 5

    def fri_params(self, degree_bits: int) -> "FriParams":
        out, db = [], degree_bits
        while db > self.final_poly_bits and db + self.rate_bits - self.cap_height > self.arity_bits:
            out.append(self.arity_bits)
            db -= self.arity_bits
        return FriParams(self, degree_bits, out)


@dataclass
class FriParams:
    config: FriConfig
    degree_bits: int
    reduction_arity_bits: List[int]

    @property
    def lde_bits(self) -> int:
        return self.degree_bits + self.config.rate_bits


@dataclass
class FriBatchInfo:
    """``FriBatchInfo { point, polynomials: Vec<FriPolynomialInfo { oracle_index, polynomial_index }> }``."""
    point: np.ndarray
    polynomials: List[Tuple[int, int]]


@dataclass
class FriQueryStep:
    evals: np.ndarray              # (arity, 2)
    merkle_proof: MerkleProof


@dataclass
class FriQueryRound:
    initial_trees_proof: List[Tuple[np.ndarray, MerkleProof]]   # per oracle: (row, proof)
    steps: List[FriQueryStep]


@dataclass
class FriProof:
    commit_phase_merkle_caps: List[MerkleCap]
    query_round_proofs: List[FriQueryRound]
    final_poly: np.ndarray         # (len, 2)
    pow_witness: int
    fri_openings: List[np.ndarray] = field(default_factory=list)   # per batch (count, 2); carried for convenience


def open_batches(batches: Sequence[FriBatchInfo], oracles: Sequence[PolynomialBatch]) -> List[np.ndarray]:
    """``OpeningSet::new(...).to_fri_openings()``: per batch, its polynomials evaluated at its point."""
    points = np.stack([np.asarray(b.point, dtype=np.uint64).reshape(2) for b in batches])
    per_oracle = [o.eval(points) for o in oracles]      # (npoints, ncols, 2) each: one pass over each oracle
    return [np.stack([per_oracle[oi][i, pi] for oi, pi in b.polynomials]) for i, b in enumerate(batches)]


def fri_proof_of_work(challenger: Challenger, config: FriConfig) -> int:
    """plonky2 ``fri_proof_of_work``; the grind runs on the device and returns the smallest witness."""
    min_leading_zeros = config.proof_of_work_bits + (64 - 64)     # F::order().bits() == 64
    state = challenger.sponge_state.copy()
    pos = len(challenger.input_buffer)
    for i, x in enumerate(challenger.input_buffer):
        state[i] = x
    witness = P2.fri_proof_of_work(state, pos, min_leading_zeros, challenger.hash_kind)
    challenger.observe_element(witness)
    response = challenger.get_challenge()
    if 64 - int(response).bit_length() < min_leading_zeros:
        raise Mp2GpuError("proof of work response does not have the required leading zeros")
    return witness


def fri_proof(oracles: Sequence[PolynomialBatch], phase: FriCommitPhase, challenger: Challenger,
              params: FriParams) -> FriProof:
    """``fri_proof``: commit phase (``fri_committed_trees``), PoW, query rounds (``fri_prover_query_rounds``)."""
    caps = []
    for arity_bits in params.reduction_arity_bits:
        cap = phase.commit_layer(arity_bits)
        challenger.observe_cap(cap)
        caps.append(cap)
        phase.fold(challenger.get_extension_challenge())
    final_poly = phase.finish()
    challenger.observe_extension_elements(final_poly)
    pow_witness = fri_proof_of_work(challenger, params.config)
    n = 1 << params.lde_bits
    x_indices = [c % n for c in challenger.get_n_challenges(params.config.num_query_rounds)]
    # one gather per tree for all rounds
    initial = [o.open(x_indices) for o in oracles]
    layer_opens, idx = [], list(x_indices)
    for i, arity_bits in enumerate(params.reduction_arity_bits):
        idx = [x >> arity_bits for x in idx]
        layer_opens.append(phase.open_layer(i, idx))
    rounds = []
    for q in range(len(x_indices)):
        init = [(rows[q], MerkleProof(sib[q])) for rows, sib in initial]
        steps = [FriQueryStep(leaves[q].reshape(-1, 2), MerkleProof(sib[q])) for leaves, sib in layer_opens]
        rounds.append(FriQueryRound(init, steps))
    return FriProof(caps, rounds, final_poly, pow_witness)


def prove_openings(batches: Sequence[FriBatchInfo], oracles: Sequence[PolynomialBatch], challenger: Challenger,
                   params: FriParams) -> FriProof:
    """``PolynomialBatch::prove_openings(instance, oracles, challenger, fri_params, timing)``."""
    alpha = challenger.get_extension_challenge()
    phase = FriCommitPhase.from_openings(oracles, [(b.point, b.polynomials) for b in batches], alpha,
                                         params.config.cap_height, challenger.hash_kind)
    try:
        return fri_proof(oracles, phase, challenger, params)
    finally:
        phase.free()
