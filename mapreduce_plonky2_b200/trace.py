"""Commitment traces of whole proofs (BASELINE configs 2, 4, 5).

The Rust prover cannot be built in this image, so a proof is *replayed* as the sequence of
polynomial-batch commitments and FRI-layer trees its ``prove()`` calls make, on synthetic data of the
right shapes.  Every ``prove()`` of degree ``n`` under ``standard_recursion_config``
(mp2-common/src/lib.rs:45-47: 135 wires, 2 challenges, quotient degree factor 8, rate_bits 3,
cap_height 4, ConstantArityBits(4, 5)) commits:

    from_values(135 x n)   wires
    from_values( 20 x n)   Z / partial products  (2 * (1 + 9))
    from_coeffs( 16 x n)   quotient chunks       (2 * 8)
    MerkleTree::new per FRI reduction layer: (8n / 16) leaves of 16*D = 32 elements, then /16 per layer

A proof of one mp2 circuit is a base ``prove()`` followed by the wrap chain down to 2^12 rows
(recursion-framework/src/universal_verifier_gadget/wrap_circuit.rs:64-115, RECURSION_THRESHOLD = 12 at
recursion-framework/src/universal_verifier_gadget/mod.rs:34).  The degrees below are ASSUMED (base 2^14,
wraps 2^13 and 2^12) until a Rust toolchain can report the real ones; a trace's proofs/s is an UPPER
BOUND on prover throughput (witness generation, quotient evaluation and openings are outside the path).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

RATE_BITS, CAP_HEIGHT, NUM_WIRES, ZS_PP_COLS, QUOTIENT_COLS, FRI_LEAF_LEN = 3, 4, 135, 20, 16, 32
ARITY_BITS, FINAL_POLY_BITS = 4, 5

TRACE_INCLUDES = "per prove(): from_values(135 cols), from_values(20), from_coeffs(16), one Merkle tree per FRI layer"

LEAF_PROOF_DEGREES = (14, 13, 12)          # assumed: base circuit + two wrap steps (BASELINE config 2)
AGGREGATION_DEGREES = (13, 12)             # assumed: 2-proof branch circuit + one wrap (BASELINE config 4)


@dataclass(frozen=True)
class Op:
    kind: str       # "from_values" | "from_coeffs" | "merkle"
    ncols: int      # columns, or leaf length for "merkle"
    n_log: int      # log2 rows, or log2 leaves for "merkle"

    @property
    def lde_elems(self) -> int:
        return self.ncols << (self.n_log + RATE_BITS) if self.kind != "merkle" else 0

    @property
    def perms(self) -> int:
        if self.kind == "merkle":
            nl = 1 << self.n_log
            return nl * ((self.ncols + 7) // 8) + nl - (1 << min(CAP_HEIGHT, self.n_log))
        N = 1 << (self.n_log + RATE_BITS)
        return N * ((self.ncols + 7) // 8) + N - (1 << CAP_HEIGHT)


def fri_reduction_arity_bits(degree_bits: int) -> List[int]:
    """plonky2 FriReductionStrategy::ConstantArityBits(4, 5).reduction_arity_bits(...)."""
    out, db = [], degree_bits
    while db > FINAL_POLY_BITS and db + RATE_BITS - CAP_HEIGHT > ARITY_BITS:
        out.append(ARITY_BITS)
        db -= ARITY_BITS
    return out


def prove_ops(degree_bits: int) -> List[Op]:
    ops = [Op("from_values", NUM_WIRES, degree_bits), Op("from_values", ZS_PP_COLS, degree_bits),
           Op("from_coeffs", QUOTIENT_COLS, degree_bits)]
    lde_bits = degree_bits + RATE_BITS
    for ab in fri_reduction_arity_bits(degree_bits):
        lde_bits -= ab
        ops.append(Op("merkle", FRI_LEAF_LEN, lde_bits))
    return ops


def proof_ops(degrees=LEAF_PROOF_DEGREES) -> List[Op]:
    return [op for d in degrees for op in prove_ops(d)]


class TraceRunner:
    """Replays proof traces on the current CUDA device, ``nstreams`` proofs in flight."""

    def __init__(self, degrees=LEAF_PROOF_DEGREES, hash_kind: int = 1, nstreams: int = 8):
        import torch

        from . import device as D

        self.torch, self.D = torch, D
        D.bind_current_device()
        self.ops = proof_ops(degrees)
        self.hash_kind = hash_kind
        self.streams = [torch.cuda.Stream() for _ in range(nstreams)]
        self.slots = [self._alloc() for _ in range(nstreams)]

    def _alloc(self):
        torch, D = self.torch, self.D
        bufs = []
        for op in self.ops:
            if op.kind == "merkle":
                nl = 1 << op.n_log
                cap_h = min(CAP_HEIGHT, op.n_log)
                leaves = torch.randint(0, 1 << 62, (nl, op.ncols), dtype=torch.int64, device="cuda")
                dig = torch.empty((max(2 * (nl - (1 << cap_h)), 1), 4), dtype=torch.int64, device="cuda")
                cap = torch.empty((1 << cap_h, 4), dtype=torch.int64, device="cuda")
                bufs.append((leaves, dig, cap, cap_h))
            else:
                cols = torch.randint(0, 1 << 62, (op.ncols, 1 << op.n_log), dtype=torch.int64, device="cuda")
                bufs.append((cols, D.CommitBuffers(op.ncols, op.n_log, RATE_BITS, CAP_HEIGHT, True)))
        return bufs

    def _one(self, slot):
        D = self.D
        for op, b in zip(self.ops, slot):
            if op.kind == "merkle":
                D.merkle_rowmajor(b[0], b[3], self.hash_kind, b[1], b[2])
            else:
                D.commit_resident(b[0], b[1], self.hash_kind, op.kind == "from_coeffs")

    def run(self, nproofs: int) -> None:
        """Launches ``nproofs`` traces round-robin over the streams (asynchronous)."""
        torch = self.torch
        for i in range(nproofs):
            with torch.cuda.stream(self.streams[i % len(self.streams)]):
                self._one(self.slots[i % len(self.slots)])

    def sync(self) -> None:
        for s in self.streams:
            s.synchronize()

    @property
    def perms_per_proof(self) -> int:
        return sum(op.perms for op in self.ops)

    @property
    def lde_elems_per_proof(self) -> int:
        return sum(op.lde_elems for op in self.ops)
