"""Byte-level wire formats of the proof objects the hot path feeds.

The reference moves proofs between tasks as ``bincode::serialize(&ProofWithVK)``
(mp2-common/src/proof.rs:41-57: ``ProofWithVK { proof: ProofWithPublicInputs<F, C, D>, vk }`` with
``#[serde(serialize_with = "serialize")]`` on ``vk``, i.e. ``serialize_bytes(vk.to_bytes())`` --
mp2-common/src/serialization/mod.rs:47-50) and pins determinism on those bytes (mp2-v1/src/api.rs:617-636).
This module restates that encoding so that a proof assembled from the GPU path can be written to bytes, parsed
back, and -- once a Rust toolchain is available -- diffed against the CPU prover's bytes:

* bincode 1.x default options: little-endian, fixed-width integers, ``Vec<T>`` = u64 length + elements,
  structs / tuples / fixed arrays = their fields in order with no framing, newtype structs transparent.
* ``GoldilocksField`` serialises as its canonical ``u64``; ``QuadraticExtension([F; 2])`` as two of them;
  ``HashOut { elements: [F; 4] }`` as four; ``MerkleCap(Vec<Hash>)``, ``MerkleProof { siblings: Vec<Hash> }``,
  ``PolynomialCoeffs { coeffs: Vec<T> }`` as length-prefixed vectors.
* plonky2 0.2.2 ``Proof`` field order: wires_cap, plonk_zs_partial_products_cap, quotient_polys_cap, openings,
  opening_proof;  ``OpeningSet``: constants, plonk_sigmas, wires, plonk_zs, plonk_zs_next, partial_products,
  quotient_polys, lookup_zs, lookup_zs_next;  ``FriProof``: commit_phase_merkle_caps, query_round_proofs,
  final_poly, pow_witness;  ``FriQueryRound``: initial_trees_proof { evals_proofs: Vec<(Vec<F>, MerkleProof)> },
  steps: Vec<FriQueryStep { evals: Vec<Ext>, merkle_proof }>.
* ``VerifierOnlyCircuitData::to_bytes`` (plonky2 ``util/serialization``): ``write_merkle_cap`` (usize length as
  u64 LE, then the hashes) followed by ``write_hash(circuit_digest)``.

STATUS: layouts restated from plonky2 0.2.2 / serde / bincode conventions, **not yet confirmed against bytes
produced by the Rust reference** (no toolchain in this image); ``tools/golden_dump`` holds the Rust program that
emits the pinning vectors and ``tests/test_reference_goldens.py`` picks them up when present.  Pure host code
(struct packing); no field arithmetic happens here.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np

from .fri import FriProof, FriQueryRound, FriQueryStep
from .plonky2 import MerkleCap, MerkleProof, Mp2GpuError

ORDER = 0xFFFFFFFF00000001


# ---- primitive writers / readers ------------------------------------------------------------------
class _W:
    def __init__(self):
        self.parts: List[bytes] = []

    def u64(self, v: int) -> None:
        self.parts.append(struct.pack("<Q", int(v)))

    def felts(self, a) -> None:
        """field elements, canonical, no length prefix"""
        x = np.ascontiguousarray(np.asarray(a, dtype=np.uint64)).reshape(-1)
        if x.size and int(x.max()) >= ORDER:
            x = np.where(x >= np.uint64(ORDER), x - np.uint64(ORDER), x)
        self.parts.append(x.astype("<u8").tobytes())

    def vec_felts(self, a) -> None:
        x = np.asarray(a, dtype=np.uint64)
        self.u64(x.shape[0] if x.ndim else 0)
        self.felts(x)

    def bytes_(self, b: bytes) -> None:
        self.u64(len(b))
        self.parts.append(bytes(b))

    def done(self) -> bytes:
        return b"".join(self.parts)


class _R:
    def __init__(self, data: bytes):
        self.d = memoryview(bytes(data))
        self.at = 0

    def take(self, n: int) -> memoryview:
        if n < 0 or self.at + n > len(self.d):
            raise Mp2GpuError("wire format: input ends inside a field (need %d bytes at offset %d of %d)" % (n, self.at, len(self.d)))
        out = self.d[self.at:self.at + n]
        self.at += n
        return out

    def u64(self) -> int:
        return struct.unpack("<Q", self.take(8))[0]

    def length(self, elem_bytes: int) -> int:
        n = self.u64()
        if n * max(elem_bytes, 1) > len(self.d) - self.at:
            raise Mp2GpuError("wire format: length prefix %d exceeds the remaining input" % n)
        return n

    def felts(self, count: int, width: int = 1) -> np.ndarray:
        a = np.frombuffer(self.take(8 * count * width), dtype="<u8").astype(np.uint64)
        if a.size and int(a.max()) >= ORDER:
            raise Mp2GpuError("wire format: non-canonical field element")
        return a.reshape(count, width) if width > 1 else a

    def vec_felts(self, width: int = 1) -> np.ndarray:
        return self.felts(self.length(8 * width), width)

    def bytes_(self) -> bytes:
        return bytes(self.take(self.length(1)))

    def finish(self) -> None:
        if self.at != len(self.d):
            raise Mp2GpuError("wire format: %d trailing bytes" % (len(self.d) - self.at))


def _w_cap(w: _W, cap: MerkleCap) -> None:
    h = np.asarray(cap.hashes, dtype=np.uint64).reshape(-1, 4)
    w.u64(h.shape[0])
    w.felts(h)


def _r_cap(r: _R) -> MerkleCap:
    return MerkleCap(r.vec_felts(4))


def _w_merkle_proof(w: _W, p: MerkleProof) -> None:
    s = np.asarray(p.siblings, dtype=np.uint64).reshape(-1, 4)
    w.u64(s.shape[0])
    w.felts(s)


def _r_merkle_proof(r: _R) -> MerkleProof:
    return MerkleProof(r.vec_felts(4))


# ---- FriProof ---------------------------------------------------------------------------------------
def _w_fri_proof(w: _W, p: FriProof) -> None:
    w.u64(len(p.commit_phase_merkle_caps))
    for cap in p.commit_phase_merkle_caps:
        _w_cap(w, cap)
    w.u64(len(p.query_round_proofs))
    for q in p.query_round_proofs:
        w.u64(len(q.initial_trees_proof))            # FriInitialTreeProof { evals_proofs }
        for evals, proof in q.initial_trees_proof:
            w.vec_felts(np.asarray(evals, dtype=np.uint64).reshape(-1))
            _w_merkle_proof(w, proof)
        w.u64(len(q.steps))
        for st in q.steps:
            e = np.asarray(st.evals, dtype=np.uint64).reshape(-1, 2)
            w.u64(e.shape[0])
            w.felts(e)
            _w_merkle_proof(w, st.merkle_proof)
    fp = np.asarray(p.final_poly, dtype=np.uint64).reshape(-1, 2)   # PolynomialCoeffs<F::Extension>
    w.u64(fp.shape[0])
    w.felts(fp)
    w.u64(int(p.pow_witness) % ORDER)


def _r_fri_proof(r: _R) -> FriProof:
    caps = [_r_cap(r) for _ in range(r.length(8))]
    rounds = []
    for _ in range(r.length(16)):
        init: List[Tuple[np.ndarray, MerkleProof]] = []
        for _ in range(r.length(16)):
            evals = r.vec_felts()
            init.append((evals, _r_merkle_proof(r)))
        steps = []
        for _ in range(r.length(16)):
            evals = r.vec_felts(2)
            steps.append(FriQueryStep(evals, _r_merkle_proof(r)))
        rounds.append(FriQueryRound(init, steps))
    final_poly = r.vec_felts(2)
    pow_witness = int(r.felts(1)[0])
    return FriProof(caps, rounds, final_poly, pow_witness)


def write_fri_proof(p: FriProof) -> bytes:
    """``bincode::serialize(&FriProof<F, C::Hasher, D>)``."""
    w = _W()
    _w_fri_proof(w, p)
    return w.done()


def read_fri_proof(data: bytes) -> FriProof:
    r = _R(data)
    p = _r_fri_proof(r)
    r.finish()
    return p


# ---- Proof / ProofWithPublicInputs / ProofWithVK ---------------------------------------------------------
OPENING_FIELDS = ("constants", "plonk_sigmas", "wires", "plonk_zs", "plonk_zs_next", "partial_products",
                  "quotient_polys", "lookup_zs", "lookup_zs_next")


def _ext0() -> np.ndarray:
    return np.zeros((0, 2), dtype=np.uint64)


@dataclass
class OpeningSet:
    """``OpeningSet<F, D>``: every committed polynomial at zeta (and Z / lookup Z at g*zeta), as extension elements."""
    constants: np.ndarray = field(default_factory=_ext0)
    plonk_sigmas: np.ndarray = field(default_factory=_ext0)
    wires: np.ndarray = field(default_factory=_ext0)
    plonk_zs: np.ndarray = field(default_factory=_ext0)
    plonk_zs_next: np.ndarray = field(default_factory=_ext0)
    partial_products: np.ndarray = field(default_factory=_ext0)
    quotient_polys: np.ndarray = field(default_factory=_ext0)
    lookup_zs: np.ndarray = field(default_factory=_ext0)
    lookup_zs_next: np.ndarray = field(default_factory=_ext0)


@dataclass
class Proof:
    wires_cap: MerkleCap
    plonk_zs_partial_products_cap: MerkleCap
    quotient_polys_cap: MerkleCap
    openings: OpeningSet
    opening_proof: FriProof


@dataclass
class ProofWithPublicInputs:
    proof: Proof
    public_inputs: np.ndarray


@dataclass
class VerifierOnlyCircuitData:
    constants_sigmas_cap: MerkleCap
    circuit_digest: np.ndarray   # 4 elements

    def to_bytes(self) -> bytes:
        """plonky2 ``VerifierOnlyCircuitData::to_bytes``: write_merkle_cap + write_hash."""
        w = _W()
        _w_cap(w, self.constants_sigmas_cap)
        w.felts(np.asarray(self.circuit_digest, dtype=np.uint64).reshape(4))
        return w.done()

    @classmethod
    def from_bytes(cls, data: bytes) -> "VerifierOnlyCircuitData":
        r = _R(data)
        cap = _r_cap(r)
        digest = r.felts(4)
        r.finish()
        return cls(cap, digest)


@dataclass
class ProofWithVK:
    """mp2-common/src/proof.rs:41-46."""
    proof: ProofWithPublicInputs
    vk: VerifierOnlyCircuitData

    def serialize(self) -> bytes:
        return write_proof_with_vk(self)

    @classmethod
    def deserialize(cls, buff: bytes) -> "ProofWithVK":
        return read_proof_with_vk(buff)


def _w_proof_with_pis(w: _W, p: ProofWithPublicInputs) -> None:
    _w_cap(w, p.proof.wires_cap)
    _w_cap(w, p.proof.plonk_zs_partial_products_cap)
    _w_cap(w, p.proof.quotient_polys_cap)
    for name in OPENING_FIELDS:
        e = np.asarray(getattr(p.proof.openings, name), dtype=np.uint64).reshape(-1, 2)
        w.u64(e.shape[0])
        w.felts(e)
    _w_fri_proof(w, p.proof.opening_proof)
    w.vec_felts(np.asarray(p.public_inputs, dtype=np.uint64).reshape(-1))


def _r_proof_with_pis(r: _R) -> ProofWithPublicInputs:
    caps = [_r_cap(r) for _ in range(3)]
    openings = OpeningSet(**{name: r.vec_felts(2) for name in OPENING_FIELDS})
    fri = _r_fri_proof(r)
    pis = r.vec_felts()
    return ProofWithPublicInputs(Proof(caps[0], caps[1], caps[2], openings, fri), pis)


def write_proof_with_public_inputs(p: ProofWithPublicInputs) -> bytes:
    """``bincode::serialize(&ProofWithPublicInputs<F, C, D>)`` = mp2-common ``serialize_proof`` (proof.rs:86-90)."""
    w = _W()
    _w_proof_with_pis(w, p)
    return w.done()


def read_proof_with_public_inputs(data: bytes) -> ProofWithPublicInputs:
    r = _R(data)
    p = _r_proof_with_pis(r)
    r.finish()
    return p


def write_proof_with_vk(p: ProofWithVK) -> bytes:
    """``ProofWithVK::serialize`` (mp2-common/src/proof.rs:49-52)."""
    w = _W()
    _w_proof_with_pis(w, p.proof)
    w.bytes_(p.vk.to_bytes())          # serialize_bytes(vk.to_bytes())
    return w.done()


def read_proof_with_vk(data: bytes) -> ProofWithVK:
    r = _R(data)
    proof = _r_proof_with_pis(r)
    vk = VerifierOnlyCircuitData.from_bytes(r.bytes_())
    r.finish()
    return ProofWithVK(proof, vk)
