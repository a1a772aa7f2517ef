"""Host mirror of plonky2's quotient-polynomial step of ``prove()`` (plonk/prover.rs ``compute_quotient_polys`` +
``PolynomialBatch::from_coeffs(all_quotient_poly_chunks, ...)``) over the C ABI entry point
``mp2gpu_quotient_polys``: the three committed batches stay on the device, only the cap (and whatever else the
caller asks for) comes back.  ctypes + numpy only.

The descriptor mirrors the ``CommonCircuitData`` fields the vanishing polynomial depends on: ``gates`` in circuit
order, ``SelectorsInfo { selector_indices, groups }``, ``num_constants`` (selectors + gate constants), the wire counts
and ``quotient_degree_factor``.  Gate kinds outside the staged subset (ArithmeticGate, ConstantGate, PublicInputGate,
NoopGate, PoseidonGate, ArithmeticExtensionGate, MulExtensionGate, BaseSumGate<B>, ReducingGate,
ReducingExtensionGate, RandomAccessGate, ExponentiationGate, PoseidonMdsGate, CosetInterpolationGate of mp2-common/src/serialization/circuit_data_serialization.rs:234-266) raise, they are never skipped.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import Mp2GpuError
from .plonky2 import POSEIDON2, MerkleCap, MerkleTree, PolynomialBatch, _arr, _col_ptrs, _ptr

GATE_KINDS = {"noop": 0, "arithmetic": 1, "constant": 2, "public_input": 3, "poseidon": 4, "arithmetic_extension": 5,
              "mul_extension": 6, "base_sum": 7, "reducing": 8, "reducing_extension": 9, "random_access": 10,
              "exponentiation": 11, "poseidon_mds": 12, "coset_interpolation": 13}


class _CGate(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("num_ops", C.c_uint32), ("selector_index", C.c_uint32),
                ("group_begin", C.c_uint32), ("group_end", C.c_uint32), ("param", C.c_uint32)]


class _CCircuit(C.Structure):
    _fields_ = [("degree_bits", C.c_uint32), ("quotient_degree_bits", C.c_uint32), ("num_challenges", C.c_uint32),
                ("num_wires", C.c_uint32), ("num_routed_wires", C.c_uint32), ("num_constants", C.c_uint32),
                ("num_selectors", C.c_uint32), ("num_gates", C.c_uint32), ("gates", C.POINTER(_CGate))]


@dataclass
class GateDesc:
    kind: str           # one of GATE_KINDS
    num_ops: int = 0    # Arithmetic(Extension)Gate / MulExtensionGate::num_ops, ConstantGate::num_consts, BaseSumGate::num_limbs
    param: int = 0      # BaseSumGate<B>: B; RandomAccessGate: bits | num_extra_constants << 8


@dataclass
class CircuitDesc:
    degree_bits: int
    num_wires: int
    num_routed_wires: int
    num_constants: int
    gates: List[GateDesc]
    selector_indices: List[int]
    groups: List[Tuple[int, int]]
    quotient_degree_bits: int = 3
    num_challenges: int = 2

    @classmethod
    def from_circuit(cls, c) -> "CircuitDesc":
        """From any object with the same attribute names (e.g. tests/plonk_ref.Circuit)."""
        return cls(c.degree_bits, c.num_wires, c.num_routed_wires, c.num_constants,
                   [GateDesc(g.kind, g.num_ops, getattr(g, "param", 0)) for g in c.gates], list(c.selector_indices), list(c.groups),
                   c.quotient_degree_bits, c.num_challenges)

    @property
    def num_selectors(self) -> int:
        return len(self.groups)

    @property
    def num_partial_products(self) -> int:
        return -(-self.num_routed_wires // (1 << self.quotient_degree_bits)) - 1

    def _c(self):
        arr = (_CGate * max(1, len(self.gates)))()
        for i, g in enumerate(self.gates):
            if g.kind not in GATE_KINDS:
                raise Mp2GpuError("gate kind %r is outside the supported subset %s" % (g.kind, sorted(GATE_KINDS)))
            a, b = self.groups[self.selector_indices[i]]
            arr[i] = _CGate(GATE_KINDS[g.kind], g.num_ops, self.selector_indices[i], a, b, g.param)
        cc = _CCircuit(self.degree_bits, self.quotient_degree_bits, self.num_challenges, self.num_wires,
                       self.num_routed_wires, self.num_constants, self.num_selectors, len(self.gates), arr)
        return cc, arr  # keep `arr` alive with the struct


def compute_quotient_polys(desc: CircuitDesc, constants_sigmas: PolynomialBatch, wires: PolynomialBatch,
                           zs_partial_products: PolynomialBatch, betas: Sequence[int], gammas: Sequence[int],
                           alphas: Sequence[int], public_inputs_hash: Sequence[int], rate_bits: int, cap_height: int,
                           hash_kind: int = POSEIDON2, keep_on_device: bool = True, fetch_leaves: bool = False,
                           fetch_digests: bool = True, fetch_chunks: bool = True) -> PolynomialBatch:
    """-> the quotient ``PolynomialBatch`` (``polynomials`` = the num_challenges * quotient_degree_factor chunks).
    The three inputs must be device-resident (``keep_on_device=True`` when they were committed)."""
    for b in (constants_sigmas, wires, zs_partial_products):
        if b._handle is None:
            raise Mp2GpuError("compute_quotient_polys needs device-resident batches (commit with keep_on_device=True)")
    nch, md, n = desc.num_challenges, 1 << desc.quotient_degree_bits, 1 << desc.degree_bits
    if not (len(betas) == len(gammas) == len(alphas) == nch):
        raise Mp2GpuError("betas / gammas / alphas must have num_challenges entries")
    cc, keep = desc._c()
    vec = lambda v: np.ascontiguousarray(np.array([int(x) for x in v], dtype=np.uint64))
    b_, g_, a_, pi = vec(betas), vec(gammas), vec(alphas), vec(public_inputs_hash)
    N, ncap, ncols = n << rate_bits, 1 << cap_height, nch * md
    chunks = np.empty((ncols, n), dtype=np.uint64) if fetch_chunks else None
    leaves = np.empty((N, ncols), dtype=np.uint64) if fetch_leaves else None
    digests = np.empty((max(2 * (N - ncap), 0), 4), dtype=np.uint64) if fetch_digests else None
    cap = np.empty((ncap, 4), dtype=np.uint64)
    handle = C.c_void_p(None)
    _lib.call("mp2gpu_quotient_polys", C.byref(cc), constants_sigmas._handle, wires._handle, zs_partial_products._handle,
              _ptr(b_), _ptr(g_), _ptr(a_), _ptr(pi) if pi.size else None, rate_bits, cap_height, hash_kind,
              _col_ptrs(chunks) if chunks is not None else None, _ptr(leaves), _ptr(digests) if digests is not None and digests.size else None, _ptr(cap),
              C.byref(handle) if keep_on_device else None)
    del keep
    tree = MerkleTree(leaves, digests, MerkleCap(cap), hash_kind)
    return PolynomialBatch(chunks, tree, desc.degree_bits, rate_bits, False, handle if keep_on_device else None, ncols)


def partial_products_and_zs(desc: CircuitDesc, constants_sigmas: PolynomialBatch, wires: PolynomialBatch,
                            betas: Sequence[int], gammas: Sequence[int], rate_bits: int, cap_height: int,
                            hash_kind: int = POSEIDON2, keep_on_device: bool = True, fetch_values: bool = True,
                            fetch_leaves: bool = False, fetch_digests: bool = False) -> PolynomialBatch:
    """plonky2 ``all_wires_permutation_partial_products`` + the second commitment of ``prove()`` on the device
    (``mp2gpu_partial_products_and_zs``).  -> the ``zs_partial_products`` batch; with ``fetch_values`` its
    ``polynomials`` are the VALUES on the subgroup, columns [Z_0.., pp(ch 0).., pp(ch 1)..]."""
    for b in (constants_sigmas, wires):
        if b._handle is None:
            raise Mp2GpuError("partial_products_and_zs needs device-resident batches (commit with keep_on_device=True)")
    nch, n = desc.num_challenges, 1 << desc.degree_bits
    if not (len(betas) == len(gammas) == nch):
        raise Mp2GpuError("betas / gammas must have num_challenges entries")
    cc, keep = desc._c()
    vec = lambda v: np.ascontiguousarray(np.array([int(x) for x in v], dtype=np.uint64))
    b_, g_ = vec(betas), vec(gammas)
    N, ncap, ncols = n << rate_bits, 1 << cap_height, nch * (1 + desc.num_partial_products)
    values = np.empty((ncols, n), dtype=np.uint64) if fetch_values else None
    leaves = np.empty((N, ncols), dtype=np.uint64) if fetch_leaves else None
    digests = np.empty((max(2 * (N - ncap), 0), 4), dtype=np.uint64) if fetch_digests else None
    cap = np.empty((ncap, 4), dtype=np.uint64)
    handle = C.c_void_p(None)
    _lib.call("mp2gpu_partial_products_and_zs", C.byref(cc), constants_sigmas._handle, wires._handle, _ptr(b_), _ptr(g_),
              rate_bits, cap_height, hash_kind, _col_ptrs(values) if values is not None else None, _ptr(leaves),
              _ptr(digests) if digests is not None and digests.size else None, _ptr(cap),
              C.byref(handle) if keep_on_device else None)
    del keep
    tree = MerkleTree(leaves, digests, MerkleCap(cap), hash_kind)
    return PolynomialBatch(values, tree, desc.degree_bits, rate_bits, False, handle if keep_on_device else None, ncols)
