"""Host mirror of plonky2's ``prove`` from the point where the witness is known (plonk/prover.rs
``prove_with_partition_witness``; reached in the reference from ``circuit_data.prove(pw)`` at
recursion-framework/src/circuit_builder.rs:308 and .../universal_verifier_gadget/wrap_circuit.rs:143), every data-path
step on the device:

    wires commitment -> challenger(circuit digest, public-inputs hash, wires cap) -> betas, gammas
    -> Z / partial-products commitment (the VALUES are host work: the caller's callback) -> alphas
    -> quotient polynomials + their commitment (mp2gpu_quotient_polys) -> zeta
    -> OpeningSet at zeta and g*zeta (mp2gpu_batch_eval) -> prove_openings (FRI: mp2gpu_fri_*)

The rows of the four batches never leave HBM.  Gate coverage is the quotient kernel's (quotient.py); lookups and
zero-knowledge blinding are not supported (the reference enables neither).  ctypes + numpy only.
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import numpy as np

from . import fri as GF
from . import plonky2 as P2
from .quotient import CircuitDesc, compute_quotient_polys, partial_products_and_zs

P = P2.ORDER


class _CProveConfig(__import__("ctypes").Structure):
    import ctypes as _C
    _fields_ = [("rate_bits", _C.c_uint32), ("cap_height", _C.c_uint32), ("hash_kind", _C.c_uint32),
                ("proof_of_work_bits", _C.c_uint32), ("num_query_rounds", _C.c_uint32), ("num_reductions", _C.c_uint32),
                ("reduction_arity_bits", _C.POINTER(_C.c_uint32))]


# Proof / OpeningSet are the wire-format classes (wire.py): a proof produced here serialises as
# bincode(ProofWithVK) / bincode(ProofWithPublicInputs) without conversion (mp2-common/src/proof.rs:41-57).
from .wire import OpeningSet, Proof, ProofWithPublicInputs, ProofWithVK, VerifierOnlyCircuitData  # noqa: E402,F401


def primitive_root_of_unity(bits: int) -> int:
    return pow(pow(7, (P - 1) >> 32, P), 1 << (32 - bits), P)


def prove(circuit: CircuitDesc, constants_sigmas: P2.PolynomialBatch, circuit_digest: Sequence[int],
          wires_values, public_inputs_hash: Sequence[int],
          zs_partial_products: Callable[[List[int], List[int]], np.ndarray] = None,
          config: GF.FriConfig = None, hash_kind: int = P2.POSEIDON2) -> Proof:
    """``constants_sigmas``: the device-resident batch committed at build time.  ``wires_values``: (num_wires, n).
    ``zs_partial_products(betas, gammas)`` -> (num_challenges * (1 + num_partial_products), n) values, laid out
    [Z_0.., partial products of challenge 0, of challenge 1, ...] (``wires_permutation_partial_products_and_zs``);
    None (the default) computes them on the device from the two resident batches (``mp2gpu_partial_products_and_zs``)."""
    config = config or GF.FriConfig()
    nch, db = circuit.num_challenges, circuit.degree_bits
    commit = lambda cols: P2.PolynomialBatch.from_values(np.asarray(cols, dtype=np.uint64), config.rate_bits, False,
                                                         config.cap_height, hash_kind=hash_kind, keep_on_device=True,
                                                         fetch_leaves=False, fetch_coeffs=False, fetch_digests=False)
    wires = commit(wires_values)
    ch = GF.Challenger(hash_kind)
    ch.observe_hash(np.asarray(circuit_digest, dtype=np.uint64))
    ch.observe_hash(np.asarray(public_inputs_hash, dtype=np.uint64))
    ch.observe_cap(wires.merkle_tree.cap)
    betas, gammas = ch.get_n_challenges(nch), ch.get_n_challenges(nch)
    if zs_partial_products is None:
        zs_pp = partial_products_and_zs(circuit, constants_sigmas, wires, betas, gammas, config.rate_bits, config.cap_height,
                                        hash_kind=hash_kind, fetch_values=False)
    else:
        zs_pp = commit(zs_partial_products(betas, gammas))
    ch.observe_cap(zs_pp.merkle_tree.cap)
    alphas = ch.get_n_challenges(nch)
    quotient = compute_quotient_polys(circuit, constants_sigmas, wires, zs_pp, betas, gammas, alphas, public_inputs_hash,
                                      config.rate_bits, config.cap_height, hash_kind=hash_kind, fetch_digests=False,
                                      fetch_chunks=False)
    ch.observe_cap(quotient.merkle_tree.cap)
    zeta = ch.get_extension_challenge()
    g = primitive_root_of_unity(db)
    g_zeta = np.array([int(zeta[0]) * g % P, int(zeta[1]) * g % P], dtype=np.uint64)
    oracles = [constants_sigmas, wires, zs_pp, quotient]
    batches = [GF.FriBatchInfo(zeta, [(o, p) for o, b in enumerate(oracles) for p in range(b.num_polys)]),
               GF.FriBatchInfo(g_zeta, [(2, p) for p in range(nch)])]
    fri_openings = GF.open_batches(batches, oracles)
    for vals in fri_openings:          # challenger.observe_openings(&openings.to_fri_openings())
        ch.observe_extension_elements(vals)
    at_zeta, nc = fri_openings[0], circuit.num_constants
    w0 = constants_sigmas.num_polys
    z0 = w0 + wires.num_polys
    q0 = z0 + zs_pp.num_polys
    openings = OpeningSet(constants=at_zeta[:nc], plonk_sigmas=at_zeta[nc:w0], wires=at_zeta[w0:z0],
                          plonk_zs=at_zeta[z0:z0 + nch], plonk_zs_next=fri_openings[1],
                          partial_products=at_zeta[z0 + nch:q0], quotient_polys=at_zeta[q0:])
    try:
        opening_proof = GF.prove_openings(batches, oracles, ch, config.fri_params(db))
    finally:
        caps = wires.merkle_tree.cap, zs_pp.merkle_tree.cap, quotient.merkle_tree.cap
        for b in (wires, zs_pp, quotient):
            b.free()
    return Proof(caps[0], caps[1], caps[2], openings, opening_proof)


def prove_native(circuit: CircuitDesc, constants_sigmas: P2.PolynomialBatch, circuit_digest: Sequence[int], wires_values,
                 public_inputs: Sequence[int], public_inputs_hash: Sequence[int], config: GF.FriConfig = None,
                 hash_kind: int = P2.POSEIDON2) -> bytes:
    """The same ``prove()`` as ONE native call (``mp2gpu_prove``, csrc/prover.cpp): -> bincode(ProofWithPublicInputs)
    bytes (mp2-common/src/proof.rs:86-90), readable with ``wire.read_proof_with_public_inputs``.  The GIL is released
    for the whole proof, so prover threads run concurrently."""
    import ctypes as C

    from . import _lib
    from .quotient import _CCircuit  # noqa: F401

    config = config or GF.FriConfig()
    if constants_sigmas._handle is None:
        raise P2.Mp2GpuError("prove_native needs a device-resident constants_sigmas batch (keep_on_device=True)")
    w = np.ascontiguousarray(np.asarray(wires_values, dtype=np.uint64))
    if w.ndim != 2 or w.shape != (circuit.num_wires, 1 << circuit.degree_bits):
        raise P2.Mp2GpuError("wires_values must be (num_wires, 2^degree_bits)")
    cc, keep = circuit._c()
    arity = np.array(config.fri_params(circuit.degree_bits).reduction_arity_bits, dtype=np.uint32)
    cfg = _CProveConfig(config.rate_bits, config.cap_height, hash_kind, config.proof_of_work_bits, config.num_query_rounds,
                        arity.size, arity.ctypes.data_as(C.POINTER(C.c_uint32)))
    vec = lambda v: np.ascontiguousarray(np.array([int(x) for x in v], dtype=np.uint64))
    dg, pis, pih = vec(circuit_digest), vec(public_inputs), vec(public_inputs_hash)
    out, ln = C.c_void_p(None), C.c_size_t(0)
    _lib.call("mp2gpu_prove", C.byref(cc), C.byref(cfg), constants_sigmas._handle, P2._ptr(dg), P2._col_ptrs(w),
              P2._ptr(pis) if pis.size else None, pis.size, P2._ptr(pih), C.byref(out), C.byref(ln))
    del keep
    try:
        return C.string_at(out.value, ln.value)
    finally:
        _lib.load().mp2gpu_free_bytes(out)
