"""Host-side mirror of the plonky2 surface the reference reaches the hot path through.

Names, argument meaning and error behaviour follow plonky2 0.2.2 (the reference's pinned crate,
Cargo.toml:63,114-117) so that the parity tests read like the reference's own:

* ``PolynomialBatch.from_values / from_coeffs / get_lde_values``  (plonky2 ``fri/oracle.rs``; reached via
  ``circuit_data.prove`` at recursion-framework/src/circuit_builder.rs:308 and ``builder.build`` at :177)
* ``MerkleTree.new / prove / get``, ``MerkleCap``, ``MerkleProof`` (plonky2 ``hash/merkle_tree.rs``; called
  directly at recursion-framework/src/universal_verifier_gadget/circuit_set.rs:189, :216)
* ``hash_no_pad / hash_or_noop / two_to_one / hash_pad / permute`` (plonky2 ``hash/hashing.rs``;
  native uses mp2-common/src/poseidon.rs:49-51, mp2-common/src/utils.rs:294-315)

Where Rust would panic (``MerkleTree::new`` with a bad ``cap_height`` ...), :class:`Mp2GpuError` is
raised.  All work is done by ``libmp2gpu.so`` through its C ABI; nothing here computes on the CPU.
The hasher is a compile-time feature in the reference (``original_poseidon``); here it is the
``hash_kind`` argument (default Poseidon2 = the reference's default ``C``).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import Mp2GpuError, u64p, u64pp

POSEIDON = 0   # PoseidonGoldilocksConfig
POSEIDON2 = 1  # Poseidon2GoldilocksConfig (default C, mp2-common/src/lib.rs:37-40)
SALT_SIZE = 4
ORDER = 0xFFFFFFFF00000001


def _arr(x, ndim=None) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(x, dtype=np.uint64))
    if ndim is not None and a.ndim != ndim:
        raise ValueError("expected a %d-d array of field elements" % ndim)
    return a


def _ptr(a: Optional[np.ndarray]):
    return a.ctypes.data_as(u64p) if a is not None else None


def _col_ptrs(a: np.ndarray):
    return (u64p * a.shape[0])(*[C.cast(a[c].ctypes.data, u64p) for c in range(a.shape[0])])


def init(device: int = 0) -> None:
    """Bind the calling thread to ``device`` (and fail loudly if it is not a usable sm_100 GPU)."""
    _lib.call("mp2gpu_init", device)


def pinned_empty(shape, dtype=np.uint64) -> np.ndarray:
    """A numpy array over page-locked host memory (``mp2gpu_host_alloc``): what a caller should hand to the host-buffer
    entry points so that copies run at PCIe speed and overlap with compute.  The memory lives as long as the array."""
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    ptr = C.c_void_p(None)
    _lib.call("mp2gpu_host_alloc", C.byref(ptr), max(nbytes, 1))
    buf = (C.c_char * max(nbytes, 1)).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    class _Owner:
        def __init__(self, p):
            self.p = p

        def __del__(self):  # pragma: no cover
            try:
                _lib.load().mp2gpu_host_free(self.p)
            except Exception:  # noqa: BLE001
                pass

    _PINNED_OWNERS[arr.__array_interface__["data"][0]] = _Owner(ptr)
    return arr


_PINNED_OWNERS = {}


def device_count() -> int:
    n = C.c_int(0)
    _lib.call("mp2gpu_device_count", C.byref(n))
    return n.value


def launch_count() -> int:
    return int(_lib.load().mp2gpu_launch_count())


# ------------------------------------------------------------------------------------------------
# Hasher
# ------------------------------------------------------------------------------------------------
def permute(states, hash_kind: int = POSEIDON2) -> np.ndarray:
    """PlonkyPermutation::permute on a batch of width-12 states (shape (..., 12))."""
    s = _arr(states).copy()
    if s.shape[-1] != 12:
        raise ValueError("states must have 12 lanes")
    _lib.call("mp2gpu_permute_batch", _ptr(s), s.size // 12, hash_kind)
    return s


def transcript_permute(state, hash_kind: int = POSEIDON2) -> np.ndarray:
    """One permutation on the HOST (``mp2gpu_transcript_permute``): the challenger's duplexing only."""
    st = np.ascontiguousarray(_arr(state).reshape(12).copy())
    _lib.call("mp2gpu_transcript_permute", _ptr(st), hash_kind)
    return st


def hash_no_pad_batch(inputs, hash_kind: int = POSEIDON2) -> np.ndarray:
    x = _arr(inputs, 2)
    out = np.zeros((x.shape[0], 4), dtype=np.uint64)
    _lib.call("mp2gpu_hash_no_pad_batch", _ptr(x) if x.size else None, x.shape[0], x.shape[1], hash_kind, _ptr(out))
    return out


def hash_no_pad(x, hash_kind: int = POSEIDON2) -> np.ndarray:
    x = _arr(x).reshape(-1)
    if x.size == 0:  # hash_no_pad(&[]) absorbs nothing: mp2-common/src/poseidon.rs:49-51
        return np.zeros(4, dtype=np.uint64)
    return hash_no_pad_batch(x.reshape(1, -1), hash_kind)[0]


def hash_pad(x, hash_kind: int = POSEIDON2) -> np.ndarray:
    """pad10*1 to a multiple of the rate (8), then hash_no_pad (circuit_set.rs:149-151)."""
    x = [int(v) for v in _arr(x).reshape(-1)] + [1]
    while (len(x) + 1) % 8:
        x.append(0)
    return hash_no_pad(np.array(x + [1], dtype=np.uint64), hash_kind)


def circuit_digest(constants_sigmas_cap, degree_bits: int, hash_kind: int = POSEIDON2, domain_separator=()) -> np.ndarray:
    """plonky2's ``circuit_digest`` of ``CircuitBuilder::build``: ``hash_no_pad(cap ‖ hash_pad(domain_separator) ‖
    [degree_bits])`` -- the first consumer of the constants/sigmas commitment's cap, and the formula the reference
    re-checks in-circuit at recursion-framework/src/universal_verifier_gadget/circuit_set.rs:136-158 (which assumes an
    empty domain separator, the default here)."""
    cap = constants_sigmas_cap.hashes if isinstance(constants_sigmas_cap, MerkleCap) else _arr(constants_sigmas_cap)
    parts = np.concatenate([_arr(cap).reshape(-1), hash_pad(np.array(list(domain_separator), dtype=np.uint64), hash_kind),
                            np.array([degree_bits], dtype=np.uint64)])
    return hash_no_pad(parts, hash_kind)


def hash_or_noop(x, hash_kind: int = POSEIDON2) -> np.ndarray:
    x = _arr(x).reshape(-1)
    if x.size <= 4:
        out = np.zeros(4, dtype=np.uint64)
        out[:x.size] = np.where(x >= np.uint64(ORDER), x - np.uint64(ORDER), x)
        return out
    return hash_no_pad(x, hash_kind)


def two_to_one_batch(a, b, hash_kind: int = POSEIDON2) -> np.ndarray:
    a, b = _arr(a, 2), _arr(b, 2)
    if a.shape != b.shape or a.shape[1] != 4:
        raise ValueError("two_to_one takes (count, 4) digests")
    out = np.zeros_like(a)
    _lib.call("mp2gpu_two_to_one_batch", _ptr(a), _ptr(b), a.shape[0], hash_kind, _ptr(out))
    return out


def two_to_one(a, b, hash_kind: int = POSEIDON2) -> np.ndarray:
    return two_to_one_batch(_arr(a).reshape(1, 4), _arr(b).reshape(1, 4), hash_kind)[0]


# ------------------------------------------------------------------------------------------------
# MerkleTree
# ------------------------------------------------------------------------------------------------
@dataclass
class MerkleCap:
    """``MerkleCap(pub Vec<H::Hash>)``: 2^cap_height digests."""
    hashes: np.ndarray  # (2^cap_height, 4)

    def height(self) -> int:
        return int(self.hashes.shape[0]).bit_length() - 1

    def __len__(self) -> int:
        return int(self.hashes.shape[0])

    def flatten(self) -> np.ndarray:
        return self.hashes.reshape(-1)


@dataclass
class MerkleProof:
    siblings: np.ndarray  # (log2(leaves) - cap_height, 4), bottom-up

    def __len__(self) -> int:
        return int(self.siblings.shape[0])


class MerkleTree:
    """``MerkleTree<F, H>{ leaves, digests, cap }`` built on the GPU."""

    def __init__(self, leaves, digests: np.ndarray, cap: MerkleCap, hash_kind: int):
        self.leaves = leaves
        self.digests = digests
        self.cap = cap
        self.hash_kind = hash_kind

    @classmethod
    def new(cls, leaves, cap_height: int, hash_kind: int = POSEIDON2) -> "MerkleTree":
        """``MerkleTree::new(leaves: Vec<Vec<F>>, cap_height)``.

        ``leaves`` is a 2-d array, or a list of 1-d arrays of differing lengths (the circuit-set tree
        pads with ``vec![F::ZERO]``).  Raises where plonky2 panics: length not a power of two,
        ``cap_height > log2(len)``."""
        ragged = not isinstance(leaves, np.ndarray) and len({len(l) for l in leaves}) > 1
        n = len(leaves)
        ncap = 1 << cap_height
        digests = np.zeros((max(2 * (n - ncap), 0), 4), dtype=np.uint64)
        cap = np.zeros((ncap, 4), dtype=np.uint64)
        if ragged:
            rows = [_arr(l).reshape(-1) for l in leaves]
            ptrs = (u64p * n)(*[C.cast(r.ctypes.data, u64p) for r in rows])
            lens = (C.c_size_t * n)(*[r.size for r in rows])
            _lib.call("mp2gpu_merkle_new_ragged", ptrs, lens, n, cap_height, hash_kind,
                      _ptr(digests) if digests.size else None, _ptr(cap))
            return cls(rows, digests, MerkleCap(cap), hash_kind)
        lv = _arr(leaves, 2)
        _lib.call("mp2gpu_merkle_new", _ptr(lv) if lv.size else None, lv.shape[0], lv.shape[1], cap_height,
                  hash_kind, _ptr(digests) if digests.size else None, _ptr(cap))
        return cls(lv, digests, MerkleCap(cap), hash_kind)

    def get(self, i: int) -> np.ndarray:
        return self.leaves[i]

    def prove(self, leaf_index: int) -> MerkleProof:
        n = len(self.leaves)
        h = (n.bit_length() - 1) - self.cap.height()
        sib = np.zeros((max(h, 1), 4), dtype=np.uint64)
        cnt = C.c_size_t(0)
        _lib.call("mp2gpu_merkle_prove", _ptr(self.digests) if self.digests.size else None, n,
                  self.cap.height(), leaf_index, _ptr(sib), C.byref(cnt))
        return MerkleProof(sib[:cnt.value])


# ---- wire format -------------------------------------------------------------------------------------
# plonky2 util/serialization Buffer::{write,read}_merkle_tree, used by the reference to (de)serialize the
# circuit-set tree (mp2-common/src/serialization/circuit_data_serialization.rs:74-89, tested at :344-370).
# Layout as recalled from plonky2 0.2.2 -- NOT yet confirmed against a real dump:
#   usize = u64 LE; leaves.len(), then per leaf: len + canonical u64 LE elements;
#   digests.len() + 4 x u64 LE each; cap height; 2^height cap hashes.
def write_merkle_tree(tree: MerkleTree) -> bytes:
    import struct

    out = [struct.pack("<Q", len(tree.leaves))]
    for leaf in tree.leaves:
        leaf = _arr(leaf).reshape(-1)
        out.append(struct.pack("<Q", leaf.size))
        out.append(leaf.astype("<u8").tobytes())
    out.append(struct.pack("<Q", int(tree.digests.shape[0])))
    out.append(np.ascontiguousarray(tree.digests).astype("<u8").tobytes())
    out.append(struct.pack("<Q", tree.cap.height()))
    out.append(np.ascontiguousarray(tree.cap.hashes).astype("<u8").tobytes())
    return b"".join(out)


def read_merkle_tree(data: bytes, hash_kind: int = POSEIDON2) -> MerkleTree:
    import struct

    off = 0

    def usize():
        nonlocal off
        v = struct.unpack_from("<Q", data, off)[0]
        off += 8
        return v

    def elems(n):
        nonlocal off
        a = np.frombuffer(data, dtype="<u8", count=n, offset=off).astype(np.uint64)
        off += 8 * n
        return a

    leaves = [elems(usize()) for _ in range(usize())]
    digests = elems(4 * usize()).reshape(-1, 4)
    cap = elems(4 << usize()).reshape(-1, 4)
    if len({l.size for l in leaves}) == 1:
        leaves = np.stack(leaves)
    return MerkleTree(leaves, digests, MerkleCap(cap), hash_kind)


# plonky2 util/serialization `Write::write_polynomial_batch` (inside CircuitData::to_bytes, which the reference
# caches parameters with: mp2-common/src/serialization/circuit_data_serialization.rs:74-150):
#   polynomials.len(); per polynomial: coeffs.len() + canonical u64 LE coefficients; the merkle tree as above;
#   degree_log, rate_bits as usize; blinding as one byte.   Layout restated from plonky2 0.2.2, unconfirmed (no
#   Rust toolchain here), like write_merkle_tree.
def write_polynomial_batch(batch: "PolynomialBatch") -> bytes:
    import struct

    polys = _arr(batch.polynomials, 2)
    out = [struct.pack("<Q", polys.shape[0])]
    for col in polys:
        out.append(struct.pack("<Q", col.size))
        out.append(col.astype("<u8").tobytes())
    out.append(write_merkle_tree(batch.merkle_tree))
    out.append(struct.pack("<QQB", batch.degree_log, batch.rate_bits, 1 if batch.blinding else 0))
    return b"".join(out)


def read_polynomial_batch(data: bytes, hash_kind: int = POSEIDON2) -> "PolynomialBatch":
    import struct

    off = 0
    (npolys,) = struct.unpack_from("<Q", data, off)
    off += 8
    cols = []
    for _ in range(npolys):
        (ln,) = struct.unpack_from("<Q", data, off)
        off += 8
        cols.append(np.frombuffer(data, dtype="<u8", count=ln, offset=off).astype(np.uint64))
        off += 8 * ln
    tree = read_merkle_tree(data[off:], hash_kind)
    off += len(write_merkle_tree(tree))
    degree_log, rate_bits, blinding = struct.unpack_from("<QQB", data, off)
    if off + 17 != len(data):
        raise Mp2GpuError("trailing bytes after the polynomial batch")
    return PolynomialBatch(np.stack(cols), tree, degree_log, rate_bits, bool(blinding))


def verify_merkle_proof_to_cap(leaf_data, leaf_index: int, cap: MerkleCap, proof: MerkleProof,
                               hash_kind: int = POSEIDON2) -> None:
    """plonky2's ``verify_merkle_proof_to_cap`` (native twin of the gadget used at
    recursion-framework/src/universal_verifier_gadget/verifier_gadget.rs:136-167)."""
    cur = hash_or_noop(leaf_data, hash_kind)
    idx = leaf_index
    for sib in proof.siblings:
        cur = two_to_one(sib, cur, hash_kind) if idx & 1 else two_to_one(cur, sib, hash_kind)
        idx >>= 1
    if not np.array_equal(cur, cap.hashes[idx]):
        raise Mp2GpuError("Invalid Merkle proof.")


# ------------------------------------------------------------------------------------------------
# PolynomialBatch
# ------------------------------------------------------------------------------------------------
def reverse_bits(x: int, bits: int) -> int:
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


class Communicator:
    """The GPUs of this process that share one wide batch (``mp2gpu_comm_init``): a power-of-two number of
    devices with peer access to each other.  Used by :meth:`PolynomialBatch.from_values_sharded`."""

    def __init__(self, devices: Sequence[int]):
        devs = (C.c_int * len(devices))(*devices)
        self._handle = C.c_void_p(None)
        _lib.call("mp2gpu_comm_init", len(devices), devs, C.byref(self._handle))
        self.devices = list(devices)

    def free(self) -> None:
        if self._handle:
            _lib.load().mp2gpu_comm_free(self._handle)
            self._handle = C.c_void_p(None)

    def __del__(self):  # pragma: no cover
        try:
            self.free()
        except Exception:
            pass


class PolynomialBatch:
    """``PolynomialBatch<F, C, D>{ polynomials, merkle_tree, degree_log, rate_bits, blinding }``."""

    def __init__(self, polynomials, merkle_tree: MerkleTree, degree_log: int, rate_bits: int, blinding: bool,
                 handle=None, num_polys: Optional[int] = None):
        self.polynomials = polynomials  # (ncols, n) coefficients (None when the caller left them on the device)
        self.num_polys = len(polynomials) if polynomials is not None else num_polys
        self.merkle_tree = merkle_tree
        self.degree_log = degree_log
        self.rate_bits = rate_bits
        self.blinding = blinding
        self._handle = handle  # device-resident copy (mp2gpu_batch*), optional

    # -- constructors ------------------------------------------------------------------------
    @classmethod
    def _commit(cls, fn: str, cols, rate_bits, blinding, cap_height, hash_kind, keep_on_device, fetch_leaves,
                fetch_coeffs=True, fetch_digests=True):
        if blinding:
            raise Mp2GpuError("blinding (salted) batches are not supported: the reference never enables "
                              "zero_knowledge (mp2-common/src/lib.rs:45-47)")
        cols = _arr(cols, 2)
        ncols, n = cols.shape
        n_log = int(n).bit_length() - 1
        if ncols == 0 or n == 0 or (1 << n_log) != n:
            raise Mp2GpuError("PolynomialValues length must be a power of two and the batch non-empty")
        N = n << rate_bits
        ncap = 1 << cap_height
        if not keep_on_device and not (fetch_coeffs and fetch_digests):
            raise Mp2GpuError("outputs can only be left on the device with keep_on_device=True")
        coeffs = np.empty((ncols, n), dtype=np.uint64) if fetch_coeffs else None
        leaves = np.empty((N, ncols), dtype=np.uint64) if fetch_leaves else None
        digests = np.empty((max(2 * (N - ncap), 0), 4), dtype=np.uint64) if fetch_digests else None
        cap = np.empty((ncap, 4), dtype=np.uint64)
        handle = C.c_void_p(None)
        _lib.call(fn, _col_ptrs(cols), ncols, n_log, rate_bits, cap_height, hash_kind,
                  _col_ptrs(coeffs) if coeffs is not None else None,
                  _ptr(leaves), _ptr(digests) if digests is not None and digests.size else None, _ptr(cap),
                  C.byref(handle) if keep_on_device else None)
        tree = MerkleTree(leaves, digests, MerkleCap(cap), hash_kind)
        return cls(coeffs, tree, n_log, rate_bits, False, handle if keep_on_device else None, ncols)

    @classmethod
    def from_values(cls, values, rate_bits: int, blinding: bool, cap_height: int, timing=None,
                    fft_root_table=None, hash_kind: int = POSEIDON2, keep_on_device: bool = False,
                    fetch_leaves: bool = True, fetch_coeffs: bool = True, fetch_digests: bool = True) -> "PolynomialBatch":
        """``PolynomialBatch::from_values(values, rate_bits, blinding, cap_height, timing, fft_root_table)``.
        ``timing`` / ``fft_root_table`` are accepted and ignored (twiddles are device resident).  With
        ``keep_on_device`` the leaves / coefficients / digests may stay in HBM (``fetch_* = False``): every later
        reader (quotient, openings, query rounds) has a device entry point; only the cap always comes back."""
        return cls._commit("mp2gpu_commit_from_values", values, rate_bits, blinding, cap_height, hash_kind,
                           keep_on_device, fetch_leaves, fetch_coeffs, fetch_digests)

    @classmethod
    def from_coeffs(cls, polynomials, rate_bits: int, blinding: bool, cap_height: int, timing=None,
                    fft_root_table=None, hash_kind: int = POSEIDON2, keep_on_device: bool = False,
                    fetch_leaves: bool = True, fetch_coeffs: bool = True, fetch_digests: bool = True) -> "PolynomialBatch":
        return cls._commit("mp2gpu_commit_from_coeffs", polynomials, rate_bits, blinding, cap_height, hash_kind,
                           keep_on_device, fetch_leaves, fetch_coeffs, fetch_digests)

    @classmethod
    def from_values_sharded(cls, comm: "Communicator", values, rate_bits: int, blinding: bool, cap_height: int,
                            hash_kind: int = POSEIDON2, from_coeffs: bool = False,
                            fetch_leaves: bool = True) -> "PolynomialBatch":
        """The same commitment computed by all the devices of ``comm`` (``mp2gpu_commit_from_values_sharded``):
        columns sharded, the exchange fused into the LDE kernel's peer stores, rows and subtrees sharded.
        Same result, same layouts as :meth:`from_values` / :meth:`from_coeffs`."""
        if blinding:
            raise Mp2GpuError("blinding (salted) batches are not supported: the reference never enables "
                              "zero_knowledge (mp2-common/src/lib.rs:45-47)")
        cols = _arr(values, 2)
        ncols, n = cols.shape
        n_log = int(n).bit_length() - 1
        if ncols == 0 or n == 0 or (1 << n_log) != n:
            raise Mp2GpuError("PolynomialValues length must be a power of two and the batch non-empty")
        N = n << rate_bits
        ncap = 1 << cap_height
        coeffs = np.empty((ncols, n), dtype=np.uint64)
        leaves = np.empty((N, ncols), dtype=np.uint64) if fetch_leaves else None
        digests = np.empty((max(2 * (N - ncap), 0), 4), dtype=np.uint64)
        cap = np.empty((ncap, 4), dtype=np.uint64)
        _lib.call("mp2gpu_commit_from_values_sharded", comm._handle, _col_ptrs(cols), ncols, n_log, rate_bits,
                  cap_height, hash_kind, 1 if from_coeffs else 0, _col_ptrs(coeffs), _ptr(leaves),
                  _ptr(digests) if digests.size else None, _ptr(cap))
        return cls(coeffs, MerkleTree(leaves, digests, MerkleCap(cap), hash_kind), n_log, rate_bits, False, None)

    # -- accessors -----------------------------------------------------------------------------
    def get_lde_values(self, index: int, step: int) -> np.ndarray:
        """Row ``reverse_bits(index * step, degree_log + rate_bits)`` of the leaves (minus salt)."""
        row = reverse_bits(index * step, self.degree_log + self.rate_bits)
        if self.merkle_tree.leaves is not None:
            return self.merkle_tree.leaves[row]
        return self.fetch_rows([row])[0]

    def fetch_rows(self, rows: Sequence[int]) -> np.ndarray:
        if self._handle is None:
            raise Mp2GpuError("batch was not kept on the device")
        idx = _arr(rows, 1)
        out = np.empty((idx.size, self.polynomials.shape[0]), dtype=np.uint64)
        _lib.call("mp2gpu_batch_fetch_rows", self._handle, _ptr(idx), idx.size, _ptr(out))
        return out

    def prove_on_device(self, leaf_index: int) -> MerkleProof:
        if self._handle is None:
            raise Mp2GpuError("batch was not kept on the device")
        h = self.degree_log + self.rate_bits - self.merkle_tree.cap.height()
        sib = np.zeros((max(h, 1), 4), dtype=np.uint64)
        cnt = C.c_size_t(0)
        _lib.call("mp2gpu_batch_prove", self._handle, leaf_index, _ptr(sib), C.byref(cnt))
        return MerkleProof(sib[:cnt.value])

    def open(self, leaf_indices: Sequence[int]):
        """Query-phase openings from the device-resident batch: (rows (q, ncols), siblings (q, h, 4))."""
        if self._handle is None:
            raise Mp2GpuError("batch was not kept on the device")
        idx = _arr(leaf_indices, 1)
        h = self.degree_log + self.rate_bits - self.merkle_tree.cap.height()
        rows = np.empty((idx.size, self.num_polys), dtype=np.uint64)
        sib = np.empty((idx.size, h, 4), dtype=np.uint64)
        _lib.call("mp2gpu_batch_open", self._handle, _ptr(idx), idx.size, _ptr(rows), _ptr(sib) if h else None)
        return rows, sib

    def eval(self, points) -> np.ndarray:
        """``p.to_extension().eval(z)`` for every polynomial and every extension point (``OpeningSet::new``):
        (npoints, ncols, 2), computed from the coefficients resident in HBM."""
        if self._handle is None:
            raise Mp2GpuError("batch was not kept on the device (keep_on_device=False)")
        pts = np.ascontiguousarray(_arr(points).reshape(-1, 2))
        out = np.zeros((pts.shape[0], self.num_polys, 2), dtype=np.uint64)
        _lib.call("mp2gpu_batch_eval", self._handle, _ptr(pts), pts.shape[0], _ptr(out))
        return out

    def free(self) -> None:
        if self._handle is not None:
            _lib.load().mp2gpu_batch_free(self._handle)
            self._handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.free()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
# FRI commit phase (plonky2 fri/prover.rs fri_committed_trees)
# ------------------------------------------------------------------------------------------------
class FriCommitPhase:
    """Device-resident state of one ``fri_committed_trees`` loop (extension elements are [a0, a1] pairs).

    ``final_poly_coeffs``: (n, 2) coefficients of the batched opening polynomial *before* ``lde(rate_bits)``
    (the zero padding is implied).  Per reduction layer: :meth:`commit_layer` = ``MerkleTree::new(chunked
    values, cap_height)``, then the caller's challenger yields ``beta``, then :meth:`fold` =
    ``reduce_with_powers`` per chunk + ``coset_fft(shift^arity)``."""

    def __init__(self, final_poly_coeffs, rate_bits: int, cap_height: int, hash_kind: int = POSEIDON2):
        c = _arr(final_poly_coeffs, 2)
        n = c.shape[0]
        n_log = int(n).bit_length() - 1
        if c.shape[1] != 2 or n == 0 or (1 << n_log) != n:
            raise Mp2GpuError("final polynomial must be (2^k, 2) extension coefficients")
        self.hash_kind, self.cap_height, self.rate_bits = hash_kind, cap_height, rate_bits
        self._h = C.c_void_p(None)
        self.num_layers = 0
        _lib.call("mp2gpu_fri_begin", _ptr(c), n_log, rate_bits, cap_height, hash_kind, C.byref(self._h))

    @classmethod
    def from_openings(cls, oracles, batches, alpha, cap_height: int, hash_kind: int = POSEIDON2, want_final_poly=False):
        """``PolynomialBatch::prove_openings`` up to ``fri_proof``: the alpha-batched quotient of the opened
        polynomials, computed from the coefficients the ``oracles`` (device-resident :class:`PolynomialBatch`
        objects, ``FRI_ORACLES`` order) hold in HBM, and kept there as this commit phase's polynomial.

        ``batches``: ``[(point [z0, z1], [(oracle_index, polynomial_index), ...]), ...]`` -- plonky2's
        ``FriBatchInfo { point, polynomials: Vec<FriPolynomialInfo> }``.  With ``want_final_poly`` the (n, 2)
        coefficients are also returned (``self.final_poly``)."""
        if not oracles or not batches:
            raise Mp2GpuError("prove_openings: no oracles / no batches")
        handles = []
        for o in oracles:
            if getattr(o, "_handle", None) is None or not o._handle:
                raise Mp2GpuError("prove_openings needs device-resident batches (keep_on_device=True)")
            handles.append(o._handle)
        harr = (C.c_void_p * len(handles))(*[h.value if isinstance(h, C.c_void_p) else h for h in handles])
        points = np.ascontiguousarray(np.concatenate([_arr(z).reshape(2) for z, _ in batches]))
        sizes = np.array([len(polys) for _, polys in batches], dtype=np.uint32)
        oi = np.array([p[0] for _, polys in batches for p in polys], dtype=np.uint32)
        pi = np.array([p[1] for _, polys in batches for p in polys], dtype=np.uint32)
        u32p = C.POINTER(C.c_uint32)
        self = cls.__new__(cls)
        self.hash_kind, self.cap_height, self.rate_bits = hash_kind, cap_height, oracles[0].rate_bits
        self._h = C.c_void_p(None)
        self.num_layers = 0
        self.final_poly = None
        out = None
        if want_final_poly:
            out = np.zeros((1 << oracles[0].degree_log, 2), dtype=np.uint64)
        _lib.call("mp2gpu_fri_begin_openings", harr, len(handles), _ptr(points), sizes.ctypes.data_as(u32p), len(batches),
                  oi.ctypes.data_as(u32p), pi.ctypes.data_as(u32p), _ptr(_arr(alpha).reshape(2)), cap_height, hash_kind,
                  _ptr(out) if out is not None else None, C.byref(self._h))
        self.final_poly = out
        return self

    def commit_layer(self, arity_bits: int) -> MerkleCap:
        cap = np.zeros((1 << self.cap_height, 4), dtype=np.uint64)
        _lib.call("mp2gpu_fri_commit_layer", self._h, arity_bits, _ptr(cap))
        self.num_layers += 1
        return MerkleCap(cap)

    def fold(self, beta) -> None:
        b = _arr(beta).reshape(2)
        _lib.call("mp2gpu_fri_fold", self._h, _ptr(b))

    def layer(self, i: int) -> MerkleTree:
        nl, ll, nd, nc = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
        _lib.call("mp2gpu_fri_layer_shape", self._h, i, C.byref(nl), C.byref(ll), C.byref(nd), C.byref(nc))
        leaves = np.zeros((nl.value, ll.value), dtype=np.uint64)
        digests = np.zeros((nd.value, 4), dtype=np.uint64)
        cap = np.zeros((nc.value, 4), dtype=np.uint64)
        _lib.call("mp2gpu_fri_fetch_layer", self._h, i, _ptr(leaves), _ptr(digests) if nd.value else None, _ptr(cap))
        return MerkleTree(leaves, digests, MerkleCap(cap), self.hash_kind)

    def open_layer(self, i: int, leaf_indices):
        """(leaves (q, leaf_len), siblings (q, h, 4)) of layer ``i`` for the query rounds."""
        nl, ll, nd, nc = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
        _lib.call("mp2gpu_fri_layer_shape", self._h, i, C.byref(nl), C.byref(ll), C.byref(nd), C.byref(nc))
        idx = _arr(leaf_indices, 1)
        h = (nl.value.bit_length() - 1) - (nc.value.bit_length() - 1)
        leaves = np.empty((idx.size, ll.value), dtype=np.uint64)
        sib = np.empty((idx.size, h, 4), dtype=np.uint64)
        _lib.call("mp2gpu_fri_open_layer", self._h, i, _ptr(idx), idx.size, _ptr(leaves), _ptr(sib) if h else None)
        return leaves, sib

    def finish(self) -> np.ndarray:
        ln = C.c_size_t(0)
        _lib.call("mp2gpu_fri_finish", self._h, None, C.byref(ln))
        out = np.zeros((ln.value, 2), dtype=np.uint64)
        _lib.call("mp2gpu_fri_finish", self._h, _ptr(out), C.byref(ln))
        return out

    def free(self) -> None:
        if self._h:
            _lib.load().mp2gpu_fri_free(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):  # pragma: no cover
        try:
            self.free()
        except Exception:
            pass


def fri_committed_trees(final_poly_coeffs, reduction_arity_bits, betas, rate_bits: int, cap_height: int,
                        hash_kind: int = POSEIDON2):
    """The whole commit phase with the challenger's betas given up front (tests / trace replay).
    Returns ``(trees, final_coeffs)`` like plonky2's function."""
    ph = FriCommitPhase(final_poly_coeffs, rate_bits, cap_height, hash_kind)
    try:
        for ab, beta in zip(reduction_arity_bits, betas):
            ph.commit_layer(ab)
            ph.fold(beta)
        return [ph.layer(i) for i in range(ph.num_layers)], ph.finish()
    finally:
        ph.free()


def fri_proof_of_work(duplex_state, witness_pos: int, min_leading_zeros: int = 16, hash_kind: int = POSEIDON2) -> int:
    """``fri_proof_of_work``: smallest PoW witness for the challenger's intermediate duplex state."""
    st = _arr(duplex_state).reshape(12)
    w = np.zeros(1, dtype=np.uint64)
    _lib.call("mp2gpu_fri_proof_of_work", _ptr(st), witness_pos, min_leading_zeros, hash_kind, _ptr(w))
    return int(w[0])
