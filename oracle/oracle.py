"""ctypes binding of the C oracle (``libmp2oracle.so``) -- test infrastructure only.

Function names follow plonky2's (``hash_no_pad``, ``two_to_one``, ``merkle_new`` ...);
the restated algorithm and its reference anchors are documented in ``mp2_oracle.h`` /
``mp2_oracle.c`` (reference call sites
recursion-framework/src/universal_verifier_gadget/circuit_set.rs:173-237).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmp2oracle.so")

POSEIDON = 0
POSEIDON2 = 1
P = 0xFFFFFFFF00000001

_u64p = C.POINTER(C.c_uint64)


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only)."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
        os.path.getmtime(os.path.join(_HERE, f)) for f in ("mp2_oracle.c", "mp2_oracle.h", "Makefile")
    ):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        L = _lib
        for name in ("orc_gl_add", "orc_gl_sub", "orc_gl_mul", "orc_gl_pow"):
            getattr(L, name).restype = C.c_uint64
            getattr(L, name).argtypes = [C.c_uint64, C.c_uint64]
        for name in ("orc_gl_inv", "orc_gl_canon"):
            getattr(L, name).restype = C.c_uint64
            getattr(L, name).argtypes = [C.c_uint64]
        L.orc_gl_root_of_unity.restype = C.c_uint64
        L.orc_gl_root_of_unity.argtypes = [C.c_uint32]
        L.orc_permute.argtypes = [C.c_uint32, _u64p]
        L.orc_hash_no_pad.argtypes = [C.c_uint32, _u64p, C.c_size_t, _u64p]
        L.orc_hash_pad.argtypes = [C.c_uint32, _u64p, C.c_size_t, _u64p]
        L.orc_hash_or_noop.argtypes = [C.c_uint32, _u64p, C.c_size_t, _u64p]
        L.orc_two_to_one.argtypes = [C.c_uint32, _u64p, _u64p, _u64p]
        L.orc_fft.argtypes = [_u64p, C.c_uint32]
        L.orc_ifft.argtypes = [_u64p, C.c_uint32]
        L.orc_coset_lde.argtypes = [_u64p, C.c_uint32, C.c_uint32, C.c_uint64, _u64p]
        L.orc_eval_naive.argtypes = [_u64p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_size_t, _u64p]
        L.orc_merkle_new.restype = C.c_int
        L.orc_merkle_new.argtypes = [_u64p, C.c_size_t, C.c_size_t, C.c_uint32, C.c_uint32, _u64p,
                                     _u64p, C.c_int]
        L.orc_merkle_prove.restype = C.c_int
        L.orc_merkle_prove.argtypes = [_u64p, C.c_size_t, C.c_uint32, C.c_size_t, _u64p]
        L.orc_merkle_verify.restype = C.c_int
        L.orc_merkle_verify.argtypes = [_u64p, C.c_size_t, C.c_size_t, _u64p, C.c_size_t,
                                        C.c_uint32, _u64p]
        L.orc_commit.restype = C.c_int
        L.orc_commit.argtypes = [C.POINTER(_u64p), C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32,
                                 C.c_uint32, C.c_int, _u64p, _u64p, _u64p, _u64p, C.c_int]
        L.orc_ext_mul.argtypes = [_u64p, _u64p, _u64p]
        L.orc_fri_fold.argtypes = [_u64p, C.c_size_t, C.c_uint32, _u64p, _u64p]
        L.orc_coset_fft_ext.argtypes = [_u64p, C.c_uint32, C.c_uint64, _u64p]
        L.orc_fri_layer_leaves.argtypes = [_u64p, C.c_uint32, C.c_uint32, _u64p]
        L.orc_fri_combine.restype = C.c_int
        L.orc_fri_combine.argtypes = [C.POINTER(_u64p), C.POINTER(C.c_uint32), C.c_size_t, _u64p, _u64p, C.c_size_t, _u64p]
        L.orc_fri_pow.restype = C.c_int
        L.orc_fri_pow.argtypes = [C.c_uint32, _u64p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, _u64p]
        L.orc_max_threads.restype = C.c_int
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u64p)


def _arr(x, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(x, dtype=np.uint64))
    return a if shape is None else a.reshape(shape)


def max_threads() -> int:
    return int(lib().orc_max_threads())


# ---- field ----
def gl_add(a, b): return int(lib().orc_gl_add(a, b))
def gl_sub(a, b): return int(lib().orc_gl_sub(a, b))
def gl_mul(a, b): return int(lib().orc_gl_mul(a, b))
def gl_pow(a, e): return int(lib().orc_gl_pow(a, e))
def gl_inv(a): return int(lib().orc_gl_inv(a))
def root_of_unity(log_n): return int(lib().orc_gl_root_of_unity(log_n))


# ---- permutations / constants ----
def poseidon_round_constants() -> np.ndarray:
    out = np.zeros(360, dtype=np.uint64)
    lib().orc_poseidon_round_constants(_p(out))
    return out


def poseidon2_round_constants() -> np.ndarray:
    out = np.zeros(118, dtype=np.uint64)
    lib().orc_poseidon2_round_constants(_p(out))
    return out


def poseidon2_diag() -> np.ndarray:
    out = np.zeros(12, dtype=np.uint64)
    lib().orc_poseidon2_diag(_p(out))
    return out


def permute(state, hash_kind=POSEIDON) -> np.ndarray:
    s = _arr(state).copy()
    assert s.shape == (12,)
    lib().orc_permute(hash_kind, _p(s))
    return s


# ---- sponge ----
def _hash(fn, x, hash_kind):
    x = _arr(x).reshape(-1)
    out = np.zeros(4, dtype=np.uint64)
    buf = x if x.size else np.zeros(1, dtype=np.uint64)
    fn(hash_kind, _p(buf), x.size, _p(out))
    return out


def hash_no_pad(x, hash_kind=POSEIDON): return _hash(lib().orc_hash_no_pad, x, hash_kind)
def hash_pad(x, hash_kind=POSEIDON): return _hash(lib().orc_hash_pad, x, hash_kind)
def hash_or_noop(x, hash_kind=POSEIDON): return _hash(lib().orc_hash_or_noop, x, hash_kind)


def two_to_one(a, b, hash_kind=POSEIDON) -> np.ndarray:
    a, b = _arr(a), _arr(b)
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_two_to_one(hash_kind, _p(a), _p(b), _p(out))
    return out


# ---- transforms (one column) ----
def fft(v) -> np.ndarray:
    v = _arr(v).copy()
    lib().orc_fft(_p(v), int(v.size).bit_length() - 1)
    return v


def ifft(v) -> np.ndarray:
    v = _arr(v).copy()
    lib().orc_ifft(_p(v), int(v.size).bit_length() - 1)
    return v


def coset_lde(coeffs, rate_bits, shift=7) -> np.ndarray:
    c = _arr(coeffs)
    out = np.zeros(c.size << rate_bits, dtype=np.uint64)
    lib().orc_coset_lde(_p(c), int(c.size).bit_length() - 1, rate_bits, shift, _p(out))
    return out


def eval_naive(coeffs, shift, w, n_out) -> np.ndarray:
    c = _arr(coeffs)
    out = np.zeros(n_out, dtype=np.uint64)
    lib().orc_eval_naive(_p(c), c.size, shift, w, n_out, _p(out))
    return out


# ---- Merkle ----
def merkle_new(leaves, cap_height, hash_kind=POSEIDON, nthreads=0):
    """leaves: (nleaves, leaf_len) -> (digests (2*(n-2^cap),4), cap (2^cap,4)).

    Raises ValueError where plonky2's MerkleTree::new would panic (non power of two,
    cap_height > log2(len))."""
    lv = _arr(leaves)
    assert lv.ndim == 2
    n, ll = lv.shape
    ncap = 1 << cap_height
    digests = np.zeros((max(2 * (n - ncap), 0), 4), dtype=np.uint64)
    cap = np.zeros((ncap, 4), dtype=np.uint64)
    dbuf = digests if digests.size else np.zeros((1, 4), dtype=np.uint64)
    rc = lib().orc_merkle_new(_p(lv), n, ll, cap_height, hash_kind, _p(dbuf), _p(cap),
                              nthreads or max_threads())
    if rc != 0:
        raise ValueError("MerkleTree::new: bad nleaves / cap_height")
    return digests, cap


def merkle_prove(digests, nleaves, cap_height, leaf_index) -> np.ndarray:
    d = _arr(digests)
    h = (int(nleaves).bit_length() - 1) - cap_height
    sib = np.zeros((max(h, 1), 4), dtype=np.uint64)
    dbuf = d if d.size else np.zeros((1, 4), dtype=np.uint64)
    rc = lib().orc_merkle_prove(_p(dbuf), nleaves, cap_height, leaf_index, _p(sib))
    if rc < 0:
        raise ValueError("MerkleTree::prove: bad arguments")
    return sib[:rc]


def merkle_verify(leaf, leaf_index, siblings, hash_kind=POSEIDON):
    leaf = _arr(leaf).reshape(-1)
    sib = _arr(siblings).reshape(-1, 4)
    root = np.zeros(4, dtype=np.uint64)
    sbuf = sib if sib.size else np.zeros((1, 4), dtype=np.uint64)
    cap_idx = lib().orc_merkle_verify(_p(leaf), leaf.size, leaf_index, _p(sbuf), sib.shape[0],
                                      hash_kind, _p(root))
    return cap_idx, root


# ---- PolynomialBatch ----
def commit(cols, rate_bits, cap_height, hash_kind=POSEIDON, from_coeffs=False, nthreads=0,
           want_leaves=True):
    """cols: (ncols, n) values (or coeffs).  Returns dict(coeffs, leaves, digests, cap)."""
    cols = _arr(cols)
    assert cols.ndim == 2
    ncols, n = cols.shape
    log_n = int(n).bit_length() - 1
    assert 1 << log_n == n
    N = n << rate_bits
    ncap = 1 << cap_height
    coeffs = np.zeros((ncols, n), dtype=np.uint64)
    leaves = np.zeros((N, ncols), dtype=np.uint64) if want_leaves else None
    digests = np.zeros((max(2 * (N - ncap), 0), 4), dtype=np.uint64)
    cap = np.zeros((ncap, 4), dtype=np.uint64)
    ptrs = (_u64p * ncols)(*[C.cast(cols[c].ctypes.data, _u64p) for c in range(ncols)])
    dbuf = digests if digests.size else np.zeros((1, 4), dtype=np.uint64)
    rc = lib().orc_commit(ptrs, ncols, log_n, rate_bits, cap_height, hash_kind,
                          1 if from_coeffs else 0, _p(coeffs),
                          _p(leaves) if want_leaves else None, _p(dbuf), _p(cap),
                          nthreads or max_threads())
    if rc != 0:
        raise ValueError("commit: bad arguments")
    return {"coeffs": coeffs, "leaves": leaves, "digests": digests, "cap": cap}


# ---- FRI commit phase (ext elements as (m, 2) arrays of [a0, a1]) ----
def ext_mul(a, b) -> np.ndarray:
    a, b = _arr(a), _arr(b)
    out = np.zeros(2, dtype=np.uint64)
    lib().orc_ext_mul(_p(a), _p(b), _p(out))
    return out


def fri_fold(coeffs, arity_bits, beta) -> np.ndarray:
    c = _arr(coeffs).reshape(-1, 2)
    out = np.zeros((c.shape[0] >> arity_bits, 2), dtype=np.uint64)
    lib().orc_fri_fold(_p(c), c.shape[0], arity_bits, _p(_arr(beta)), _p(out))
    return out


def coset_fft_ext(coeffs, shift) -> np.ndarray:
    c = _arr(coeffs).reshape(-1, 2)
    out = np.zeros_like(c)
    lib().orc_coset_fft_ext(_p(c), int(c.shape[0]).bit_length() - 1, shift, _p(out))
    return out


def fri_layer_leaves(values, arity_bits) -> np.ndarray:
    v = _arr(values).reshape(-1, 2)
    out = np.zeros_like(v)
    lib().orc_fri_layer_leaves(_p(v), int(v.shape[0]).bit_length() - 1, arity_bits, _p(out))
    return out.reshape(v.shape[0] >> arity_bits, 2 << arity_bits)


def fri_committed_trees(coeffs, values, arity_bits_list, betas, cap_height, hash_kind=POSEIDON, rate_bits=3):
    """plonky2 fri_committed_trees with the challenger's betas supplied by the caller.
    Returns ([(leaves, digests, cap) per layer], final_coeffs truncated by rate_bits)."""
    coeffs, values = _arr(coeffs).reshape(-1, 2), _arr(values).reshape(-1, 2)
    shift, trees = 7, []
    for ab, beta in zip(arity_bits_list, betas):
        leaves = fri_layer_leaves(values, ab)
        digests, cap = merkle_new(leaves, min(cap_height, int(leaves.shape[0]).bit_length() - 1), hash_kind)
        trees.append((leaves, digests, cap))
        coeffs = fri_fold(coeffs, ab, beta)
        shift = gl_pow(shift, 1 << ab)
        values = coset_fft_ext(coeffs, shift)
    return trees, coeffs[:coeffs.shape[0] >> rate_bits]


def fri_combine(batches, alpha) -> np.ndarray:
    """``batches``: [(point [z0, z1], [coefficient column, ...]), ...] -> (n, 2) final polynomial of prove_openings."""
    cols = [_arr(c) for _, polys in batches for c in polys]
    n = cols[0].size
    assert all(c.size == n for c in cols)
    ptrs = (_u64p * len(cols))(*[_p(c) for c in cols])
    sizes = (C.c_uint32 * len(batches))(*[len(polys) for _, polys in batches])
    points = np.concatenate([_arr(z).reshape(2) for z, _ in batches])
    out = np.zeros((n, 2), dtype=np.uint64)
    rc = lib().orc_fri_combine(ptrs, sizes, len(batches), _p(points), _p(_arr(alpha)), n, _p(out))
    if rc != 0:
        raise ValueError("orc_fri_combine: bad arguments")
    return out


def fri_pow(state, pos, min_leading_zeros, hash_kind=POSEIDON, start=0, limit=1 << 40):
    st = _arr(state)
    w = np.zeros(1, dtype=np.uint64)
    rc = lib().orc_fri_pow(hash_kind, _p(st), pos, min_leading_zeros, start, limit, _p(w))
    if rc != 0:
        raise ValueError("no proof-of-work witness below the limit")
    return int(w[0])
