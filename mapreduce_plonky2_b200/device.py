"""Device-pointer stages of the commitment for data already resident in HBM.

PyTorch is used only as plumbing here: device memory (``torch.int64`` tensors viewed as u64 field
elements), streams and, in ``sharded.py``, ``torch.distributed``.  All arithmetic is done by the
hand-written kernels of ``libmp2gpu.so`` reached through the ``mp2gpu_dev_*`` C ABI, launched on
torch's current stream so that ``torch.cuda.Event`` timings bracket them.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

POSEIDON, POSEIDON2 = 0, 1


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _chk(t: torch.Tensor, name: str) -> int:
    if t.dtype != torch.int64 or not t.is_cuda or not t.is_contiguous():
        raise ValueError("%s must be a contiguous CUDA int64 tensor (u64 field elements)" % name)
    return t.data_ptr()


def bind_current_device() -> None:
    _lib.call("mp2gpu_init", torch.cuda.current_device())


class CommitBuffers:
    """Caller-owned device buffers of one commitment (allocated once, reused every step)."""

    def __init__(self, ncols: int, n_log: int, rate_bits: int, cap_height: int, want_leaves: bool = True,
                 device=None):
        n, N = 1 << n_log, (1 << n_log) << rate_bits
        ncap = 1 << cap_height
        kw = dict(dtype=torch.int64, device=device or torch.device("cuda", torch.cuda.current_device()))
        self.shape = (ncols, n_log, rate_bits, cap_height)
        self.coeffs = torch.empty((ncols, n), **kw)
        self.lde = torch.empty((ncols, N), **kw)            # leaf-ordered, column-major
        self.leaves = torch.empty((N, ncols), **kw) if want_leaves else None  # plonky2's row-major leaves
        self.digests = torch.empty((max(2 * (N - ncap), 1), 4), **kw)
        self.cap = torch.empty((ncap, 4), **kw)


def commit_resident(cols: torch.Tensor, bufs: CommitBuffers, hash_kind: int = POSEIDON2,
                    from_coeffs: bool = False) -> None:
    """PolynomialBatch::from_values / from_coeffs on resident inputs; asynchronous on the current stream."""
    ncols, n_log, rate_bits, cap_height = bufs.shape
    if tuple(cols.shape) != (ncols, 1 << n_log):
        raise ValueError("cols must be (ncols, n)")
    _lib.call("mp2gpu_dev_commit", _chk(cols, "cols"), ncols, n_log, rate_bits, cap_height, hash_kind,
              1 if from_coeffs else 0, _chk(bufs.coeffs, "coeffs"), _chk(bufs.lde, "lde"),
              _chk(bufs.leaves, "leaves") if bufs.leaves is not None else None, _chk(bufs.digests, "digests"),
              _chk(bufs.cap, "cap"), _stream_ptr())


def intt(values: torch.Tensor, coeffs: torch.Tensor) -> None:
    ncols, n = values.shape
    _lib.call("mp2gpu_dev_intt", _chk(values, "values"), n, _chk(coeffs, "coeffs"), n, ncols,
              n.bit_length() - 1, _stream_ptr())


def coset_lde(coeffs: torch.Tensor, lde: torch.Tensor, rate_bits: int, shard_log: int = 0) -> None:
    """coeffs (ncols, n) -> leaf-ordered column-major LDE.  With ``shard_log = log2(G)`` ``lde`` has shape
    (G, ncols, N/G): block g holds the leaves of row-shard g (the all-to-all payload for rank g)."""
    ncols, n = coeffs.shape
    N = n << rate_bits
    if shard_log:
        G = 1 << shard_log
        if tuple(lde.shape) != (G, ncols, N // G):
            raise ValueError("sharded lde must be (G, ncols, N/G)")
        lde_stride, shard_stride = N // G, ncols * (N // G)
    else:
        if tuple(lde.shape) != (ncols, N):
            raise ValueError("lde must be (ncols, N)")
        lde_stride, shard_stride = N, 0
    _lib.call("mp2gpu_dev_coset_lde", _chk(coeffs, "coeffs"), n, _chk(lde, "lde"), lde_stride, ncols,
              n.bit_length() - 1, rate_bits, shard_log, shard_stride, _stream_ptr())


def coset_lde_peer(coeffs: torch.Tensor, shard_ptrs, n_loc: int, rate_bits: int, first_shard: int = 0,
                   scratch: Optional[torch.Tensor] = None) -> None:
    """coeffs (ncols, n) -> LDE whose shard g is stored at device address ``shard_ptrs[g]`` (column c at
    + c*n_loc elements): the peers' receive buffers, i.e. the all-to-all happens in the kernel's store.
    ``first_shard``: destination written first (pass the caller's rank: the ranks then never share a target).
    ``scratch``: local (ncols * N)-element buffer for the four-step intermediate (reused across calls)."""
    import ctypes as C

    ncols, n = coeffs.shape
    G = len(shard_ptrs)
    arr = (C.c_void_p * G)(*[C.c_void_p(int(p)) for p in shard_ptrs])
    if scratch is not None and scratch.numel() < ncols * (n << rate_bits):
        raise ValueError("scratch must hold ncols * N elements")
    _lib.call("mp2gpu_dev_coset_lde_peer", _chk(coeffs, "coeffs"), n, arr, n_loc, ncols, n.bit_length() - 1,
              rate_bits, G.bit_length() - 1, first_shard, _chk(scratch, "scratch") if scratch is not None else None,
              _stream_ptr())


def merkle_colmajor(lde: torch.Tensor, cap_height: int, hash_kind: int, leaves: Optional[torch.Tensor],
                    digests: torch.Tensor, cap: torch.Tensor) -> None:
    """lde (ncols, nleaves) leaf-ordered column-major -> row-major leaves (optional), digests, cap."""
    ncols, nleaves = lde.shape
    _lib.call("mp2gpu_dev_merkle_colmajor", _chk(lde, "lde"), nleaves, ncols, nleaves, cap_height, hash_kind,
              _chk(leaves, "leaves") if leaves is not None else None, _chk(digests, "digests"), _chk(cap, "cap"),
              _stream_ptr())


def merkle_colmajor_leaves(lde: torch.Tensor, cap_height: int, hash_kind: int, leaf_begin: int, leaf_end: int,
                           leaves: Optional[torch.Tensor], digests: torch.Tensor, cap: torch.Tensor) -> None:
    """Leaf digests (and optional row-major rows) of leaves [leaf_begin, leaf_end) only -- the first half of
    :func:`merkle_colmajor`, for callers that overlap host copies of one leaf block with the hashing of the next."""
    ncols, nleaves = lde.shape
    _lib.call("mp2gpu_dev_merkle_colmajor_leaves", _chk(lde, "lde"), nleaves, ncols, nleaves, cap_height, hash_kind,
              leaf_begin, leaf_end, _chk(leaves, "leaves") if leaves is not None else None, _chk(digests, "digests"),
              _chk(cap, "cap"), _stream_ptr())


def merkle_levels(nleaves: int, cap_height: int, hash_kind: int, digests: torch.Tensor, cap: torch.Tensor) -> None:
    """Inner levels + cap once every leaf digest is in place (second half of :func:`merkle_colmajor`)."""
    _lib.call("mp2gpu_dev_merkle_levels", nleaves, cap_height, hash_kind, _chk(digests, "digests"), _chk(cap, "cap"),
              _stream_ptr())


def merkle_rowmajor(leaves: torch.Tensor, cap_height: int, hash_kind: int, digests: torch.Tensor,
                    cap: torch.Tensor) -> None:
    nleaves, leaf_len = leaves.shape
    _lib.call("mp2gpu_dev_merkle_rowmajor", _chk(leaves, "leaves"), nleaves, leaf_len, cap_height, hash_kind,
              _chk(digests, "digests"), _chk(cap, "cap"), _stream_ptr())


def canonicalize(src: torch.Tensor, dst: torch.Tensor) -> None:
    _lib.call("mp2gpu_dev_canonicalize", _chk(src, "src"), _chk(dst, "dst"), src.numel(), _stream_ptr())


# ---- measurement hooks ---------------------------------------------------------------------------
def profile_enable(on: bool) -> None:
    _lib.call("mp2gpu_profile_enable", 1 if on else 0)


def profile_report() -> dict:
    """{kernel_name: (launches, total_ms)} since the last report (synchronises the device)."""
    import ctypes as C

    buf = C.create_string_buffer(1 << 16)
    _lib.call("mp2gpu_profile_report", buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.split()
        out[name] = (int(cnt), float(ms))
    return out


def int_pipe_peak() -> dict:
    """Live IMAD issue-rate probe: the Poseidon roofline denominator."""
    import ctypes as C

    a, b, c = C.c_double(0), C.c_double(0), C.c_double(0)
    _lib.call("mp2gpu_debug_int_pipe_peak", C.byref(a), C.byref(b), C.byref(c))
    return {"imad_per_clk_per_sm": a.value, "sm_clock_mhz": b.value, "t_imad_per_s": c.value}
