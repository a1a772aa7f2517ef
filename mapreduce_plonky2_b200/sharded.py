"""One wide batch across the GPUs of a node:

    columns sharded -> per-rank iNTT + coset LDE -> ONE exchange (all-to-all: column shards become
    row shards aligned with the cap subtrees) -> per-rank leaf hashing + subtrees -> all-gather of the
    cap (<= 16 x 32 B).

Rank g of G owns columns [g*c/G, (g+1)*c/G) before the exchange and leaves [g*N/G, (g+1)*N/G) after
it.  Leaf L is LDE row bitrev(L), and the LDE kernel already writes its output leaf-ordered and
blocked by destination rank (``shard_log``), so the all-to-all payload for rank s is one contiguous
block and what arrives is directly the column-major input of the leaf-hash kernel: no pack/unpack
kernels on either side.  A rank's digests / cap entries are contiguous slices of the global
``MerkleTree.digests`` / ``cap`` (plonky2 splits digests into 2^cap_height per-subtree chunks).

``torch.distributed`` is plumbing only (NCCL over NVLink on GPUs; gloo in the CPU tests); the
arithmetic is delegated to an *engine*: :class:`CudaEngine` (libmp2gpu.so kernels) in production,
an oracle-backed stand-in inside ``tests/`` to exercise this file's index logic without a GPU.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist


class _StageTimer:
    """Optional per-stage CUDA-event timing of one commit_sharded call (``MP2_SHARDED_TIMING=1``): the marks are
    appended to ``scratch["timing"]`` as (label, event) lists, one list per call; :func:`timing_report` turns them
    into milliseconds after a synchronize.  Diagnostic only (round-1 open item: two stalled peer-exchange samples
    with unchanged kernel times); no effect when the variable is unset."""

    def __init__(self, scratch):
        self.on = bool(os.environ.get("MP2_SHARDED_TIMING")) and torch.cuda.is_available() and scratch is not None
        if self.on:
            self.marks = []
            scratch.setdefault("timing", []).append(self.marks)

    def mark(self, label):
        if self.on:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.marks.append((label, ev))


def timing_report(scratch) -> list:
    """[{stage: ms, ...} per recorded call]; call after ``torch.cuda.synchronize()``."""
    out = []
    for marks in (scratch or {}).get("timing", []):
        out.append({"%s->%s" % (a[0], b[0]): a[1].elapsed_time(b[1]) for a, b in zip(marks, marks[1:])})
    return out


class CudaEngine:
    """Stages implemented by the hand-written kernels (device pointers on the current CUDA stream)."""

    def __init__(self):
        from . import device

        self.d = device
        device.bind_current_device()

    def empty(self, shape):
        return torch.empty(shape, dtype=torch.int64, device=torch.device("cuda", torch.cuda.current_device()))

    def intt(self, values, coeffs):
        self.d.intt(values, coeffs)

    def canonical_copy(self, src, dst):
        self.d.canonicalize(src, dst)

    def coset_lde(self, coeffs, lde, rate_bits, shard_log):
        self.d.coset_lde(coeffs, lde, rate_bits, shard_log)

    def coset_lde_peer(self, coeffs, shard_ptrs, n_loc, rate_bits, first_shard=0, scratch=None):
        self.d.coset_lde_peer(coeffs, shard_ptrs, n_loc, rate_bits, first_shard, scratch)

    def merkle_colmajor(self, lde, cap_height, hash_kind, leaves, digests, cap):
        self.d.merkle_colmajor(lde, cap_height, hash_kind, leaves, digests, cap)

    def merkle_colmajor_leaves(self, lde, cap_height, hash_kind, lb, le, leaves, digests, cap):
        self.d.merkle_colmajor_leaves(lde, cap_height, hash_kind, lb, le, leaves, digests, cap)

    def merkle_levels(self, nleaves, cap_height, hash_kind, digests, cap):
        self.d.merkle_levels(nleaves, cap_height, hash_kind, digests, cap)


@dataclass
class HostOutputs:
    """Pinned host buffers that receive this rank's outputs while the commitment is still running (the end-to-end
    form of the sharded call): coefficients travel back under the LDE, each block of leaf rows under the hashing
    of the next block, digests and cap at the end -- all on ``copy_stream``.  The caller synchronises."""
    coeffs: Optional[torch.Tensor]
    leaves: Optional[torch.Tensor]
    digests: Optional[torch.Tensor]
    cap: Optional[torch.Tensor]
    copy_stream: "torch.cuda.Stream"
    chunks: int = 8


class PeerExchange:
    """Receive buffer of this rank in symmetric memory, mapped into every peer over NVLink/NVSwitch.

    With it the column-shard -> row-shard exchange is not a separate collective: every rank's LDE kernel
    stores shard ``g`` of its output directly into rank ``g``'s receive buffer (``mp2gpu_dev_coset_lde_peer``),
    bracketed by two device-side barriers on the compute stream.  NCCL stays for the 512-byte cap gather."""

    def __init__(self, G: int, c_loc: int, n_loc: int, group=None):
        import torch.distributed._symmetric_memory as symm

        dev = torch.device("cuda", torch.cuda.current_device())
        self.recv = symm.empty((G, c_loc, n_loc), dtype=torch.int64, device=dev)
        self.hdl = symm.rendezvous(self.recv, group if group is not None else dist.group.WORLD)
        block_bytes = c_loc * n_loc * 8
        # where MY block (my columns) lands inside rank g's receive buffer
        self.shard_ptrs = [int(self.hdl.buffer_ptrs[g]) + self.hdl.rank * block_bytes for g in range(G)]
        self.shape = (G, c_loc, n_loc)

    def barrier(self):
        self.hdl.barrier()


@dataclass
class ShardedBatch:
    """This rank's part of the PolynomialBatch."""
    coeffs: torch.Tensor            # (c/G, n)   coefficients of the local columns
    leaves: Optional[torch.Tensor]  # (N/G, c)   rows [g*N/G, (g+1)*N/G) of MerkleTree.leaves
    digests: torch.Tensor           # slice [g*D/G, (g+1)*D/G) of MerkleTree.digests, D = 2*(N - 2^cap)
    cap: torch.Tensor               # (2^cap, 4) the whole cap (all-gathered)
    rank: int
    world: int


def _log2(x: int) -> int:
    l = x.bit_length() - 1
    if x <= 0 or (1 << l) != x:
        raise ValueError("%d is not a power of two" % x)
    return l


def commit_sharded(cols_local: torch.Tensor, ncols_total: int, rate_bits: int, cap_height: int, hash_kind: int,
                   engine, group=None, from_coeffs: bool = False, want_leaves: bool = True,
                   scratch: Optional[dict] = None, exchange: str = "auto",
                   host_out: Optional[HostOutputs] = None) -> ShardedBatch:
    """PolynomialBatch::from_values / from_coeffs of one (ncols_total x n) batch over ``group``.

    ``cols_local``: this rank's columns, shape (ncols_total / G, n).  Requirements: G is a power of two,
    G divides ncols_total, and G <= 2^cap_height (every rank owns whole cap subtrees).
    ``scratch`` may hold reusable buffers (keys: coeffs, send, recv, leaves, digests, cap_local, cap).
    ``exchange``: "peer" = the LDE kernel stores straight into the peers' receive buffers (symmetric memory over
    NVLink), no all-to-all -- the default on GPUs ("auto"); "nccl" = LDE into a send buffer + ``all_to_all_single``
    (what "auto" picks for CPU tensors, i.e. the gloo tests).
    ``host_out``: also stream this rank's outputs into pinned host buffers, overlapped with the compute."""
    G = dist.get_world_size(group)
    g = dist.get_rank(group)
    glog = _log2(G)
    if exchange == "auto":
        exchange = "peer" if cols_local.is_cuda else "nccl"
    c_loc, n = cols_local.shape
    n_log = _log2(n)
    if c_loc * G != ncols_total:
        raise ValueError("columns must be split evenly: %d local x %d ranks != %d" % (c_loc, G, ncols_total))
    if glog > cap_height:
        raise ValueError("world size %d exceeds the number of cap subtrees 2^%d" % (G, cap_height))
    N = n << rate_bits
    n_loc = N >> glog                      # leaves per rank
    ncap_loc = (1 << cap_height) >> glog   # cap entries per rank
    ndig_loc = 2 * (n_loc - ncap_loc)      # digests per rank
    sc = scratch if scratch is not None else {}

    def buf(name, shape):
        t = sc.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            t = engine.empty(shape)
            sc[name] = t
        return t

    tm = _StageTimer(scratch)
    tm.mark("start")
    # 1. per-column work on the local column shard
    coeffs = buf("coeffs", (c_loc, n))
    if from_coeffs:
        engine.canonical_copy(cols_local, coeffs)
    else:
        engine.intt(cols_local, coeffs)
    def to_host(dst, src):
        # dst <- src on the copy stream, ordered after everything queued on the compute stream so far
        if host_out is None or dst is None:
            return
        ev = torch.cuda.Event()
        ev.record()
        host_out.copy_stream.wait_event(ev)
        with torch.cuda.stream(host_out.copy_stream):
            dst.copy_(src, non_blocking=True)

    to_host(host_out.coeffs if host_out else None, coeffs)
    if exchange == "peer" and G > 1:
        # 2+3 fused: every rank's LDE kernel writes block s of its output into rank s's receive buffer
        ex = sc.get("peer_exchange")
        if ex is None or ex.shape != (G, c_loc, n_loc):
            ex = PeerExchange(G, c_loc, n_loc, group)
            sc["peer_exchange"] = ex
        tm.mark("intt")
        ex.barrier()   # peers are done reading what the previous step put into my buffer
        tm.mark("barrier1")
        # rank g stores to rank g first, then g+1, ...: at any moment the ranks target different peers
        # the four-step intermediate lives in a buffer that persists across calls (same shape as the NCCL path's
        # send buffer): allocating its 8*c_loc*N bytes from the pool on every call stalled the first step after idle
        mid = buf("send", (G, c_loc, n_loc))
        engine.coset_lde_peer(coeffs, ex.shard_ptrs, n_loc, rate_bits, g, scratch=mid)
        tm.mark("lde_peer")
        ex.barrier()   # every block of my receive buffer has landed
        tm.mark("barrier2")
        recv = ex.recv
    else:
        # 2. LDE written leaf-ordered and blocked by destination rank: send[s] = (c_loc, n_loc) block for rank s
        send = buf("send", (G, c_loc, n_loc))
        tm.mark("intt")
        engine.coset_lde(coeffs, send, rate_bits, glog)
        tm.mark("lde")
        # 3. the one exchange of the path
        if G > 1:
            recv = buf("recv", (G, c_loc, n_loc))
            dist.all_to_all_single(recv, send, group=group)
            tm.mark("all_to_all")
        else:
            recv = send
    # recv[s][j] is column s*c_loc + j restricted to my leaves: a (ncols_total, n_loc) column-major LDE
    lde_rows = recv.view(ncols_total, n_loc)
    # 4. my leaves, my subtrees
    leaves = buf("leaves", (n_loc, ncols_total)) if want_leaves else None
    digests = buf("digests", (max(ndig_loc, 1), 4))
    cap_local = buf("cap_local", (ncap_loc, 4))
    nchunks = host_out.chunks if (host_out is not None and host_out.leaves is not None and leaves is not None
                                  and n_loc >= (1 << 16)) else 1
    if nchunks > 1:
        # block j of the rows goes to the host while block j+1 is hashed
        for j in range(nchunks):
            lb, le = j * (n_loc // nchunks), (j + 1) * (n_loc // nchunks)
            engine.merkle_colmajor_leaves(lde_rows, cap_height - glog, hash_kind, lb, le, leaves, digests, cap_local)
            to_host(host_out.leaves[lb:le], leaves[lb:le])
        engine.merkle_levels(n_loc, cap_height - glog, hash_kind, digests, cap_local)
    else:
        engine.merkle_colmajor(lde_rows, cap_height - glog, hash_kind, leaves, digests, cap_local)
        if host_out is not None and leaves is not None:
            to_host(host_out.leaves, leaves)
    to_host(host_out.digests if host_out else None, digests[:ndig_loc])
    tm.mark("merkle")
    # 5. everyone gets the whole cap
    cap = buf("cap", (1 << cap_height, 4))
    if G > 1:
        dist.all_gather_into_tensor(cap, cap_local, group=group)
    else:
        cap.copy_(cap_local)
    tm.mark("cap")
    to_host(host_out.cap if host_out else None, cap)
    return ShardedBatch(coeffs, leaves, digests[:ndig_loc], cap, g, G)
