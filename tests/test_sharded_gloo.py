"""Multi-rank logic of ``mapreduce_plonky2_b200.sharded`` on CPU: world_size 2 and 4 over gloo, with an
oracle-backed engine standing in for the CUDA kernels (the oracle is the checker here; the product
engine is CudaEngine).  Checks that the column-shard -> all-to-all -> row-shard plan reassembles to
exactly the single-process PolynomialBatch."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleEngine:
    """numpy/oracle implementation of the engine interface (TEST ONLY)."""

    def __init__(self):
        import oracle as O

        self.O = O

    def empty(self, shape):
        return torch.zeros(shape, dtype=torch.int64)

    @staticmethod
    def _u(t):
        return t.numpy().view(np.uint64)

    def intt(self, values, coeffs):
        for c in range(values.shape[0]):
            self._u(coeffs)[c] = self.O.ifft(self._u(values)[c])

    def canonical_copy(self, src, dst):
        v = self._u(src)
        self._u(dst)[...] = np.where(v >= np.uint64(self.O.P), v - np.uint64(self.O.P), v)

    def coset_lde(self, coeffs, lde, rate_bits, shard_log):
        c_loc, n = coeffs.shape
        N = n << rate_bits
        bits = N.bit_length() - 1
        rev = np.array([int(format(i, "0%db" % bits)[::-1], 2) if bits else 0 for i in range(N)])
        out = self._u(lde).reshape(-1)
        G = 1 << shard_log
        n_loc = N // G
        for c in range(c_loc):
            nat = self.O.coset_lde(self._u(coeffs)[c], rate_bits)
            leaf_ordered = nat[rev]  # leaf L holds LDE row bitrev(L)
            for s in range(G):
                base = (s * c_loc + c) * n_loc
                out[base:base + n_loc] = leaf_ordered[s * n_loc:(s + 1) * n_loc]

    def merkle_colmajor(self, lde, cap_height, hash_kind, leaves, digests, cap):
        rows = np.ascontiguousarray(self._u(lde).T)
        d, c = self.O.merkle_new(rows, cap_height, hash_kind, nthreads=1)
        if leaves is not None:
            self._u(leaves)[...] = rows
        if d.size:
            self._u(digests)[:d.shape[0]] = d
        self._u(cap)[...] = c


def _worker(rank, world, port, ncols, n_log, rate_bits, cap_height, kind, from_coeffs, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from util import field_elems

    from mapreduce_plonky2_b200.sharded import commit_sharded

    cols = field_elems(0x5EED, (ncols, 1 << n_log))
    c_loc = ncols // world
    mine = torch.from_numpy(cols[rank * c_loc:(rank + 1) * c_loc].view(np.int64).copy())
    res = commit_sharded(mine, ncols, rate_bits, cap_height, kind, OracleEngine(), from_coeffs=from_coeffs)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), coeffs=res.coeffs.numpy().view(np.uint64),
             leaves=res.leaves.numpy().view(np.uint64), digests=res.digests.numpy().view(np.uint64),
             cap=res.cap.numpy().view(np.uint64))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,ncols,n_log,rate_bits,cap_height,kind,from_coeffs", [
    (2, 6, 4, 3, 4, 0, False),
    (2, 10, 3, 1, 1, 1, True),   # G == number of cap subtrees: one subtree per rank
    (4, 8, 3, 2, 2, 1, False),
    (4, 12, 5, 3, 4, 0, False),
])
def test_sharded_equals_single_process(tmp_path, oracle, world, ncols, n_log, rate_bits, cap_height, kind, from_coeffs):
    from util import field_elems

    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ncols, n_log, rate_bits, cap_height, kind, from_coeffs, str(tmp_path)),
             nprocs=world, join=True)
    cols = field_elems(0x5EED, (ncols, 1 << n_log))
    ref = oracle.commit(cols, rate_bits, cap_height, kind, from_coeffs)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    assert np.array_equal(np.concatenate([p["coeffs"] for p in parts]), ref["coeffs"])
    assert np.array_equal(np.concatenate([p["leaves"] for p in parts]), ref["leaves"])
    assert np.array_equal(np.concatenate([p["digests"] for p in parts]), ref["digests"])
    for p in parts:
        assert np.array_equal(p["cap"], ref["cap"])  # every rank holds the whole cap


def test_sharded_rejects_bad_world(oracle):
    from mapreduce_plonky2_b200 import sharded

    with pytest.raises(ValueError):
        sharded._log2(6)
