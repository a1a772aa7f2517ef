"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/mp2gpu.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mp2gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mp2gpu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from mapreduce_plonky2_b200 import _lib

    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "libmp2gpu.so does not export %s" % s
    # the ctypes table binds exactly the declared surface
    assert sorted(_lib.SIGNATURES) == syms


def test_header_cites_reference_interfaces():
    text = open(os.path.join(ROOT, "include", "mp2gpu.h")).read()
    for anchor in ("circuit_set.rs:189", "circuit_builder.rs:177,308", "gnark-utils/src/lib.rs:17-52"):
        assert anchor in text


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import numpy as np

    import mapreduce_plonky2_b200 as G

    with pytest.raises(G.Mp2GpuError, match="no usable CUDA device|no CPU fallback"):
        G.PolynomialBatch.from_values(np.zeros((2, 4), dtype=np.uint64), 1, False, 0)
    with pytest.raises(G.Mp2GpuError):
        G.MerkleTree.new(np.zeros((4, 5), dtype=np.uint64), 0)


def test_error_strings_are_freed_and_pure_host_entry_points_work():
    """mp2gpu_merkle_prove is index arithmetic on host memory: usable (and testable) without a GPU."""
    import numpy as np

    import oracle as O
    from mapreduce_plonky2_b200 import plonky2 as P2

    leaves = np.arange(64 * 6, dtype=np.uint64).reshape(64, 6)
    d, cap = O.merkle_new(leaves, 2, 0)
    mt = P2.MerkleTree(leaves, d, P2.MerkleCap(cap), 0)
    for i in (0, 13, 63):
        assert np.array_equal(mt.prove(i).siblings, O.merkle_prove(d, 64, 2, i))
    with pytest.raises(P2.Mp2GpuError, match="out of range"):
        mt.prove(64)
    assert ctypes.c_char_p(P2._lib.load().mp2gpu_version()).value.startswith(b"0.")


@pytest.mark.parametrize("n,cap", [(16, 0), (32, 3), (8, 3)])
def test_merkle_tree_wire_format_round_trip(n, cap):
    """Mirror of the rstest at mp2-common/src/serialization/circuit_data_serialization.rs:344-370 (valid cases):
    write_merkle_tree -> read_merkle_tree is the identity, on an oracle-built tree (host-only code path)."""
    import numpy as np

    import oracle as O
    from mapreduce_plonky2_b200 import plonky2 as P2

    leaves = (np.arange(n * 7, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)).reshape(n, 7) % np.uint64(O.P)
    d, c = O.merkle_new(leaves, cap, 1)
    tree = P2.MerkleTree(leaves, d, P2.MerkleCap(c), 1)
    blob = P2.write_merkle_tree(tree)
    assert len(blob) == 8 + n * (8 + 7 * 8) + 8 + d.shape[0] * 32 + 8 + (1 << cap) * 32
    back = P2.read_merkle_tree(blob, 1)
    assert np.array_equal(back.leaves, leaves) and np.array_equal(back.digests, d)
    assert np.array_equal(back.cap.hashes, c) and back.cap.height() == cap
    assert P2.write_merkle_tree(back) == blob


def test_polynomial_batch_wire_format_round_trip():
    """write_polynomial_batch -> read_polynomial_batch is the identity on an oracle-built batch (host-only code path);
    the byte count follows plonky2's Write::write_polynomial_batch field by field."""
    import numpy as np

    import oracle as O
    from mapreduce_plonky2_b200 import plonky2 as P2
    from util import field_elems

    ncols, n_log, r, cap = 5, 4, 3, 2
    n, N = 1 << n_log, 1 << (n_log + r)
    res = O.commit(field_elems(0xB17E, (ncols, n)), r, cap, 1)
    tree = P2.MerkleTree(res["leaves"], res["digests"], P2.MerkleCap(res["cap"]), 1)
    batch = P2.PolynomialBatch(res["coeffs"], tree, n_log, r, False)
    blob = P2.write_polynomial_batch(batch)
    tree_bytes = 8 + N * (8 + ncols * 8) + 8 + res["digests"].shape[0] * 32 + 8 + (1 << cap) * 32
    assert len(blob) == 8 + ncols * (8 + n * 8) + tree_bytes + 8 + 8 + 1
    back = P2.read_polynomial_batch(blob, 1)
    assert np.array_equal(back.polynomials, res["coeffs"]) and np.array_equal(back.merkle_tree.leaves, res["leaves"])
    assert np.array_equal(back.merkle_tree.digests, res["digests"]) and np.array_equal(back.merkle_tree.cap.hashes, res["cap"])
    assert (back.degree_log, back.rate_bits, back.blinding) == (n_log, r, False)
    assert P2.write_polynomial_batch(back) == blob
    with pytest.raises(P2.Mp2GpuError, match="trailing"):
        P2.read_polynomial_batch(blob + b"\0", 1)
