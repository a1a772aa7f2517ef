"""GPU parity tests: the CUDA path, called through the C ABI (ctypes mirror in
``mapreduce_plonky2_b200.plonky2``), against the CPU oracle and the committed golden fixtures.
Bit-exact (integer work): every comparison is ``np.array_equal``.

Run on the B200 box:  python -m pytest tests -m gpu -x -q
"""
import json
import os

import numpy as np
import pytest

from util import GOLDEN_DIR, P, field_elems, hexlist, unhex

pytestmark = pytest.mark.gpu

KATS = json.load(open(os.path.join(GOLDEN_DIR, "kats.json")))
COMMITS = json.load(open(os.path.join(GOLDEN_DIR, "commit_small.json")))["cases"]
MERKLE = json.load(open(os.path.join(GOLDEN_DIR, "merkle_small.json")))
CAPS = json.load(open(os.path.join(GOLDEN_DIR, "config1_caps.json")))["cases"]


@pytest.fixture(scope="module")
def G():
    import mapreduce_plonky2_b200 as g

    g.init(0)  # raises Mp2GpuError if the extension or the GPU is missing -- no fallback
    return g


def u2(rows):
    return np.array([[int(x, 16) for x in r] for r in rows], dtype=np.uint64)


# ---------------------------------------------------------------- permutations / sponge
def test_permutation_published_kats(G):
    ins = {"zeros": np.zeros(12, dtype=np.uint64), "iota": np.arange(12, dtype=np.uint64),
           "neg_one": np.full(12, P - 1, dtype=np.uint64)}
    for kat in KATS["poseidon_perm"]:
        assert np.array_equal(G.permute(ins[kat["in"]], G.POSEIDON), unhex(kat["out"]))
    for kat in KATS["poseidon2_perm"]:
        assert np.array_equal(G.permute(ins[kat["in"]], G.POSEIDON2), unhex(kat["out"]))


@pytest.mark.parametrize("kind", [0, 1])
def test_permutation_random_and_noncanonical(G, oracle, kind):
    states = field_elems(0xABC0 + kind, (4096, 12), canonical=False)  # any u64, incl. >= p
    states[0] = 2**64 - 1
    states[1] = P
    got = G.permute(states, kind)
    for i in list(range(64)) + [4095]:
        assert np.array_equal(got[i], oracle.permute(states[i], kind)), i
    assert (got < np.uint64(P)).all()


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("length", [1, 4, 5, 7, 8, 9, 16, 17, 32, 135, 256])
def test_hash_no_pad_and_noop(G, oracle, kind, length):
    x = field_elems(77 + length, (33, length), canonical=False)
    got = G.hash_no_pad_batch(x, kind)
    for i in (0, 1, 32):
        assert np.array_equal(got[i], oracle.hash_no_pad(x[i], kind))
        assert np.array_equal(G.hash_or_noop(x[i], kind), oracle.hash_or_noop(x[i], kind))


@pytest.mark.parametrize("kind", [0, 1])
def test_two_to_one_hash_pad_empty(G, oracle, kind):
    a, b = field_elems(5, (100, 4)), field_elems(6, (100, 4))
    got = G.two_to_one_batch(a, b, kind)
    for i in (0, 50, 99):
        assert np.array_equal(got[i], oracle.two_to_one(a[i], b[i], kind))
    assert hexlist(G.hash_pad([], kind)) == MERKLE["misc"]["hash_pad_empty"][str(kind)]
    assert not G.hash_no_pad([], kind).any()  # mp2-common/src/poseidon.rs:49-51


# ---------------------------------------------------------------- MerkleTree::new
@pytest.mark.parametrize("idx", range(len(MERKLE["cases"])))
def test_merkle_tree_golden(G, idx):
    case = MERKLE["cases"][idx]
    kind, nl, cap = case["hash_kind"], case["nleaves"], case["cap_height"]
    rows = [np.array([int(x, 16) for x in r], dtype=np.uint64) for r in case["leaves"]]
    leaves = np.stack(rows) if isinstance(case["leaf_len"], int) else rows  # ragged circuit-set shape
    mt = G.MerkleTree.new(leaves, cap, kind)
    if case["digests"]:
        assert np.array_equal(mt.digests, u2(case["digests"]))
    assert np.array_equal(mt.cap.hashes, u2(case["cap"]))
    for i, sib in case["proofs"].items():
        proof = mt.prove(int(i))
        want = u2(sib) if sib else np.zeros((0, 4), dtype=np.uint64)
        assert np.array_equal(proof.siblings, want)
        G.verify_merkle_proof_to_cap(rows[int(i)], int(i), mt.cap, proof, kind)


def test_merkle_tree_panics_like_plonky2(G):
    leaves = field_elems(1, (8, 5))
    with pytest.raises(G.Mp2GpuError, match="cap_height"):
        G.MerkleTree.new(leaves, 4)
    with pytest.raises(G.Mp2GpuError, match="power of two"):
        G.MerkleTree.new(leaves[:6], 0)
    with pytest.raises(G.Mp2GpuError, match="power of two"):
        G.MerkleTree.new(leaves[:3], 0)
    mt = G.MerkleTree.new(leaves, 3)  # tree is all cap
    assert mt.digests.size == 0 and len(mt.cap) == 8
    assert len(mt.prove(5)) == 0
    with pytest.raises(G.Mp2GpuError):
        G.verify_merkle_proof_to_cap(leaves[1], 2, mt.cap, mt.prove(2))  # wrong leaf -> "Invalid Merkle proof."


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("shape", [(1 << 12, 32, 4), (1 << 9, 32, 4), (1 << 5, 32, 4), (1 << 13, 7, 0), (2, 135, 1),
                                   (1 << 10, 4, 2), (1 << 10, 1, 0), (1 << 15, 8, 15)])
def test_merkle_tree_vs_oracle(G, oracle, kind, shape):
    """Incl. the FRI commit-phase shapes: leaves of 16*D = 32 elements, cap 4."""
    nl, ll, cap = shape
    leaves = field_elems(nl * 31 + ll, (nl, ll), canonical=False)
    mt = G.MerkleTree.new(leaves, cap, kind)
    d, c = oracle.merkle_new(leaves, cap, kind)
    assert np.array_equal(mt.cap.hashes, c)
    assert np.array_equal(mt.digests, d)
    for i in (0, nl // 2 + 1 if nl > 2 else 1, nl - 1):
        assert np.array_equal(mt.prove(i).siblings, oracle.merkle_prove(d, nl, cap, i))


# ---------------------------------------------------------------- PolynomialBatch
@pytest.mark.parametrize("case", COMMITS, ids=lambda c: "k%d_%dx2^%d_r%d_cap%d%s" % (
    c["hash_kind"], c["ncols"], c["log_n"], c["rate_bits"], c["cap_height"], "_coeffs" if c["from_coeffs"] else ""))
def test_polynomial_batch_golden(G, case):
    ctor = G.PolynomialBatch.from_coeffs if case["from_coeffs"] else G.PolynomialBatch.from_values
    pb = ctor(u2(case["cols"]), case["rate_bits"], False, case["cap_height"], None, None, hash_kind=case["hash_kind"])
    assert np.array_equal(pb.polynomials, u2(case["coeffs"]))
    assert np.array_equal(pb.merkle_tree.leaves, u2(case["leaves"]))
    if case["digests"]:
        assert np.array_equal(pb.merkle_tree.digests, u2(case["digests"]))
    assert np.array_equal(pb.merkle_tree.cap.hashes, u2(case["cap"]))
    assert pb.degree_log == case["log_n"] and pb.rate_bits == case["rate_bits"] and pb.blinding is False


def _check_batch(G, oracle, cols, rate_bits, cap, kind, from_coeffs=False):
    ctor = G.PolynomialBatch.from_coeffs if from_coeffs else G.PolynomialBatch.from_values
    pb = ctor(cols, rate_bits, False, cap, hash_kind=kind)
    ref = oracle.commit(cols, rate_bits, cap, kind, from_coeffs)
    assert np.array_equal(pb.polynomials, ref["coeffs"]), "coefficients differ"
    assert np.array_equal(pb.merkle_tree.leaves, ref["leaves"]), "LDE leaves differ"
    assert np.array_equal(pb.merkle_tree.digests, ref["digests"]), "digests differ"
    assert np.array_equal(pb.merkle_tree.cap.hashes, ref["cap"]), "cap differs"
    return pb, ref


@pytest.mark.parametrize("ncols", [1, 3, 4, 5, 8, 9, 20, 135])
@pytest.mark.parametrize("log_n", [0, 1, 2, 5, 9, 12])
def test_polynomial_batch_shapes(G, oracle, ncols, log_n):
    cols = field_elems(1000 * ncols + log_n, (ncols, 1 << log_n), canonical=(ncols % 2 == 0))
    kind = (ncols + log_n) % 2
    for cap in sorted({0, min(4, log_n + 3), log_n + 3}):
        _check_batch(G, oracle, cols, 3, cap, kind)


@pytest.mark.parametrize("rate_bits,log_n,ncols", [(0, 6, 7), (1, 10, 16), (2, 11, 5), (4, 7, 9), (3, 13, 6), (3, 14, 3)])
def test_polynomial_batch_rates_and_from_coeffs(G, oracle, rate_bits, log_n, ncols):
    cols = field_elems(31 * log_n + rate_bits, (ncols, 1 << log_n))
    _check_batch(G, oracle, cols, rate_bits, min(4, log_n + rate_bits), 0, from_coeffs=False)
    _check_batch(G, oracle, cols, rate_bits, min(4, log_n + rate_bits), 1, from_coeffs=True)


@pytest.mark.parametrize("log_n", [15, 16, 17])
def test_polynomial_batch_two_pass_sizes(G, oracle, log_n):
    """n > 2^14 takes the two-pass (four-step) transform path."""
    cols = field_elems(0x2A55 + log_n, (3, 1 << log_n))
    _check_batch(G, oracle, cols, 3, 4, 1)
    _check_batch(G, oracle, cols[:2], 1, 2, 0, from_coeffs=True)
    if log_n == 15:     # the host entry point finishes the LDE coset by coset for 4..16 cosets: cover 4 and 16 too
        _check_batch(G, oracle, cols[:2], 2, 3, 1)
        _check_batch(G, oracle, cols[:1], 4, 0, 0, from_coeffs=True)


def test_intt_lde_large_degree_tiles(G, oracle):
    """n = 2^21 and 2^23 exercise the four-step split with a != b and b close to the tile size (device stages,
    one column checked against the oracle)."""
    import torch

    from mapreduce_plonky2_b200 import device as D

    D.bind_current_device()
    for log_n in (21, 23):
        col = field_elems(0x7711 + log_n, (1, 1 << log_n))
        vals = torch.from_numpy(col.view(np.int64)).cuda()
        coeffs = torch.empty_like(vals)
        lde = torch.empty((1, 1 << (log_n + 1)), dtype=torch.int64, device="cuda")
        D.intt(vals, coeffs)
        D.coset_lde(coeffs, lde, 1)
        torch.cuda.synchronize()
        c_ref = oracle.ifft(col[0])
        assert np.array_equal(coeffs.cpu().numpy().view(np.uint64)[0], c_ref)
        nat = oracle.coset_lde(c_ref, 1)
        got = lde.cpu().numpy().view(np.uint64)[0]
        idx = np.array([0, 1, 2, 12345, (1 << (log_n + 1)) - 1])
        rev = np.array([G.reverse_bits(int(i), log_n + 1) for i in idx])
        assert np.array_equal(got[idx], nat[rev])


def test_edge_batches_config1_shape(G, oracle):
    """all-zero, all p-1 and column j == j batches (SURVEY.md 8(d) edge inputs)."""
    n = 1 << 10
    for cols in (np.zeros((9, n), dtype=np.uint64), np.full((9, n), P - 1, dtype=np.uint64),
                 np.repeat(np.arange(9, dtype=np.uint64)[:, None], n, axis=1)):
        _check_batch(G, oracle, cols, 3, 4, 0)


@pytest.mark.parametrize("case", CAPS, ids=lambda c: "%s_k%d" % (c["name"], c["hash_kind"]))
def test_config1_full_size_caps(G, case):
    """BASELINE config 1 (2^14 x 135, r=3, cap 4) and the two sibling commitments of every prove(),
    against caps frozen from the CPU oracle."""
    cols = field_elems(case["seed"], (case["ncols"], 1 << case["log_n"]))
    ctor = G.PolynomialBatch.from_coeffs if case["from_coeffs"] else G.PolynomialBatch.from_values
    pb = ctor(cols, case["rate_bits"], False, case["cap_height"], hash_kind=case["hash_kind"])
    assert hexlist(pb.merkle_tree.cap.hashes) == case["cap"]
    assert hexlist(np.bitwise_xor.reduce(pb.merkle_tree.digests, axis=0)) == case["digest_xor"]
    assert hexlist(np.bitwise_xor.reduce(pb.polynomials, axis=1)[:8]) == case["coeff_xor"]


def test_config1_full_size_vs_oracle_and_properties(G, oracle):
    cols = field_elems(0x6D7032, (135, 1 << 14))
    pb, ref = _check_batch(G, oracle, cols, 3, 4, 0)
    N = 1 << 17
    # size-independent properties: every sampled Merkle proof re-hashes to the cap ...
    for i in (0, 1, 12345, N - 1):
        G.verify_merkle_proof_to_cap(pb.merkle_tree.get(i), i, pb.merkle_tree.cap, pb.merkle_tree.prove(i), 0)
    # ... the LDE restricted to ... is linear: commit(a + b) leaves == leaves(a) + leaves(b) mod p
    a, b = cols[:2], cols[2:4]
    s = ((a.astype(object) + b.astype(object)) % P).astype(np.uint64)
    la = G.PolynomialBatch.from_values(a, 3, False, 4, hash_kind=0).merkle_tree.leaves
    lb = G.PolynomialBatch.from_values(b, 3, False, 4, hash_kind=0).merkle_tree.leaves
    ls = G.PolynomialBatch.from_values(s, 3, False, 4, hash_kind=0).merkle_tree.leaves
    assert np.array_equal(((la.astype(object) + lb.astype(object)) % P).astype(np.uint64), ls)
    # get_lde_values(index, step): rate_bits-strided rows are the original values on the shifted coset...
    # row bitrev(i * 8) is P(7 * w_n^i); check it against the oracle's leaves
    for i in (0, 1, 777):
        assert np.array_equal(pb.get_lde_values(i, 8), ref["leaves"][G.reverse_bits(i * 8, 17)])


def test_device_resident_handle(G, oracle):
    cols = field_elems(99, (20, 1 << 10))
    pb = G.PolynomialBatch.from_values(cols, 3, False, 4, hash_kind=1, keep_on_device=True, fetch_leaves=False)
    ref = oracle.commit(cols, 3, 4, 1)
    rows = [0, 5, 8191, 4097]
    assert np.array_equal(pb.fetch_rows(rows), ref["leaves"][rows])
    assert np.array_equal(pb.get_lde_values(3, 8), ref["leaves"][G.reverse_bits(24, 13)])
    for i in rows:
        assert np.array_equal(pb.prove_on_device(i).siblings, oracle.merkle_prove(ref["digests"], 1 << 13, 4, i))
    with pytest.raises(G.Mp2GpuError):
        pb.fetch_rows([1 << 13])
    # query-phase openings: 28 indices at once, rows + proofs gathered on the device
    q = [int(x) for x in field_elems(3, (28,)) % np.uint64(1 << 13)]
    got_rows, got_sib = pb.open(q)
    assert np.array_equal(got_rows, ref["leaves"][q])
    for j, i in enumerate(q):
        assert np.array_equal(got_sib[j], oracle.merkle_prove(ref["digests"], 1 << 13, 4, i))
    pb.free()


def test_bad_arguments(G):
    cols = field_elems(1, (3, 8))
    with pytest.raises(G.Mp2GpuError, match="cap_height"):
        G.PolynomialBatch.from_values(cols, 1, False, 5)
    with pytest.raises(G.Mp2GpuError, match="blinding"):
        G.PolynomialBatch.from_values(cols, 1, True, 1)
    with pytest.raises(G.Mp2GpuError, match="power of two"):
        G.PolynomialBatch.from_values(cols[:, :6], 1, False, 1)
    with pytest.raises(G.Mp2GpuError, match="hash_kind"):
        G.PolynomialBatch.from_values(cols, 1, False, 1, hash_kind=7)


def test_empty_and_degenerate_inputs(G, oracle):
    """Empty leaves (hash_or_noop of [] is the zero digest), single-leaf trees, one-row polynomials."""
    empty = np.zeros((8, 0), dtype=np.uint64)
    mt = G.MerkleTree.new(empty, 1, 0)
    d, c = oracle.merkle_new(np.zeros((8, 1), dtype=np.uint64), 1, 0)   # [0] and [] hash to the same no-op digest
    assert np.array_equal(mt.digests, d) and np.array_equal(mt.cap.hashes, c)
    one = G.MerkleTree.new(field_elems(4, (1, 9)), 0, 1)                 # a single leaf: the cap is its hash
    assert one.digests.size == 0 and np.array_equal(one.cap.hashes[0], oracle.hash_no_pad(one.leaves[0], 1))
    assert len(one.prove(0)) == 0
    with pytest.raises(G.Mp2GpuError, match="non-empty|no polynomials|power of two"):
        G.PolynomialBatch.from_values(np.zeros((0, 8), dtype=np.uint64), 1, False, 0)
    # degree-0 polynomials (n = 1): the LDE is constant, every leaf row equals the value row
    cols = field_elems(8, (6, 1))
    pb, ref = _check_batch(G, oracle, cols, 3, 2, 0)
    assert (pb.merkle_tree.leaves == cols[:, 0]).all()
    assert G.hash_no_pad_batch(np.zeros((0, 5), dtype=np.uint64), 0).shape == (0, 4)


def test_concurrent_callers(G, oracle):
    """prove() is called from several threads at once (SURVEY.md 8(b) Threading)."""
    import threading

    cols = [field_elems(500 + t, (12, 1 << 9)) for t in range(4)]
    out = [None] * 4

    def work(t):
        G.init(0)
        out[t] = G.PolynomialBatch.from_values(cols[t], 3, False, 4, hash_kind=t % 2).merkle_tree.cap.hashes

    th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    for t in range(4):
        assert np.array_equal(out[t], oracle.commit(cols[t], 3, 4, t % 2, want_leaves=False)["cap"])


@pytest.mark.parametrize("log_n,nshards", [(10, 8), (12, 2), (15, 8), (16, 4)])
def test_peer_store_lde_equals_sharded_lde_on_one_gpu(G, log_n, nshards):
    """mp2gpu_dev_coset_lde_peer with all "peer" buffers on this GPU: for every first_shard rotation the shards
    land exactly where the shard_log layout of mp2gpu_dev_coset_lde puts them (single- and two-pass sizes)."""
    import torch

    from mapreduce_plonky2_b200 import device as D

    D.bind_current_device()
    ncols, r = 3, 3
    n, N = 1 << log_n, 1 << (log_n + r)
    n_loc = N // nshards
    coeffs = torch.from_numpy(field_elems(0x9EE7 + log_n, (ncols, n)).view(np.int64)).cuda()
    want = torch.empty((nshards, ncols, n_loc), dtype=torch.int64, device="cuda")
    D.coset_lde(coeffs, want, r, nshards.bit_length() - 1)
    for first in range(nshards):
        got = torch.zeros((nshards, ncols, n_loc), dtype=torch.int64, device="cuda")
        ptrs = [got[g].data_ptr() for g in range(nshards)]
        D.coset_lde_peer(coeffs, ptrs, n_loc, r, first)
        torch.cuda.synchronize()
        assert torch.equal(got, want), first
    with pytest.raises(Exception, match="first_shard"):
        D.coset_lde_peer(coeffs, ptrs, n_loc, r, nshards)


def test_circuit_digest_on_device(G, oracle):
    """hash_no_pad(cap ‖ hash_pad(&[]) ‖ [degree_bits]) (circuit_set.rs:136-158) through the device hashers."""
    cap = field_elems(0xC1C, (16, 4))
    for kind in (0, 1):
        parts = np.concatenate([cap.reshape(-1), oracle.hash_pad(np.zeros(0, dtype=np.uint64), kind),
                                np.array([12], dtype=np.uint64)])
        assert np.array_equal(G.circuit_digest(cap, 12, kind), oracle.hash_no_pad(parts, kind))


def test_trim_releases_and_everything_still_works(oracle):
    """mp2gpu_trim drops the thread's block cache, the device's tables and the idle pool memory; the next commitment
    rebuilds what it needs and is still bit-exact."""
    import mapreduce_plonky2_b200 as G
    from mapreduce_plonky2_b200 import _lib

    G.init(0)
    cols = field_elems(0x7717, (9, 1 << 9))
    a = G.PolynomialBatch.from_values(cols, 3, False, 4, hash_kind=G.POSEIDON)
    _lib.call("mp2gpu_trim")
    b = G.PolynomialBatch.from_values(cols, 3, False, 4, hash_kind=G.POSEIDON)
    ref = oracle.commit(cols, 3, 4, 0)
    for pb in (a, b):
        assert np.array_equal(pb.merkle_tree.cap.hashes, ref["cap"]) and np.array_equal(pb.merkle_tree.leaves, ref["leaves"])


def test_grouped_levels_with_pipelined_digest_copy(oracle):
    """N = 2^20 leaves: the host entry point builds the cap subtrees in 4 groups and copies each group's digest chunk out
    while the next is built (api.cu commit_host); the result must be the plain tree."""
    import mapreduce_plonky2_b200 as G

    G.init(0)
    cols = field_elems(0x6E0, (5, 1 << 17))
    for kind in (G.POSEIDON, G.POSEIDON2):
        pb = G.PolynomialBatch.from_values(cols, 3, False, 4, hash_kind=kind, fetch_leaves=False)
        ref = oracle.commit(cols, 3, 4, kind, want_leaves=False)
        assert np.array_equal(pb.merkle_tree.digests, ref["digests"])
        assert np.array_equal(pb.merkle_tree.cap.hashes, ref["cap"])
        assert np.array_equal(pb.polynomials, ref["coeffs"])
