"""FRI commit phase on the GPU (mp2gpu_fri_* through the ctypes mirror) against the CPU oracle's restatement
of plonky2's fri_committed_trees: per-layer leaves, digests and caps, and the final polynomial."""
import numpy as np
import pytest

from util import P, field_elems

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("degree_bits", [6, 9, 12, 13, 14])
def test_fri_committed_trees_matches_oracle(oracle, kind, degree_bits):
    import mapreduce_plonky2_b200 as G
    from mapreduce_plonky2_b200.trace import fri_reduction_arity_bits

    G.init(0)
    rate_bits, cap_height = 3, 4
    n = 1 << degree_bits
    coeffs = field_elems(0xF21 + degree_bits, (n, 2), canonical=(degree_bits % 2 == 0))
    arities = fri_reduction_arity_bits(degree_bits)      # ConstantArityBits(4, 5)
    betas = field_elems(0xBE7A + degree_bits, (len(arities), 2))
    trees, final = G.fri_committed_trees(coeffs, arities, betas, rate_bits, cap_height, kind)
    # oracle: lde(rate_bits) + coset_fft, then the loop
    padded = np.zeros((n << rate_bits, 2), dtype=np.uint64)
    padded[:n] = np.where(coeffs >= np.uint64(P), coeffs - np.uint64(P), coeffs)
    values = oracle.coset_fft_ext(padded, 7)
    ref_trees, ref_final = oracle.fri_committed_trees(padded, values, arities, betas, cap_height, kind, rate_bits)
    assert len(trees) == len(ref_trees) == len(arities)
    for t, (lv, dg, cap) in zip(trees, ref_trees):
        assert np.array_equal(t.leaves, lv)
        assert np.array_equal(t.digests, dg)
        assert np.array_equal(t.cap.hashes, cap)
        assert t.leaves.shape[1] == 32      # 16 * D elements per leaf
    assert np.array_equal(final, ref_final)
    # Merkle proofs of a layer tree verify (query phase reads these)
    t0 = trees[0]
    for i in (0, len(t0.leaves) - 1):
        G.verify_merkle_proof_to_cap(t0.get(i), i, t0.cap, t0.prove(i), kind)


def test_fri_layer_openings(oracle):
    import mapreduce_plonky2_b200 as G

    G.init(0)
    coeffs = field_elems(0x0FE2, (1 << 10, 2))
    ph = G.FriCommitPhase(coeffs, 3, 4, 1)
    ph.commit_layer(4)
    tree = ph.layer(0)
    idx = [0, 5, 511, 300]
    leaves, sib = ph.open_layer(0, idx)
    assert np.array_equal(leaves, tree.leaves[idx])
    for j, i in enumerate(idx):
        assert np.array_equal(sib[j], oracle.merkle_prove(tree.digests, 512, 4, i))
    ph.free()


def test_fri_errors():
    import mapreduce_plonky2_b200 as G

    G.init(0)
    ph = G.FriCommitPhase(np.ones((64, 2), dtype=np.uint64), 3, 4, 0)
    with pytest.raises(G.Mp2GpuError, match="committed layer"):
        ph.fold([1, 2])
    with pytest.raises(G.Mp2GpuError, match="arity"):
        ph.commit_layer(0)
    ph.commit_layer(4)
    ph.fold([3, 4])
    assert ph.finish().shape == (4, 2)
    ph.free()
    with pytest.raises(G.Mp2GpuError):
        G.FriCommitPhase(np.ones((6, 2), dtype=np.uint64), 3, 4, 0)


@pytest.mark.parametrize("kind", [0, 1])
def test_fri_proof_of_work_smallest_witness(oracle, kind):
    import mapreduce_plonky2_b200 as G

    G.init(0)
    for seed, pos, bits in ((1, 0, 8), (2, 3, 12), (3, 7, 16)):
        state = field_elems(0x90 + seed, (12,))
        w = G.fri_proof_of_work(state, pos, bits, kind)
        assert w == oracle.fri_pow(state, pos, bits, kind)
        # the response really has the leading zeros
        st = state.copy()
        st[pos] = w
        resp = int(oracle.permute(st, kind)[7])
        assert resp < 1 << (64 - bits)
