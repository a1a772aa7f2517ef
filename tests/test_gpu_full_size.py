"""BASELINE-size checks through size-independent properties (the oracle cannot redo 2^20 x 256 in seconds):
the wide batch is committed device-resident, then sampled rows / columns / proofs are checked against the
oracle's definitions."""
import numpy as np
import pytest

from util import P, field_elems

pytestmark = pytest.mark.gpu


def test_wide_batch_2_20_x_256_properties(oracle):
    import torch

    import mapreduce_plonky2_b200 as G
    from mapreduce_plonky2_b200 import device as D

    if torch.cuda.get_device_properties(0).total_memory < 60 * 2**30:
        pytest.skip("needs ~40 GB of device memory")
    G.init(0)
    D.bind_current_device()
    n_log, ncols, r, cap_h, kind = 20, 256, 3, 4, 0
    n, N = 1 << n_log, 1 << (n_log + r)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(0x6D7033)
    cols = torch.randint(0, 1 << 62, (ncols, n), dtype=torch.int64, device="cuda", generator=gen)
    bufs = D.CommitBuffers(ncols, n_log, r, cap_h, True)
    D.commit_resident(cols, bufs, kind, False)
    torch.cuda.synchronize()
    cap = bufs.cap.cpu().numpy().view(np.uint64)

    # (1) coefficients of sampled columns == oracle ifft of the same column
    for c in (0, 137, 255):
        col = cols[c].cpu().numpy().view(np.uint64)
        assert np.array_equal(bufs.coeffs[c].cpu().numpy().view(np.uint64), oracle.ifft(col)), c

    # (2) sampled leaf rows are the polynomials evaluated at 7 * w_N^bitrev(L)  (A.3), by Horner on the CPU
    wN = oracle.root_of_unity(n_log + r)
    sample_cols = [0, 1, 128, 255]
    coeffs_h = {c: bufs.coeffs[c].cpu().numpy().view(np.uint64) for c in sample_cols}
    rows = [0, 1, 5, N // 2 + 3, N - 1, 0x5A5A5A]
    for L in rows:
        row = bufs.leaves[L].cpu().numpy().view(np.uint64)
        x = oracle.gl_mul(7, oracle.gl_pow(wN, G.reverse_bits(L, n_log + r)))
        for c in sample_cols:
            assert int(row[c]) == int(oracle.eval_naive(coeffs_h[c], x, 1, 1)[0]), (L, c)
        # the column-major LDE and the row-major leaves hold the same values
        assert np.array_equal(bufs.lde[:, L].cpu().numpy().view(np.uint64), row)

    # (3) Merkle proofs read from the device digests re-hash (oracle hashing) to the cap
    h = n_log + r - cap_h
    per = 2 * ((1 << h) - 1)
    dig = bufs.digests
    for L in rows:
        sub, pair = L >> h, L & ((1 << h) - 1)
        sib = []
        for i in range(h):
            parity, pair = pair & 1, pair >> 1
            q = (pair << (i + 1)) + (1 << i) - 1
            sib.append(dig[sub * per + 2 * q + (1 - parity)].cpu().numpy().view(np.uint64))
        cap_idx, root = oracle.merkle_verify(bufs.leaves[L].cpu().numpy().view(np.uint64), L, np.stack(sib), kind)
        assert cap_idx == sub and np.array_equal(root, cap[cap_idx]), L

    # (4) every digest is canonical
    top = bufs.digests.cpu().numpy().view(np.uint64)
    assert (top < np.uint64(P)).all()


def test_linearity_at_config1_size(oracle):
    """commit(a) + commit(b) == commit(a + b) leaf by leaf (mod p), 2^14 x 135."""
    import mapreduce_plonky2_b200 as G

    G.init(0)
    a = field_elems(1, (135, 1 << 14))
    b = field_elems(2, (135, 1 << 14))
    s = ((a.astype(object) + b.astype(object)) % P).astype(np.uint64)
    la = G.PolynomialBatch.from_values(a, 3, False, 4, hash_kind=1).merkle_tree.leaves
    lb = G.PolynomialBatch.from_values(b, 3, False, 4, hash_kind=1).merkle_tree.leaves
    ls = G.PolynomialBatch.from_values(s, 3, False, 4, hash_kind=1).merkle_tree.leaves
    sl = slice(0, 1 << 17, 257)
    assert np.array_equal(((la[sl].astype(object) + lb[sl].astype(object)) % P).astype(np.uint64), ls[sl])
