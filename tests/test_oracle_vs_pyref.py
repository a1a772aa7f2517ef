"""C oracle vs. the independent pure-Python restatement and the committed golden fixtures."""
import json
import os
import random

import numpy as np
import pytest

import pyref as R
from util import GOLDEN_DIR, P, field_elems, hexlist, unhex

COMMITS = json.load(open(os.path.join(GOLDEN_DIR, "commit_small.json")))["cases"]
MERKLE = json.load(open(os.path.join(GOLDEN_DIR, "merkle_small.json")))
CAPS = json.load(open(os.path.join(GOLDEN_DIR, "config1_caps.json")))["cases"]


def u2(rows):
    return np.array([[int(x, 16) for x in r] for r in rows], dtype=np.uint64)


@pytest.mark.parametrize("case", COMMITS, ids=lambda c: "k%d_%dx2^%d_r%d_cap%d%s" % (
    c["hash_kind"], c["ncols"], c["log_n"], c["rate_bits"], c["cap_height"], "_coeffs" if c["from_coeffs"] else ""))
def test_commit_matches_golden(oracle, case):
    out = oracle.commit(u2(case["cols"]), case["rate_bits"], case["cap_height"], case["hash_kind"],
                        case["from_coeffs"])
    assert np.array_equal(out["coeffs"], u2(case["coeffs"]))
    assert np.array_equal(out["leaves"], u2(case["leaves"]))
    if case["digests"]:
        assert np.array_equal(out["digests"], u2(case["digests"]))
    else:
        assert out["digests"].size == 0
    assert np.array_equal(out["cap"], u2(case["cap"]))


@pytest.mark.parametrize("idx", range(len(MERKLE["cases"])))
def test_merkle_matches_golden(oracle, idx):
    case = MERKLE["cases"][idx]
    kind, nl, cap = case["hash_kind"], case["nleaves"], case["cap_height"]
    rows = [[int(x, 16) for x in r] for r in case["leaves"]]
    if isinstance(case["leaf_len"], int):
        digests, capv = oracle.merkle_new(np.array(rows, dtype=np.uint64), cap, kind)
    else:
        # ragged circuit-set leaves (4-element digests + [0] padding): hash_or_noop zero-pads both to
        # the same 4-element no-op digest, so padding the short leaves with zeros is equivalent
        rows4 = [r + [0] * (4 - len(r)) for r in rows]
        digests, capv = oracle.merkle_new(np.array(rows4, dtype=np.uint64), cap, kind)
    if case["digests"]:
        assert np.array_equal(digests, u2(case["digests"]))
    assert np.array_equal(capv, u2(case["cap"]))
    for i, sib in case["proofs"].items():
        got = oracle.merkle_prove(digests, nl, cap, int(i))
        want = u2(sib) if sib else np.zeros((0, 4), dtype=np.uint64)
        assert np.array_equal(got, want)
        leaf = rows[int(i)]
        cap_idx, root = oracle.merkle_verify(np.array(leaf, dtype=np.uint64), int(i), got, kind)
        assert np.array_equal(root, capv[cap_idx])


def test_hash_pad_and_empty(oracle):
    for k in (0, 1):
        assert hexlist(oracle.hash_pad([], k)) == MERKLE["misc"]["hash_pad_empty"][str(k)]
    assert hexlist(oracle.hash_no_pad([], 0)) == MERKLE["misc"]["hash_no_pad_empty"] == ["0" * 16] * 4


def test_sponge_overwrite_semantics(oracle):
    """A short last chunk overwrites only state[0..len) (mp2-common/src/poseidon.rs:151-171)."""
    rng = random.Random(3)
    for k in (0, 1):
        for ln in (1, 4, 5, 7, 8, 9, 15, 16, 17, 135):
            x = [rng.randrange(P) for _ in range(ln)]
            assert [int(v) for v in oracle.hash_no_pad(x, k)] == R.hash_no_pad(x, k)
            assert [int(v) for v in oracle.hash_or_noop(x, k)] == R.hash_or_noop(x, k)
        a = [rng.randrange(P) for _ in range(4)]
        b = [rng.randrange(P) for _ in range(4)]
        assert [int(v) for v in oracle.two_to_one(a, b, k)] == R.two_to_one(a, b, k)
        assert np.array_equal(oracle.two_to_one(a, b, k), oracle.hash_no_pad(a + b, k))


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5, 7])
def test_transforms_by_definition(oracle, log_n):
    n = 1 << log_n
    v = field_elems(100 + log_n, (n,))
    coeffs = oracle.ifft(v)
    assert [int(x) for x in coeffs] == R.ifft([int(x) for x in v])
    assert np.array_equal(oracle.fft(coeffs), v)
    for r in (0, 1, 3):
        lde = oracle.coset_lde(coeffs, r)
        w = oracle.root_of_unity(log_n + r)
        assert np.array_equal(lde, oracle.eval_naive(coeffs, 7, w, n << r))
        if log_n <= 5:
            assert [int(x) for x in lde] == R.coset_lde([int(x) for x in coeffs], r)
    # shift-1 LDE restricted to the subgroup (every 2^r-th point) gives the values back
    lde1 = oracle.coset_lde(coeffs, 2, shift=1)
    assert np.array_equal(lde1[::4], v)


def test_merkle_new_rejects_like_plonky2(oracle):
    leaves = field_elems(1, (8, 5))
    with pytest.raises(ValueError):
        oracle.merkle_new(leaves, 4)           # cap_height > log2(len)
    with pytest.raises(ValueError):
        oracle.merkle_new(leaves[:6], 0)       # not a power of two (cases (3,0),(6,0) of the serde rstest)
    d, cap = oracle.merkle_new(leaves, 3)      # tree is all cap
    assert d.size == 0 and cap.shape == (8, 4)
    assert np.array_equal(cap[5], oracle.hash_or_noop(leaves[5]))


def test_closed_form_prove_vs_explicit_tree(oracle):
    rng = random.Random(11)
    for log_n in range(0, 7):
        for cap in range(0, log_n + 1):
            n = 1 << log_n
            leaves = [[rng.randrange(P) for _ in range(6)] for _ in range(n)]
            _, capv, trees = R.merkle_new(leaves, cap, 0)
            digests, cap_c = oracle.merkle_new(np.array(leaves, dtype=np.uint64), cap, 0)
            assert cap_c.tolist() == capv
            for i in sorted({0, n // 2, n - 1, rng.randrange(n)}):
                want = R.merkle_prove_from_tree(trees, n, cap, i)
                got = oracle.merkle_prove(digests, n, cap, i)
                assert got.tolist() == want


def test_thread_count_does_not_change_results(oracle):
    cols = field_elems(5, (7, 64))
    a = oracle.commit(cols, 3, 2, 0, nthreads=1)
    b = oracle.commit(cols, 3, 2, 0, nthreads=5)
    for k in a:
        assert np.array_equal(a[k], b[k])


@pytest.mark.parametrize("case", [c for c in CAPS if c["ncols"] <= 20], ids=lambda c: "%s_k%d" % (c["name"], c["hash_kind"]))
def test_config1_sibling_caps_regression(oracle, case):
    cols = field_elems(case["seed"], (case["ncols"], 1 << case["log_n"]))
    res = oracle.commit(cols, case["rate_bits"], case["cap_height"], case["hash_kind"], case["from_coeffs"],
                        want_leaves=False)
    assert hexlist(res["cap"]) == case["cap"]
    assert hexlist(np.bitwise_xor.reduce(res["digests"], axis=0)) == case["digest_xor"]


def test_fri_commit_phase_oracle_vs_definition(oracle):
    """oracle.fri_committed_trees (FFT-based) == pyref (Horner in GF(p^2), explicit fold), two layers."""
    rng = random.Random(9)
    n_log, r = 6, 3
    m = 1 << (n_log + r)
    coeffs = [(rng.randrange(P), rng.randrange(P)) if j < (1 << n_log) else (0, 0) for j in range(m)]
    arities = [4, 2]
    betas = [(rng.randrange(P), rng.randrange(P)) for _ in arities]
    for kind in (0, 1):
        ref_trees, ref_final = R.fri_committed_trees(coeffs, arities, betas, 2, kind, r)
        c = np.array(coeffs, dtype=np.uint64)
        trees, final = oracle.fri_committed_trees(c, oracle.coset_fft_ext(c, 7), arities,
                                                  np.array(betas, dtype=np.uint64), 2, kind, r)
        for (lv, dg, cap), rt in zip(trees, ref_trees):
            assert lv.tolist() == rt["leaves"] and dg.tolist() == rt["digests"] and cap.tolist() == rt["cap"]
        assert final.tolist() == [list(x) for x in ref_final]
    # the extension's generator of the 2^33 subgroup squares to the base field's 2-adic generator, so
    # FFTs commute with the field inclusion (QuadraticExtension<GoldilocksField>::EXT_POWER_OF_TWO_GENERATOR)
    g = (0, 15659105665374529263)
    assert R.ext_mul(g, g) == (1753635133440165772, 0)


def test_prove_openings_combination_oracle_vs_definition(oracle):
    """oracle.fri_combine (plonky2's loop: reduce_polys_base, divide_by_linear, shift_poly, +=) evaluated at random
    points equals the defining formula sum_i alpha^(k_i) (F_i(x) - F_i(z_i)) / (x - z_i) from pyref."""
    rng = random.Random(0x0FE)
    for n, sizes in ((1, [1]), (2, [2, 1]), (16, [5, 2]), (32, [3, 1, 4])):
        batches = [((rng.randrange(P), rng.randrange(P)), [[rng.randrange(P) for _ in range(n)] for _ in range(c)])
                   for c in sizes]
        alpha = (rng.randrange(P), rng.randrange(P))
        out = oracle.fri_combine([(np.array(z, dtype=np.uint64), [np.array(f, dtype=np.uint64) for f in polys])
                                  for z, polys in batches], np.array(alpha, dtype=np.uint64))
        assert out.shape == (n, 2) and not out[-1].any()
        coeffs = [(int(a), int(b)) for a, b in out]
        for _ in range(3):
            x = (rng.randrange(P), rng.randrange(P))
            assert R.ext_horner(coeffs, x) == R.fri_combined_eval(batches, alpha, x)
    # an extension inverse really is one
    a = (rng.randrange(P), rng.randrange(P))
    assert R.ext_mul(a, R.ext_inv(a)) == (1, 0)


def test_random_shapes_oracle_vs_definition(oracle):
    """Differential check over randomly drawn small shapes (both hashers, from_values and from_coeffs, leaf lengths on
    both sides of the no-op and rate boundaries, non-canonical inputs): the C oracle's whole PolynomialBatch equals
    the pure-Python definition, and every Merkle proof read through the closed-form indices verifies."""
    rng = random.Random(0xD1FF)
    for trial in range(24):
        kind = trial & 1
        ncols = rng.choice([1, 2, 3, 4, 5, 7, 8, 9, 13, 16, 17])
        log_n = rng.randrange(0, 4)
        rate_bits = rng.randrange(0, 3)
        cap = rng.randrange(0, log_n + rate_bits + 1)
        from_coeffs = bool(rng.getrandbits(1))
        n = 1 << log_n
        cols = [[rng.randrange(P) for _ in range(n)] for _ in range(ncols)]
        if trial % 5 == 0:
            cols[0][0] = P + rng.randrange(2**32 - 1)          # non-canonical input
        want = R.commit(cols, rate_bits, cap, kind, from_coeffs)
        got = oracle.commit(np.array(cols, dtype=np.uint64), rate_bits, cap, kind, from_coeffs=from_coeffs)
        shape = (ncols, log_n, rate_bits, cap, kind, from_coeffs)
        assert got["coeffs"].tolist() == want["coeffs"], shape
        assert got["leaves"].tolist() == want["leaves"], shape
        assert got["digests"].tolist() == want["digests"], shape
        assert got["cap"].tolist() == want["cap"], shape
        N = n << rate_bits
        i = rng.randrange(N)
        sib = oracle.merkle_prove(got["digests"], N, cap, i)
        assert R.verify_merkle_proof_to_cap(want["leaves"][i], i, want["cap"], sib.tolist(), kind), shape
