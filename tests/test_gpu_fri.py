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


# ---- prove_openings: the alpha-batched quotient (mp2gpu_fri_begin_openings) ----
def _oracles(G, degree_bits, widths, kind, seed):
    n = 1 << degree_bits
    return [G.PolynomialBatch.from_coeffs(list(field_elems(seed + 7 * k, (w, n))), 3, False, min(4, degree_bits + 3),
                                          hash_kind=kind, keep_on_device=True, fetch_leaves=False)
            for k, w in enumerate(widths)]


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("degree_bits,widths", [(0, (2, 1)), (1, (3, 2)), (5, (3, 4, 2)), (10, (9, 17, 5, 4)),
                                                 (11, (5, 3)), (14, (85, 135, 20, 16))])
def test_prove_openings_final_poly_matches_oracle(oracle, kind, degree_bits, widths):
    """plonky2's instance shape: every polynomial of every oracle at zeta, the first two polynomials of the
    third (or last) oracle -- the Zs -- again at g*zeta."""
    import mapreduce_plonky2_b200 as G

    G.init(0)
    oracles = _oracles(G, degree_bits, widths, kind, 0x09E0 + degree_bits)
    zeta, gzeta, alpha = (field_elems(0x2E7A + i + degree_bits, 2) for i in range(3))
    all_polys = [(o, p) for o, w in enumerate(widths) for p in range(w)]
    z_oracle = min(2, len(widths) - 1)
    zs = [(z_oracle, p) for p in range(min(2, widths[z_oracle]))]
    batches = [(zeta, all_polys), (gzeta, zs)]
    ph = G.FriCommitPhase.from_openings(oracles, batches, alpha, min(4, degree_bits), kind, want_final_poly=True)
    ref = oracle.fri_combine([(z, [oracles[o].polynomials[p] for o, p in polys]) for z, polys in batches], alpha)
    assert np.array_equal(ph.final_poly, ref)
    assert not ph.final_poly[-1].any()            # divide_by_linear drops a degree; push(ZERO) pads it back
    if degree_bits >= 5:
        # the device-resident polynomial feeds the commit phase exactly like host coefficients would
        arity = 4 if degree_bits >= 4 else 1
        cap = ph.commit_layer(arity)
        ph2 = G.FriCommitPhase(ref, 3, min(4, degree_bits), kind)
        assert np.array_equal(cap.hashes, ph2.commit_layer(arity).hashes)
        ph2.free()
    ph.free()
    for o in oracles:
        o.free()


def test_prove_openings_three_batches_and_noncanonical_challenges(oracle):
    """ReducingFactor::shift_poly multiplies what has been accumulated by alpha^(size of the NEW batch): three
    batches of different sizes pin that exponent; challenges >= p are reduced."""
    import mapreduce_plonky2_b200 as G

    G.init(0)
    oracles = _oracles(G, 12, (6, 3), 1, 0x7B3)
    pts = [np.array([P + 5, 3], dtype=np.uint64), field_elems(0x51, 2), field_elems(0x52, 2)]
    alpha = np.array([2**64 - 1, P + 1], dtype=np.uint64)
    batches = [(pts[0], [(0, 0), (0, 1), (0, 2), (1, 2), (0, 5)]), (pts[1], [(1, 0)]), (pts[2], [(0, 3), (1, 1), (0, 3)])]
    ph = G.FriCommitPhase.from_openings(oracles, batches, alpha, 4, 1, want_final_poly=True)
    ref = oracle.fri_combine([(z % np.uint64(P), [oracles[o].polynomials[p] for o, p in polys]) for z, polys in batches],
                             alpha % np.uint64(P))
    assert np.array_equal(ph.final_poly, ref)
    ph.free()
    for o in oracles:
        o.free()


def test_prove_openings_large_degree_evaluation_identity(oracle):
    """n = 2^18 (two levels of the quotient scan): (X - z) * quotient + F(z) == F at a random point, with the
    evaluations done by the oracle's extension Horner -- no O(n) oracle pass over 2^18 x c."""
    import mapreduce_plonky2_b200 as G
    import pyref

    G.init(0)
    degree_bits, w = 18, 3
    oracles = _oracles(G, degree_bits, (w,), 1, 0x18AB)
    z, alpha, x = (tuple(int(v) for v in field_elems(0x90 + i, 2)) for i in range(3))
    ph = G.FriCommitPhase.from_openings(oracles, [(np.array(z, dtype=np.uint64), [(0, p) for p in range(w)])],
                                        np.array(alpha, dtype=np.uint64), 4, 1, want_final_poly=True)
    q = ph.final_poly

    def horner_np(coeffs_ext, pt):      # vectorised-free but O(n) python ints: 2^18 steps is ~1 s
        acc = (0, 0)
        for a, b in zip(coeffs_ext[::-1, 0].tolist(), coeffs_ext[::-1, 1].tolist()):
            acc = pyref.ext_mul(acc, pt)
            acc = ((acc[0] + a) % P, (acc[1] + b) % P)
        return acc

    comp = np.zeros((1 << degree_bits, 2), dtype=object)
    pw = (1, 0)
    for p in range(w):
        col = oracles[0].polynomials[p].astype(object)
        comp[:, 0] = (comp[:, 0] + col * pw[0]) % P
        comp[:, 1] = (comp[:, 1] + col * pw[1]) % P
        pw = pyref.ext_mul(pw, alpha)
    fx, fz, qx = horner_np(comp, x), horner_np(comp, z), horner_np(q.astype(object), x)
    lhs = pyref.ext_mul(qx, ((x[0] - z[0]) % P, (x[1] - z[1]) % P))
    assert ((lhs[0] + fz[0]) % P, (lhs[1] + fz[1]) % P) == fx
    ph.free()
    oracles[0].free()


def test_prove_openings_errors():
    import mapreduce_plonky2_b200 as G

    G.init(0)
    a = _oracles(G, 6, (3,), 0, 1)[0]
    b = _oracles(G, 7, (3,), 0, 2)[0]
    host_only = G.PolynomialBatch.from_coeffs(list(field_elems(3, (2, 64))), 3, False, 4, hash_kind=0)
    z = np.array([1, 2], dtype=np.uint64)
    with pytest.raises(G.Mp2GpuError, match="share degree"):
        G.FriCommitPhase.from_openings([a, b], [(z, [(0, 0)])], z, 4, 0)
    with pytest.raises(G.Mp2GpuError, match="polynomial_index"):
        G.FriCommitPhase.from_openings([a], [(z, [(0, 3)])], z, 4, 0)
    with pytest.raises(G.Mp2GpuError, match="oracle_index"):
        G.FriCommitPhase.from_openings([a], [(z, [(1, 0)])], z, 4, 0)
    with pytest.raises(G.Mp2GpuError, match="empty batch"):
        G.FriCommitPhase.from_openings([a], [(z, [])], z, 4, 0)
    with pytest.raises(G.Mp2GpuError, match="device-resident"):
        G.FriCommitPhase.from_openings([host_only], [(z, [(0, 0)])], z, 4, 0)
    a.free()
    b.free()


# ---- whole FRI prover: openings -> prove_openings -> fri_proof, against the oracle-built proof and the verifier ----
@pytest.mark.parametrize("kind,degree_bits,widths", [(0, 6, (3, 5, 4, 2)), (1, 6, (2, 9, 3)), (1, 10, (3, 2)),
                                                      (1, 12, (5, 7, 3, 2))])
def test_fri_proof_matches_oracle_and_verifies(oracle, kind, degree_bits, widths):
    import fri_ref
    import pyref
    import mapreduce_plonky2_b200 as G
    from mapreduce_plonky2_b200 import fri as GF

    G.init(0)
    n = 1 << degree_bits
    rounds, pow_bits = 5, 6
    coeff_sets = [field_elems(0xF0 + 3 * k + degree_bits, (w, n)) for k, w in enumerate(widths)]
    zeta, gzeta = (tuple(int(v) for v in field_elems(0x5E7A + i, 2)) for i in range(2))
    ref_batches = fri_ref.plonky2_instance(widths, zeta, gzeta)
    commits, ref_openings, ref = fri_ref.oracle_fri_proof(oracle, coeff_sets, ref_batches, degree_bits, kind,
                                                          pow_bits=pow_bits, num_query_rounds=rounds)
    # device side
    oracles = [G.PolynomialBatch.from_coeffs(list(c), 3, False, 4, hash_kind=kind, keep_on_device=True, fetch_leaves=False)
               for c in coeff_sets]
    batches = [GF.FriBatchInfo(np.array(z, dtype=np.uint64), polys) for z, polys in ref_batches]
    openings = GF.open_batches(batches, oracles)
    for got, want in zip(openings, ref_openings):
        assert got.tolist() == [list(v) for v in want]
    ch = GF.Challenger(kind)
    for o in oracles:
        ch.observe_cap(o.merkle_tree.cap)
    for vals in openings:
        ch.observe_extension_elements(vals)
    params = GF.FriConfig(proof_of_work_bits=pow_bits, num_query_rounds=rounds).fri_params(degree_bits)
    assert params.reduction_arity_bits == fri_ref.arity_schedule(degree_bits)
    proof = GF.prove_openings(batches, oracles, ch, params)
    # bit-exact against the proof assembled from the oracle's pieces
    assert len(proof.commit_phase_merkle_caps) == len(ref["caps"])
    for cap, rc in zip(proof.commit_phase_merkle_caps, ref["caps"]):
        assert np.array_equal(cap.hashes, rc)
    assert np.array_equal(proof.final_poly, ref["final_poly"])
    assert proof.pow_witness == ref["pow_witness"]
    assert len(proof.query_round_proofs) == rounds
    for rnd, rr in zip(proof.query_round_proofs, ref["rounds"]):
        for (row, mp), (rrow, rsib) in zip(rnd.initial_trees_proof, rr["initial"]):
            assert np.array_equal(row, rrow) and np.array_equal(mp.siblings, rsib)
        for st, (rev, rsib) in zip(rnd.steps, rr["steps"]):
            assert np.array_equal(st.evals, rev) and np.array_equal(st.merkle_proof.siblings, rsib)
    # GPU proof -> bincode bytes -> parse -> identical proof -> identical bytes (wire.py, mp2-common/src/proof.rs:41-57)
    from mapreduce_plonky2_b200 import wire as W

    data = W.write_fri_proof(proof)
    back = W.read_fri_proof(data)
    assert W.write_fri_proof(back) == data
    assert back.pow_witness == proof.pow_witness and np.array_equal(back.final_poly, proof.final_poly)
    for ra, rb in zip(back.query_round_proofs, proof.query_round_proofs):
        for (ea, pa), (eb, pb) in zip(ra.initial_trees_proof, rb.initial_trees_proof):
            assert np.array_equal(ea, eb) and np.array_equal(pa.siblings, pb.siblings)
        for sa, sb in zip(ra.steps, rb.steps):
            assert np.array_equal(sa.evals, sb.evals) and np.array_equal(sa.merkle_proof.siblings, sb.merkle_proof.siblings)
    # and accepted by the by-definition verifier with its own transcript
    vch = pyref.Challenger(kind)
    caps = [o.merkle_tree.cap.hashes.tolist() for o in oracles]
    fri_ref.transcript_head(vch.observe, caps, [v.tolist() for v in openings])
    p = {"caps": [c.hashes.tolist() for c in proof.commit_phase_merkle_caps], "final_poly": proof.final_poly.tolist(),
         "pow_witness": proof.pow_witness,
         "rounds": [{"initial": [(r.tolist(), m.siblings.tolist()) for r, m in rnd.initial_trees_proof],
                     "steps": [(s.evals.tolist(), s.merkle_proof.siblings.tolist()) for s in rnd.steps]}
                    for rnd in proof.query_round_proofs]}
    pyref.verify_fri_proof(ref_batches, [[tuple(v) for v in vals.tolist()] for vals in openings], caps, p, vch,
                           degree_bits, params.reduction_arity_bits, pow_bits=pow_bits, kind=kind)
    for o in oracles:
        o.free()


def test_batch_eval_matches_horner(oracle):
    """mp2gpu_batch_eval over short, non-multiple-of-256 and long polynomials, base and extension points."""
    import pyref
    import mapreduce_plonky2_b200 as G

    G.init(0)
    for degree_bits, w in ((0, 2), (3, 3), (8, 2), (9, 2), (13, 1)):
        n = 1 << degree_bits
        cols = field_elems(0xE7A1 + degree_bits, (w, n))
        b = G.PolynomialBatch.from_coeffs(list(cols), 3, False, min(4, degree_bits + 3), hash_kind=1,
                                          keep_on_device=True, fetch_leaves=False)
        pts = np.array([[5, 0], [P + 2, 2**64 - 1], field_elems(0x99, 2).tolist()], dtype=np.uint64)
        got = b.eval(pts)
        assert got.shape == (3, w, 2)
        for pi, z in enumerate(pts.tolist()):
            for c in range(w):
                want = pyref.ext_horner([(int(v), 0) for v in cols[c]], (z[0] % P, z[1] % P))
                assert tuple(got[pi, c].tolist()) == want
        b.free()


def test_gpu_equals_fri_golden():
    """tests/golden/fri_small.json (computed by definition in tests/pyref.py): the device's prove_openings polynomial,
    the commit-phase caps / digests and the remaining coefficients."""
    import fri_ref
    import mapreduce_plonky2_b200 as G

    G.init(0)
    for c in fri_ref.load_fri_golden():
        kind = c["hash_kind"]
        oracles = [G.PolynomialBatch.from_coeffs(list(o), c["rate_bits"], False, 4, hash_kind=kind, keep_on_device=True,
                                                 fetch_leaves=False) for o in c["oracles"]]
        ph = G.FriCommitPhase.from_openings(oracles, c["batches"], c["alpha"], c["cap_height"], kind, want_final_poly=True)
        assert np.array_equal(ph.final_poly, c["final_poly"])
        for ab, beta, want_cap in zip(c["arity_bits"], c["betas"], c["layer_caps"]):
            cap = ph.commit_layer(ab)
            assert np.array_equal(cap.hashes, want_cap)
            ph.fold(beta)
        for i, want_xor in enumerate(c["layer_digests_xor"]):
            assert np.array_equal(np.bitwise_xor.reduce(ph.layer(i).digests, axis=0), want_xor)
        assert np.array_equal(ph.finish(), c["final_coeffs"])
        ph.free()
        for o in oracles:
            o.free()
