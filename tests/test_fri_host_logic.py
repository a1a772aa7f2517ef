"""Host-side FRI glue of mapreduce_plonky2_b200.fri on the CPU: the challenger, the reduction schedule, the PoW
hand-off and the query-round assembly, with every device primitive replaced by an oracle-backed stand-in (the
same pattern as tests/test_sharded_gloo.py).  The GPU tests run the same code over the real library."""
import numpy as np
import pytest

import fri_ref
import pyref
from util import P, field_elems


@pytest.fixture
def cpu_primitives(monkeypatch, oracle):
    """Route the two library calls the glue makes directly (permute, PoW grind) to the oracle."""
    from mapreduce_plonky2_b200 import plonky2 as P2

    def permute(states, hash_kind=1):
        st = np.array(states, dtype=np.uint64).reshape(-1, 12).copy()
        for row in st:
            row[:] = oracle.permute(row, hash_kind)
        return st

    def hash_no_pad_batch(inputs, hash_kind=1):
        return np.stack([oracle.hash_no_pad(np.asarray(row, dtype=np.uint64), hash_kind) for row in inputs])

    monkeypatch.setattr(P2, "hash_no_pad_batch", hash_no_pad_batch)
    monkeypatch.setattr(P2, "permute", permute)
    monkeypatch.setattr(P2, "fri_proof_of_work",
                        lambda state, pos, bits=16, kind=1: oracle.fri_pow(np.array(state, dtype=np.uint64), pos, bits, kind))
    return oracle


class _OracleBatch:
    """Stand-in for a device-resident PolynomialBatch: open() / eval() / merkle_tree.cap from the oracle's commitment."""

    def __init__(self, oracle, coeffs, kind):
        from mapreduce_plonky2_b200.plonky2 import MerkleCap

        self.o, self.kind, self.polynomials = oracle, kind, coeffs
        self.c = oracle.commit(coeffs, 3, 4, kind, from_coeffs=True)
        self.degree_log, self.rate_bits = int(coeffs.shape[1]).bit_length() - 1, 3
        self.merkle_tree = type("T", (), {"cap": MerkleCap(self.c["cap"])})()

    def open(self, idx):
        N = self.c["leaves"].shape[0]
        return (self.c["leaves"][list(idx)], np.stack([self.o.merkle_prove(self.c["digests"], N, 4, int(i)) for i in idx]))

    def eval(self, points):
        return np.array([[pyref.ext_horner([(int(c), 0) for c in col], tuple(int(v) % P for v in z)) for col in self.polynomials]
                         for z in points], dtype=np.uint64)


class _OraclePhase:
    """Stand-in for FriCommitPhase: the oracle's fri_committed_trees, one layer at a time."""

    def __init__(self, oracle, final_poly, kind):
        from mapreduce_plonky2_b200.plonky2 import MerkleCap

        self.o, self.kind, self.MerkleCap = oracle, kind, MerkleCap
        n = final_poly.shape[0]
        self.coeffs = np.zeros((n << 3, 2), dtype=np.uint64)
        self.coeffs[:n] = final_poly
        self.shift, self.layers, self.ab = 7, [], None

    def commit_layer(self, arity_bits):
        leaves = self.o.fri_layer_leaves(self.o.coset_fft_ext(self.coeffs, self.shift), arity_bits)
        digests, cap = self.o.merkle_new(leaves, 4, self.kind)
        self.layers.append((leaves, digests))
        self.ab = arity_bits
        return self.MerkleCap(cap)

    def fold(self, beta):
        self.coeffs = self.o.fri_fold(self.coeffs, self.ab, np.asarray(beta, dtype=np.uint64))
        self.shift = pow(self.shift, 1 << self.ab, P)

    def finish(self):
        return self.coeffs[:self.coeffs.shape[0] >> 3]

    def open_layer(self, i, idx):
        leaves, digests = self.layers[i]
        return leaves[list(idx)], np.stack([self.o.merkle_prove(digests, leaves.shape[0], 4, int(x)) for x in idx])


def test_challenger_matches_definition(cpu_primitives):
    from mapreduce_plonky2_b200 import fri as GF

    rng = np.random.default_rng(5)
    for kind in (0, 1):
        a, b = GF.Challenger(kind), pyref.Challenger(kind)
        for step in range(40):
            if rng.integers(0, 3):
                xs = [int(v) for v in rng.integers(0, 2**31, int(rng.integers(1, 11)), dtype=np.uint64)]
                xs = [x * 0x1_0000_0001 % P for x in xs]
                xs[0] += P if step % 5 == 0 and xs[0] < 2**64 - P else 0          # non-canonical input is reduced
                a.observe_elements(np.array(xs, dtype=np.uint64))
                b.observe(xs)
            else:
                assert a.get_challenge() == b.challenge()
        assert tuple(int(v) for v in a.get_extension_challenge()) == b.ext_challenge()
        assert len(a.input_buffer) < 8


def test_reduction_schedule_and_params():
    from mapreduce_plonky2_b200 import fri as GF

    cfg = GF.FriConfig()
    for bits in range(3, 21):
        assert cfg.fri_params(bits).reduction_arity_bits == fri_ref.arity_schedule(bits)
    assert cfg.fri_params(14).lde_bits == 17 and cfg.num_query_rounds == 28 and cfg.proof_of_work_bits == 16


@pytest.mark.parametrize("kind,degree_bits,widths", [(1, 6, (3, 4, 2)), (0, 10, (2, 3))])
def test_fri_proof_glue_equals_reference_flow(cpu_primitives, kind, degree_bits, widths):
    """fri.open_batches + the transcript + fri.fri_proof over oracle-backed stand-ins == tests/fri_ref.py's flow."""
    from mapreduce_plonky2_b200 import fri as GF

    oracle = cpu_primitives
    n = 1 << degree_bits
    rounds, pow_bits = 4, 5
    coeff_sets = [field_elems(0xAB0 + 5 * k + degree_bits, (w, n)) for k, w in enumerate(widths)]
    zeta, gzeta = (tuple(int(v) for v in field_elems(0x77A + i, 2)) for i in range(2))
    ref_batches = fri_ref.plonky2_instance(widths, zeta, gzeta)
    _, ref_openings, ref = fri_ref.oracle_fri_proof(oracle, coeff_sets, ref_batches, degree_bits, kind, pow_bits=pow_bits,
                                                   num_query_rounds=rounds)
    oracles = [_OracleBatch(oracle, c, kind) for c in coeff_sets]
    batches = [GF.FriBatchInfo(np.array(z, dtype=np.uint64), polys) for z, polys in ref_batches]
    openings = GF.open_batches(batches, oracles)
    assert [v.tolist() for v in openings] == [[list(x) for x in vals] for vals in ref_openings]
    ch = GF.Challenger(kind)
    for o in oracles:
        ch.observe_cap(o.merkle_tree.cap)
    for v in openings:
        ch.observe_extension_elements(v)
    alpha = ch.get_extension_challenge()
    final = oracle.fri_combine([(np.array(z, dtype=np.uint64), [coeff_sets[o][p] for o, p in polys]) for z, polys in ref_batches], alpha)
    params = GF.FriConfig(proof_of_work_bits=pow_bits, num_query_rounds=rounds).fri_params(degree_bits)
    proof = GF.fri_proof(oracles, _OraclePhase(oracle, final, kind), ch, params)
    assert proof.pow_witness == ref["pow_witness"]
    assert np.array_equal(proof.final_poly, ref["final_poly"])
    assert [c.hashes.tolist() for c in proof.commit_phase_merkle_caps] == [c.tolist() for c in ref["caps"]]
    for rnd, rr in zip(proof.query_round_proofs, ref["rounds"]):
        for (row, mp), (rrow, rsib) in zip(rnd.initial_trees_proof, rr["initial"]):
            assert np.array_equal(row, rrow) and np.array_equal(mp.siblings, rsib)
        for st, (rev, rsib) in zip(rnd.steps, rr["steps"]):
            assert np.array_equal(st.evals, rev) and np.array_equal(st.merkle_proof.siblings, rsib)


def test_circuit_digest_formula(cpu_primitives):
    """circuit_digest = hash_no_pad(cap ‖ hash_pad(&[]) ‖ [degree_bits]) -- the native side of
    recursion-framework/src/universal_verifier_gadget/circuit_set.rs:136-158, against pyref's sponge."""
    from mapreduce_plonky2_b200 import plonky2 as P2

    cap = field_elems(0xC1C, (16, 4))
    for kind in (0, 1):
        want = pyref.hash_no_pad([int(x) for x in cap.reshape(-1)] + list(pyref.hash_pad([], kind)) + [12], kind)
        assert [int(x) for x in P2.circuit_digest(cap, 12, kind)] == [int(x) for x in want]
        assert [int(x) for x in P2.circuit_digest(P2.MerkleCap(cap), 12, kind)] == [int(x) for x in want]
        other = P2.circuit_digest(cap, 13, kind)
        assert [int(x) for x in other] != [int(x) for x in want]
