"""GPU parity of mp2gpu_partial_products_and_zs (plonky2 plonk/prover.rs `all_wires_permutation_partial_products`
+ the prover's second commitment) through the C ABI: the values equal the by-definition restatement
(tests/plonk_ref.zs_partial_products: per-wire quotients, chunk products, running product over the rows), the
commitment equals from_values of the same columns, and the batch feeds mp2gpu_quotient_polys unchanged."""
import random

import numpy as np
import pytest

import plonk_ref as PR

pytestmark = pytest.mark.gpu
P = PR.P


def _batches(G, inst, rate_bits, cap, kind):
    mk = lambda cols: G.PolynomialBatch.from_values(np.array(cols, dtype=np.uint64), rate_bits, False, cap, hash_kind=kind,
                                                    keep_on_device=True, fetch_leaves=False)
    return mk(inst.constants + inst.sigmas), mk(inst.wires)


@pytest.mark.parametrize("seed,degree_bits,two_groups,kind,qbits,with_poseidon", [
    (1, 3, False, 0, 3, False), (2, 4, True, 1, 3, False), (3, 5, True, 0, 2, False), (4, 6, False, 1, 3, True),
    (5, 10, True, 1, 3, True), (6, 11, True, 0, 1, False)])
def test_partial_products_and_zs_match_the_definition(seed, degree_bits, two_groups, kind, qbits, with_poseidon):
    import mapreduce_plonky2_b200 as G
    from mapreduce_plonky2_b200 import quotient as Q

    G.init(0)
    rng = random.Random(0x9292 + seed)
    inst = PR.synthetic_instance(seed, degree_bits=degree_bits, two_groups=two_groups, with_poseidon=with_poseidon)
    c = inst.circuit
    c.quotient_degree_bits = qbits
    betas, gammas = ([rng.randrange(P) for _ in range(c.num_challenges)] for _ in range(2))
    rate_bits, cap = 3, min(4, degree_bits + 3)
    b_cs, b_w = _batches(G, inst, rate_bits, cap, kind)
    desc = Q.CircuitDesc.from_circuit(c)
    got = Q.partial_products_and_zs(desc, b_cs, b_w, betas, gammas, rate_bits, cap, hash_kind=kind, fetch_leaves=True,
                                    fetch_digests=True)
    want = np.array(PR.zs_partial_products(inst, betas, gammas), dtype=np.uint64)
    assert got.polynomials.shape == want.shape == (c.num_challenges * (1 + c.num_partial_products), c.n)
    assert np.array_equal(got.polynomials, want)
    # the commitment is from_values of those columns
    ref = G.PolynomialBatch.from_values(want, rate_bits, False, cap, hash_kind=kind)
    assert np.array_equal(got.merkle_tree.cap.hashes, ref.merkle_tree.cap.hashes)
    assert np.array_equal(got.merkle_tree.leaves, ref.merkle_tree.leaves)
    assert np.array_equal(got.merkle_tree.digests, ref.merkle_tree.digests)
    # and the resident batch is what the quotient step consumes
    alphas = [rng.randrange(P) for _ in range(c.num_challenges)]
    q1 = Q.compute_quotient_polys(desc, b_cs, b_w, got, betas, gammas, alphas, inst.public_inputs_hash, rate_bits, cap,
                                  hash_kind=kind)
    b_z = G.PolynomialBatch.from_values(want, rate_bits, False, cap, hash_kind=kind, keep_on_device=True, fetch_leaves=False)
    q2 = Q.compute_quotient_polys(desc, b_cs, b_w, b_z, betas, gammas, alphas, inst.public_inputs_hash, rate_bits, cap,
                                  hash_kind=kind)
    assert np.array_equal(q1.polynomials, q2.polynomials)
    assert np.array_equal(q1.merkle_tree.cap.hashes, q2.merkle_tree.cap.hashes)
    for b in (b_cs, b_w, got, b_z, q1, q2):
        b.free()


def test_partial_products_errors():
    import mapreduce_plonky2_b200 as G
    from mapreduce_plonky2_b200 import quotient as Q

    G.init(0)
    inst = PR.synthetic_instance(3, degree_bits=4)
    c = inst.circuit
    b_cs, b_w = _batches(G, inst, 3, 4, 1)
    desc = Q.CircuitDesc.from_circuit(c)
    with pytest.raises(G.Mp2GpuError, match="num_challenges entries"):
        Q.partial_products_and_zs(desc, b_cs, b_w, [1], [2, 3], 3, 4)
    junk = G.PolynomialBatch.from_values(np.array(inst.wires[:3], dtype=np.uint64), 3, False, 4, hash_kind=1,
                                         keep_on_device=True, fetch_leaves=False)
    with pytest.raises(G.Mp2GpuError, match="wires batch must hold"):
        Q.partial_products_and_zs(desc, b_cs, junk, [1, 2], [3, 4], 3, 4)
    with pytest.raises(G.Mp2GpuError, match="constants_sigmas batch must hold"):
        Q.partial_products_and_zs(desc, junk, b_w, [1, 2], [3, 4], 3, 4)
    junk.free()
    # a zero denominator: gamma = -(w + beta * sigma) at row 0, wire 0 (plonky2's batch inversion panics here)
    beta = 5
    gamma = (-(inst.wires[0][0] + beta * inst.sigmas[0][0])) % P
    with pytest.raises(G.Mp2GpuError, match="zero denominator"):
        Q.partial_products_and_zs(desc, b_cs, b_w, [beta, 7], [gamma, 9], 3, 4)
    b_cs.free()
    b_w.free()
