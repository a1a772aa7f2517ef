"""Host logic of the proof-trace replay: FRI reduction schedule and commitment shapes per prove()."""
from mapreduce_plonky2_b200 import trace as T


def test_fri_reduction_schedule_matches_constant_arity_bits_4_5():
    # plonky2 ConstantArityBits(4, 5) with rate_bits 3, cap_height 4 (standard_recursion_config)
    assert T.fri_reduction_arity_bits(14) == [4, 4, 4]
    assert T.fri_reduction_arity_bits(13) == [4, 4]
    assert T.fri_reduction_arity_bits(12) == [4, 4]
    assert T.fri_reduction_arity_bits(5) == []


def test_prove_ops_shapes():
    ops = T.prove_ops(14)
    assert [(o.kind, o.ncols, o.n_log) for o in ops] == [
        ("from_values", 135, 14), ("from_values", 20, 14), ("from_coeffs", 16, 14),
        ("merkle", 32, 13), ("merkle", 32, 9), ("merkle", 32, 5)]
    # 2 228 224 + 131 056 permutations for the wires commitment
    assert ops[0].perms == 2228224 + 131056
    assert ops[0].lde_elems == 17694720
    assert ops[1].perms == 393216 + 131056
    assert ops[2].perms == 262144 + 131056


def test_leaf_proof_trace_is_three_proves():
    ops = T.proof_ops()
    assert sum(1 for o in ops if o.kind == "from_values" and o.ncols == 135) == 3
    assert [o.n_log for o in ops if o.ncols == 135] == [14, 13, 12]
