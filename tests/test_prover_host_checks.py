"""Host-side argument checks of the prover mirrors (no device call is reached): prove_native / partial_products_and_zs /
compute_quotient_polys refuse batches that are not device-resident, wrong challenge counts, wrong witness shapes and gate
kinds outside the supported set -- errors, never silent fallbacks."""
import numpy as np
import pytest


class _FakeBatch:
    """What the checks look at: a PolynomialBatch that was NOT kept on the device."""
    _handle = None
    num_polys = 3


def _desc():
    from mapreduce_plonky2_b200 import quotient as Q
    return Q.CircuitDesc(4, 11, 8, 3, [Q.GateDesc("arithmetic", 2), Q.GateDesc("noop")], [0, 0], [(0, 2)], 3, 2)


def test_descriptor_properties_and_unknown_gate():
    from mapreduce_plonky2_b200 import quotient as Q
    from mapreduce_plonky2_b200._lib import Mp2GpuError
    d = _desc()
    assert d.num_selectors == 1 and d.num_partial_products == 0
    assert Q.CircuitDesc(4, 135, 80, 4, [], [], [], 3, 2).num_partial_products == 9
    cc, keep = d._c()
    assert cc.num_gates == 2 and keep[0].kind == Q.GATE_KINDS["arithmetic"] and keep[1].group_end == 2
    d.gates[1] = Q.GateDesc("lookup_table")
    with pytest.raises(Mp2GpuError, match="outside the supported subset"):
        d._c()
    assert Q.GATE_KINDS["coset_interpolation"] == 13 and len(set(Q.GATE_KINDS.values())) == len(Q.GATE_KINDS) == 14


def test_non_resident_batches_and_bad_shapes_are_refused():
    from mapreduce_plonky2_b200 import prover as GP
    from mapreduce_plonky2_b200 import quotient as Q
    from mapreduce_plonky2_b200._lib import Mp2GpuError
    d, fake = _desc(), _FakeBatch()
    with pytest.raises(Mp2GpuError, match="device-resident"):
        Q.partial_products_and_zs(d, fake, fake, [1, 2], [3, 4], 3, 4)
    with pytest.raises(Mp2GpuError, match="device-resident"):
        Q.compute_quotient_polys(d, fake, fake, fake, [1, 2], [3, 4], [5, 6], [0] * 4, 3, 4)
    with pytest.raises(Mp2GpuError, match="device-resident"):
        GP.prove_native(d, fake, [1, 2, 3, 4], np.zeros((11, 16), dtype=np.uint64), [], [0] * 4)
    resident = _FakeBatch()
    resident._handle = object()
    with pytest.raises(Mp2GpuError, match="num_challenges entries"):
        Q.partial_products_and_zs(d, resident, resident, [1], [3, 4], 3, 4)
    with pytest.raises(Mp2GpuError, match="num_challenges entries"):
        Q.compute_quotient_polys(d, resident, resident, resident, [1, 2], [3, 4], [5], [0] * 4, 3, 4)
    with pytest.raises(Mp2GpuError, match="num_wires, 2\\^degree_bits"):
        GP.prove_native(d, resident, [1, 2, 3, 4], np.zeros((10, 16), dtype=np.uint64), [], [0] * 4)


def test_fri_params_schedule_matches_the_constant_arity_strategy():
    from mapreduce_plonky2_b200 import fri as GF
    cfg = GF.FriConfig()
    assert [cfg.fri_params(d).reduction_arity_bits for d in (5, 12, 13, 14, 20)] == [[], [4, 4], [4, 4], [4, 4, 4], [4, 4, 4, 4]]
    assert cfg.fri_params(14).lde_bits == 17
