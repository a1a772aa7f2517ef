"""Device self-test of the Goldilocks primitives (csrc/gl.cuh, csrc/dft.cuh) through the C ABI: every result of
add / add-canonical / sub / mul / sqr / mul-add / reduce128 / x^7 / the shift twiddles over ~1M operand pairs
(all combinations of 48 corner values around 0, 2^32, 2^63, p, 2^64, plus pseudo-random ones) must equal
128-bit arithmetic by definition.  Mirrors plonky2_field's goldilocks_field arithmetic tests."""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu

NAMES = ["add", "add_canonical", "sub", "mul", "sqr", "mul_add", "reduce128", "pow7", "mul_2^24", "mul_2^48", "mul_2^72",
         "mul_2^(12j)", "dft8", "dft16"]


def test_field_primitives_exact_on_corner_and_random_operands():
    import mapreduce_plonky2_b200 as G
    from mapreduce_plonky2_b200 import _lib

    G.init(0)
    bad = (C.c_uint64 * len(NAMES))()
    _lib.call("mp2gpu_debug_field_selftest", bad, len(NAMES))
    wrong = {n: int(b) for n, b in zip(NAMES, bad) if b}
    assert not wrong, "mismatches vs 128-bit reference: %r" % wrong
