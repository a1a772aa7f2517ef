"""Device self-test of the Goldilocks primitives (csrc/gl.cuh, csrc/dft.cuh) through the C ABI: every result of
add / add-canonical / sub / mul / sqr / mul-add / reduce128 / x^7 / the shift twiddles over ~1M operand pairs
(all combinations of 48 corner values around 0, 2^32, 2^63, p, 2^64, plus pseudo-random ones) must equal
128-bit arithmetic by definition.  Mirrors plonky2_field's goldilocks_field arithmetic tests."""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu

NAMES = ["add", "add_canonical", "sub", "mul", "sqr", "mul_add", "reduce128", "pow7", "mul_2^24", "mul_2^48", "mul_2^72",
         "mul_2^(12j)"]


def test_field_primitives_exact_on_corner_and_random_operands():
    import mapreduce_plonky2_b200 as G
    from mapreduce_plonky2_b200 import _lib

    G.init(0)
    bad = (C.c_uint64 * len(NAMES))()
    _lib.call("mp2gpu_debug_field_selftest", bad, len(NAMES))
    wrong = {n: int(b) for n, b in zip(NAMES, bad) if b}
    assert not wrong, "mismatches vs 128-bit reference: %r" % wrong


@pytest.mark.parametrize("log_points", [3, 4])
def test_shift_twiddle_butterflies_equal_the_dft_by_definition(log_points):
    """csrc/dft.cuh's 8- and 16-point butterflies on caller data vs sum_k x_k w^(f k) in Python integers
    (w = plonky2's primitive_root_of_unity(log_points); output position j holds frequency bitrev(j))."""
    import numpy as np

    import mapreduce_plonky2_b200 as G
    from mapreduce_plonky2_b200 import _lib
    from util import P, splitmix64

    G.init(0)
    npts, count = 1 << log_points, 64
    x = splitmix64(0xD0F7 + log_points, npts * count)            # raw u64: non-canonical values included
    x[:npts] = np.uint64(P - 1)
    x[npts:2 * npts] = np.uint64(2**64 - 1)
    io = x.copy()
    _lib.call("mp2gpu_debug_dft", io.ctypes.data_as(_lib.u64p), log_points, count)
    w = pow(7, (P - 1) >> log_points, P)
    assert w == pow(2, 39 * (64 >> log_points), P)                # w_64 = 2^39
    for t in range(count):
        xs = [int(v) % P for v in x[t * npts:(t + 1) * npts]]
        for pos in range(npts):
            f = int(format(pos, "0%db" % log_points)[::-1], 2)
            want = sum(v * pow(w, f * k, P) for k, v in enumerate(xs)) % P
            assert int(io[t * npts + pos]) == want, (t, pos)
