"""Host-side model of the 22|21|21-bit limb-plane arithmetic of csrc/poseidon.cuh (Poseidon's linear layers,
plonky2 hash/poseidon.rs mds_layer -- SURVEY.md 8(a) a6): the re-normalisation identity and the 32-bit bounds the
kernel relies on.  Pure Python integers; no GPU."""
import random

P = 0xFFFFFFFF00000001
M22, M21 = (1 << 22) - 1, (1 << 21) - 1
ROW_SUM = 264  # circ(17,15,41,16,2,28,13,13,39,18,34,20) sums to 256, row 0 adds diag 8


def renorm3(o0, o1, o2):
    """pos_renorm3, with the kernel's 32-bit wrap made explicit (asserted never to happen)."""
    t1 = o1 + (o0 >> 22)
    t2 = o2 + (t1 >> 21)
    assert t1 < 1 << 32 and t2 < 1 << 32
    ov = t2 >> 21
    l0 = (o0 & M22) - ov + 0x400001
    l1 = (t1 & M21) + ov * 1024 + 0x1FFBFF
    l2 = (t2 & M21) + 0x1FFFFF
    assert 0 <= l0 < 1 << 32 and 0 <= l1 < 1 << 32 and 0 <= l2 < 1 << 32
    return l0, l1, l2


def value(l0, l1, l2):
    return (l0 + (l1 << 22) + (l2 << 43)) % P


def test_bias_is_a_multiple_of_p():
    assert (0x400001 + (0x1FFBFF << 22) + (0x1FFFFF << 43)) % P == 0


def test_renorm_preserves_the_value_mod_p():
    rng = random.Random(0x6D7032)
    corners = [0, 1, M22, M22 + 1, M21, M21 + 1, (1 << 31) - 1, 1 << 31, ROW_SUM * ((1 << 23) + 2), (1 << 32) - (1 << 12)]
    cases = [(a, b, c) for a in corners for b in corners for c in corners]
    cases += [(rng.randrange(1 << 32 - 1), rng.randrange(1 << 32 - 1), rng.randrange(1 << 32 - 1)) for _ in range(20000)]
    for o in cases:
        if o[1] + (o[0] >> 22) >= 1 << 32 or o[2] + ((o[1] + (o[0] >> 22)) >> 21) >= 1 << 32:
            continue
        assert value(*renorm3(*o)) == value(*o)


def test_bounds_are_a_fixed_point():
    """Worst-case limbs after a renorm never let the next plane output reach 2^32, for ever."""
    lmax = (M22, M21, M21)  # fresh split of a u64
    for _ in range(64):
        omax = tuple(ROW_SUM * l for l in lmax)
        assert all(o < (1 << 32) - (1 << 12) for o in omax)
        # the largest limbs a renorm can produce from outputs <= omax
        t1 = omax[1] + (omax[0] >> 22)
        t2 = omax[2] + (t1 >> 21)
        ov = t2 >> 21
        new = (M22 + 0x400001, M21 + ov * 1024 + 0x1FFBFF, M21 + 0x1FFFFF)
        lmax = tuple(max(a, b) for a, b in zip(lmax, new))
    assert lmax[0] < (1 << 23) + 2 and lmax[1] < 1 << 23 and lmax[2] < 1 << 22
