"""Host-side model of the FP64 linear layers of csrc/poseidon.cuh (poseidon_permute_f64; plonky2 hash/poseidon.rs
mds_layer / partial rounds -- SURVEY.md 8(a) a6).  Python floats are IEEE doubles, so every operation below is the
operation the kernel executes; the model asserts that no intermediate leaves the 53-bit exact range, that the
biased conversions always see non-negative integers, and that the result equals the by-definition permutation.
No GPU."""
import os
import random
import re
import struct

import pyref as R

P = R.P
HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mapreduce_plonky2_b200", "csrc",
                   "poseidon_constants.h")
D52 = float(1 << 52)
K84 = 1.5 * 2.0 ** 84
I32 = 2.0 ** -32


def table(name):
    txt = open(HDR).read()
    m = re.search(r"#define %s_LIST \\\n((?:.*\\\n)*.*)\n" % name, txt)
    body = m.group(1).replace("\\", " ")
    return [tok for tok in re.split(r"[,\s]+", body) if tok]


def dtable(name):
    return [float(t) for t in table(name)]


def utable(name):
    return [int(t.rstrip("UL"), 16) for t in table(name)]


def exact(x):
    assert x == int(x) and abs(x) < 2.0 ** 53, x
    return x


def u32_to_d(w):
    bits = (0x43300000 << 32) | w
    return exact(struct.unpack("<d", struct.pack("<Q", bits))[0] - D52)


def merge_d(tA, tB):
    """pos_merge_d: words of the biased doubles -> loose u64 (32-bit wrap made explicit)."""
    qa, qb = (struct.unpack("<Q", struct.pack("<d", t))[0] for t in (tA, tB))
    aL, hA, bL, hB = qa & 0xFFFFFFFF, qa >> 32, qb & 0xFFFFFFFF, qb >> 32
    assert hA >= 0x43300000 and hB >= 0x43300000 and hA - 0x43300000 < 1 << 20 and hB - 0x43300000 < 1 << 20
    u = (hA + hB + 0x79A00000) & 0xFFFFFFFF
    bh = (hB + 0xBCD00000) & 0xFFFFFFFF
    hi = bL + u
    c = hi >> 32
    hi &= 0xFFFFFFFF
    lo = aL - bh
    brw = 1 if lo < 0 else 0
    lo &= 0xFFFFFFFF
    hi = hi - brw
    brw2 = 1 if hi < 0 else 0
    hi &= 0xFFFFFFFF
    k = c - brw2
    full = (hi << 32 | lo) + k * 0xFFFFFFFF
    assert 0 <= full < 1 << 64  # one fold is enough
    return full


def mds_plane_d(x):
    a = [exact(x[k] + x[k + 6]) for k in range(6)]
    b = [exact(x[k] - x[k + 6]) for k in range(6)]
    aa = [exact(a[k] + a[k + 3]) for k in range(3)]
    ab = [exact(a[k] - a[k + 3]) for k in range(3)]
    S = exact(exact(aa[0] + aa[1]) + aa[2])
    q = [exact(S + aa[2]), exact(S + aa[0]), exact(S + aa[1])]
    yab = [exact(ab[2] * 8.0 - ab[0] - ab[1] * 2.0), exact(-(ab[0] * 8.0) - ab[1] - ab[2] * 2.0),
           exact(ab[0] * 2.0 - ab[1] * 8.0 - ab[2])]
    ya = [exact(q[k] * 16.0 + yab[k]) for k in range(3)] + [exact(q[k] * 16.0 - yab[k]) for k in range(3)]
    f = [2, -4, 16, 1, -1, -1]  # negacyclic-6 block, as spelled out term by term in the kernel
    yb = []
    for k in range(6):
        acc = 0.0
        for j in range(6):
            # y_k = sum_j f[(k - j) mod 6] * b_j * (-1 if j > k)
            coef = f[(k - j) % 6] * (-1 if j > k else 1)
            acc = exact(acc + coef * b[j])
        yb.append(acc)
    y = [exact(ya[k] + yb[k]) for k in range(6)] + [exact(ya[k] - yb[k]) for k in range(6)]
    y[0] = exact(x[0] * 8.0 + y[0])
    return y


def mds_by_definition(s):
    return [(sum(s[(i + row) % 12] * R.POS_CIRC[i] for i in range(12)) + s[row] * R.POS_DIAG[row]) % P for row in range(12)]


def renorm_d(A, B):
    cA = (A + K84) - K84
    A1 = exact(A - cA)
    B1 = exact(cA * I32 + B)
    cB = (B1 + K84) - K84
    B2 = exact(B1 - cB)
    return exact(A1 - cB * I32), exact(cB * I32 + B2)


def poseidon_f64(state, track=None, f64_full=False):
    """poseidon_permute_f64; f64_full mirrors MP2_POSEIDON_F64_FULL (0 = the shipped default: integer planes in the
    full rounds, whose result is by definition `mds(state) + constants`, FP64 planes in the partial rounds)."""
    rc = utable("MP2_POSEIDON_RC")
    dbias, t0bias = dtable("MP2_POSEIDON_DBIAS"), dtable("MP2_POSEIDON_T0BIAS")
    exbias, r4 = dtable("MP2_POSEIDON_EXBIAS"), dtable("MP2_POSEIDON_R4D")
    s = [(v + rc[i]) % P for i, v in enumerate(state)]
    A = B = None
    for phase in range(2):
        r0 = 26 if phase else 0
        for k in range(4):
            s = [pow(v, 7, P) for v in s]
            if not f64_full:
                nxt = (rc + [0] * 12)[12 * (r0 + k + 1):12 * (r0 + k + 2)]
                s = [(a + b) % P for a, b in zip(mds_by_definition(s), nxt)]
                continue
            A = [u32_to_d(v & 0xFFFFFFFF) for v in s]
            B = [u32_to_d(v >> 32) for v in s]
            YA, YB = mds_plane_d(A), mds_plane_d(B)
            if phase == 0 and k == 3:
                A = [None] + [exact(YA[i] + r4[2 * i]) for i in range(1, 12)]
                B = [None] + [exact(YB[i] + r4[2 * i + 1]) for i in range(1, 12)]
                s[0] = merge_d(YA[0] + dbias[24 * 4], YB[0] + dbias[24 * 4 + 1])
            else:
                o = 24 * (r0 + k + 1)
                s = [merge_d(YA[i] + dbias[o + 2 * i], YB[i] + dbias[o + 2 * i + 1]) for i in range(12)]
        if phase == 0:
            if not f64_full:
                A = [None] + [u32_to_d(v & 0xFFFFFFFF) for v in s[1:]]
                B = [None] + [u32_to_d(v >> 32) for v in s[1:]]
            parity = 0 if f64_full else 1
            s0 = s[0]
            for r in range(4, 26):
                s0 = pow(s0, 7, P)
                A[0], B[0] = u32_to_d(s0 & 0xFFFFFFFF), u32_to_d(s0 >> 32)
                YA, YB = mds_plane_d(A), mds_plane_d(B)
                if track is not None:
                    track["y0"] = max(track.get("y0", 0), abs(YA[0]), abs(YB[0]))
                s0 = merge_d(exact(YA[0] + t0bias[2 * (r - 4)]), exact(YB[0] + t0bias[2 * (r - 4) + 1]))
                A, B = list(YA), list(YB)
                if r % 2 == parity:
                    for i in range(1, 12):
                        A[i], B[i] = renorm_d(A[i], B[i])
                        assert abs(A[i]) <= 2 ** 31 + 2 ** 19 and abs(B[i]) <= 2 ** 31 + 2 ** 19
            s[0] = s0
            if track is not None:
                track["exit"] = max(track.get("exit", 0), max(abs(v) for v in A[1:] + B[1:]))
            for i in range(1, 12):
                s[i] = merge_d(exact(A[i] + exbias[2 * i]), exact(B[i] + exbias[2 * i + 1]))
    return [v % P for v in s]


def test_plane_offsets_are_multiples_of_p():
    t0 = utable("MP2_POSEIDON_PARTIAL_T0")
    t0bias = dtable("MP2_POSEIDON_T0BIAS")
    for k, v in enumerate(t0):
        a, b = int(t0bias[2 * k] - D52), int(t0bias[2 * k + 1] - D52)
        assert (a + (b << 32)) % P == v % P


def test_f64_model_equals_the_permutation():
    rng = random.Random(0xF64)
    cases = [[0] * 12, [P - 1] * 12, [(1 << 64) - 1] * 12, list(range(12))]
    cases += [[rng.randrange(1 << 64) for _ in range(12)] for _ in range(40)]
    for full in (False, True):
        track = {}
        for st in cases:
            assert poseidon_f64(st, track, full) == R.poseidon([v % P for v in st])
        assert track["y0"] < 2.0 ** 48.5 and track["exit"] < 2.0 ** 39.3


def test_worst_case_magnitudes_fit_the_mantissa():
    """All-maximal planes (every input at its bound, all signs aligned) stay below 2^53 through two un-normalised
    layers: the row sum of |coefficients| is 272 and no intermediate of the decomposition exceeds it."""
    x = [2.0 ** 31 + 2.0 ** 19] * 12
    y = mds_plane_d(x)                      # asserts exactness internally
    assert max(y) <= 272 * x[0]
    z = mds_plane_d([max(y)] * 12)
    assert max(z) < 2.0 ** 48 and max(z) <= 272 * 272 * x[0]
    # entry state of the partial rounds: 41-bit planes, one layer, then the renormalisation
    w = mds_plane_d([272.0 * 2 ** 32 + 2 ** 32] * 12)
    assert max(w) < 2.0 ** 49
