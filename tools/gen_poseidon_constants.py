"""Emits mapreduce_plonky2_b200/csrc/poseidon_constants.h: the static round-constant tables the CUDA
library ships with.  Constants are regenerated from first principles by the pure-Python generators
(ChaCha8Rng seed 0 for Poseidon, Grain LFSR for Poseidon2 -- SURVEY.md A.6/A.7); the test-suite
checks the library against the oracle, which regenerates them independently in C.

    python tools/gen_poseidon_constants.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyref as R  # noqa: E402


def table(name, vals, per_line=4):
    """Emits NAME_LEN and a NAME_LIST initialiser list usable for __constant__ and host arrays."""
    out = ["#define %s_LEN %d" % (name, len(vals)), "#define %s_LIST \\" % name]
    rows = [" ".join("0x%016xULL," % v for v in vals[i:i + per_line]) for i in range(0, len(vals), per_line)]
    out.append(" \\\n".join("    " + r for r in rows))
    return "\n".join(out)


def table32(name, vals, per_line=6):
    out = ["#define %s_LEN %d" % (name, len(vals)), "#define %s_LIST \\" % name]
    rows = [" ".join("0x%08xu," % v for v in vals[i:i + per_line]) for i in range(0, len(vals), per_line)]
    out.append(" \\\n".join("    " + r for r in rows))
    return "\n".join(out)


def limbs3(v):
    """22 | 21 | 21 bit limbs used by the 32-bit IMAD MDS (poseidon.cuh)."""
    return [v & 0x3FFFFF, (v >> 22) & 0x1FFFFF, v >> 43]


def tabled(name, vals, per_line=4):
    """Integer-valued doubles (all < 2^53, exact) for the FP64 linear layers."""
    assert all(0 <= v < 1 << 53 for v in vals)
    out = ["#define %s_LEN %d" % (name, len(vals)), "#define %s_LIST \\" % name]
    rows = [" ".join("%d.0," % v for v in vals[i:i + per_line]) for i in range(0, len(vals), per_line)]
    out.append(" \\\n".join("    " + r for r in rows))
    return "\n".join(out)


def plane_offsets(bits):
    """(OA, OB) with OA + 2^32*OB = k*p, both ~2^bits: added to signed plane values |.| < 2^bits so that the
    biased conversions see non-negative integers.  OA = k + j*2^32, OB = k*(2^32 - 1) - j with j = k = 2^(bits-32)."""
    k = j = 1 << (bits - 32)
    oa, ob = k + (j << 32), k * ((1 << 32) - 1) - j
    assert (oa + (ob << 32)) % R.P == 0 and oa > 1 << (bits - 1) and ob > 1 << (bits - 1)
    return oa, ob


def poseidon_f64_tables():
    """Bias tables of poseidon_permute_f64 (poseidon.cuh): 2^52 + offset + 32-bit word of the constant."""
    rc = R.poseidon_round_constants() + [0] * 12
    t0, d = poseidon_partial_constants()
    B52 = 1 << 52
    words = lambda v: (v & 0xFFFFFFFF, v >> 32)
    dbias = [B52 + w for v in rc for w in words(v)]
    oa, ob = plane_offsets(49)      # lane 0 inside the partial rounds: |Y| < 2^48.4
    t0bias = [x for v in t0 for x in (B52 + oa + words(v)[0], B52 + ob + words(v)[1])]
    oa, ob = plane_offsets(40)      # leaving the partial rounds: |A|, |B| < 2^39.2
    exbias = [x for v in d for x in (B52 + oa + words(v)[0], B52 + ob + words(v)[1])]
    r4 = [w for v in rc[48:60] for w in words(v)]
    return dbias, t0bias, exbias, r4


def poseidon_partial_constants():
    """Constants of the 22 partial rounds pushed through the linear layers so that only lane 0 receives one per
    round (plus one correction vector when the partial rounds end).  With x_r the true state before the S-box
    of round r and y_r = x_r - d_r the state the kernel carries (d_r[0] = 0, d_4 = 0):
        x_{r+1} = M z + c_{r+1},  z = x_r with lane 0 S-boxed
                = M z' + t_{r+1},  z' = y_r with lane 0 S-boxed,  t_{r+1} = M (0, d_r[1..]) + c_{r+1}
        y_{r+1} = M z' + t_{r+1}[0] e_0,   d_{r+1} = (0, t_{r+1}[1..]).
    Returns (T0[22] = t_5[0] .. t_26[0], D[12] = d_26)."""
    P = R.P
    rc = R.poseidon_round_constants()

    def mds(v):
        return [(sum(v[(i + r) % 12] * R.POS_CIRC[i] for i in range(12)) + v[r] * R.POS_DIAG[r]) % P for r in range(12)]

    d = [0] * 12
    t0 = []
    for r in range(4, 26):
        c_next = rc[12 * (r + 1):12 * (r + 2)]
        t = [(a + b) % P for a, b in zip(mds([0] + d[1:]), c_next)]
        t0.append(t[0])
        d = [0] + t[1:]
    # self-check: the rewritten schedule equals the naive permutation
    import random
    rng = random.Random(7)
    for _ in range(5):
        x = [rng.randrange(P) for _ in range(12)]
        s = list(x)
        for r in range(4):
            s = mds([pow((v + rc[12 * r + i]) % P, 7, P) for i, v in enumerate(s)])
        s = [(v + rc[48 + i]) % P for i, v in enumerate(s)]          # constants of round 4 (added by round 3's layer)
        for k, r in enumerate(range(4, 26)):
            s[0] = pow(s[0], 7, P)
            s = mds(s)
            s[0] = (s[0] + t0[k]) % P
        s = [(v + dd) % P for v, dd in zip(s, d)]
        for r in range(26, 30):
            s = mds([pow(v, 7, P) for v in s])
            if r < 29:
                s = [(v + rc[12 * (r + 1) + i]) % P for i, v in enumerate(s)]
        assert s == R.poseidon(x), "partial-round constant folding is wrong"
    return t0, d


def main():
    rc = R.poseidon_round_constants() + [0] * 12   # padded: "MDS then add the NEXT round's constants"
    rc3 = [l for v in rc for l in limbs3(v)]
    # Poseidon2: constants added right after a linear layer = the NEXT external round's constants.
    # slot 0: before round 0 (after the initial M_E); slots 1..3 after external rounds 0..2; slot 4 (after
    # round 3) and slot 8 (after round 7) are zero; slots 5..7 after external rounds 4..6.  RC_ext[4] is added
    # explicitly when the state leaves the internal rounds.
    p2 = R.poseidon2_round_constants()
    ext = [p2[12 * r:12 * r + 12] for r in range(4)] + [p2[70 + 12 * r:70 + 12 * r + 12] for r in range(4)]
    zero = [0] * 12
    slots = [ext[0], ext[1], ext[2], ext[3], zero, ext[5], ext[6], ext[7], zero]
    p2_rc3 = [l for slot in slots for v in slot for l in limbs3(v)]
    body = [
        "// GENERATED by tools/gen_poseidon_constants.py -- do not edit.",
        "// Poseidon: RC[k] = ChaCha8Rng::seed_from_u64(0).gen_range(0..p) (plonky2 0.2.2 ALL_ROUND_CONSTANTS).",
        "// Poseidon2: Horizen-Labs Goldilocks t=12 instance (Grain LFSR round constants, MAT_DIAG12_M_1).",
        "#pragma once",
        table("MP2_POSEIDON_RC", R.poseidon_round_constants()),
        "// the same constants (plus one all-zero round) as 22|21|21-bit limbs, index (12*round + lane)*3 + limb",
        table32("MP2_POSEIDON_RC3", rc3),
        table("MP2_POSEIDON2_RC", R.poseidon2_round_constants()),
        "// Poseidon partial rounds: lane-0 constants t_5[0]..t_26[0] and the correction vector d_26 (see generator)",
        table("MP2_POSEIDON_PARTIAL_T0", poseidon_partial_constants()[0]),
        table("MP2_POSEIDON_PARTIAL_D", poseidon_partial_constants()[1]),
        "// FP64 linear layers: conversion biases (2^52 + offset + constant word), see poseidon_permute_f64",
        tabled("MP2_POSEIDON_DBIAS", poseidon_f64_tables()[0]),
        tabled("MP2_POSEIDON_T0BIAS", poseidon_f64_tables()[1]),
        tabled("MP2_POSEIDON_EXBIAS", poseidon_f64_tables()[2]),
        tabled("MP2_POSEIDON_R4D", poseidon_f64_tables()[3]),
        "// Poseidon2 external-round constants as 22|21|21-bit limbs, index ((12*slot + lane)*3 + limb); see generator",
        table32("MP2_POSEIDON2_RC3", p2_rc3),
        "// the same slot constants as FP64 conversion biases: (12*slot + lane)*2 + word -> 2^52 + 32-bit word (p2_external_rc_d)",
        tabled("MP2_POSEIDON2_DBIAS", [(1 << 52) + w for slot in slots for v in slot for w in (v & 0xFFFFFFFF, v >> 32)]),
        table("MP2_POSEIDON2_DIAG", R.P2_DIAG),
        "",
    ]
    path = os.path.join(ROOT, "mapreduce_plonky2_b200", "csrc", "poseidon_constants.h")
    with open(path, "w") as f:
        f.write("\n".join(body))
    print("wrote", path)


if __name__ == "__main__":
    main()
