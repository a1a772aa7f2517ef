"""Generates the golden fixtures in this directory.  Run from the repo root:

    python tests/golden/make_golden.py

* ``kats.json``         -- PUBLISHED known-answer vectors, typed in from plonky2 0.2.2's own Poseidon
                           ``test_vectors`` / round-constant table and the Horizen-Labs Poseidon2
                           Goldilocks t=12 KAT.  Not computed by our code.
* ``commit_small.json`` -- whole PolynomialBatch outputs for tiny shapes, computed by the pure-Python
                           restatement ``tests/pyref.py`` (independent of the C oracle and of CUDA).
* ``merkle_small.json`` -- MerkleTree::new outputs incl. the circuit-set shape of
                           recursion-framework/src/universal_verifier_gadget/circuit_set.rs:173-191
                           (4-element digests padded with vec![F::ZERO], cap_height 0), from pyref.
* ``fri_small.json``    -- prove_openings' final polynomial (pyref.fri_combine: composition, exact division by
                           X - z, alpha weights -- by definition) and the FRI commit phase over it (layer caps, final
                           coefficients; pyref.fri_committed_trees: Horner evaluation, explicit fold) for tiny instances.
* ``config1_caps.json`` -- Merkle caps of the BASELINE config-1 shapes on seeded inputs, frozen from the
                           C oracle (self-golden: regression pin, not an external anchor).

The reference itself cannot be executed here (Rust, no toolchain), so nothing below imports it.
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import numpy as np  # noqa: E402
import pyref as R  # noqa: E402
from util import field_elems, hexlist  # noqa: E402


def hx(lst):
    return ["%016x" % v for v in lst]


def hx2(rows):
    return [hx(r) for r in rows]


def kats():
    return {
        "source": "plonky2 0.2.2 poseidon_goldilocks.rs test_vectors + ALL_ROUND_CONSTANTS anchors; "
                  "HorizenLabs poseidon2 goldilocks t=12 KAT (SURVEY.md A.6/A.7)",
        "poseidon_rc_first4": ["b585f766f2144405", "7746a55f43921ad7", "b2fb0d31cee799b4", "0f6760a4803427d7"],
        "poseidon_rc_12": "86287821f722c881",
        "poseidon_rc_last4": ["4543d9df5476d3cb", "f172d73e004fc90d", "dfd1c4febcc81238", "bc8dfb627fe558fc"],
        "poseidon_perm": [
            {"in": "zeros", "out": "3c18a9786cb0b359 c4055e3364a246c3 7953db0ab48808f4 c71603f33a1144ca d7709673896996dc 46a84e87642f44ed d032648251ee0b3c 1c687363b207df62 df8565563e8045fe 40f5b37ff4254dae d070f637b431067c 1792b1c4342109d7".split()},
            {"in": "iota", "out": "d64e1e3efc5b8e9e 53666633020aaa47 d40285597c6a8825 613a4f81e81231d2 414754bfebd051f0 cb1f8980294a023f 6eb2a9e4d54a9d0f 1902bc3af467e056 f045d5eafdc6021f e4150f77caaa3be5 c9bfd01d39b50cce 5c0a27fcb0e1459b".split()},
            {"in": "neg_one", "out": "be0085cfc57a8357 d95af71847d05c09 cf55a13d33c1c953 95803a74f4530e82 fcd99eb30a135df1 e095905e913a3029 de0392461b42919b 7d3260e24e81d031 10d3d0465d9deaa0 a87571083dfc2a47 e18263681e9958f8 e28e96f1ae5e60d3".split()},
        ],
        "poseidon2_rc_first_row": "13dcf33aba214f46 30b3b654a1da6d83 1fc634ada6159b56 937459964dc03466 edd2ef2ca7949924 ede9affde0e22f68 8515b9d6bac9282d 6b5c07b4e9e900d8 1ec66368838c8a08 9042367d80d1fbab 400283564a3c3799 4a00be0466bca75e".split(),
        "poseidon2_rc_first_internal": "4adf842aa75d4316",
        "poseidon2_perm": [
            {"in": "iota", "out": "01eaef96bdf1c0c1 1f0d2cc525b2540c 6282c1dfe1e0358d e780d721f698e1e6 280c0b6f753d833b 1b942dd5023156ab 43f0df3fcccb8398 e8e8190585489025 56bdbf72f77ada22 7911c32bf9dcd705 ec467926508fbe67 6a50450ddf85a6ed".split()},
        ],
        "goldilocks": {"p": "ffffffff00000001", "generator": 7, "two_adic_generator": "1753635133440165772",
                        "omega_64": str(2**39), "omega_2^17": "12380578893860276750"},
    }


def commit_small():
    rng = random.Random(0x6D7032)
    cases = []
    # (ncols, log_n, rate_bits, cap_height, from_coeffs)
    shapes = [(4, 3, 1, 1, False), (5, 3, 1, 1, False), (9, 2, 2, 0, False), (3, 1, 3, 4, False),
              (1, 0, 3, 2, False), (17, 2, 1, 2, True), (12, 3, 3, 4, False), (8, 3, 0, 1, True)]
    for kind in (0, 1):
        for (c, ln, r, cap, fc) in shapes:
            n = 1 << ln
            cols = [[rng.randrange(R.P) for _ in range(n)] for _ in range(c)]
            if c == 9:  # non-canonical inputs (>= p) must be accepted and canonicalised
                cols[0][0] = R.P + 5
                cols[1][1] = 2**64 - 1
            out = R.commit(cols, r, cap, kind, fc)
            cases.append({"hash_kind": kind, "ncols": c, "log_n": ln, "rate_bits": r, "cap_height": cap,
                          "from_coeffs": fc, "cols": hx2(cols), "coeffs": hx2(out["coeffs"]),
                          "leaves": hx2(out["leaves"]), "digests": hx2(out["digests"]), "cap": hx2(out["cap"])})
    return {"generator": "tests/pyref.py (pure Python, by definition)", "cases": cases}


def merkle_small():
    rng = random.Random(0x6D7033)
    cases = []
    for kind in (0, 1):
        # shapes of mp2-common/src/serialization/circuit_data_serialization.rs:344-370 that are valid
        # powers of two, plus leaf lengths around the rate/no-op boundaries
        for (nl, ll, cap) in [(16, 7, 0), (32, 3, 3), (8, 4, 0), (8, 5, 1), (4, 8, 2), (4, 9, 0), (2, 16, 1),
                              (1, 20, 0), (16, 32, 4), (64, 1, 0)]:
            leaves = [[rng.randrange(R.P) for _ in range(ll)] for _ in range(nl)]
            digests, cap_v, trees = R.merkle_new(leaves, cap, kind)
            proofs = {str(i): hx2(R.merkle_prove_from_tree(trees, nl, cap, i)) for i in {0, nl // 3, nl - 1}}
            cases.append({"hash_kind": kind, "nleaves": nl, "leaf_len": ll, "cap_height": cap,
                          "leaves": hx2(leaves), "digests": hx2(digests), "cap": hx2(cap_v), "proofs": proofs})
        # circuit-set shape: 42 digests padded to 64 leaves with [0], cap 0 (circuit_set.rs:173-191, :296-371)
        leaves = [[rng.randrange(R.P) for _ in range(4)] for _ in range(42)] + [[0]] * 22
        digests, cap_v, trees = R.merkle_new(leaves, 0, kind)
        proofs = {str(i): hx2(R.merkle_prove_from_tree(trees, 64, 0, i)) for i in (0, 17, 41)}
        cases.append({"hash_kind": kind, "nleaves": 64, "leaf_len": "ragged(4|1)", "cap_height": 0,
                      "leaves": hx2(leaves), "digests": hx2(digests), "cap": hx2(cap_v), "proofs": proofs})
    # hash_pad(&[]) domain separator and the circuit-digest formula inputs (circuit_set.rs:136-158)
    misc = {"hash_pad_empty": {str(k): hx(R.hash_pad([], k)) for k in (0, 1)},
            "hash_no_pad_empty": hx(R.hash_no_pad([], 0))}
    return {"generator": "tests/pyref.py", "cases": cases, "misc": misc}


def fri_small():
    rng = random.Random(0x6D7036)
    cases = []
    for kind, degree_bits, widths, arities in ((0, 5, (3, 2), [4]), (1, 6, (2, 3, 1), [4])):
        n = 1 << degree_bits
        oracles = [[[rng.randrange(R.P) for _ in range(n)] for _ in range(w)] for w in widths]
        zeta, gzeta, alpha = ((rng.randrange(R.P), rng.randrange(R.P)) for _ in range(3))
        polys0 = [(o, p) for o, w in enumerate(widths) for p in range(w)]
        polys1 = [(len(widths) - 1, 0), (0, 1)]
        batches = [(zeta, polys0), (gzeta, polys1)]
        final = R.fri_combine([(z, [oracles[o][p] for o, p in polys]) for z, polys in batches], alpha)
        betas = [(rng.randrange(R.P), rng.randrange(R.P)) for _ in arities]
        padded = final + [(0, 0)] * (n * 7)
        layers, rest = R.fri_committed_trees(padded, arities, betas, 2, kind, 3)
        cases.append({"hash_kind": kind, "degree_bits": degree_bits, "rate_bits": 3, "cap_height": 2,
                      "oracles": [hx2(o) for o in oracles], "points": [hx(zeta), hx(gzeta)], "alpha": hx(alpha),
                      "batches": [polys0, polys1], "arity_bits": arities, "betas": hx2(betas),
                      "final_poly": hx2(final), "layer_caps": [hx2(l["cap"]) for l in layers],
                      "layer_digests_xor": [hx([__import__("functools").reduce(lambda a, b: a ^ b, col) for col in zip(*l["digests"])]) if l["digests"] else [] for l in layers],
                      "final_coeffs": hx2(rest)})
    return {"generator": "tests/pyref.py (fri_combine, fri_committed_trees: by definition)", "cases": cases}


def config1_caps():
    import oracle as O
    out = {"generator": "oracle/mp2_oracle.c (self-golden, regression pin)", "cases": []}
    for kind in (0, 1):
        for (name, c, fc, seed) in [("wires_135_from_values", 135, False, 0x6D7032),
                                    ("zs_pp_20_from_values", 20, False, 0x6D7034),
                                    ("quotient_16_from_coeffs", 16, True, 0x6D7035)]:
            cols = field_elems(seed, (c, 1 << 14))
            res = O.commit(cols, 3, 4, kind, fc, want_leaves=False)
            out["cases"].append({"name": name, "hash_kind": kind, "ncols": c, "log_n": 14, "rate_bits": 3,
                                 "cap_height": 4, "from_coeffs": fc, "seed": seed,
                                 "cap": hexlist(res["cap"]),
                                 "digest_xor": hexlist(np.bitwise_xor.reduce(res["digests"], axis=0)),
                                 "coeff_xor": hexlist(np.bitwise_xor.reduce(res["coeffs"], axis=1)[:8])})
    return out


if __name__ == "__main__":
    only = sys.argv[1:]
    for name, fn in (("kats", kats), ("commit_small", commit_small), ("merkle_small", merkle_small),
                     ("fri_small", fri_small), ("config1_caps", config1_caps)):
        if only and name not in only:
            continue
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(fn(), f, indent=0, separators=(",", ":"))
        print("wrote", name)
