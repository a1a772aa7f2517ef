"""Reference FRI prover for the tests: the CPU oracle's pieces (commitments, the prove_openings combination, the
commit phase, Merkle paths) driven by pyref's challenger, producing the proof in the plain-dict form that
pyref.verify_fri_proof consumes.  Mirrors plonky2's prove_openings -> fri_proof flow (fri/oracle.rs, fri/prover.rs)."""
import numpy as np

import pyref
from util import P


def arity_schedule(degree_bits, rate_bits=3, cap_height=4, arity_bits=4, final_poly_bits=5):
    out, db = [], degree_bits
    while db > final_poly_bits and db + rate_bits - cap_height > arity_bits:
        out.append(arity_bits)
        db -= arity_bits
    return out


def plonky2_instance(widths, zeta, gzeta):
    """Every polynomial at zeta; the first two polynomials of oracle min(2, last) again at g*zeta."""
    zo = min(2, len(widths) - 1)
    return [(zeta, [(o, p) for o, w in enumerate(widths) for p in range(w)]),
            (gzeta, [(zo, p) for p in range(min(2, widths[zo]))])]


def transcript_head(challenger_observe, caps, openings):
    """What precedes FRI in the transcript here: the oracles' caps, then the opened values per batch."""
    for cap in caps:
        challenger_observe([int(x) for h in cap for x in h])
    for vals in openings:
        challenger_observe([int(x) for v in vals for x in v])


def oracle_openings(coeff_sets, batches):
    return [[pyref.ext_horner([(int(c), 0) for c in coeff_sets[o][p]], tuple(int(v) for v in z)) for o, p in polys]
            for z, polys in batches]


def oracle_fri_proof(oracle, coeff_sets, batches, degree_bits, kind, rate_bits=3, cap_height=4, pow_bits=16,
                     num_query_rounds=28):
    """coeff_sets: per oracle (ncols, n) canonical coefficients.  Returns (commitments, openings, proof dict)."""
    commits = [oracle.commit(c, rate_bits, cap_height, kind, from_coeffs=True) for c in coeff_sets]
    openings = oracle_openings(coeff_sets, batches)
    ch = pyref.Challenger(kind)
    transcript_head(ch.observe, [c["cap"] for c in commits], openings)
    alpha = ch.ext_challenge()
    final = oracle.fri_combine([(np.array(z, dtype=np.uint64), [coeff_sets[o][p] for o, p in polys]) for z, polys in batches],
                               np.array(alpha, dtype=np.uint64))
    n = 1 << degree_bits
    arities = arity_schedule(degree_bits, rate_bits, cap_height)
    coeffs = np.zeros((n << rate_bits, 2), dtype=np.uint64)
    coeffs[:n] = final
    values = oracle.coset_fft_ext(coeffs, 7)
    shift, trees, caps = 7, [], []
    for ab in arities:
        leaves = oracle.fri_layer_leaves(values, ab)
        digests, cap = oracle.merkle_new(leaves, cap_height, kind)
        trees.append((leaves, digests))
        caps.append(cap)
        ch.observe([int(x) for h in cap for x in h])
        beta = ch.ext_challenge()
        coeffs = oracle.fri_fold(coeffs, ab, np.array(beta, dtype=np.uint64))
        shift = pow(shift, 1 << ab, P)
        values = oracle.coset_fft_ext(coeffs, shift)
    final_poly = coeffs[:coeffs.shape[0] >> rate_bits]
    ch.observe([int(x) for c in final_poly for x in c])
    state = list(ch.state)
    for i, x in enumerate(ch.inp):
        state[i] = x
    witness = oracle.fri_pow(np.array(state, dtype=np.uint64), len(ch.inp), pow_bits, kind)
    ch.observe([witness])
    response = ch.challenge()
    assert 64 - response.bit_length() >= pow_bits
    N = n << rate_bits
    rounds = []
    for _ in range(num_query_rounds):
        x = ch.challenge() % N
        initial = [(c["leaves"][x], oracle.merkle_prove(c["digests"], N, cap_height, x)) for c in commits]
        steps, nl = [], N
        for (leaves, digests), ab in zip(trees, arities):
            x >>= ab
            nl >>= ab
            steps.append((leaves[x].reshape(-1, 2), oracle.merkle_prove(digests, nl, cap_height, x)))
        rounds.append({"initial": initial, "steps": steps})
    proof = {"caps": caps, "final_poly": final_poly, "pow_witness": witness, "rounds": rounds}
    return commits, openings, proof


def load_fri_golden():
    """tests/golden/fri_small.json decoded to numpy: list of dicts with oracles [(w, n) arrays], batches
    [(point (2,), [(oracle_index, polynomial_index)])], alpha (2,), arity_bits, betas (k, 2), final_poly (n, 2),
    layer_caps [(4, 4)], layer_digests_xor [(4,)], final_coeffs (m, 2)."""
    import json
    import os

    from util import GOLDEN_DIR, unhex

    with open(os.path.join(GOLDEN_DIR, "fri_small.json")) as f:
        raw = json.load(f)["cases"]
    out = []
    for c in raw:
        oracles = [unhex([x for col in o for x in col]).reshape(len(o), -1) for o in c["oracles"]]
        out.append({
            "hash_kind": c["hash_kind"], "degree_bits": c["degree_bits"], "rate_bits": c["rate_bits"],
            "cap_height": c["cap_height"], "oracles": oracles,
            "batches": [(unhex(z), [tuple(p) for p in polys]) for z, polys in zip(c["points"], c["batches"])],
            "alpha": unhex(c["alpha"]), "arity_bits": list(c["arity_bits"]),
            "betas": unhex([x for b in c["betas"] for x in b]).reshape(-1, 2),
            "final_poly": unhex([x for r in c["final_poly"] for x in r]).reshape(-1, 2),
            "layer_caps": [unhex([x for h in cap for x in h]).reshape(-1, 4) for cap in c["layer_caps"]],
            "layer_digests_xor": [unhex(d) for d in c["layer_digests_xor"]],
            "final_coeffs": unhex([x for r in c["final_coeffs"] for x in r]).reshape(-1, 2),
        })
    return out
