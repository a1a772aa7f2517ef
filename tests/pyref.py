"""Second, independent restatement of the hot path in pure Python integers (slow; tiny sizes).

Test infrastructure only.  It shares NO code with ``oracle/mp2_oracle.c``: transforms are evaluated
by definition (Horner, O(n^2)), the Merkle tree is built top-down from SURVEY.md A.4's recursive
description, and ``prove`` walks an explicit tree instead of using the closed-form indices.  Used to
cross-validate the C oracle and to generate ``tests/golden/*.json`` (``tests/golden/make_golden.py``).
Reference anchors: recursion-framework/src/universal_verifier_gadget/circuit_set.rs:173-237,
mp2-common/src/poseidon.rs:136-172.
"""
from __future__ import annotations

P = 2**64 - 2**32 + 1
M32 = 0xFFFFFFFF


def root_of_unity(log_n: int) -> int:
    return pow(pow(7, (P - 1) >> 32, P), 1 << (32 - log_n), P)


# ---------------- Poseidon constants: ChaCha8Rng::seed_from_u64(0).gen_range(0..p) -------------
def _rotl(x, r):
    return ((x << r) | (x >> (32 - r))) & M32


def _chacha8_block(key, ctr):
    init = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key) + [ctr & M32, ctr >> 32, 0, 0]
    x = list(init)

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & M32; x[d] = _rotl(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & M32; x[b] = _rotl(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & M32; x[d] = _rotl(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & M32; x[b] = _rotl(x[b] ^ x[c], 7)

    for _ in range(4):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(a + b) & M32 for a, b in zip(x, init)]


def poseidon_round_constants():
    st, key = 0, []
    for _ in range(8):
        st = (st * 6364136223846793005 + 11634580027462260723) % 2**64
        xs = (((st >> 18) ^ st) >> 27) & M32
        rot = st >> 59
        key.append(((xs >> rot) | (xs << (32 - rot))) & M32 if rot else xs)
    out, ctr, words = [], 0, []
    while len(out) < 360:
        if len(words) < 2:
            words += _chacha8_block(key, ctr); ctr += 1
        v = words[0] | (words[1] << 32); words = words[2:]
        m = v * P
        if (m & (2**64 - 1)) <= P - 1:
            out.append(m >> 64)
    return out


POS_CIRC = [17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]
POS_DIAG = [8] + [0] * 11
_POS_RC = None


def poseidon(state):
    global _POS_RC
    if _POS_RC is None:
        _POS_RC = poseidon_round_constants()
    s = [x % P for x in state]
    for r in range(30):
        s = [(s[i] + _POS_RC[12 * r + i]) % P for i in range(12)]
        if r < 4 or r >= 26:
            s = [pow(x, 7, P) for x in s]
        else:
            s[0] = pow(s[0], 7, P)
        s = [(sum(s[(i + row) % 12] * POS_CIRC[i] for i in range(12)) + s[row] * POS_DIAG[row]) % P
             for row in range(12)]
    return s


# ---------------- Poseidon2 (HL Goldilocks t=12): Grain LFSR constants ----------------
def poseidon2_round_constants():
    bits = []
    for val, width in ((1, 2), (0, 4), (64, 12), (12, 12), (8, 10), (22, 10)):
        bits += [(val >> i) & 1 for i in range(width - 1, -1, -1)]
    bits += [1] * 30

    def step():
        nb = bits[62] ^ bits[51] ^ bits[38] ^ bits[23] ^ bits[13] ^ bits[0]
        bits.pop(0); bits.append(nb)
        return nb

    for _ in range(160):
        step()

    def rbit():
        while True:
            a = step(); c = step()
            if a:
                return c

    out = []
    while len(out) < 118:
        v = 0
        for _ in range(64):
            v = (v << 1) | rbit()
        if v < P:
            out.append(v)
    return out


P2_DIAG = [0xc3b6c08e23ba9300, 0xd84b5de94a324fb6, 0x0d0c371c5b35b84f, 0x7964f570e7188037,
           0x5daf18bbd996604b, 0x6743bc47b9595257, 0x5528b9362c59bb70, 0xac45e25b7127b68b,
           0xa2077d7dfbb606b5, 0xf3faac6faee378ae, 0x0c6388b51545e883, 0xd27dbb6944917b60]
_M4 = [[5, 7, 1, 3], [4, 6, 1, 1], [1, 3, 5, 7], [1, 1, 4, 6]]
_P2_RC = None


def _p2_ext(s):
    y = []
    for c in range(0, 12, 4):
        y += [sum(_M4[r][k] * s[c + k] for k in range(4)) % P for r in range(4)]
    return [(y[i] + sum(y[4 * k + (i % 4)] for k in range(3))) % P for i in range(12)]


def poseidon2(state):
    global _P2_RC
    if _P2_RC is None:
        _P2_RC = poseidon2_round_constants()
    rc = list(_P2_RC)
    s = _p2_ext([x % P for x in state])
    for _ in range(4):
        s = _p2_ext([pow((s[i] + rc[i]) % P, 7, P) for i in range(12)]); rc = rc[12:]
    for _ in range(22):
        s[0] = pow((s[0] + rc[0]) % P, 7, P); rc = rc[1:]
        tot = sum(s) % P
        s = [(s[i] * P2_DIAG[i] + tot) % P for i in range(12)]
    for _ in range(4):
        s = _p2_ext([pow((s[i] + rc[i]) % P, 7, P) for i in range(12)]); rc = rc[12:]
    return s


def permute(state, kind=0):
    return poseidon2(state) if kind == 1 else poseidon(state)


# ---------------- sponge wrapper (A.5) ----------------
def hash_no_pad(x, kind=0):
    st = [0] * 12
    for off in range(0, len(x), 8):
        chunk = x[off:off + 8]
        st[:len(chunk)] = [v % P for v in chunk]
        st = permute(st, kind)
    return st[:4]


def hash_pad(x, kind=0):
    x = list(x) + [1]
    while (len(x) + 1) % 8:
        x.append(0)
    return hash_no_pad(x + [1], kind)


def hash_or_noop(x, kind=0):
    if len(x) <= 4:
        return [v % P for v in x] + [0] * (4 - len(x))
    return hash_no_pad(x, kind)


def two_to_one(a, b, kind=0):
    return permute([v % P for v in a] + [v % P for v in b] + [0] * 4, kind)[:4]


# ---------------- transforms by definition (A.2) ----------------
def horner(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % P
    return acc


def ifft(values):
    """coeffs c with values[i] = sum_j c_j w^(ij): c_j = n^-1 sum_i values[i] w^(-ij)."""
    n = len(values)
    w_inv = pow(root_of_unity(n.bit_length() - 1), P - 2, P)
    n_inv = pow(n, P - 2, P)
    return [horner(values, pow(w_inv, j, P)) * n_inv % P for j in range(n)]


def coset_lde(coeffs, rate_bits, shift=7):
    N = len(coeffs) << rate_bits
    w = root_of_unity(N.bit_length() - 1)
    return [horner(coeffs, shift * pow(w, i, P) % P) for i in range(N)]


def bitrev(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


# ---------------- Merkle tree (A.4), explicit tree + plonky2 digest layout ----------------
def _fill(buf, lo, hi, leaves, kind, nodes, path):
    """buf[lo:hi] is this subtree's digest slice; returns its root digest."""
    if hi == lo:
        d = hash_or_noop(leaves[0], kind)
        nodes[path] = d
        return d
    mid = (lo + hi) // 2
    half = len(leaves) // 2
    ld = _fill(buf, lo, mid - 1, leaves[:half], kind, nodes, path + "0")
    rd = _fill(buf, mid + 1, hi, leaves[half:], kind, nodes, path + "1")
    buf[mid - 1] = ld   # last digest of the left half
    buf[mid] = rd       # first digest of the right half
    d = two_to_one(ld, rd, kind)
    nodes[path] = d
    return d


def merkle_new(leaves, cap_height, kind=0):
    n = len(leaves)
    log_n = n.bit_length() - 1
    if n == 0 or 1 << log_n != n or cap_height > log_n:
        raise ValueError("MerkleTree::new: bad nleaves / cap_height")
    ncap = 1 << cap_height
    sub = n // ncap
    per = 2 * (sub - 1)
    digests = [None] * (ncap * per)
    cap, trees = [], []
    for s in range(ncap):
        nodes = {}
        cap.append(_fill(digests, s * per, (s + 1) * per, leaves[s * sub:(s + 1) * sub], kind, nodes, ""))
        trees.append(nodes)
    return digests, cap, trees


def merkle_prove_from_tree(trees, n, cap_height, leaf_index):
    """Siblings bottom-up, read from the explicit node map (no closed-form indices)."""
    h = (n.bit_length() - 1) - cap_height
    nodes = trees[leaf_index >> h]
    path = format(leaf_index & ((1 << h) - 1), "0%db" % h) if h else ""
    sib = []
    for d in range(h, 0, -1):
        p = path[:d]
        sib.append(nodes[p[:-1] + ("1" if p[-1] == "0" else "0")])
    return sib


# ---------------- PolynomialBatch ----------------
def commit(cols, rate_bits, cap_height, kind=0, from_coeffs=False):
    n = len(cols[0])
    coeffs = [[v % P for v in c] if from_coeffs else ifft(c) for c in cols]
    lde = [coset_lde(c, rate_bits) for c in coeffs]
    N = n << rate_bits
    bits = N.bit_length() - 1
    leaves = [[lde[c][bitrev(i, bits)] for c in range(len(cols))] for i in range(N)]
    digests, cap, _ = merkle_new(leaves, cap_height, kind)
    return {"coeffs": coeffs, "leaves": leaves, "digests": digests, "cap": cap}


# ---------------- FRI commit phase (plonky2 fri_committed_trees), by definition ----------------
def ext_mul(a, b):
    """GF(p^2) = F[X]/(X^2 - 7)."""
    return ((a[0] * b[0] + 7 * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def ext_pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = ext_mul(r, a)
        a = ext_mul(a, a)
        e >>= 1
    return r


def ext_horner(coeffs, x):
    acc = (0, 0)
    for c in reversed(coeffs):
        acc = ext_mul(acc, x)
        acc = ((acc[0] + c[0]) % P, (acc[1] + c[1]) % P)
    return acc


def ext_inv(a):
    """1/(a0 + a1 X) = (a0 - a1 X) / (a0^2 - 7 a1^2)."""
    norm_inv = pow((a[0] * a[0] - 7 * a[1] * a[1]) % P, P - 2, P)
    return (a[0] * norm_inv % P, (P - a[1]) * norm_inv % P)


def fri_combined_eval(batches, alpha, x):
    """prove_openings' final polynomial evaluated at the extension point x, from its defining formula
        sum_i alpha^(k_i) (F_i(x) - F_i(z_i)) / (x - z_i),  F_i = sum_j alpha^j f_ij,
    where batch i is multiplied by alpha^(number of polynomials in the batches after it... that were reduced
    after it): k_i = sum of the sizes of batches i+1.. (ReducingFactor::shift_poly).
    ``batches``: [(z, [coefficient list, ...]), ...]."""
    total = (0, 0)
    sizes = [len(polys) for _, polys in batches]
    for i, (z, polys) in enumerate(batches):
        def big_f(pt):
            acc, pw = (0, 0), (1, 0)
            for f in polys:
                v = ext_horner([(c % P, 0) for c in f], pt)
                t = ext_mul(v, pw)
                acc = ((acc[0] + t[0]) % P, (acc[1] + t[1]) % P)
                pw = ext_mul(pw, alpha)
            return acc
        fx, fz = big_f(x), big_f(z)
        num = ((fx[0] - fz[0]) % P, (fx[1] - fz[1]) % P)
        den = ext_inv(((x[0] - z[0]) % P, (x[1] - z[1]) % P))
        term = ext_mul(ext_mul(num, den), ext_pow(alpha, sum(sizes[i + 1:])))
        total = ((total[0] + term[0]) % P, (total[1] + term[1]) % P)
    return total


def fri_combine(batches, alpha):
    """prove_openings' final polynomial as a list of n extension coefficients, from the definition: for each batch
    the composition F_i = sum_j alpha^j f_ij, its exact quotient by (X - z_i) (the remainder F_i(z_i) is dropped;
    one zero coefficient pads the quotient back to n), and the sum of the quotients weighted by alpha^(number of
    polynomials in the later batches)."""
    n = len(batches[0][1][0])
    sizes = [len(polys) for _, polys in batches]
    total = [(0, 0)] * n
    for i, (z, polys) in enumerate(batches):
        comp, pw = [(0, 0)] * n, (1, 0)
        for f in polys:
            comp = [((c[0] + v % P * pw[0]) % P, (c[1] + v % P * pw[1]) % P) for c, v in zip(comp, f)]
            pw = ext_mul(pw, alpha)
        # long division by (X - z): q_(n-2) = a_(n-1), q_(k-1) = a_k + z q_k
        q = [(0, 0)] * n
        carry = (0, 0)
        for k in range(n - 1, 0, -1):
            carry = ((comp[k][0] + ext_mul(carry, z)[0]) % P, (comp[k][1] + ext_mul(carry, z)[1]) % P)
            q[k - 1] = carry
        weight = ext_pow(alpha, sum(sizes[i + 1:]))
        total = [((t[0] + ext_mul(x, weight)[0]) % P, (t[1] + ext_mul(x, weight)[1]) % P) for t, x in zip(total, q)]
    return total


def fri_committed_trees(coeffs, arity_bits_list, betas, cap_height, kind=0, rate_bits=3):
    """coeffs: list of ext pairs (zero-padded LDE length).  Every layer is computed from the definition:
    values[i] = P(shift * w^i) by Horner in the extension field; the fold is
    P(x) = sum_i x^i P_i(x^r)  ->  sum_i beta^i P_i(x)."""
    shift, out = 7, []
    for ab, beta in zip(arity_bits_list, betas):
        m = len(coeffs)
        bits = m.bit_length() - 1
        w = root_of_unity(bits)
        values = [ext_horner(coeffs, (shift * pow(w, i, P) % P, 0)) for i in range(m)]
        rev = [values[bitrev(i, bits)] for i in range(m)]
        arity = 1 << ab
        leaves = [[x for v in rev[i:i + arity] for x in v] for i in range(0, m, arity)]
        digests, cap, _ = merkle_new(leaves, min(cap_height, (m >> ab).bit_length() - 1), kind)
        out.append({"leaves": leaves, "digests": digests, "cap": cap})
        folded = []
        for j in range(m >> ab):
            acc, bp = (0, 0), (1, 0)
            for t in range(arity):
                term = ext_mul(coeffs[(j << ab) + t], bp)
                acc = ((acc[0] + term[0]) % P, (acc[1] + term[1]) % P)
                bp = ext_mul(bp, beta)
            folded.append(acc)
        coeffs = folded
        shift = pow(shift, arity, P)
    return out, coeffs[:len(coeffs) >> rate_bits]


# ---------------- FRI verifier (plonky2 fri/verifier.rs, iop/challenger.rs), by definition ----------------
class Challenger:
    """Overwrite-mode duplex sponge of the Fiat-Shamir transcript; outputs are popped from the end."""

    def __init__(self, kind=0):
        self.kind, self.state, self.inp, self.out = kind, [0] * 12, [], []

    def observe(self, xs):
        for x in xs:
            self.out = []
            self.inp.append(int(x) % P)
            if len(self.inp) == 8:
                self.duplex()

    def duplex(self):
        for i, x in enumerate(self.inp):
            self.state[i] = x
        self.inp = []
        self.state = permute(self.state, self.kind)
        self.out = list(self.state[:8])

    def challenge(self):
        if self.inp or not self.out:
            self.duplex()
        return self.out.pop()

    def ext_challenge(self):
        a = self.challenge()
        return (a, self.challenge())


def verify_merkle_proof_to_cap(leaf, index, cap, siblings, kind=0):
    cur = hash_or_noop(list(leaf), kind)
    for sib in siblings:
        cur = two_to_one(cur, list(sib), kind) if index & 1 == 0 else two_to_one(list(sib), cur, kind)
        index >>= 1
    return list(cur) == list(cap[index])


def _ext_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def _ext_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def _reduce_with_alpha(values, alpha):
    """ReducingFactor::reduce: sum_j alpha^j v_j."""
    acc = (0, 0)
    for v in reversed(values):
        acc = _ext_add(ext_mul(acc, alpha), v)
    return acc


def fri_compute_evaluation(x, x_index_within_coset, arity_bits, evals, beta):
    """Interpolate {(x g^i, P(x g^i))} over the coset of x and evaluate at beta (Lagrange, by definition)."""
    arity = 1 << arity_bits
    g = root_of_unity(arity_bits)
    ordered = [evals[bitrev(i, arity_bits)] for i in range(arity)]
    coset_start = x * pow(g, arity - bitrev(x_index_within_coset, arity_bits), P) % P
    pts = [(coset_start * pow(g, i, P) % P, 0) for i in range(arity)]
    total = (0, 0)
    for i in range(arity):
        num, den = (1, 0), (1, 0)
        for j in range(arity):
            if j != i:
                num = ext_mul(num, _ext_sub(beta, pts[j]))
                den = ext_mul(den, _ext_sub(pts[i], pts[j]))
        total = _ext_add(total, ext_mul(ordered[i], ext_mul(num, ext_inv(den))))
    return total


def verify_fri_proof(batches, openings, initial_caps, proof, challenger, degree_bits, arity_bits_list,
                     rate_bits=3, pow_bits=16, kind=0):
    """plonky2 ``verify_fri_proof`` with the challenges re-derived from the transcript.

    batches: [(point, [(oracle_index, polynomial_index), ...])]; openings: per batch the claimed values;
    initial_caps: per oracle its cap; proof: dict(caps, final_poly, pow_witness, rounds=[dict(initial=[(row, siblings)],
    steps=[(evals, siblings)])]).  ``challenger`` has already observed whatever precedes FRI (caps, openings).
    Returns None, raises AssertionError with the failing check otherwise."""
    alpha = challenger.ext_challenge()
    betas = []
    for cap in proof["caps"]:
        challenger.observe([x for h in cap for x in h])
        betas.append(challenger.ext_challenge())
    challenger.observe([x for c in proof["final_poly"] for x in c])
    challenger.observe([proof["pow_witness"]])
    pow_response = challenger.challenge()
    assert 64 - pow_response.bit_length() >= pow_bits, "proof of work"
    log_n = degree_bits + rate_bits
    n = 1 << log_n
    assert len(proof["final_poly"]) == 1 << (degree_bits - sum(arity_bits_list))
    x_indices = [challenger.challenge() % n for _ in proof["rounds"]]
    reduced_openings = [_reduce_with_alpha([tuple(v) for v in vals], alpha) for vals in openings]
    for x_index, rnd in zip(x_indices, proof["rounds"]):
        for (row, sib), cap in zip(rnd["initial"], initial_caps):
            assert verify_merkle_proof_to_cap(row, x_index, cap, sib, kind), "initial tree proof"
        subgroup_x = 7 * pow(root_of_unity(log_n), bitrev(x_index, log_n), P) % P
        # fri_combine_initial
        total, count = (0, 0), 0
        for (point, polys), red in zip(batches, reduced_openings):
            evals = [(int(rnd["initial"][o][0][p]) % P, 0) for o, p in polys]
            num = _ext_sub(_reduce_with_alpha(evals, alpha), red)
            den = _ext_sub((subgroup_x, 0), tuple(point))
            total = ext_mul(total, ext_pow(alpha, len(polys)))      # alpha.shift(sum): count of THIS batch's reduce
            total = _ext_add(total, ext_mul(num, ext_inv(den)))
        old_eval = total
        for i, ab in enumerate(arity_bits_list):
            evals, sib = rnd["steps"][i]
            evals = [tuple(int(v) for v in e) for e in evals]
            coset_index, within = x_index >> ab, x_index & ((1 << ab) - 1)
            assert evals[within] == old_eval, "consistency with the previous layer (layer %d)" % i
            old_eval = fri_compute_evaluation(subgroup_x, within, ab, evals, betas[i])
            assert verify_merkle_proof_to_cap([x for e in evals for x in e], coset_index, proof["caps"][i], sib, kind), \
                "layer tree proof"
            subgroup_x = pow(subgroup_x, 1 << ab, P)
            x_index = coset_index
        final = ext_horner([tuple(int(v) for v in c) for c in proof["final_poly"]], (subgroup_x, 0))
        assert final == old_eval, "final polynomial evaluation"
