"""CPU ORACLE (test infrastructure, NOT the product): plonky2 0.2.2 `compute_quotient_polys` restated
(plonk/prover.rs + plonk/vanishing_poly.rs `eval_vanishing_poly_base_batch`).

The prover evaluates, at every point x_i = g * w_{8n}^i of the coset LDE the three committed batches already hold
(`get_lde_values(i, step)` = leaf bitrev(i * step)):

    terms_i = [ L_0(x)(Z_c(x) - 1) ]_c  ++  [ check_partial_products(...) ]_c  ++  gate constraints
    q_c(x_i) = reduce_with_powers(terms_i, alpha_c) / Z_H(x_i)

then `coset_ifft` turns each q_c into coefficients and cuts it into `quotient_degree_factor` chunks of n, which are
committed with `from_coeffs`.  Gate set restated here (the staged subset of
mp2-common/src/serialization/circuit_data_serialization.rs:234-266): ArithmeticGate, ConstantGate, PublicInputGate,
NoopGate, PoseidonGate, ArithmeticExtensionGate, MulExtensionGate, BaseSumGate<B>, ReducingGate, ReducingExtensionGate, RandomAccessGate,
ExponentiationGate, PoseidonMdsGate, CosetInterpolationGate behind plonky2's selector filters; no lookups, no blinding (the reference never enables zero_knowledge).

Pinned by definition, not by the Rust prover (absent): tests/plonk_ref.py restates the VERIFIER's
`eval_vanishing_poly` + final identity, and the quotients computed here must pass it at random points
(tests/test_quotient_oracle.py) -- prove, then verify, as every `run_circuit` test of the reference does.
`desc` is any object with the attributes of tests/plonk_ref.Circuit / mapreduce_plonky2_b200.quotient.CircuitDesc.
"""
from __future__ import annotations

import numpy as np

from . import oracle as O

P = O.P
UNUSED_SELECTOR = (1 << 32) - 1
COSET_SHIFT = 7


_POS_CIRC = (17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20)
_POS_RC = None


def _poseidon_gate(w):
    """PoseidonGate::eval_unfiltered (plonky2 gates/poseidon.rs), naive round structure: wires input 0..11, output
    12..23, swap 24, delta 25..28, S-box inputs of full rounds 1..3 at 29.., of the 22 partial rounds at 65.., of the
    last four full rounds at 87..; 123 constraints (swap bit, 4 deltas, 36 + 22 + 48 S-box inputs, 12 outputs)."""
    global _POS_RC
    if _POS_RC is None:
        _POS_RC = [int(v) for v in O.poseidon_round_constants()]
    rc = _POS_RC
    sw = w[24]
    out = [sw * (sw - 1) % P]
    st = list(w[:12])
    for i in range(4):
        out.append((sw * (w[i + 4] - w[i]) - w[25 + i]) % P)
        st[i], st[i + 4] = (w[i] + w[25 + i]) % P, (w[i + 4] - w[25 + i]) % P
    for r in range(30):
        st = [(st[i] + rc[12 * r + i]) % P for i in range(12)]
        if r < 4 or r >= 26:
            if r:
                base = 29 + 12 * (r - 1) if r < 4 else 87 + 12 * (r - 26)
                out.extend((st[i] - w[base + i]) % P for i in range(12))
                st = list(w[base:base + 12])
            st = [pow(v, 7, P) for v in st]
        else:
            out.append((st[0] - w[65 + r - 4]) % P)
            st[0] = pow(w[65 + r - 4], 7, P)
        st = [(sum(st[(i + row) % 12] * _POS_CIRC[i] for i in range(12)) + (8 * st[0] if row == 0 else 0)) % P
              for row in range(12)]
    out.extend((st[i] - w[12 + i]) % P for i in range(12))
    return out


_CI_TABLES = {}


def _coset_interpolation_gate(w, bits, degree):
    """CosetInterpolationGate::eval_unfiltered_base_one (plonky2 gates/coset_interpolation.rs), D = 2: barycentric
    interpolation of the 2^bits extension values (wires 1 + 2k) at shifted_evaluation_point, folded point by point --
    eval <- eval (x - w^k) + value_k prod weight_k, prod <- prod (x - w^k) -- and cut after `degree` points, then every
    `degree - 1`, where the running (eval, prod) pair is pinned to intermediate wires.  The barycentric weight of the
    point w^k of the order-m subgroup is 1 / prod_{j != k} (w^k - w^j) = w^k / m (the derivative of X^m - 1)."""
    m = 1 << bits
    if bits not in _CI_TABLES:
        g = O.root_of_unity(bits)
        dom = [pow(g, k, P) for k in range(m)]
        m_inv = pow(m, P - 2, P)
        _CI_TABLES[bits] = (dom, [x * m_inv % P for x in dom])
    dom, wts = _CI_TABLES[bits]
    ni = (m - 2) // (degree - 1)
    at_point, at_value, at_inter = 1 + 2 * m, 3 + 2 * m, 5 + 2 * m
    at_shifted = at_inter + 4 * ni
    x0, x1 = w[at_shifted], w[at_shifted + 1]
    cons = [(w[at_point] - x0 * w[0]) % P, (w[at_point + 1] - x1 * w[0]) % P]
    e0, e1, p0, p1 = 0, 0, 1, 0
    cut = degree                      # index of the first point of the next run
    for k in range(m):
        if k == cut:                  # pin the running pair to the next intermediate wires and continue from them
            i = (k - degree) // (degree - 1)
            ie, ip = at_inter + 2 * i, at_inter + 2 * (ni + i)
            cons += [(w[ie] - e0) % P, (w[ie + 1] - e1) % P, (w[ip] - p0) % P, (w[ip + 1] - p1) % P]
            e0, e1, p0, p1 = w[ie], w[ie + 1], w[ip], w[ip + 1]
            cut += degree - 1
        t0 = (x0 - dom[k]) % P        # term = x - w^k  (its second component is x1)
        v0, v1 = w[1 + 2 * k], w[2 + 2 * k]
        q0, q1 = p0 * wts[k] % P, p1 * wts[k] % P
        e0, e1 = ((e0 * t0 + 7 * e1 * x1) + (v0 * q0 + 7 * v1 * q1)) % P, ((e0 * x1 + e1 * t0) + (v0 * q1 + v1 * q0)) % P
        p0, p1 = (p0 * t0 + 7 * p1 * x1) % P, (p0 * x1 + p1 * t0) % P
    cons += [(w[at_value] - e0) % P, (w[at_value + 1] - e1) % P]
    return cons


def _gate_constraints(desc, local_constants, local_wires, pi_hash):
    out = [0] * desc.num_gate_constraints
    many = desc.num_selectors > 1
    gc = local_constants[desc.num_selectors:]
    for g, gate in enumerate(desc.gates):
        sel = desc.selector_indices[g]
        s = local_constants[sel]
        a, b = desc.groups[sel]
        filt = 1
        for j in range(a, b):
            if j != g:
                filt = filt * (j - s) % P
        if many:
            filt = filt * (UNUSED_SELECTOR - s) % P
        if gate.kind == "arithmetic":
            cons = [(local_wires[4 * i + 3] - (gc[0] * local_wires[4 * i] * local_wires[4 * i + 1] + gc[1] * local_wires[4 * i + 2])) % P
                    for i in range(gate.num_ops)]
        elif gate.kind == "constant":
            cons = [(gc[i] - local_wires[i]) % P for i in range(gate.num_ops)]
        elif gate.kind == "public_input":
            cons = [(local_wires[i] - pi_hash[i]) % P for i in range(4)]
        elif gate.kind == "noop":
            cons = []
        elif gate.kind == "poseidon":
            cons = _poseidon_gate(local_wires)
        elif gate.kind == "arithmetic_extension":   # gates/arithmetic_extension.rs, D = 2: out - (c0 m0 m1 + c1 addend)
            cons = []
            for i in range(gate.num_ops):
                w = local_wires[8 * i:8 * i + 8]
                cons.append((w[6] - (gc[0] * (w[0] * w[2] + 7 * w[1] * w[3]) + gc[1] * w[4])) % P)
                cons.append((w[7] - (gc[0] * (w[0] * w[3] + w[1] * w[2]) + gc[1] * w[5])) % P)
        elif gate.kind == "mul_extension":          # gates/multiplication_extension.rs: out - c0 m0 m1
            cons = []
            for i in range(gate.num_ops):
                w = local_wires[6 * i:6 * i + 6]
                cons.append((w[4] - gc[0] * (w[0] * w[2] + 7 * w[1] * w[3])) % P)
                cons.append((w[5] - gc[0] * (w[0] * w[3] + w[1] * w[2])) % P)
        elif gate.kind in ("reducing", "reducing_extension"):   # gates/reducing{,_extension}.rs: acc*alpha + coeff - next acc
            ext = gate.kind == "reducing_extension"
            w, nco = local_wires, gate.num_ops
            start_accs = 6 + (2 * nco if ext else nco)
            a0, a1 = w[4], w[5]
            cons = []
            for i in range(nco):
                c0, c1 = (w[6 + 2 * i], w[7 + 2 * i]) if ext else (w[6 + i], 0)
                n0, n1 = (w[0], w[1]) if i == nco - 1 else (w[start_accs + 2 * i], w[start_accs + 2 * i + 1])
                cons.append((a0 * w[2] + 7 * a1 * w[3] + c0 - n0) % P)
                cons.append((a0 * w[3] + a1 * w[2] + c1 - n1) % P)
                a0, a1 = n0, n1
        elif gate.kind == "exponentiation":         # gates/exponentiation.rs ExponentiationGate{num_power_bits = num_ops}
            w, nb = local_wires, gate.num_ops
            cons = []
            for i in range(nb):
                prev = 1 if i == 0 else w[2 + nb + i - 1] ** 2 % P
                b = w[1 + (nb - 1 - i)]
                cons.append((prev * (b * w[0] + 1 - b) - w[2 + nb + i]) % P)
            cons.append((w[1 + nb] - w[2 + 2 * nb - 1]) % P)
        elif gate.kind == "poseidon_mds":           # gates/poseidon_mds.rs: output - MDS(input), 12 extension elements
            w, cons = local_wires, []
            for r in range(12):
                for comp in range(2):
                    acc = sum(w[2 * ((i + r) % 12) + comp] * _POS_CIRC[i] for i in range(12)) + (8 * w[comp] if r == 0 else 0)
                    cons.append((w[24 + 2 * r + comp] - acc) % P)
        elif gate.kind == "random_access":          # gates/random_access.rs: bits = param & 0xFF, extra constants = param >> 8
            bits, copies, nx = gate.param & 0xFF, gate.num_ops, gate.param >> 8
            vec = 1 << bits
            routed = (2 + vec) * copies + nx
            w, cons = local_wires, []
            for cp in range(copies):
                b0 = (2 + vec) * cp
                bs = [w[routed + cp * bits + i] for i in range(bits)]
                cons += [b * (b - 1) % P for b in bs]
                cons.append((sum(b << i for i, b in enumerate(bs)) - w[b0]) % P)
                items = list(w[b0 + 2:b0 + 2 + vec])
                for b in bs:
                    items = [(items[2 * k] + b * (items[2 * k + 1] - items[2 * k])) % P for k in range(len(items) // 2)]
                cons.append((items[0] - w[b0 + 1]) % P)
            cons += [(gc[i] - w[(2 + vec) * copies + i]) % P for i in range(nx)]
        elif gate.kind == "coset_interpolation":    # gates/coset_interpolation.rs: subgroup_bits = num_ops, degree = param
            cons = _coset_interpolation_gate(local_wires, gate.num_ops, gate.param)
        elif gate.kind == "base_sum":               # gates/base_sum.rs BaseSumGate<B>{num_limbs}: B = gate.param
            limbs = local_wires[1:1 + gate.num_ops]
            total = sum(l * pow(gate.param, i, P) for i, l in enumerate(limbs)) % P
            cons = [(total - local_wires[0]) % P]
            for l in limbs:
                pr = 1
                for k in range(gate.param):
                    pr = pr * (l - k) % P
                cons.append(pr)
        else:
            raise ValueError("gate kind %r is outside the staged subset" % gate.kind)
        for i, v in enumerate(cons):
            out[i] = (out[i] + filt * v) % P
    return out


def compute_quotient_polys(desc, constants_sigmas_coeffs, wires_coeffs, zs_pp_coeffs, betas, gammas, alphas, pi_hash):
    """-> (num_challenges * quotient_degree_factor, n) coefficient chunks, challenge-major (what `from_coeffs` commits).
    Inputs: coefficient matrices (ncols, n) of the three committed batches; constants_sigmas columns are
    [selectors, gate constants, sigmas]; zs_pp columns are [Z_c..., partial products of c = 0, of c = 1, ...]."""
    n = 1 << desc.degree_bits
    qb = desc.quotient_degree_bits
    N, md, nch, R, npp = n << qb, 1 << qb, desc.num_challenges, desc.num_routed_wires, desc.num_partial_products
    lde = lambda m: np.stack([O.coset_lde(np.asarray(col, dtype=np.uint64), qb) for col in m])   # values at g*w_N^i, natural i
    cs, wi, zp = lde(constants_sigmas_coeffs), lde(wires_coeffs), lde(zs_pp_coeffs)
    w_N = O.root_of_unity(desc.degree_bits + qb)
    k_is = [pow(COSET_SHIFT, j, P) for j in range(R)]
    # ZeroPolyOnCoset: Z_H(g w_N^i) = g^n * w_{2^qb}^(i mod 2^qb) - 1
    g_pow_n = pow(COSET_SHIFT, n, P)
    w_rate = O.root_of_unity(qb)
    zh = [(g_pow_n * pow(w_rate, i, P) - 1) % P for i in range(md)]
    zh_inv = [pow(v, P - 2, P) for v in zh]
    n_inv_dummy = None  # noqa: F841 (L_0 uses n * (x - 1), inverted per point)
    out = np.zeros((nch, N), dtype=np.uint64)
    x = 1
    for i in range(N):
        sx = COSET_SHIFT * x % P                      # shifted_x
        i_next = (i + md) % N                         # next_step = 2^quotient_degree_bits
        local_constants = [int(v) for v in cs[:desc.num_constants, i]]
        s_sigmas = [int(v) for v in cs[desc.num_constants:desc.num_constants + R, i]]
        local_wires = [int(v) for v in wi[:, i]]
        l_0 = zh[i % md] * pow(n * (sx - 1) % P, P - 2, P) % P
        z1, pp_terms = [], []
        for c in range(nch):
            z_x, z_gx = int(zp[c, i]), int(zp[c, i_next])
            z1.append(l_0 * (z_x - 1) % P)
            accs = [z_x] + [int(zp[nch + c * npp + k, i]) for k in range(npp)] + [z_gx]
            for q, s in enumerate(range(0, R, md)):
                pn = pd = 1
                for j in range(s, min(s + md, R)):
                    pn = pn * ((local_wires[j] + betas[c] * (k_is[j] * sx % P) + gammas[c]) % P) % P
                    pd = pd * ((local_wires[j] + betas[c] * s_sigmas[j] + gammas[c]) % P) % P
                pp_terms.append((accs[q] * pn - accs[q + 1] * pd) % P)
        terms = z1 + pp_terms + _gate_constraints(desc, local_constants, local_wires, pi_hash)
        for c in range(nch):
            acc = 0
            for t in reversed(terms):
                acc = (acc * alphas[c] + t) % P
            out[c, i] = acc * zh_inv[i % md] % P
        x = x * w_N % P
    # coset_ifft(g): interpolate on the subgroup, then undo the shift coefficient by coefficient
    chunks = []
    g_inv = pow(COSET_SHIFT, P - 2, P)
    for c in range(nch):
        co = O.ifft(out[c])
        sc, scaled = 1, np.zeros(N, dtype=np.uint64)
        for j in range(N):
            scaled[j] = int(co[j]) * sc % P
            sc = sc * g_inv % P
        # quotient_poly.trim_to_len(quotient_degree) is a no-op at length 8n; chunks(degree)
        chunks.extend(scaled[k * n:(k + 1) * n] for k in range(md))
    return np.stack(chunks)
