// Quotient-polynomial evaluation on the device: plonky2 0.2.2 `compute_quotient_polys` (plonk/prover.rs) with
// `eval_vanishing_poly_base_batch` (plonk/vanishing_poly.rs) -- SURVEY.md 8(f) row 3.  Reached from every
// `circuit_data.prove(pw)` of the reference (recursion-framework/src/circuit_builder.rs:308,
// .../universal_verifier_gadget/wrap_circuit.rs:143) between the second and the third commitment; it is the last
// host-side reader of `merkle_tree.leaves`, so with it on the device the LDE rows of the three batches never leave
// HBM (the 17.2 GB / 141 MB D2H of the drop-in call).
//
// One thread per point x_i = g * w_{N_q}^i of the quotient coset, N_q = n * 2^quotient_degree_bits:
//   * the rows it needs are leaves bitrev(i * step) of the batches' own LDE (`get_lde_values(i, step)`); because
//     step = 2^(rate_bits - quotient_degree_bits), those are exactly the FIRST N_q leaves, in order -- thread L reads
//     row L of each batch (row-major leaves) and is point i = bitrev(L);
//   * terms = [L_0(x)(Z_c(x) - 1)]_c ++ [partial-product checks]_c ++ gate constraints (selector-filtered);
//     q_c(x_i) = (sum_j terms_j alpha_c^j) / Z_H(x_i), written to natural position i;
//   * then the library's own iNTT (size N_q) + coefficient scaling by g^-j = `coset_ifft`, and the N_q coefficients
//     of challenge c ARE its 2^qb chunks of n, back to back -- the chunk matrix is committed in place with
//     `from_coeffs` (the prover's quotient_polys_commitment).
// Gate set: the staged subset of mp2-common/src/serialization/circuit_data_serialization.rs:234-266 that
// oracle/quotient.py restates -- ArithmeticGate, ConstantGate, PublicInputGate, NoopGate, PoseidonGate,
// ArithmeticExtensionGate, MulExtensionGate, BaseSumGate<B>, ReducingGate, ReducingExtensionGate, RandomAccessGate,
// ExponentiationGate, PoseidonMdsGate, CosetInterpolationGate; anything else is an error.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/mp2gpu.h"
#include "internal.h"
#include "gl.cuh"
#include "poseidon.cuh"

namespace mp2 {
namespace {

constexpr u32 kMaxGates = 32, kMaxChallenges = 4, kMaxQuotientBits = 4, kMaxGateConstraints = 256;
constexpr u64 kUnusedSelector = 0xFFFFFFFFull;

struct QGate {
  u32 kind, num_ops, selector, group_begin, group_end, param, nc;  // nc = number of constraints of the gate
  u32 aux;                                                          // CosetInterpolationGate: offset of its tables in gate_tab
};
struct QParams {
  u32 n_log, qb, nch, num_wires, R, num_constants, num_selectors, npp, num_gates, gate_term_base, nterms;
  u32 cs_cols, wi_cols, zp_cols;
  const u64 *cs, *wi, *zp;          // the three batches: row-major leaves (N_lde rows each, first N_q used), or -- COLMAJOR --
                                    // their leaf-ordered column-major LDE (column c at + c * stride), read coalesced
  size_t cs_stride, wi_stride, zp_stride;
  const u64 *apow;                  // nch x nterms: alpha_c^j
  const u64 *k_is;                  // R coset shifts 7^j
  const u64 *gate_tab;              // per-gate constant tables (CosetInterpolationGate: 2^bits subgroup points, then weights)
  u64 betas[kMaxChallenges], gammas[kMaxChallenges], pi_hash[4];
  u64 zh[1u << kMaxQuotientBits], zh_inv[1u << kMaxQuotientBits];
  u64 w_nq, n_field;                // w_{N_q};  n as a field element
  QGate gates[kMaxGates];
};

static __constant__ u32 c_circ[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
GL_DEV u64 gl_neg(u64 a) { return gl_sub(0, a); }
GL_DEV u64 gl_inv(u64 a) { return gl_pow(a, GL_P - 2); }

// one batch row seen from one thread: element c of leaf L
template <bool COLMAJOR>
struct Row {
  const u64 *p;
  size_t stride;
  GL_DEV u64 operator[](u32 c) const { return COLMAJOR ? p[(size_t)c * stride] : p[c]; }
  GL_DEV Row operator+(u32 c) const { return Row{COLMAJOR ? p + (size_t)c * stride : p + c, stride}; }
};

// Folds `count` buffered constraint values into the per-challenge sums: inner[c] += v_k * alpha_c^(first + k), `ap`
// already pointing at alpha_c^first of challenge 0.  Deliberately NOT inlined: it is reached from ~150 constraint sites.
constexpr u32 kRing = 16;
__device__ __noinline__ void q_fold_ring(const u64 *ring, u32 count, const u64 *ap, u32 nterms, u32 nch, u64 *inner) {
#pragma unroll 1
  for (u32 k = 0; k < count; k++) {
    const u64 v = ring[k];
#pragma unroll 1
    for (u32 c = 0; c < nch; c++) inner[c] = gl_mul_add(v, ap[c * nterms + k], inner[c]);
  }
}

template <bool COLMAJOR>
__global__ void __launch_bounds__(128) k_quotient_points(const __grid_constant__ QParams P, u64 *__restrict__ out) {
  const u32 nq_log = P.n_log + P.qb;
  const u32 L = blockIdx.x * blockDim.x + threadIdx.x;
  if (L >= (1u << nq_log)) return;
  const u32 i = brev_bits(L, nq_log);
  const u32 md = 1u << P.qb;
  const u32 L_next = brev_bits((i + md) & ((1u << nq_log) - 1), nq_log);  // next_step = 2^quotient_degree_bits
  typedef Row<COLMAJOR> R;
  const R cs{COLMAJOR ? P.cs + L : P.cs + (size_t)L * P.cs_cols, P.cs_stride};
  const R wi{COLMAJOR ? P.wi + L : P.wi + (size_t)L * P.wi_cols, P.wi_stride};
  const R zp{COLMAJOR ? P.zp + L : P.zp + (size_t)L * P.zp_cols, P.zp_stride};
  const R zp_next{COLMAJOR ? P.zp + L_next : P.zp + (size_t)L_next * P.zp_cols, P.zp_stride};
  const u64 sx = gl_mul(kCosetShift, gl_pow(P.w_nq, i));  // shifted_x
  const u64 zh = P.zh[i & (md - 1)];
  const u64 l_0 = gl_mul(zh, gl_inv(gl_mul(P.n_field, gl_sub(sx, 1))));
  u64 acc[kMaxChallenges];
#pragma unroll
  for (u32 c = 0; c < kMaxChallenges; c++) acc[c] = 0;
  auto add_term = [&](u32 j, u64 t) {
#pragma unroll
    for (u32 c = 0; c < kMaxChallenges; c++)
      if (c < P.nch) acc[c] = gl_mul_add(t, P.apow[c * P.nterms + j], acc[c]);
  };
  // vanishing_z_1_terms, then the partial-product checks challenge by challenge
  for (u32 c = 0; c < P.nch; c++) add_term(c, gl_mul(l_0, gl_sub(zp[c], 1)));
  const u32 nchunks = P.npp + 1;
  const R sig = cs + P.num_constants;
  for (u32 c = 0; c < P.nch; c++) {
    const u64 beta = P.betas[c], gamma = P.gammas[c];
    u64 prev = zp[c];  // accs[q]: Z(x), partial products..., Z(g x)
    for (u32 q = 0; q < nchunks; q++) {
      const u64 next = q + 1 < nchunks ? zp[P.nch + c * P.npp + q] : zp_next[c];
      u64 pn = 1, pd = 1;
      const u32 j1 = min((q + 1) * md, P.R);
      for (u32 j = q * md; j < j1; j++) {
        const u64 w = gl_add(wi[j], gamma);
        pn = gl_mul(pn, gl_mul_add(beta, gl_mul(P.k_is[j], sx), w));
        pd = gl_mul(pd, gl_mul_add(beta, sig[j], w));
      }
      add_term(P.nch + c * nchunks + q, gl_sub(gl_mul(prev, pn), gl_mul(next, pd)));
      prev = next;
    }
  }
  // evaluate_gate_constraints_base_batch: sum_g filter_g * sum_i alpha^(base + i) * constraint_{g,i}
  const R gc = cs + P.num_selectors;
  // A gate's constraint values pass through a 16-entry per-thread ring (128 B: stays in L1) that a non-inlined helper
  // folds into the alpha-weighted sums whenever it is full; every gate emits its constraints in index order.
  // (History: folding at every constraint site -- 4 multiply-adds inlined ~150 times -- made the kernel 13 k SASS
  // instructions with `no_instruction` its second stall; a whole-gate scratch array of 256 values fixed that but put
  // 2 KB per thread in local memory: 227 MB for the resident threads, more than L2 -- ncu: 384 MB written and 1.14 GB
  // read for 2^17 points whose inputs are 0.25 GB, `long_scoreboard` the top stall; profiles/r2z_prove_kernels.summary.txt.)
  u64 ring[kRing];
  u64 inner[kMaxChallenges];
  const u64 *ap0 = P.apow + P.gate_term_base;
  for (u32 g = 0; g < P.num_gates; g++) {
    const QGate gate = P.gates[g];
    const u64 s = cs[gate.selector];
    u64 filt = 1;
    for (u32 j = gate.group_begin; j < gate.group_end; j++)
      if (j != g) filt = gl_mul(filt, gl_sub((u64)j, s));
    if (P.num_selectors > 1) filt = gl_mul(filt, gl_sub(kUnusedSelector, s));
#pragma unroll
    for (u32 c = 0; c < kMaxChallenges; c++) inner[c] = 0;
    auto cons = [&](u32 idx, u64 v) {
      ring[idx & (kRing - 1)] = v;
      if ((idx & (kRing - 1)) == kRing - 1) q_fold_ring(ring, kRing, ap0 + (idx - (kRing - 1)), P.nterms, P.nch, inner);
    };
    if (gate.kind == MP2GPU_GATE_ARITHMETIC) {
      const u64 c0 = gc[0], c1 = gc[1];
      for (u32 op = 0; op < gate.num_ops; op++) {
        const R w = wi + 4 * op;
        cons(op, gl_sub(w[3], gl_mul_add(gl_mul(c0, w[0]), w[1], gl_mul(c1, w[2]))));
      }
    } else if (gate.kind == MP2GPU_GATE_CONSTANT) {
      for (u32 k = 0; k < gate.num_ops; k++) cons(k, gl_sub(gc[k], wi[k]));
    } else if (gate.kind == MP2GPU_GATE_PUBLIC_INPUT) {
      for (u32 k = 0; k < 4; k++) cons(k, gl_sub(wi[k], P.pi_hash[k]));
    } else if (gate.kind == MP2GPU_GATE_ARITHMETIC_EXT || gate.kind == MP2GPU_GATE_MUL_EXT) {
      // D = 2, X^2 = 7: (a0 + a1 X)(b0 + b1 X) = (a0 b0 + 7 a1 b1) + (a0 b1 + a1 b0) X
      const bool arith = gate.kind == MP2GPU_GATE_ARITHMETIC_EXT;
      const u32 stride = arith ? 8 : 6, out_at = arith ? 6 : 4;
      const u64 c0 = gc[0], c1 = arith ? gc[1] : 0;
      for (u32 op = 0; op < gate.num_ops; op++) {
        const R w = wi + stride * op;
        const u64 p0 = gl_mul_add(gl_mul(7, w[1]), w[3], gl_mul(w[0], w[2]));
        const u64 p1 = gl_mul_add(w[0], w[3], gl_mul(w[1], w[2]));
        u64 r0 = gl_mul(c0, p0), r1 = gl_mul(c0, p1);
        if (arith) {
          r0 = gl_mul_add(c1, w[4], r0);
          r1 = gl_mul_add(c1, w[5], r1);
        }
        cons(2 * op, gl_sub(w[out_at], r0));
        cons(2 * op + 1, gl_sub(w[out_at + 1], r1));
      }
    } else if (gate.kind == MP2GPU_GATE_REDUCING || gate.kind == MP2GPU_GATE_REDUCING_EXT) {
      const bool ext = gate.kind == MP2GPU_GATE_REDUCING_EXT;
      const u32 nco = gate.num_ops, start_accs = 6 + (ext ? 2 * nco : nco);
      const u64 al0 = wi[2], al1 = wi[3];
      u64 a0 = wi[4], a1 = wi[5];
      for (u32 k = 0; k < nco; k++) {
        const u64 c0 = ext ? wi[6 + 2 * k] : wi[6 + k], c1 = ext ? wi[7 + 2 * k] : 0;
        const u32 at = k == nco - 1 ? 0 : start_accs + 2 * k;
        const u64 n0 = wi[at], n1 = wi[at + 1];
        // acc * alpha + coeff - next, in GF(p^2)
        const u64 t0 = gl_add(gl_mul_add(gl_mul(7, a1), al1, gl_mul(a0, al0)), c0);
        const u64 t1 = gl_add(gl_mul_add(a0, al1, gl_mul(a1, al0)), c1);
        cons(2 * k, gl_sub(t0, n0));
        cons(2 * k + 1, gl_sub(t1, n1));
        a0 = n0;
        a1 = n1;
      }
    } else if (gate.kind == MP2GPU_GATE_COSET_INTERPOLATION) {
      // barycentric interpolation at the shifted point, folded point by point in GF(p^2); the running (eval, prod) pair
      // is pinned to intermediate wires after `deg` points and then every `deg - 1`
      const u32 bits = gate.num_ops, deg = gate.param, npts = 1u << bits, ni = (npts - 2) / (deg - 1);
      const u64 *dom = P.gate_tab + gate.aux, *wt = dom + npts;
      const u32 at_point = 1 + 2 * npts, at_value = at_point + 2, at_inter = at_point + 4, at_shifted = at_inter + 4 * ni;
      const u64 shift = wi[0], x0 = wi[at_shifted], x1 = wi[at_shifted + 1];
      cons(0, gl_sub(wi[at_point], gl_mul(x0, shift)));
      cons(1, gl_sub(wi[at_point + 1], gl_mul(x1, shift)));
      const u64 x1w = gl_mul(7, x1);
      u64 e0 = 0, e1 = 0, p0 = 1, p1 = 0;
      u32 ci = 2, cut = deg, run = 0;
#pragma unroll 1
      for (u32 k = 0; k < npts; k++) {
        if (k == cut) {
          const u32 ie = at_inter + 2 * run, ip = at_inter + 2 * (ni + run);
          const u64 w0 = wi[ie], w1 = wi[ie + 1], w2 = wi[ip], w3 = wi[ip + 1];
          cons(ci, gl_sub(w0, e0));
          cons(ci + 1, gl_sub(w1, e1));
          cons(ci + 2, gl_sub(w2, p0));
          cons(ci + 3, gl_sub(w3, p1));
          ci += 4;
          e0 = w0; e1 = w1; p0 = w2; p1 = w3;
          cut += deg - 1;
          run++;
        }
        const u64 t0 = gl_sub(x0, dom[k]);                      // term = (t0, x1)
        const u64 v0 = wi[1 + 2 * k], v1 = wi[2 + 2 * k];
        const u64 q0 = gl_mul(p0, wt[k]), q1 = gl_mul(p1, wt[k]);
        // eval * term + value * (prod * weight);  prod * term
        const u64 n0 = gl_add(gl_mul_add(e1, x1w, gl_mul(e0, t0)), gl_mul_add(gl_mul(7, v1), q1, gl_mul(v0, q0)));
        const u64 n1 = gl_add(gl_mul_add(e0, x1, gl_mul(e1, t0)), gl_mul_add(v0, q1, gl_mul(v1, q0)));
        const u64 r0 = gl_mul_add(p1, x1w, gl_mul(p0, t0));
        const u64 r1 = gl_mul_add(p0, x1, gl_mul(p1, t0));
        e0 = n0; e1 = n1; p0 = r0; p1 = r1;
      }
      cons(ci, gl_sub(wi[at_value], e0));
      cons(ci + 1, gl_sub(wi[at_value + 1], e1));
    } else if (gate.kind == MP2GPU_GATE_EXPONENTIATION) {
      const u32 nb = gate.num_ops;
      const u64 base = wi[0];
      const R bits = wi + 1, iv = wi + (2 + nb);
      for (u32 k = 0; k < nb; k++) {
        const u64 prev = k == 0 ? 1 : gl_sqr(iv[k - 1]);
        const u64 b = bits[nb - 1 - k];
        // b * base + (1 - b)
        const u64 sel = gl_add(gl_mul(b, base), gl_sub(1, b));
        cons(k, gl_sub(gl_mul(prev, sel), iv[k]));
      }
      cons(nb, gl_sub(wi[1 + nb], iv[nb - 1]));
    } else if (gate.kind == MP2GPU_GATE_POSEIDON_MDS) {
#pragma unroll 1
      for (u32 r = 0; r < 12; r++)
#pragma unroll 1
        for (u32 comp = 0; comp < 2; comp++) {
          u64 acc = r == 0 ? gl_mul(8, wi[comp]) : 0;
#pragma unroll 1
          for (u32 k = 0; k < 12; k++) {
            u32 src = k + r;
            src = src >= 12 ? src - 12 : src;
            acc = gl_mul_add(wi[2 * src + comp], (u64)c_circ[k], acc);
          }
          cons(2 * r + comp, gl_sub(wi[24 + 2 * r + comp], acc));
        }
    } else if (gate.kind == MP2GPU_GATE_RANDOM_ACCESS) {
      const u32 bits = gate.param & 0xFF, copies = gate.num_ops, nx = gate.param >> 8, vec = 1u << bits;
      const u32 routed = (2 + vec) * copies + nx;
      u32 ci = 0;
      for (u32 cp = 0; cp < copies; cp++) {
        const R w = wi + (2 + vec) * cp, bs = wi + (routed + cp * bits);
        u64 rec = 0;
        for (u32 k = 0; k < bits; k++) cons(ci++, gl_mul(bs[k], gl_sub(bs[k], 1)));
        for (u32 k = bits; k-- > 0;) rec = gl_add(gl_add(rec, rec), bs[k]);
        cons(ci++, gl_sub(rec, w[0]));
        // fold the list pairwise by the bits; vec <= 64 elements live in registers / local memory
        u64 items[64];
        for (u32 k = 0; k < vec; k++) items[k] = w[2 + k];
        u32 len = vec;
        for (u32 b = 0; b < bits; b++) {
          len >>= 1;
          for (u32 k = 0; k < len; k++) items[k] = gl_mul_add(bs[b], gl_sub(items[2 * k + 1], items[2 * k]), items[2 * k]);
        }
        cons(ci++, gl_sub(items[0], w[1]));
      }
      for (u32 k = 0; k < nx; k++) cons(ci++, gl_sub(gc[k], wi[(2 + vec) * copies + k]));
    } else if (gate.kind == MP2GPU_GATE_BASE_SUM) {
      const u64 base = gate.param;
      u64 acc = 0;
      for (u32 k = gate.num_ops; k-- > 0;) acc = gl_mul_add(acc, base, wi[1 + k]);  // reduce_with_powers(limbs, B)
      cons(0, gl_sub(acc, wi[0]));
      for (u32 k = 0; k < gate.num_ops; k++) {
        const u64 l = wi[1 + k];
        u64 pr = l;
        for (u32 j = 1; j < gate.param; j++) pr = gl_mul(pr, gl_sub(l, (u64)j));
        cons(1 + k, pr);
      }
    } else if (gate.kind == MP2GPU_GATE_POSEIDON) {
      // PoseidonGate::eval_unfiltered (plonky2 gates/poseidon.rs), naive round structure: every S-box input except
      // round 0's is a wire the state is overwritten with (the constraint polynomials equal those of plonky2's fast
      // partial-round form: same functions of the wires).  Wire map: see MP2GPU_GATE_POSEIDON in mp2gpu.h.
      u64 st[12];
      const u64 sw = wi[24];
      cons(0, gl_mul(sw, gl_sub(sw, 1)));
#pragma unroll
      for (u32 i = 0; i < 4; i++) {
        const u64 lhs = wi[i], rhs = wi[i + 4], d = wi[25 + i];
        cons(1 + i, gl_sub(gl_mul(sw, gl_sub(rhs, lhs)), d));
        st[i] = gl_add(lhs, d);
        st[i + 4] = gl_sub(rhs, d);
      }
#pragma unroll
      for (u32 i = 8; i < 12; i++) st[i] = wi[i];
#pragma unroll
      for (u32 i = 0; i < 12; i++) st[i] = gl_add_c(st[i], c_pos_rc[i]);
      u32 ci = 5;
#pragma unroll 1
      for (u32 r = 0; r < 30; r++) {
        if (r < 4 || r >= 26) {
          if (r != 0) {
            const R sb = wi + (r < 4 ? 29 + 12 * (r - 1) : 87 + 12 * (r - 26));
#pragma unroll
            for (u32 i = 0; i < 12; i++) {
              cons(ci + i, gl_sub(st[i], sb[i]));
              st[i] = sb[i];
            }
            ci += 12;
          }
#pragma unroll
          for (u32 i = 0; i < 12; i++) st[i] = gl_pow7(st[i]);
        } else {
          const u64 sb = wi[65 + r - 4];
          cons(ci++, gl_sub(st[0], sb));
          st[0] = gl_pow7(sb);
        }
        pos_mds_rc(st, c_pos_rc3 + 36 * (r + 1));  // linear layer + the constants of round r + 1 (zeros after round 29)
      }
#pragma unroll
      for (u32 i = 0; i < 12; i++) cons(ci + i, gl_sub(st[i], wi[12 + i]));
    }
    if (gate.nc & (kRing - 1)) q_fold_ring(ring, gate.nc & (kRing - 1), ap0 + (gate.nc & ~(kRing - 1)), P.nterms, P.nch, inner);
#pragma unroll
    for (u32 c = 0; c < kMaxChallenges; c++)
      if (c < P.nch) acc[c] = gl_mul_add(filt, inner[c], acc[c]);
  }
  const u64 zi = P.zh_inv[i & (md - 1)];
#pragma unroll
  for (u32 c = 0; c < kMaxChallenges; c++)
    if (c < P.nch) out[((size_t)c << nq_log) + i] = gl_mul(acc[c], zi);
}

// coset_ifft's second half: coefficient j *= g^-j
__global__ void k_coset_unshift(u64 *__restrict__ coeffs, u32 len_log, size_t total, u64 g_inv) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const u32 j = (u32)(t & (((size_t)1 << len_log) - 1));
  coeffs[t] = gl_canon(gl_mul(coeffs[t], gl_pow(g_inv, j)));
}

}  // namespace

Status commit_device_columns(const u64 *d_cols, size_t ncols, u32 n_log, u32 rate_bits, u32 cap_height, u32 hash_kind,
                             int from_coeffs, uint64_t *const *coeffs_out, uint64_t *leaves_out, uint64_t *digests_out,
                             uint64_t *cap_out, mp2gpu_batch **handle_out, cudaStream_t st);

Status quotient_polys(const mp2gpu_circuit *ci, const mp2gpu_batch *bcs, const mp2gpu_batch *bwi, const mp2gpu_batch *bzp,
                      const uint64_t *betas, const uint64_t *gammas, const uint64_t *alphas, const uint64_t *pi_hash,
                      u32 rate_bits, u32 cap_height, u32 hash_kind, uint64_t *const *chunks_out, uint64_t *leaves_out,
                      uint64_t *digests_out, uint64_t *cap_out, mp2gpu_batch **handle_out) {
  if (!ci || !bcs || !bwi || !bzp || !betas || !gammas || !alphas || !cap_out) return "quotient_polys: null argument";
  if (!ci->gates && ci->num_gates) return "quotient_polys: null gate list";
  const u32 n_log = ci->degree_bits, qb = ci->quotient_degree_bits, nch = ci->num_challenges, R = ci->num_routed_wires;
  if (nch == 0 || nch > kMaxChallenges) return "quotient_polys: num_challenges must be 1.." + std::to_string(kMaxChallenges);
  if (qb == 0 || qb > kMaxQuotientBits) return "quotient_polys: quotient_degree_bits must be 1.." + std::to_string(kMaxQuotientBits);
  if (ci->num_gates > kMaxGates) return "quotient_polys: more than " + std::to_string(kMaxGates) + " gate types";
  if (R == 0 || R > ci->num_wires) return "quotient_polys: num_routed_wires out of range";
  if (ci->num_selectors > ci->num_constants) return "quotient_polys: num_selectors > num_constants";
  for (const mp2gpu_batch *b : {bcs, bwi, bzp}) {
    if (b->n_log != n_log) return "quotient_polys: batch degree differs from the circuit's degree_bits";
    if (b->rate_bits < qb) return "quotient_polys: quotient_degree_bits exceeds a batch's rate_bits (max_quotient_degree_factor <= 2^rate_bits)";
    if (b->device != bcs->device) return "quotient_polys: the three batches live on different devices";
    if (!b->leaves && !b->lde) return "quotient_polys: batch holds no LDE rows";
  }
  const u32 md = 1u << qb, npp = (R + md - 1) / md - 1;
  if (bcs->ncols != ci->num_constants + R) return "quotient_polys: constants_sigmas batch must hold num_constants + num_routed_wires columns";
  if (bwi->ncols != ci->num_wires) return "quotient_polys: wires batch must hold num_wires columns";
  if (bzp->ncols != (size_t)nch * (1 + npp)) return "quotient_polys: zs_partial_products batch must hold num_challenges * (1 + num_partial_products) columns";
  QParams P = {};
  std::vector<u64> gate_tab;
  u32 ngc = 0, max_gate_constants = 0;
  for (u32 g = 0; g < ci->num_gates; g++) {
    const mp2gpu_gate &s = ci->gates[g];
    QGate &d = P.gates[g];
    d.kind = s.kind;
    d.num_ops = s.num_ops;
    d.selector = s.selector_index;
    d.group_begin = s.group_begin;
    d.group_end = s.group_end;
    d.param = s.param;
    if (s.selector_index >= ci->num_selectors) return "quotient_polys: gate selector_index out of range";
    if (s.group_begin > g || s.group_end <= g || s.group_end > ci->num_gates) return "quotient_polys: gate is outside its selector group";
    u32 nc = 0, nk = 0;
    switch (s.kind) {
      case MP2GPU_GATE_NOOP: break;
      case MP2GPU_GATE_ARITHMETIC:
        nc = s.num_ops;
        nk = 2;
        if (4 * s.num_ops > ci->num_wires) return "quotient_polys: ArithmeticGate ops exceed the wires";
        break;
      case MP2GPU_GATE_CONSTANT:
        nc = nk = s.num_ops;
        if (s.num_ops > ci->num_wires) return "quotient_polys: ConstantGate consts exceed the wires";
        break;
      case MP2GPU_GATE_PUBLIC_INPUT:
        nc = 4;
        if (!pi_hash) return "quotient_polys: PublicInputGate needs public_inputs_hash";
        if (ci->num_wires < 4) return "quotient_polys: PublicInputGate needs 4 wires";
        break;
      case MP2GPU_GATE_ARITHMETIC_EXT:
        nc = 2 * s.num_ops;
        nk = 2;
        if (8 * s.num_ops > ci->num_wires) return "quotient_polys: ArithmeticExtensionGate ops exceed the wires";
        break;
      case MP2GPU_GATE_MUL_EXT:
        nc = 2 * s.num_ops;
        nk = 1;
        if (6 * s.num_ops > ci->num_wires) return "quotient_polys: MulExtensionGate ops exceed the wires";
        break;
      case MP2GPU_GATE_BASE_SUM:
        nc = 1 + s.num_ops;
        if (s.param < 2 || s.param > 16) return "quotient_polys: BaseSumGate base must be 2..16";
        if (1 + s.num_ops > ci->num_wires) return "quotient_polys: BaseSumGate limbs exceed the wires";
        break;
      case MP2GPU_GATE_REDUCING:
      case MP2GPU_GATE_REDUCING_EXT: {
        const bool ext = s.kind == MP2GPU_GATE_REDUCING_EXT;
        nc = 2 * s.num_ops;
        if (s.num_ops == 0) return "quotient_polys: ReducingGate without coefficients";
        if (6 + (ext ? 2 : 1) * s.num_ops + 2 * (s.num_ops - 1) > ci->num_wires) return "quotient_polys: ReducingGate exceeds the wires";
        break;
      }
      case MP2GPU_GATE_RANDOM_ACCESS: {
        const u32 bits = s.param & 0xFF, nx = s.param >> 8;
        if (bits == 0 || bits > 6) return "quotient_polys: RandomAccessGate bits must be 1..6";
        nc = s.num_ops * (bits + 2) + nx;
        nk = nx;
        if ((2 + (1u << bits)) * s.num_ops + nx + s.num_ops * bits > ci->num_wires) return "quotient_polys: RandomAccessGate exceeds the wires";
        break;
      }
      case MP2GPU_GATE_EXPONENTIATION:
        nc = s.num_ops + 1;
        if (s.num_ops == 0) return "quotient_polys: ExponentiationGate without power bits";
        if (2 + 2 * s.num_ops > ci->num_wires) return "quotient_polys: ExponentiationGate exceeds the wires";
        break;
      case MP2GPU_GATE_COSET_INTERPOLATION: {
        const u32 bits = s.num_ops, deg = s.param;
        if (bits == 0 || bits > 6) return "quotient_polys: CosetInterpolationGate subgroup_bits must be 1..6";
        const u32 npts = 1u << bits;
        if (deg < 2 || deg > npts) return "quotient_polys: CosetInterpolationGate degree must be 2..2^subgroup_bits";
        const u32 ni = (npts - 2) / (deg - 1);
        nc = 4 + 4 * ni;
        if (5 + 2 * npts + 2 * (2 * ni + 1) > ci->num_wires) return "quotient_polys: CosetInterpolationGate exceeds the wires";
        d.aux = (u32)gate_tab.size();
        {  // subgroup points w^k and barycentric weights 1 / prod_{j != k} (w^k - w^j), by definition
          std::vector<u64> dom(npts);
          const u64 g = h_root_of_unity(bits);
          u64 x = 1;
          for (u32 k = 0; k < npts; k++) { dom[k] = x; x = h_mul(x, g); }
          gate_tab.insert(gate_tab.end(), dom.begin(), dom.end());
          for (u32 k = 0; k < npts; k++) {
            u64 den = 1;
            for (u32 j = 0; j < npts; j++)
              if (j != k) den = h_mul(den, dom[k] >= dom[j] ? dom[k] - dom[j] : dom[k] + (kP - dom[j]));  // (a sum with kP first would wrap)
            gate_tab.push_back(h_inv(den));
          }
        }
        break;
      }
      case MP2GPU_GATE_POSEIDON_MDS:
        nc = 24;
        if (ci->num_wires < 48) return "quotient_polys: PoseidonMdsGate needs 48 wires";
        break;
      case MP2GPU_GATE_POSEIDON:
        nc = 123;
        if (ci->num_wires < 135) return "quotient_polys: PoseidonGate needs 135 wires";
        break;
      default:
        return "quotient_polys: gate kind " + std::to_string(s.kind) + " is outside the supported subset (noop, arithmetic, constant, public_input, poseidon, arithmetic_extension, mul_extension, base_sum, reducing, reducing_extension, random_access, exponentiation, poseidon_mds, coset_interpolation)";
    }
    if (nc > kMaxGateConstraints) return "quotient_polys: a gate has more than " + std::to_string(kMaxGateConstraints) + " constraints";
    d.nc = nc;
    ngc = std::max(ngc, nc);
    max_gate_constants = std::max(max_gate_constants, nk);
  }
  if (ci->num_selectors + max_gate_constants > ci->num_constants) return "quotient_polys: gate constants exceed num_constants";
  DeviceScope scope(bcs->device);
  cudaStream_t st;
  MP2_TRY(ctx_stream(&st));
  const u32 nq_log = n_log + qb;
  if (nq_log > 26) return "quotient_polys: quotient domain larger than 2^26";
  const size_t Nq = (size_t)1 << nq_log, n = (size_t)1 << n_log;
  P.n_log = n_log; P.qb = qb; P.nch = nch; P.num_wires = ci->num_wires; P.R = R; P.num_constants = ci->num_constants;
  P.num_selectors = ci->num_selectors; P.npp = npp; P.num_gates = ci->num_gates;
  P.gate_term_base = nch + nch * (npp + 1);
  P.nterms = P.gate_term_base + ngc;
  P.cs_cols = (u32)bcs->ncols; P.wi_cols = (u32)bwi->ncols; P.zp_cols = (u32)bzp->ncols;
  // all three kept their column-major LDE: coalesced reads (MP2_QUOTIENT_ROWMAJOR=1 forces the row-major path: tests)
  const char *force_rows = getenv("MP2_QUOTIENT_ROWMAJOR");
  const bool colmajor = bcs->lde && bwi->lde && bzp->lde && !(force_rows && *force_rows == '1' && bcs->leaves && bwi->leaves && bzp->leaves);
  if (!colmajor && (!bcs->leaves || !bwi->leaves || !bzp->leaves)) return "quotient_polys: batches hold no common layout";
  P.cs = colmajor ? bcs->lde : bcs->leaves; P.wi = colmajor ? bwi->lde : bwi->leaves; P.zp = colmajor ? bzp->lde : bzp->leaves;
  P.cs_stride = (size_t)1 << (bcs->n_log + bcs->rate_bits);
  P.wi_stride = (size_t)1 << (bwi->n_log + bwi->rate_bits);
  P.zp_stride = (size_t)1 << (bzp->n_log + bzp->rate_bits);
  for (u32 c = 0; c < nch; c++) { P.betas[c] = betas[c] % kP; P.gammas[c] = gammas[c] % kP; }
  for (u32 k = 0; k < 4; k++) P.pi_hash[k] = pi_hash ? pi_hash[k] % kP : 0;
  // ZeroPolyOnCoset: Z_H(g w^i) = g^n w_{2^qb}^(i mod 2^qb) - 1
  const u64 g_pow_n = h_pow(kCosetShift, n), w_rate = h_root_of_unity(qb);
  u64 wr = 1;
  for (u32 j = 0; j < md; j++) {
    const u64 gw = h_mul(g_pow_n, wr);
    P.zh[j] = gw ? gw - 1 : kP - 1;  // (a sum with kP would wrap the u64)
    P.zh_inv[j] = h_inv(P.zh[j]);
    wr = h_mul(wr, w_rate);
  }
  P.w_nq = h_root_of_unity(nq_log);
  P.n_field = (u64)n % kP;
  std::vector<u64> tab((size_t)nch * P.nterms + R + gate_tab.size());
  for (u32 c = 0; c < nch; c++) {
    u64 a = 1;
    for (u32 j = 0; j < P.nterms; j++) {
      tab[(size_t)c * P.nterms + j] = a;
      a = h_mul(a, alphas[c] % kP);
    }
  }
  u64 k = 1;
  for (u32 j = 0; j < R; j++) {  // get_unique_coset_shifts: powers of the multiplicative generator
    tab[(size_t)nch * P.nterms + j] = k;
    k = h_mul(k, kCosetShift);
  }
  for (size_t j = 0; j < gate_tab.size(); j++) tab[(size_t)nch * P.nterms + R + j] = gate_tab[j];
  DevBuf d_tab, d_q;
  MP2_TRY(d_tab.alloc(tab.size(), st));
  MP2_TRY(d_q.alloc((size_t)nch * Nq, st));
  MP2_CUDA(cudaMemcpyAsync(d_tab.p, tab.data(), tab.size() * sizeof(u64), cudaMemcpyHostToDevice, st));
  P.apow = d_tab.p;
  P.k_is = d_tab.p + (size_t)nch * P.nterms;
  P.gate_tab = P.k_is + R;
  {
    ProfScope _p("k_quotient_points", st);
    if (colmajor) k_quotient_points<true><<<(unsigned)((Nq + 127) / 128), 128, 0, st>>>(P, d_q.p);
    else k_quotient_points<false><<<(unsigned)((Nq + 127) / 128), 128, 0, st>>>(P, d_q.p);
  }
  MP2_LAUNCH_CHECK();
  // the source vector `tab` must outlive the asynchronous upload
  MP2_CUDA(cudaStreamSynchronize(st));
  // coset_ifft: values on g<w> (natural order) -> coefficients; in place is not supported by the transform, so a
  // second buffer takes the coefficients, which are then the chunk matrix (nch * 2^qb columns of n)
  DevBuf d_coeffs;
  MP2_TRY(d_coeffs.alloc((size_t)nch * Nq, st));
  MP2_TRY(ntt_intt(d_q.p, Nq, d_coeffs.p, Nq, nch, nq_log, st));
  {
    const size_t total = (size_t)nch * Nq;
    ProfScope _p("k_coset_unshift", st);
    k_coset_unshift<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_coeffs.p, nq_log, total, h_inv(kCosetShift));
  }
  MP2_LAUNCH_CHECK();
  return commit_device_columns(d_coeffs.p, (size_t)nch * md, n_log, rate_bits, cap_height, hash_kind, 1, chunks_out, leaves_out,
                               digests_out, cap_out, handle_out, st);
}

}  // namespace mp2

extern "C" const char *mp2gpu_quotient_polys(const mp2gpu_circuit *circuit, const mp2gpu_batch *constants_sigmas,
                                             const mp2gpu_batch *wires, const mp2gpu_batch *zs_partial_products,
                                             const uint64_t *betas, const uint64_t *gammas, const uint64_t *alphas,
                                             const uint64_t *public_inputs_hash, uint32_t rate_bits, uint32_t cap_height,
                                             uint32_t hash_kind, uint64_t *const *chunks_out, uint64_t *leaves_out,
                                             uint64_t *digests_out, uint64_t *cap_out, mp2gpu_batch **quotient_batch_out) {
  mp2::Status s;
  try {
    s = mp2::quotient_polys(circuit, constants_sigmas, wires, zs_partial_products, betas, gammas, alphas, public_inputs_hash,
                            rate_bits, cap_height, hash_kind, chunks_out, leaves_out, digests_out, cap_out, quotient_batch_out);
  } catch (const std::exception &e) {
    s = std::string("exception: ") + e.what();
  }
  if (s.empty()) return nullptr;
  char *m = (char *)malloc(s.size() + 1);
  if (m) memcpy(m, s.c_str(), s.size() + 1);
  return m;
}
