// Width-12 Poseidon and Poseidon2 permutations over Goldilocks, one permutation per thread with the
// whole state in registers and the round constants in constant memory.
//
// Replaces plonky2::hash::poseidon{,_goldilocks}::Poseidon::poseidon (PoseidonGoldilocksConfig) and
// poseidon2_plonky2's Poseidon2 permutation (Poseidon2GoldilocksConfig, the reference's default C:
// mp2-common/src/lib.rs:37-40) -- SURVEY.md 8(a) a6/a7, Appendix A.6/A.7.
//
// Inputs are "loose" (any u64); outputs of the *_permute functions are loose too -- callers
// canonicalise what they store.
#pragma once
#include "gl.cuh"
#include "poseidon_constants.h"

// Tuning knobs (see profiles/): MP2_ROUND_BARRIER keeps the warps of a CTA in the same round so the
// instruction stream (far larger than the instruction caches) is fetched once per CTA, not per warp.
#ifndef MP2_ROUND_BARRIER
#define MP2_ROUND_BARRIER 1
#endif
// SYNC is a template flag of the permutations: only kernels in which every thread of the CTA runs the
// same number of permutations may set it.
#define MP2_ROUND_SYNC() do { if (SYNC && MP2_ROUND_BARRIER) __syncthreads(); } while (0)

#define MP2_HASH_POSEIDON 0u
#define MP2_HASH_POSEIDON2 1u

// Round constants in constant memory (one copy per translation unit including this header).
static __constant__ u64 c_pos_rc[MP2_POSEIDON_RC_LEN] = {MP2_POSEIDON_RC_LIST};
// ... and as 22|21|21-bit limbs, padded with one all-zero round so "MDS, then add the NEXT round's
// constants" needs no special last round.  Index (12*round + lane)*3 + limb.
static __constant__ u32 c_pos_rc3[MP2_POSEIDON_RC3_LEN] = {MP2_POSEIDON_RC3_LIST};
static __constant__ u64 c_p2_rc[MP2_POSEIDON2_RC_LEN] = {MP2_POSEIDON2_RC_LIST};
static __constant__ u64 c_p2_diag[MP2_POSEIDON2_DIAG_LEN] = {MP2_POSEIDON2_DIAG_LIST};

// ------------------------------------------------------------------------------------------------
// Poseidon: MDS = circ(17,15,41,16,2,28,13,13,39,18,34,20) + diag(8,0,...,0)
// out[r] = sum_i s[(i+r)%12]*CIRC[i] + s[r]*DIAG[r]           (plonky2 mds_row_shf)
//
// IMAD.WIDE is a quarter-rate instruction on this part, so the linear layer is done with plain
// 32-bit IMADs instead: every state element is cut into three limbs of 22|21|21 bits, each limb
// plane goes through the (6-bit entry, row sum 264) matrix with 32-bit multiply-adds that cannot
// overflow (2^22 * 264 + 2^22 < 2^31), and the planes are put back together with one fold of the
// <= 10 overflow bits (2^64 = eps).  The accumulators start from the limbs of the next round's
// constant, so the constant layer is free.
// ------------------------------------------------------------------------------------------------
GL_DEV void pos_split3(u64 x, u32 &l0, u32 &l1, u32 &l2) {
  u32 lo = lo32(x), hi = hi32(x);
  l0 = lo & 0x3FFFFFu;
  l1 = __funnelshift_r(lo, hi, 22) & 0x1FFFFFu;
  l2 = hi >> 11;
}
// Shift amounts read from constant memory: with literal shifts ptxas recognises "x << 22, x >> 10"
// as the two halves of x * 2^22 and emits IMAD.WIDE -- exactly the quarter-rate instruction this
// formulation exists to avoid.  A register shift amount keeps it a pair of SHFs on the alu pipe.
static __constant__ u32 c_merge_sh[4] = {22, 10, 11, 21};

// o0 + o1*2^22 + o2*2^43 with o_k < 2^31  ->  loose u64.  The <= 10 bits above 2^64 (ov) are folded
// with 2^64 = eps:  {lo,hi} + ov*eps = {lo,hi} + {-ov, ov - (ov != 0)}, one possible carry, folded again.
GL_DEV u64 pos_merge3(u32 o0, u32 o1, u32 o2) {
  u32 lo, hi;
  const u32 s22 = c_merge_sh[0], s10 = c_merge_sh[1], s11 = c_merge_sh[2], s21 = c_merge_sh[3];
  asm("{\n\t.reg .u32 a, b, c, ov, t0, t1, cy, m;\n\t"
      "shl.b32 a, %3, %5;\n\tshr.u32 b, %3, %6;\n\tshl.b32 c, %4, %7;\n\tshr.u32 ov, %4, %8;\n\t"
      "add.cc.u32 %0, %2, a;\n\taddc.cc.u32 %1, b, c;\n\taddc.u32 ov, ov, 0;\n\t"
      "sub.cc.u32 t0, 0, ov;\n\tsubc.u32 t1, ov, 0;\n\t"
      "add.cc.u32 %0, %0, t0;\n\taddc.cc.u32 %1, %1, t1;\n\taddc.u32 cy, 0, 0;\n\t"
      "sub.u32 m, 0, cy;\n\t"
      "add.cc.u32 %0, %0, m;\n\taddc.u32 %1, %1, 0;\n\t}"
      : "=r"(lo), "=r"(hi)
      : "r"(o0), "r"(o1), "r"(o2), "r"(s22), "r"(s10), "r"(s11), "r"(s21));
  return pack64(lo, hi);
}

// One 32-bit limb plane through the circulant: y[r] = sum_i CIRC[i]*x[(i+r)%12] (+ 8*x[0] on row 0)
// + rc[r].  The matrix was chosen by plonky2 so that its cyclic convolution splits, over
// z^12-1 = (z^3-1)(z^3+1)(z^6+1), into three small (nega)cyclic blocks whose constants are all
// +-powers of two:  d = reversed CIRC;  (d_k + d_{k+6})/2 = [15,24,18,17,40,14] splits again into
// the cyclic-3 block [16,32,16] and the negacyclic-3 block [-1,-8,2];  (d_k - d_{k+6})/2 =
// [2,-4,16,1,-1,-1] is the negacyclic-6 block.  ~90 shifts/adds per plane instead of 144
// multiply-adds; ptxas spreads them over the alu and fma pipes.  Arithmetic wraps mod 2^32; the
// results themselves are < 2^31 (tests/test_oracle_vs_pyref.py pins the same decomposition).
GL_DEV void pos_mds_plane(const u32 (&x)[12], const u32 *rc, int rc_stride, u32 (&y)[12]) {
  u32 a[6], b[6];
#pragma unroll
  for (int k = 0; k < 6; k++) {
    a[k] = x[k] + x[k + 6];
    b[k] = x[k] - x[k + 6];
  }
  u32 aa[3], ab[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    aa[k] = a[k] + a[k + 3];
    ab[k] = a[k] - a[k + 3];
  }
  const u32 S = aa[0] + aa[1] + aa[2];
  const u32 yaa0 = (S + aa[2]) << 4, yaa1 = (S + aa[0]) << 4, yaa2 = (S + aa[1]) << 4;
  const u32 yab0 = (ab[2] << 3) - ab[0] - (ab[1] << 1);
  const u32 yab1 = 0u - (ab[0] << 3) - ab[1] - (ab[2] << 1);
  const u32 yab2 = (ab[0] << 1) - (ab[1] << 3) - ab[2];
  const u32 ya[6] = {yaa0 + yab0, yaa1 + yab1, yaa2 + yab2, yaa0 - yab0, yaa1 - yab1, yaa2 - yab2};
  // negacyclic-6 with f = [2,-4,16,1,-1,-1]
  const u32 yb[6] = {
      (b[0] << 1) + b[1] + b[2] - b[3] - (b[4] << 4) + (b[5] << 2),
      (b[1] << 1) + b[2] + b[3] - b[4] - (b[5] << 4) - (b[0] << 2),
      (b[2] << 1) + b[3] + b[4] - b[5] + (b[0] << 4) - (b[1] << 2),
      (b[3] << 1) + b[4] + b[5] + b[0] + (b[1] << 4) - (b[2] << 2),
      (b[4] << 1) + b[5] - b[0] + b[1] + (b[2] << 4) - (b[3] << 2),
      (b[5] << 1) - b[0] - b[1] + b[2] + (b[3] << 4) - (b[4] << 2)};
#pragma unroll
  for (int k = 0; k < 6; k++) {
    y[k] = ya[k] + yb[k] + rc[k * rc_stride];
    y[k + 6] = ya[k] - yb[k] + rc[(k + 6) * rc_stride];
  }
  y[0] += x[0] << 3;  // DIAG[0] = 8
}

// s <- MDS*s + (round constants rc3[0..36), index 3*lane + limb)
GL_DEV void pos_mds_rc(u64 (&s)[12], const u32 *rc3) {
  u32 l0[12], l1[12], l2[12], o0[12], o1[12], o2[12];
#pragma unroll
  for (int i = 0; i < 12; i++) pos_split3(s[i], l0[i], l1[i], l2[i]);
  pos_mds_plane(l0, rc3, 3, o0);
  pos_mds_plane(l1, rc3 + 1, 3, o1);
  pos_mds_plane(l2, rc3 + 2, 3, o2);
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = pos_merge3(o0[i], o1[i], o2[i]);
}

// Naive schedule (A.6): 30 x { +RC ; S-box (all lanes | lane 0) ; MDS }, 4 full + 22 partial + 4 full.
template <bool SYNC>
GL_DEV void poseidon_permute(u64 (&s)[12]) {
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl_add_c(s[i], c_pos_rc[i]);
  int r = 0;
#pragma unroll 1
  for (; r < 4; r++) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_pow7(s[i]);
    pos_mds_rc(s, c_pos_rc3 + 36 * (r + 1));
    MP2_ROUND_SYNC();
  }
#pragma unroll 1
  for (; r < 26; r++) {
    s[0] = gl_pow7(s[0]);
    pos_mds_rc(s, c_pos_rc3 + 36 * (r + 1));
    MP2_ROUND_SYNC();
  }
#pragma unroll 1
  for (; r < 30; r++) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_pow7(s[i]);
    pos_mds_rc(s, c_pos_rc3 + 36 * (r + 1));
    MP2_ROUND_SYNC();
  }
}

// ------------------------------------------------------------------------------------------------
// Poseidon2 (Horizen-Labs Goldilocks t = 12):  M_E ; 4 x {+RC, S, M_E} ; 22 x {+rc on lane 0, S on
// lane 0, M_I} ; 4 x {+RC, S, M_E}
// ------------------------------------------------------------------------------------------------
// M_E = circ(2*M4, M4, M4), M4 = [[5,7,1,3],[4,6,1,1],[1,3,5,7],[1,1,4,6]]
GL_DEV void p2_external(u64 (&s)[12]) {
#pragma unroll
  for (int c = 0; c < 12; c += 4) {
    u64 x0 = s[c], x1 = s[c + 1], x2 = s[c + 2], x3 = s[c + 3];
    u64 t0 = gl_add(x0, x1), t1 = gl_add(x2, x3);
    u64 t2 = gl_add(gl_add(x1, x1), t1), t3 = gl_add(gl_add(x3, x3), t0);
    u64 t1_2 = gl_add(t1, t1), t0_2 = gl_add(t0, t0);
    u64 t4 = gl_add(gl_add(t1_2, t1_2), t3), t5 = gl_add(gl_add(t0_2, t0_2), t2);
    s[c] = gl_add(t3, t5);
    s[c + 1] = t5;
    s[c + 2] = gl_add(t2, t4);
    s[c + 3] = t4;
  }
  u64 col[4];
#pragma unroll
  for (int l = 0; l < 4; l++) col[l] = gl_add(gl_add(s[l], s[4 + l]), s[8 + l]);
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl_add(s[i], col[i % 4]);
}

// M_I: out[i] = s[i]*mu_i + sum(s)
GL_DEV void p2_internal(u64 (&s)[12]) {
  u64 sum = s[0];
#pragma unroll
  for (int i = 1; i < 12; i++) sum = gl_add(sum, s[i]);
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl_add(gl_mul(s[i], c_p2_diag[i]), sum);
}

template <bool SYNC>
GL_DEV void poseidon2_permute(u64 (&s)[12]) {
  p2_external(s);
  const u64 *rc = c_p2_rc;
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_pow7(gl_add_c(s[i], rc[i]));
    p2_external(s);
    rc += 12;
    MP2_ROUND_SYNC();
  }
#pragma unroll 1
  for (int r = 0; r < 22; r++) {
    s[0] = gl_pow7(gl_add_c(s[0], rc[0]));
    p2_internal(s);
    rc += 1;
    MP2_ROUND_SYNC();
  }
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_pow7(gl_add_c(s[i], rc[i]));
    p2_external(s);
    rc += 12;
    MP2_ROUND_SYNC();
  }
}

template <u32 KIND, bool SYNC = false>
GL_DEV void permute(u64 (&s)[12]) {
  if (KIND == MP2_HASH_POSEIDON2) poseidon2_permute<SYNC>(s);
  else poseidon_permute<SYNC>(s);
}
