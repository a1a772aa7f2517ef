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
// instruction stream (far larger than the instruction caches) is fetched once per CTA, not per warp.  It paid
// 22 % on the round-1 kernel; with the partial rounds on the FP64 pipe (shorter code, 5 CTAs/SM) it costs 2 %
// (profiles/r2_poseidon_f64.txt) and is off.
#ifndef MP2_ROUND_BARRIER
#define MP2_ROUND_BARRIER 0
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
// Partial rounds: constants pushed through the linear layers (tools/gen_poseidon_constants.py) -- lane 0
// gets c_pos_t0[r - 4] after the layer of round r, lanes 1..11 get c_pos_d once when the partial rounds end.
static __constant__ u64 c_pos_t0[MP2_POSEIDON_PARTIAL_T0_LEN] = {MP2_POSEIDON_PARTIAL_T0_LIST};
static __constant__ u64 c_pos_d[MP2_POSEIDON_PARTIAL_D_LEN] = {MP2_POSEIDON_PARTIAL_D_LIST};
static __constant__ u64 c_p2_rc[MP2_POSEIDON2_RC_LEN] = {MP2_POSEIDON2_RC_LIST};
static __constant__ u64 c_p2_diag[MP2_POSEIDON2_DIAG_LEN] = {MP2_POSEIDON2_DIAG_LIST};
static __constant__ u32 c_p2_rc3[MP2_POSEIDON2_RC3_LEN] = {MP2_POSEIDON2_RC3_LIST};

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
      : "=&r"(lo), "=&r"(hi)
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
template <bool RC>
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
    y[k] = ya[k] + yb[k] + (RC ? rc[k * rc_stride] : 0u);
    y[k + 6] = ya[k] - yb[k] + (RC ? rc[(k + 6) * rc_stride] : 0u);
  }
  y[0] += x[0] << 3;  // DIAG[0] = 8
}

// s <- MDS*s + (round constants rc3[0..36), index 3*lane + limb)
GL_DEV void pos_mds_rc(u64 (&s)[12], const u32 *rc3) {
  u32 l0[12], l1[12], l2[12], o0[12], o1[12], o2[12];
#pragma unroll
  for (int i = 0; i < 12; i++) pos_split3(s[i], l0[i], l1[i], l2[i]);
  pos_mds_plane<true>(l0, rc3, 3, o0);
  pos_mds_plane<true>(l1, rc3 + 1, 3, o1);
  pos_mds_plane<true>(l2, rc3 + 2, 3, o2);
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = pos_merge3(o0[i], o1[i], o2[i]);
}

// Carry-propagates the three plane outputs of one lane back into limbs that may enter the next MDS.
//   value = o0 + o1*2^22 + o2*2^43 = (o0 & M22) + (t1 & M21)*2^22 + (t2 & M21)*2^43 + ov*2^64,
//   t1 = o1 + (o0 >> 22), t2 = o2 + (t1 >> 21), ov = t2 >> 21 (<= 2^11),  and 2^64 = 2^32 - 1 = 2^10*2^22 - 1:
//   +ov*2^10 on limb 1, -ov on limb 0.  Limb 0 would go negative, so a multiple of p is added in limb form:
//   p + (2^22, -1, 0) + (0, 2^21, -1) = (2^22 + 1, 2^21 - 2^10 - 1, 2^21 - 1)  (limb weights 1, 2^22, 2^43),
//   which keeps every limb non-negative with no borrow logic.  Resulting limbs are < 2^23 + 2, so the next
//   plane outputs stay below 264 * 2^23.01 < 2^32 (tests/test_limb_planes.py pins the identity and the bounds).
// The round-1 form borrowed conditionally (12 instructions per lane); this one is 10.
GL_DEV void pos_renorm3(u32 o0, u32 o1, u32 o2, u32 &l0, u32 &l1, u32 &l2) {
  const u32 t1 = o1 + (o0 >> 22);
  const u32 t2 = o2 + (t1 >> 21);
  const u32 ov = t2 >> 21;
  l0 = (o0 & 0x3FFFFFu) - ov + 0x400001u;
  l1 = (t1 & 0x1FFFFFu) + (ov * 1024u + 0x1FFBFFu);
  l2 = (t2 & 0x1FFFFFu) + 0x1FFFFFu;
}

// S-box on all 12 lanes.  Fully unrolled this is ~1000 instructions per round and the permutation
// outgrows the instruction caches (ncu: `no_instruction` was the top stall); rolled up as 3 x 4 lanes
// with a register rotation (24 moves per trip) it is a third of the code for 7 % more issue slots.
// Per permutation: Poseidon (whose partial rounds moved to the FP64 pipe and shrank) runs faster unrolled, Poseidon2
// faster rolled (profiles/r2_poseidon_f64.txt).
#ifndef MP2_POS_SBOX_ROLLED
#define MP2_POS_SBOX_ROLLED 0
#endif
#ifndef MP2_P2_SBOX_ROLLED
#define MP2_P2_SBOX_ROLLED 1
#endif
template <bool ROLLED>
GL_DEV void sbox_layer(u64 (&s)[12]) {
  if (ROLLED) {
#pragma unroll 1
    for (int it = 0; it < 3; it++) {
      const u64 t0 = gl_pow7(s[0]), t1 = gl_pow7(s[1]), t2 = gl_pow7(s[2]), t3 = gl_pow7(s[3]);
#pragma unroll
      for (int i = 0; i < 8; i++) s[i] = s[i + 4];
      s[8] = t0;
      s[9] = t1;
      s[10] = t2;
      s[11] = t3;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_pow7(s[i]);
  }
}

// Naive schedule (A.6): 30 x { +RC ; S-box (all lanes | lane 0) ; MDS }, 4 full + 22 partial + 4 full.
// In the 22 partial rounds only lane 0 passes through the S-box, so lanes 1..11 stay in limb form
// from one linear layer to the next (re-normalised, never merged) and only lane 0 is merged/split; their
// round constants are pushed through the linear layers, leaving one constant per round on lane 0 and one
// correction vector at the end (the algebra and its self-check are in tools/gen_poseidon_constants.py).
template <bool SYNC>
GL_DEV void poseidon_permute_int(u64 (&s)[12]) {
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl_add_c(s[i], c_pos_rc[i]);
  // The two groups of four full rounds share ONE copy of the full-round code (phase loop): with
  // separate copies the three loop bodies (2 x 14 KB + 8 KB) no longer fit the 32 KB instruction cache
  // level when the resident CTAs are in different parts of the permutation.
#pragma unroll 1
  for (int phase = 0; phase < 2; phase++) {
    const int r0 = phase ? 26 : 0;
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
      sbox_layer<MP2_POS_SBOX_ROLLED != 0>(s);
      pos_mds_rc(s, c_pos_rc3 + 36 * (r0 + k + 1));
      MP2_ROUND_SYNC();
    }
    if (phase == 0) {
      u32 l0[12], l1[12], l2[12];
#pragma unroll
      for (int i = 1; i < 12; i++) pos_split3(s[i], l0[i], l1[i], l2[i]);
      u64 s0 = s[0];
#pragma unroll 1
      for (int r = 4; r < 26; r++) {
        s0 = gl_pow7(s0);
        pos_split3(s0, l0[0], l1[0], l2[0]);
        u32 o0[12], o1[12], o2[12];
        pos_mds_plane<false>(l0, nullptr, 0, o0);
        pos_mds_plane<false>(l1, nullptr, 0, o1);
        pos_mds_plane<false>(l2, nullptr, 0, o2);
        s0 = gl_add_c(pos_merge3(o0[0], o1[0], o2[0]), c_pos_t0[r - 4]);  // the only constant of the round
#pragma unroll
        for (int i = 1; i < 12; i++) pos_renorm3(o0[i], o1[i], o2[i], l0[i], l1[i], l2[i]);
        MP2_ROUND_SYNC();
      }
      s[0] = s0;
#pragma unroll
      for (int i = 1; i < 12; i++) s[i] = gl_add_c(pos_merge3(l0[i], l1[i], l2[i]), c_pos_d[i]);
    }
  }
}


// ------------------------------------------------------------------------------------------------
// The same permutation with the linear layers on the FP64 pipe (MP2_POSEIDON_F64, the default).
//
// B200 has a full-rate FP64 pipe (tools/intpipe_peak.cu: DADD 63, DFMA 57-59 thread-instr/clk/SM) that the
// integer formulation leaves idle while the alu and fma-heavy pipes are both ~80 % busy.  A double holds
// 53-bit integers exactly, so a state element is cut into TWO 32-bit planes (the words of the u64: no
// split arithmetic at all), each plane goes through the circulant as exact double additions / FMAs by small
// powers of two (|y| <= 272 |x|: 32-bit inputs give 41-bit outputs), and the result is read back with the
// 2^52 bias trick: y + (2^52 + c) has the integer y + c in its low mantissa words, so the round constant
// rides on the conversion.  The two 41-bit planes are folded into a loose u64 with 2^64 = 2^32 - 1 (11
// integer instructions, pos_merge_d).
//   In the partial rounds lanes 1..11 never leave the FP64 domain: they stay as signed plane values and are
// carry-normalised to |.| <= 2^31 + 2^16 only every SECOND round (41 -> 49 bits of a 53-bit mantissa), with
// the round-to-multiple-of-2^32 trick t = (x + 1.5*2^84) - 1.5*2^84 -- 9 FP64 instructions per lane and
// renormalisation, none on the integer pipes.  Lane 0 is the only value converted per round.
//   Offsets that keep the converted values non-negative are multiples of p in plane form:
//   OA + 2^32 OB = k p  for  OA = k + j 2^32, OB = k (2^32 - 1) - j   (tests/test_limb_planes.py).
// ------------------------------------------------------------------------------------------------
#define MP2_D_2_52 4503599627370496.0
// u32 -> double, exact: the word becomes the low mantissa bits of 2^52 + w
GL_DEV double pos_u32_to_d(u32 w) { return __hiloint2double(0x43300000, (int)w) - MP2_D_2_52; }

// tA, tB: plane values already biased (2^52 + a, 2^52 + b, 0 <= a, b < 2^52)  ->  loose u64 = a + 2^32 b mod p.
//   a + 2^32 b = aL + 2^32 (aH + bL) + 2^64 bH = (aL - bH) + 2^32 (aH + bH + bL)   (2^64 = 2^32 - 1)
// with the net carry k in {-1, 0, 1} folded once as k*eps (same argument as gl_reduce128w: aH, bH < 2^20).
GL_DEV u64 pos_merge_d(double tA, double tB) {
  u32 lo, hi;
  asm("{\n\t.reg .u32 u, bh, c, k, m0, m1;\n\t"
      "add.u32 u, %3, %5;\n\tadd.u32 u, u, 0x79A00000;\n\t"   // aH + bH  (hi words carry 0x43300000 each)
      "add.u32 bh, %5, 0xBCD00000;\n\t"                       // bH
      "add.cc.u32 %1, %4, u;\n\taddc.u32 c, 0, 0;\n\t"
      "sub.cc.u32 %0, %2, bh;\n\tsubc.cc.u32 %1, %1, 0;\n\tsubc.u32 k, c, 0;\n\t"
      "neg.s32 m0, k;\n\tshr.s32 m1, k, 1;\n\t"
      "add.cc.u32 %0, %0, m0;\n\taddc.u32 %1, %1, m1;\n\t}"
      : "=&r"(lo), "=&r"(hi)
      : "r"((u32)__double2loint(tA)), "r"((u32)__double2hiint(tA)), "r"((u32)__double2loint(tB)),
        "r"((u32)__double2hiint(tB)));
  return pack64(lo, hi);
}

// One plane through the circulant (same CRT decomposition as pos_mds_plane, shifts written as products).
GL_DEV void pos_mds_plane_d(const double (&x)[12], double (&y)[12]) {
  double a[6], b[6];
#pragma unroll
  for (int k = 0; k < 6; k++) {
    a[k] = x[k] + x[k + 6];
    b[k] = x[k] - x[k + 6];
  }
  double aa[3], ab[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    aa[k] = a[k] + a[k + 3];
    ab[k] = a[k] - a[k + 3];
  }
  const double S = aa[0] + aa[1] + aa[2];
  const double q0 = S + aa[2], q1 = S + aa[0], q2 = S + aa[1];  // yaa_k = 16 q_k
  const double yab0 = ab[2] * 8.0 - ab[0] - ab[1] * 2.0;
  const double yab1 = -(ab[0] * 8.0) - ab[1] - ab[2] * 2.0;
  const double yab2 = ab[0] * 2.0 - ab[1] * 8.0 - ab[2];
  const double ya[6] = {q0 * 16.0 + yab0, q1 * 16.0 + yab1, q2 * 16.0 + yab2,
                        q0 * 16.0 - yab0, q1 * 16.0 - yab1, q2 * 16.0 - yab2};
  const double yb[6] = {
      b[0] * 2.0 + b[1] + b[2] - b[3] - b[4] * 16.0 + b[5] * 4.0,
      b[1] * 2.0 + b[2] + b[3] - b[4] - b[5] * 16.0 - b[0] * 4.0,
      b[2] * 2.0 + b[3] + b[4] - b[5] + b[0] * 16.0 - b[1] * 4.0,
      b[3] * 2.0 + b[4] + b[5] + b[0] + b[1] * 16.0 - b[2] * 4.0,
      b[4] * 2.0 + b[5] - b[0] + b[1] + b[2] * 16.0 - b[3] * 4.0,
      b[5] * 2.0 - b[0] - b[1] + b[2] + b[3] * 16.0 - b[4] * 4.0};
#pragma unroll
  for (int k = 0; k < 6; k++) {
    y[k] = ya[k] + yb[k];
    y[k + 6] = ya[k] - yb[k];
  }
  y[0] = x[0] * 8.0 + y[0];  // DIAG[0] = 8
}

// Carry-normalises one lane's planes: |A|, |B| < 2^50 in, |A|, |B| <= 2^31 + 2^19 out, same value mod p.
GL_DEV void pos_renorm_d(double &A, double &B) {
  const double K = 29014219670751100192948224.0;  // 1.5 * 2^84: ulp 2^32
  const double I32 = 2.3283064365386962890625e-10;  // 2^-32
  const double cA = (A + K) - K;      // A rounded to a multiple of 2^32
  const double A1 = A - cA;
  const double B1 = cA * I32 + B;
  const double cB = (B1 + K) - K;
  const double B2 = B1 - cB;
  // 2^32 * cB = 2^64 * ov, ov = cB / 2^32, and 2^64 = 2^32 - 1
  B = cB * I32 + B2;
  A = A1 - cB * I32;
}

// Biases of the plane -> integer conversions (tools/gen_poseidon_constants.py):
//   c_pos_dbias[(12*round + lane)*2 + plane] = 2^52 + word `plane` of the constant added after the layer of
//     round `round` - 1 (full rounds; round 30 = zeros);
//   c_pos_t0bias[2*(r-4) + plane] = 2^52 + offset + word of t_{r+1}[0] (partial rounds, lane 0);
//   c_pos_exbias[2*lane + plane]  = 2^52 + offset + word of d_26[lane] (leaving the partial rounds).
#ifndef MP2_POSEIDON_F64_FULL
#define MP2_POSEIDON_F64_FULL 0
#endif
static __constant__ double c_pos_t0bias[MP2_POSEIDON_T0BIAS_LEN] = {MP2_POSEIDON_T0BIAS_LIST};
static __constant__ double c_pos_exbias[MP2_POSEIDON_EXBIAS_LEN] = {MP2_POSEIDON_EXBIAS_LIST};
#if MP2_POSEIDON_F64_FULL  // only the all-FP64 variant needs the per-round biases (6 KB of constant memory)
static __constant__ double c_pos_dbias[MP2_POSEIDON_DBIAS_LEN] = {MP2_POSEIDON_DBIAS_LIST};
static __constant__ double c_pos_r4[MP2_POSEIDON_R4D_LEN] = {MP2_POSEIDON_R4D_LIST};  // round-4 constants as plane doubles
#endif

// MP2_POSEIDON_F64_FULL = 1: the eight full rounds use the FP64 planes too; 0 (default): they keep the 22|21|21-bit integer
// planes and only the 22 partial rounds (where FP64 removes the per-round renormalisation) run on the FP64 pipe.
template <bool SYNC>
GL_DEV void poseidon_permute_f64(u64 (&s)[12]) {
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl_add_c(s[i], c_pos_rc[i]);
  double A[12], B[12];
#pragma unroll 1
  for (int phase = 0; phase < 2; phase++) {
    const int r0 = phase ? 26 : 0;
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
      sbox_layer<MP2_POS_SBOX_ROLLED != 0>(s);
#if MP2_POSEIDON_F64_FULL
      double YA[12], YB[12];
#pragma unroll
      for (int i = 0; i < 12; i++) {
        A[i] = pos_u32_to_d(lo32(s[i]));
        B[i] = pos_u32_to_d(hi32(s[i]));
      }
      pos_mds_plane_d(A, YA);
      pos_mds_plane_d(B, YB);
      if (phase == 0 && k == 3) {
        // entering the partial rounds: lanes 1..11 stay in plane form (+ the constants of round 4)
#pragma unroll
        for (int i = 1; i < 12; i++) {
          A[i] = YA[i] + c_pos_r4[2 * i];
          B[i] = YB[i] + c_pos_r4[2 * i + 1];
        }
        const double *bias = c_pos_dbias + 24 * 4;
        s[0] = pos_merge_d(YA[0] + bias[0], YB[0] + bias[1]);
      } else {
        const double *bias = c_pos_dbias + 24 * (r0 + k + 1);
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = pos_merge_d(YA[i] + bias[2 * i], YB[i] + bias[2 * i + 1]);
      }
#else
      pos_mds_rc(s, c_pos_rc3 + 36 * (r0 + k + 1));
#endif
      MP2_ROUND_SYNC();
    }
    if (phase == 0) {
#if !MP2_POSEIDON_F64_FULL
#pragma unroll
      for (int i = 1; i < 12; i++) {
        A[i] = pos_u32_to_d(lo32(s[i]));
        B[i] = pos_u32_to_d(hi32(s[i]));
      }
#endif
      // plane magnitudes: F64_FULL enters with 41-bit planes (renormalise after rounds 4, 6, ..., 24; leave with
      // 39 bits), the integer full rounds hand over 32-bit planes (renormalise after rounds 5, 7, ..., 25)
      constexpr int kRenormParity = MP2_POSEIDON_F64_FULL ? 0 : 1;
      // (Splitting the layer as M (0, x_1..x_11) + z0 * column 0, so that the lanes' FP64 work overlaps the serial
      // x^7 chain of lane 0, was measured and is slower, 2.11 vs 2.08 ms at config 1 -- profiles/r2_poseidon_f64.txt:
      // the kernel is issue-bound, not latency-bound.)
      u64 s0 = s[0];
#pragma unroll 1
      for (int r = 4; r < 26; r++) {
        s0 = gl_pow7(s0);
        A[0] = pos_u32_to_d(lo32(s0));
        B[0] = pos_u32_to_d(hi32(s0));
        double YA[12], YB[12];
        pos_mds_plane_d(A, YA);
        pos_mds_plane_d(B, YB);
        s0 = pos_merge_d(YA[0] + c_pos_t0bias[2 * (r - 4)], YB[0] + c_pos_t0bias[2 * (r - 4) + 1]);
#pragma unroll
        for (int i = 1; i < 12; i++) {
          A[i] = YA[i];
          B[i] = YB[i];
        }
        if ((r & 1) == kRenormParity) {
#pragma unroll
          for (int i = 1; i < 12; i++) pos_renorm_d(A[i], B[i]);
        }
        MP2_ROUND_SYNC();
      }
      s[0] = s0;
#pragma unroll
      for (int i = 1; i < 12; i++) s[i] = pos_merge_d(A[i] + c_pos_exbias[2 * i], B[i] + c_pos_exbias[2 * i + 1]);
    }
  }
}

#ifndef MP2_POSEIDON_F64
#define MP2_POSEIDON_F64 1
#endif
template <bool SYNC>
GL_DEV void poseidon_permute(u64 (&s)[12]) {
#if MP2_POSEIDON_F64
  poseidon_permute_f64<SYNC>(s);
#else
  poseidon_permute_int<SYNC>(s);
#endif
}

// ------------------------------------------------------------------------------------------------
// Poseidon2 (Horizen-Labs Goldilocks t = 12):  M_E ; 4 x {+RC, S, M_E} ; 22 x {+rc on lane 0, S on
// lane 0, M_I} ; 4 x {+RC, S, M_E}
// ------------------------------------------------------------------------------------------------
// M_E = circ(2*M4, M4, M4), M4 = [[5,7,1,3],[4,6,1,1],[1,3,5,7],[1,1,4,6]]; row sums <= 64, so the same
// 22|21|21-bit limb planes as Poseidon's MDS apply: one plane is ~40 32-bit adds / shifted adds.
GL_DEV void p2_ext_plane(const u32 (&x)[12], const u32 *rc, u32 (&y)[12]) {
  u32 t[12];
#pragma unroll
  for (int c = 0; c < 12; c += 4) {
    const u32 x0 = x[c], x1 = x[c + 1], x2 = x[c + 2], x3 = x[c + 3];
    const u32 t0 = x0 + x1, t1 = x2 + x3;
    const u32 t2 = (x1 << 1) + t1, t3 = (x3 << 1) + t0;
    const u32 t4 = (t1 << 2) + t3, t5 = (t0 << 2) + t2;
    t[c] = t3 + t5;
    t[c + 1] = t5;
    t[c + 2] = t2 + t4;
    t[c + 3] = t4;
  }
#pragma unroll
  for (int l = 0; l < 4; l++) {
    const u32 col = t[l] + t[4 + l] + t[8 + l];
#pragma unroll
    for (int k = 0; k < 3; k++) y[4 * k + l] = t[4 * k + l] + col + rc[3 * (4 * k + l)];
  }
}

// s <- M_E*s + constants of slot `slot` (see tools/gen_poseidon_constants.py)
GL_DEV void p2_external_rc(u64 (&s)[12], int slot) {
  u32 l0[12], l1[12], l2[12], o0[12], o1[12], o2[12];
#pragma unroll
  for (int i = 0; i < 12; i++) pos_split3(s[i], l0[i], l1[i], l2[i]);
  const u32 *rc3 = c_p2_rc3 + 36 * slot;
  p2_ext_plane(l0, rc3, o0);
  p2_ext_plane(l1, rc3 + 1, o1);
  p2_ext_plane(l2, rc3 + 2, o2);
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = pos_merge3(o0[i], o1[i], o2[i]);
}

// The same layer on the FP64 pipe (MP2_POSEIDON2_F64, default): two exact 32-bit planes as doubles (row sums <= 64:
// 38-bit outputs), the slot constant folded into the 2^52 conversion bias, planes merged by pos_merge_d.  The integer
// Poseidon2 kernel is alu-bound (ncu: alu 78 %, fma-heavy 70 %, fp64 idle, profiles/r2z_poseidon2_leafhash.summary.txt);
// this moves ~200 of the ~350 instructions of each of the nine external layers off the integer pipes.
#ifndef MP2_POSEIDON2_F64
#define MP2_POSEIDON2_F64 1
#endif
static __constant__ double c_p2_dbias[MP2_POSEIDON2_DBIAS_LEN] = {MP2_POSEIDON2_DBIAS_LIST};
GL_DEV void p2_ext_plane_d(const double (&x)[12], double (&y)[12]) {
  double t[12];
#pragma unroll
  for (int c = 0; c < 12; c += 4) {
    const double x0 = x[c], x1 = x[c + 1], x2 = x[c + 2], x3 = x[c + 3];
    const double t0 = x0 + x1, t1 = x2 + x3;
    const double t2 = x1 * 2.0 + t1, t3 = x3 * 2.0 + t0;
    const double t4 = t1 * 4.0 + t3, t5 = t0 * 4.0 + t2;
    t[c] = t3 + t5;
    t[c + 1] = t5;
    t[c + 2] = t2 + t4;
    t[c + 3] = t4;
  }
#pragma unroll
  for (int l = 0; l < 4; l++) {
    const double col = t[l] + t[4 + l] + t[8 + l];
#pragma unroll
    for (int k = 0; k < 3; k++) y[4 * k + l] = t[4 * k + l] + col;
  }
}
GL_DEV void p2_external_rc_d(u64 (&s)[12], int slot) {
  double A[12], B[12], YA[12], YB[12];
#pragma unroll
  for (int i = 0; i < 12; i++) {
    A[i] = pos_u32_to_d(lo32(s[i]));
    B[i] = pos_u32_to_d(hi32(s[i]));
  }
  p2_ext_plane_d(A, YA);
  p2_ext_plane_d(B, YB);
  const double *bias = c_p2_dbias + 24 * slot;
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = pos_merge_d(YA[i] + bias[2 * i], YB[i] + bias[2 * i + 1]);
}

// M_I: out[i] = s[i]*mu_i + sum(s).  The sum is accumulated in 96 bits and reduced once.
GL_DEV void p2_internal(u64 (&s)[12]) {
  u32 a0 = lo32(s[0]), a1 = hi32(s[0]), a2 = 0;
#pragma unroll
  for (int i = 1; i < 12; i++)
    asm("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.u32 %2, %2, 0;"
        : "+r"(a0), "+r"(a1), "+r"(a2)
        : "r"(lo32(s[i])), "r"(hi32(s[i])));
  const u64 sum = gl_reduce128w(a0, a1, a2, 0u);
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl_mul_add(s[i], c_p2_diag[i], sum);  // the sum rides on the product
}

template <bool SYNC>
GL_DEV void poseidon2_permute(u64 (&s)[12]) {
#if MP2_POSEIDON2_F64
#define P2_EXTERNAL p2_external_rc_d
#else
#define P2_EXTERNAL p2_external_rc
#endif
  P2_EXTERNAL(s, 0);
  const u64 *rc = c_p2_rc + 48;
#pragma unroll 1
  for (int phase = 0; phase < 2; phase++) {  // one copy of the external-round code, see poseidon_permute
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
      sbox_layer<MP2_P2_SBOX_ROLLED != 0>(s);
      P2_EXTERNAL(s, 4 * phase + k + 1);  // slots 1..4, then 5..8 (tools/gen_poseidon_constants.py)
      MP2_ROUND_SYNC();
    }
    if (phase == 0) {
#pragma unroll 1
      for (int r = 0; r < 22; r++) {
        s[0] = gl_pow7(gl_add_c(s[0], rc[r]));
        p2_internal(s);
        MP2_ROUND_SYNC();
      }
#pragma unroll
      for (int i = 0; i < 12; i++) s[i] = gl_add_c(s[i], rc[22 + i]);
    }
  }
}

template <u32 KIND, bool SYNC = false>
GL_DEV void permute(u64 (&s)[12]) {
  if (KIND == MP2_HASH_POSEIDON2) poseidon2_permute<SYNC>(s);
  else poseidon_permute<SYNC>(s);
}
