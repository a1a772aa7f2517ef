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

#define MP2_HASH_POSEIDON 0u
#define MP2_HASH_POSEIDON2 1u

// Round constants in constant memory (one copy per translation unit including this header).
// The Poseidon table is padded with 12 zeros so "MDS then add the NEXT round's constants" needs no
// special last round.
static __constant__ u64 c_pos_rc[MP2_POSEIDON_RC_LEN + 12] = {MP2_POSEIDON_RC_LIST};
static __constant__ u64 c_p2_rc[MP2_POSEIDON2_RC_LEN] = {MP2_POSEIDON2_RC_LIST};
static __constant__ u64 c_p2_diag[MP2_POSEIDON2_DIAG_LEN] = {MP2_POSEIDON2_DIAG_LIST};

// ------------------------------------------------------------------------------------------------
// Poseidon: MDS = circ(17,15,41,16,2,28,13,13,39,18,34,20) + diag(8,0,...,0)
// out[r] = sum_i s[(i+r)%12]*CIRC[i] + s[r]*DIAG[r]           (plonky2 mds_row_shf)
//
// The state is split into 32-bit halves; each half is multiplied by the (<= 6 bit) matrix entries
// and accumulated in a 64-bit register with one IMAD.WIDE.U32 per term.  The accumulators start
// from the halves of the next round's constant, so the constant layer is free.  With
// A = sum lo_i*c_i + rc_lo < 2^41 and B = sum hi_i*c_i + rc_hi < 2^41 the result is A + 2^32*B.
// ------------------------------------------------------------------------------------------------
GL_DEV u64 pos_reduce_ab(u64 A, u64 B) {
  // 2^32*B = lo32(B)*2^32 + hi32(B)*2^64 = lo32(B)*2^32 + hi32(B)*eps
  u64 u = mad_wide(hi32(B), GL_EPS, A);  // < 2^42
  u64 r = u + ((u64)lo32(B) << 32);
  return r < u ? r + GL_EPS : r;  // wrapped r < 2^42: no second carry
}

template <int R>
GL_DEV u64 pos_mds_row(const u32 (&lo)[12], const u32 (&hi)[12], u64 rc) {
  constexpr u32 CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
  u64 A = lo32(rc), B = hi32(rc);
#pragma unroll
  for (int i = 0; i < 12; i++) {
    A = mad_wide(lo[(i + R) % 12], CIRC[i], A);
    B = mad_wide(hi[(i + R) % 12], CIRC[i], B);
  }
  if (R == 0) {
    A = mad_wide(lo[0], 8u, A);
    B = mad_wide(hi[0], 8u, B);
  }
  return pos_reduce_ab(A, B);
}

// s <- MDS*s + rc[0..12]
GL_DEV void pos_mds_rc(u64 (&s)[12], const u64 *rc) {
  u32 lo[12], hi[12];
#pragma unroll
  for (int i = 0; i < 12; i++) {
    lo[i] = lo32(s[i]);
    hi[i] = hi32(s[i]);
  }
  s[0] = pos_mds_row<0>(lo, hi, rc[0]);
  s[1] = pos_mds_row<1>(lo, hi, rc[1]);
  s[2] = pos_mds_row<2>(lo, hi, rc[2]);
  s[3] = pos_mds_row<3>(lo, hi, rc[3]);
  s[4] = pos_mds_row<4>(lo, hi, rc[4]);
  s[5] = pos_mds_row<5>(lo, hi, rc[5]);
  s[6] = pos_mds_row<6>(lo, hi, rc[6]);
  s[7] = pos_mds_row<7>(lo, hi, rc[7]);
  s[8] = pos_mds_row<8>(lo, hi, rc[8]);
  s[9] = pos_mds_row<9>(lo, hi, rc[9]);
  s[10] = pos_mds_row<10>(lo, hi, rc[10]);
  s[11] = pos_mds_row<11>(lo, hi, rc[11]);
}

// Naive schedule (A.6): 30 x { +RC ; S-box (all lanes | lane 0) ; MDS }, 4 full + 22 partial + 4 full.
GL_DEV void poseidon_permute(u64 (&s)[12]) {
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl_add_c(s[i], c_pos_rc[i]);
  int r = 0;
#pragma unroll 1
  for (; r < 4; r++) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_pow7(s[i]);
    pos_mds_rc(s, c_pos_rc + 12 * (r + 1));
  }
#pragma unroll 1
  for (; r < 26; r++) {
    s[0] = gl_pow7(s[0]);
    pos_mds_rc(s, c_pos_rc + 12 * (r + 1));
  }
#pragma unroll 1
  for (; r < 30; r++) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_pow7(s[i]);
    pos_mds_rc(s, c_pos_rc + 12 * (r + 1));
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

GL_DEV void poseidon2_permute(u64 (&s)[12]) {
  p2_external(s);
  const u64 *rc = c_p2_rc;
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_pow7(gl_add_c(s[i], rc[i]));
    p2_external(s);
    rc += 12;
  }
#pragma unroll 1
  for (int r = 0; r < 22; r++) {
    s[0] = gl_pow7(gl_add_c(s[0], rc[0]));
    p2_internal(s);
    rc += 1;
  }
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_pow7(gl_add_c(s[i], rc[i]));
    p2_external(s);
    rc += 12;
  }
}

template <u32 KIND>
GL_DEV void permute(u64 (&s)[12]) {
  if (KIND == MP2_HASH_POSEIDON2) poseidon2_permute(s);
  else poseidon_permute(s);
}
