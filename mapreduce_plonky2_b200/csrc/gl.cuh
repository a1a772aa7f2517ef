// Goldilocks field arithmetic for sm_100a: p = 2^64 - 2^32 + 1, eps = 2^32 - 1 = 2^64 mod p.
// Replaces plonky2_field::goldilocks_field (field order also stated at
// mp2-common/src/group_hashing/utils.rs:51).  Everything is built from 32-bit IMAD.WIDE / IADD3
// chains: the B200 has no 64-bit integer multiplier, so a 64x64->128 product is four
// mad.wide.u32 and the reduction uses 2^64 = eps, 2^96 = -1 (mod p).
//
// Value conventions used by the kernels:
//   "canonical"  x <  p          -- what is written to memory that leaves the library
//   "loose"      x <  2^64       -- any u64; every routine here accepts loose inputs
#pragma once
#include <cstdint>

typedef unsigned long long u64;
typedef unsigned int u32;

#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPS 0xFFFFFFFFu

#define GL_DEV __device__ __forceinline__

GL_DEV u64 gl_canon(u64 a) { return a >= GL_P ? a - GL_P : a; }

GL_DEV u32 lo32(u64 x) { return (u32)x; }
GL_DEV u32 hi32(u64 x) { return (u32)(x >> 32); }
GL_DEV u64 pack64(u32 lo, u32 hi) { return ((u64)hi << 32) | lo; }

// a*b + c with 32-bit a, b and 64-bit c: one IMAD.WIDE.U32
GL_DEV u64 mad_wide(u32 a, u32 b, u64 c) {
  u64 d;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
  return d;
}
GL_DEV u64 mul_wide(u32 a, u32 b) {
  u64 d;
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(d) : "r"(a), "r"(b));
  return d;
}

// loose + loose -> loose.  2^64 = eps, so a carry out is folded back as +eps; a second carry is
// possible only when both inputs are >= 2^64 - 2^32, and is folded the same way.
GL_DEV u64 gl_add(u64 a, u64 b) {
  u64 s = a + b;
  if (s < a) {
    u64 t = s + GL_EPS;
    s = t < s ? t + GL_EPS : t;
  }
  return s;
}
// canonical/loose a, CANONICAL b -> loose  (single fold is enough: s_wrapped <= a - 2^32 + 1... )
GL_DEV u64 gl_add_c(u64 a, u64 b_canonical) {
  u64 s = a + b_canonical;
  return s < a ? s + GL_EPS : s;
}
// loose - loose -> loose
GL_DEV u64 gl_sub(u64 a, u64 b) {
  u64 d = a - b;
  if (a < b) {
    u64 t = d - GL_EPS;       // borrow: -2^64 = -eps
    d = t > d ? t - GL_EPS : t;  // second borrow possible only for loose, near-zero results
  }
  return d;
}

// x = lo + 2^64 * hi (hi < 2^32)  ->  loose.   2^64 = eps.
GL_DEV u64 gl_reduce96(u64 lo, u32 hi) {
  u64 t = mul_wide(hi, GL_EPS);  // hi*eps < 2^64 - 2^33 + 1
  u64 r = lo + t;
  return r < lo ? r + GL_EPS : r;  // wrapped r <= lo - 2^33, cannot carry twice
}

// x = lo + 2^64 * hi (full 128 bit) -> loose.   hi = hh*2^32 + hl : 2^96 = -1, 2^64 = eps.
GL_DEV u64 gl_reduce128(u64 lo, u64 hi) {
  u32 hh = hi32(hi), hl = lo32(hi);
  u64 t0 = lo - hh;
  if (lo < hh) t0 -= GL_EPS;  // wrapped value >= 2^64 - 2^32 + 1 > eps: no second borrow
  u64 t1 = mul_wide(hl, GL_EPS);
  u64 r = t0 + t1;
  return r < t0 ? r + GL_EPS : r;
}

// full 64x64 -> 128 product from four 32x32 IMAD.WIDE
GL_DEV void mul64x64(u64 a, u64 b, u64 &lo, u64 &hi) {
  u32 a0 = lo32(a), a1 = hi32(a), b0 = lo32(b), b1 = hi32(b);
  u64 t0 = mul_wide(a0, b0);
  u64 t1 = mad_wide(a0, b1, (u64)hi32(t0));   // <= (2^32-1)^2 + 2^32-1 : no overflow
  u64 t2 = mad_wide(a1, b0, (u64)lo32(t1));   // same bound
  u64 t3 = mad_wide(a1, b1, (u64)hi32(t1));   // <= (2^32-1)^2 + 2(2^32-1) = 2^64-1 after next add
  t3 += hi32(t2);
  lo = pack64(lo32(t0), lo32(t2));
  hi = t3;
}

// loose * loose -> loose
GL_DEV u64 gl_mul(u64 a, u64 b) {
  u64 lo, hi;
  mul64x64(a, b, lo, hi);
  return gl_reduce128(lo, hi);
}
GL_DEV u64 gl_sqr(u64 a) { return gl_mul(a, a); }

// x^7: 4 multiplications (S-box of both permutations)
GL_DEV u64 gl_pow7(u64 x) {
  u64 x2 = gl_sqr(x);
  u64 x3 = gl_mul(x2, x);
  u64 x4 = gl_sqr(x2);
  return gl_mul(x3, x4);
}

GL_DEV u64 gl_pow(u64 a, u64 e) {
  u64 r = 1;
  while (e) {
    if (e & 1) r = gl_mul(r, a);
    a = gl_sqr(a);
    e >>= 1;
  }
  return r;
}

GL_DEV u32 brev_bits(u32 x, u32 bits) { return bits ? (__brev(x) >> (32 - bits)) : 0u; }
