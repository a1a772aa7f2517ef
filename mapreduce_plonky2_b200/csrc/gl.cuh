// Goldilocks field arithmetic for sm_100a: p = 2^64 - 2^32 + 1, eps = 2^32 - 1 = 2^64 mod p.
// Replaces plonky2_field::goldilocks_field (field order also stated at
// mp2-common/src/group_hashing/utils.rs:51).
//
// The B200 has no 64-bit integer multiplier: a 64x64->128 product is four IMAD.WIDE.U32, and on this
// part IMAD.WIDE / IMAD.HI issue at ~1/3.2 of the plain 32-bit IMAD rate (measured by
// tools/intpipe_peak.cu: 20 vs 63 thread-instr/clk/SM).  Additions, subtractions and reductions are
// written as explicit carry chains (add.cc/addc -> IADD3/IADD3.X), which ptxas spreads over the alu
// and fma pipes, and reductions use 2^64 = eps, 2^96 = -1 (mod p) with no multiply.  Every asm output
// that is written before the block's last input read is early-clobber ("=&r").
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
GL_DEV u64 pack64(u32 lo, u32 hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}

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
// (Reading the carry of an add chain with `subc m, 0, 0` to get -carry in one instruction -- 8 instead of 10 --
// is NOT usable: ptxas 12.9 miscompiles the mixed add.cc -> subc chain; the device self-test counted 514 470
// wrong sums out of 1 048 576 with it, profiles/r2_field_variants.txt.)
GL_DEV u64 gl_add(u64 a, u64 b) {
  u32 r0, r1;
  asm("{\n\t.reg .u32 c, m;\n\t"
      "add.cc.u32 %0, %2, %4;\n\taddc.cc.u32 %1, %3, %5;\n\taddc.u32 c, 0, 0;\n\t"
      "sub.u32 m, 0, c;\n\t"                                     // eps if carry
      "add.cc.u32 %0, %0, m;\n\taddc.cc.u32 %1, %1, 0;\n\taddc.u32 c, 0, 0;\n\t"
      "sub.u32 m, 0, c;\n\t"
      "add.cc.u32 %0, %0, m;\n\taddc.u32 %1, %1, 0;\n\t}"
      : "=&r"(r0), "=&r"(r1)
      : "r"(lo32(a)), "r"(hi32(a)), "r"(lo32(b)), "r"(hi32(b)));
  return pack64(r0, r1);
}
// loose a + CANONICAL b -> loose: the wrapped sum is <= p - 2, so one fold is enough
GL_DEV u64 gl_add_c(u64 a, u64 b_canonical) {
  u32 r0, r1;
  asm("{\n\t.reg .u32 c, m;\n\t"
      "add.cc.u32 %0, %2, %4;\n\taddc.cc.u32 %1, %3, %5;\n\taddc.u32 c, 0, 0;\n\t"
      "sub.u32 m, 0, c;\n\t"
      "add.cc.u32 %0, %0, m;\n\taddc.u32 %1, %1, 0;\n\t}"
      : "=&r"(r0), "=&r"(r1)
      : "r"(lo32(a)), "r"(hi32(a)), "r"(lo32(b_canonical)), "r"(hi32(b_canonical)));
  return pack64(r0, r1);
}
// loose - loose -> loose (borrow: -2^64 = -eps; a second borrow only for near-zero wrapped results)
GL_DEV u64 gl_sub(u64 a, u64 b) {
  u32 r0, r1;
  asm("{\n\t.reg .u32 m;\n\t"
      "sub.cc.u32 %0, %2, %4;\n\tsubc.cc.u32 %1, %3, %5;\n\tsubc.u32 m, 0, 0;\n\t"  // m = eps if borrow
      "sub.cc.u32 %0, %0, m;\n\tsubc.cc.u32 %1, %1, 0;\n\tsubc.u32 m, 0, 0;\n\t"
      "sub.cc.u32 %0, %0, m;\n\tsubc.u32 %1, %1, 0;\n\t}"
      : "=&r"(r0), "=&r"(r1)
      : "r"(lo32(a)), "r"(hi32(a)), "r"(lo32(b)), "r"(hi32(b)));
  return pack64(r0, r1);
}

// w0 + w1*2^32 + w2*2^64 + w3*2^96  ->  loose.        2^64 = eps = 2^32 - 1, 2^96 = -1:
//   value = {w0, w1} + {0, w2} - (w2 + w3)
// computed as a 64-bit add (carry c) and a 64-bit subtract of the 33-bit a = w2 + w3 (borrow b); the net
// k = c - b in {-1, 0, 1} is folded once as k*eps = {-k, k >> 1}.  No second fold can be needed:
//   k = +1: the true value is <= (2^64 - 1) + (2^32 - 1)*2^32, so wrapped + eps <= 2^64 - 2;
//   k = -1: the true value is >= -(2^33 - 2), so wrapped - eps >= 2^64 - 2^33 - 2^32 + 3 > 0.
// 11 instructions (the round-1 form -- borrow fold, w2*eps, carry fold -- was 13).
GL_DEV u64 gl_reduce128w(u32 w0, u32 w1, u32 w2, u32 w3) {
  u32 r0, r1;
  asm("{\n\t.reg .u32 x1, k, a0, a1, m0, m1;\n\t"
      "add.cc.u32 x1, %3, %4;\n\taddc.u32 k, 0, 0;\n\t"
      "add.cc.u32 a0, %4, %5;\n\taddc.u32 a1, 0, 0;\n\t"
      "sub.cc.u32 %0, %2, a0;\n\tsubc.cc.u32 %1, x1, a1;\n\tsubc.u32 k, k, 0;\n\t"
      "neg.s32 m0, k;\n\tshr.s32 m1, k, 1;\n\t"
      "add.cc.u32 %0, %0, m0;\n\taddc.u32 %1, %1, m1;\n\t}"
      : "=&r"(r0), "=&r"(r1)
      : "r"(w0), "r"(w1), "r"(w2), "r"(w3));
  return pack64(r0, r1);
}
GL_DEV u64 gl_reduce128(u64 lo, u64 hi) { return gl_reduce128w(lo32(lo), hi32(lo), lo32(hi), hi32(hi)); }

// x = lo + 2^64 * hi (hi < 2^32)  ->  loose
GL_DEV u64 gl_reduce96(u64 lo, u32 hi) { return gl_reduce128w(lo32(lo), hi32(lo), hi, 0u); }

// loose * loose -> loose.  The 128-bit product is left to the compiler: its lowering chains the carries
// through the multiplier itself (IMAD.WIDE.U32 with a carry-out predicate, IMAD.WIDE.U32.X with a carry-in),
// 4 IMAD.WIDE + 3 adds/moves, where four separate mul.wide + two add.cc chains needed 4 + 6.
GL_DEV u64 gl_mul(u64 a, u64 b) {
  const unsigned __int128 p = (unsigned __int128)a * b;
  return gl_reduce128((u64)p, (u64)(p >> 64));
}
// loose * loose + loose -> loose: the addend rides on the 128-bit product (a*b + c < 2^128), one reduction
GL_DEV u64 gl_mul_add(u64 a, u64 b, u64 c) {
  const unsigned __int128 p = (unsigned __int128)a * b + c;
  return gl_reduce128((u64)p, (u64)(p >> 64));
}
// loose^2 -> loose: three IMAD.WIDE (the cross product is added twice).  The compiler's own a*a keeps four
// wide multiplies, and the quarter-rate multiplier is what bounds the S-box (selftest.cu's probe: the fma pipe
// saturates first), so the square is spelled out.
GL_DEV u64 gl_sqr(u64 a) {
  u32 w0, w1, w2, w3;
  asm("{\n\t.reg .u64 p00, p01, p11;\n\t.reg .u32 h00, l01, h01, l11, h11;\n\t"
      "mul.wide.u32 p00, %4, %4;\n\tmul.wide.u32 p01, %4, %5;\n\tmul.wide.u32 p11, %5, %5;\n\t"
      "mov.b64 {%0, h00}, p00;\n\tmov.b64 {l01, h01}, p01;\n\tmov.b64 {l11, h11}, p11;\n\t"
      "add.cc.u32 %1, h00, l01;\n\taddc.cc.u32 %2, h01, l11;\n\taddc.u32 %3, h11, 0;\n\t"
      "add.cc.u32 %1, %1, l01;\n\taddc.cc.u32 %2, %2, h01;\n\taddc.u32 %3, %3, 0;\n\t}"
      : "=&r"(w0), "=&r"(w1), "=&r"(w2), "=&r"(w3)
      : "r"(lo32(a)), "r"(hi32(a)));
  return gl_reduce128w(w0, w1, w2, w3);
}

// x^7: 2 squarings + 2 multiplications (S-box of both permutations)
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
