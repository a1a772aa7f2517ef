// Small decimation-in-frequency DFTs over Goldilocks whose twiddles are powers of two (shifts, no multiplier):
// w_8 = -2^24, w_4 = 2^48, 2^96 = -1 (plonky2's primitive_root_of_unity(3) = 2^120).
// Building blocks of the radix-8 passes in ntt.cu (plonky2_field fft_classic) and
// exercised by the device self-test (selftest.cu).
#pragma once
#include "gl.cuh"

// ---- multiplications by the powers of two that are roots of unity ---------------------------------
// x * 2^24: an 88-bit value {w0, w1, w2}
GL_DEV u64 gl_mul_2_24(u64 x) {
  u32 x0 = lo32(x), x1 = hi32(x);
  return gl_reduce128w(x0 << 24, __funnelshift_l(x0, x1, 24), x1 >> 8, 0u);
}
// x * 2^48: a 112-bit value {0, w1, w2, w3}
GL_DEV u64 gl_mul_2_48(u64 x) {
  u32 x0 = lo32(x), x1 = hi32(x);
  return gl_reduce128w(0u, x0 << 16, __funnelshift_l(x0, x1, 16), x1 >> 16);
}
// x * 2^72 = w2*2^64 + w3*2^96 + w4*2^128 with {w2,w3,w4} = x << 8;  2^64 = eps, 2^96 = -1,
// 2^128 = -2^32:   r = w2*eps - {w3, w4}   (a borrow is folded as -eps; the wrapped value is huge)
GL_DEV u64 gl_mul_2_72(u64 x) {
  u32 x0 = lo32(x), x1 = hi32(x);
  u32 w2 = x0 << 8, w3 = __funnelshift_l(x0, x1, 8), w4 = x1 >> 24;
  u32 r0, r1;
  asm("{\n\t.reg .u32 t0, t1, m;\n\t"
      "sub.cc.u32 t0, 0, %2;\n\tsubc.u32 t1, %2, 0;\n\t"                       // {t0,t1} = w2*eps
      "sub.cc.u32 %0, t0, %3;\n\tsubc.cc.u32 %1, t1, %4;\n\tsubc.u32 m, 0, 0;\n\t"
      "sub.cc.u32 %0, %0, m;\n\tsubc.u32 %1, %1, 0;\n\t}"
      : "=&r"(r0), "=&r"(r1)
      : "r"(w2), "r"(w3), "r"(w4));
  return pack64(r0, r1);
}

// x * 2^K for a compile-time 0 < K < 96 that is not a multiple of 32 (the 16th roots of unity are +-2^(12 j))
template <int K>
GL_DEV u64 gl_mul_pow2(u64 x) {
  static_assert(K > 0 && K < 96 && K % 32 != 0, "shift must be in (0, 96) and not a whole word");
  const u32 x0 = lo32(x), x1 = hi32(x);
  constexpr int S = K % 32;
  const u32 a = x0 << S, b = __funnelshift_l(x0, x1, S), c = x1 >> (32 - S);  // x << S = {a, b, c}
  if (K < 32) return gl_reduce128w(a, b, c, 0u);
  if (K < 64) return gl_reduce128w(0u, a, b, c);
  // K >= 64: a*2^64 + b*2^96 + c*2^128 = a*eps - {b, c}   (2^96 = -1, 2^128 = -2^32), as in gl_mul_2_72
  u32 r0, r1;
  asm("{\n\t.reg .u32 t0, t1, m;\n\t"
      "sub.cc.u32 t0, 0, %2;\n\tsubc.u32 t1, %2, 0;\n\t"
      "sub.cc.u32 %0, t0, %3;\n\tsubc.cc.u32 %1, t1, %4;\n\tsubc.u32 m, 0, 0;\n\t"
      "sub.cc.u32 %0, %0, m;\n\tsubc.u32 %1, %1, 0;\n\t}"
      : "=&r"(r0), "=&r"(r1)
      : "r"(a), "r"(b), "r"(c));
  return pack64(r0, r1);
}

// ---- small DFTs with shift twiddles; output position j holds frequency bitrev(j) -------------------
GL_DEV void gl_dft2(u64 (&x)[2]) {
  u64 a = gl_add(x[0], x[1]), b = gl_sub(x[0], x[1]);
  x[0] = a;
  x[1] = b;
}
GL_DEV void gl_dft4(u64 (&x)[4]) {  // w_4 = 2^48
  u64 a0 = gl_add(x[0], x[2]), a1 = gl_add(x[1], x[3]);
  u64 b0 = gl_sub(x[0], x[2]), b1 = gl_mul_2_48(gl_sub(x[1], x[3]));
  x[0] = gl_add(a0, a1);
  x[1] = gl_sub(a0, a1);
  x[2] = gl_add(b0, b1);
  x[3] = gl_sub(b0, b1);
}
GL_DEV void gl_dft8(u64 (&x)[8]) {  // w_8 = -2^24, w_8^2 = 2^48, w_8^3 = -2^72
  u64 a0 = gl_add(x[0], x[4]), a1 = gl_add(x[1], x[5]), a2 = gl_add(x[2], x[6]), a3 = gl_add(x[3], x[7]);
  u64 b0 = gl_sub(x[0], x[4]);
  u64 b1 = gl_mul_2_24(gl_sub(x[5], x[1]));
  u64 b2 = gl_mul_2_48(gl_sub(x[2], x[6]));
  u64 b3 = gl_mul_2_72(gl_sub(x[7], x[3]));
  u64 c0 = gl_add(a0, a2), c1 = gl_add(a1, a3), d0 = gl_sub(a0, a2), d1 = gl_mul_2_48(gl_sub(a1, a3));
  u64 e0 = gl_add(b0, b2), e1 = gl_add(b1, b3), f0 = gl_sub(b0, b2), f1 = gl_mul_2_48(gl_sub(b1, b3));
  x[0] = gl_add(c0, c1);
  x[1] = gl_sub(c0, c1);
  x[2] = gl_add(d0, d1);
  x[3] = gl_sub(d0, d1);
  x[4] = gl_add(e0, e1);
  x[5] = gl_sub(e0, e1);
  x[6] = gl_add(f0, f1);
  x[7] = gl_sub(f0, f1);
}
// 16 points: w_16 = -2^60 (plonky2's primitive_root_of_unity(4) = 2^156), so
//   w_16^k, k = 0..7:  1, -2^60, -2^24, 2^84, 2^48, 2^12, -2^72, -2^36
// (a negative sign is taken by swapping the operands of the subtraction that feeds the shift).
GL_DEV void gl_dft16(u64 (&x)[16]) {
  u64 lo[8], hi[8];
#pragma unroll
  for (int i = 0; i < 8; i++) lo[i] = gl_add(x[i], x[i + 8]);
  hi[0] = gl_sub(x[0], x[8]);
  hi[1] = gl_mul_pow2<60>(gl_sub(x[9], x[1]));
  hi[2] = gl_mul_pow2<24>(gl_sub(x[10], x[2]));
  hi[3] = gl_mul_pow2<84>(gl_sub(x[3], x[11]));
  hi[4] = gl_mul_pow2<48>(gl_sub(x[4], x[12]));
  hi[5] = gl_mul_pow2<12>(gl_sub(x[5], x[13]));
  hi[6] = gl_mul_pow2<72>(gl_sub(x[14], x[6]));
  hi[7] = gl_mul_pow2<36>(gl_sub(x[15], x[7]));
  gl_dft8(lo);
  gl_dft8(hi);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    x[i] = lo[i];
    x[i + 8] = hi[i];
  }
}
template <int RHO>
struct Dft;
template <>
struct Dft<1> {
  static GL_DEV void run(u64 (&x)[2]) { gl_dft2(x); }
};
template <>
struct Dft<2> {
  static GL_DEV void run(u64 (&x)[4]) { gl_dft4(x); }
};
template <>
struct Dft<3> {
  static GL_DEV void run(u64 (&x)[8]) { gl_dft8(x); }
};
template <>
struct Dft<4> {
  static GL_DEV void run(u64 (&x)[16]) { gl_dft16(x); }
};

