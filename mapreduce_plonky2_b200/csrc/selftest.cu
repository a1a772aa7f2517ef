// Device-side self-test of the Goldilocks primitives in gl.cuh (plonky2_field::goldilocks_field)
// against 128-bit integer arithmetic evaluated by definition on the same device, plus a register-only
// throughput probe of the two arithmetic inner loops (S-box layer, radix-8 butterfly + twiddles).
//
// Why it exists: the carry-chain formulations are exact only through case analyses about rare carries
// (a second fold, a net borrow); random data almost never reaches those cases, so they are driven here
// with every combination of "corner" words around 0, 2^32, 2^63, p and 2^64.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mp2gpu.h"
#include "internal.h"
#include "gl.cuh"
#include "dft.cuh"

namespace mp2 {
namespace {

typedef unsigned __int128 u128;

__device__ u64 ref_mod(u128 x) { return (u64)(x % (u128)GL_P); }

__device__ u64 corner64(u32 i, u64 seed) {
  // 48 hand-picked values, then pseudo-random ones
  const u64 P = GL_P;
  const u64 tab[48] = {0, 1, 2, 3, 0x7FFFFFFFull, 0x80000000ull, 0xFFFFFFFEull, 0xFFFFFFFFull, 0x100000000ull,
                       0x100000001ull, 0x1FFFFFFFFull, 0x200000000ull, 0x7FFFFFFFFFFFFFFFull, 0x8000000000000000ull,
                       0x8000000000000001ull, P - 0x100000000ull, P - 2, P - 1, P, P + 1, P + 2, P + 0x7FFFFFFFull,
                       P + 0xFFFFFFFDull, P + 0xFFFFFFFEull, 0xFFFFFFFF00000000ull, 0xFFFFFFFE00000000ull,
                       0xFFFFFFFEFFFFFFFFull, 0xFFFFFFFF7FFFFFFFull, 0xFFFFFFFF80000000ull, 0xFFFFFFFFFFFFFFFDull,
                       0xFFFFFFFFFFFFFFFEull, 0xFFFFFFFFFFFFFFFFull, 0x00000001FFFFFFFFull, 0x0000000100000000ull,
                       0x7FFFFFFF00000000ull, 0x7FFFFFFFFFFFFFFFull, 0x80000000FFFFFFFFull, 0x8000000000000000ull,
                       0xFFFFFFFF00000002ull, 0xFFFFFFFF0000FFFFull, 0x0000FFFF00000000ull, 0x00000000FFFF0000ull,
                       0xAAAAAAAAAAAAAAAAull, 0x5555555555555555ull, 0xFFFF0000FFFF0000ull, 0x0000FFFF0000FFFFull,
                       0xFFFFFFFE00000001ull, 0xFFFFFFFE00000002ull};
  if (i < 48) return tab[i];
  u64 z = seed + (u64)i * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ u32 corner32(u32 i) {
  const u32 tab[8] = {0u, 1u, 2u, 0x7FFFFFFFu, 0x80000000u, 0xFFFFFFFDu, 0xFFFFFFFEu, 0xFFFFFFFFu};
  return tab[i & 7];
}

enum { T_ADD, T_ADDC, T_SUB, T_MUL, T_SQR, T_MULADD, T_REDUCE, T_POW7, T_SHIFT24, T_SHIFT48, T_SHIFT72, T_POW2, T_COUNT };

// one thread per (i, j) pair of an NV x NV grid of corner/random values
__global__ void k_field_selftest(u32 nv, u64 seed, unsigned long long *bad) {
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nv * nv) return;
  const u32 i = t / nv, j = t % nv;
  const u64 a = corner64(i, seed), b = corner64(j, seed), c = corner64((i * 7 + j * 13) % nv, seed ^ 0x5555);
  const u64 bc = b >= GL_P ? b - GL_P : b;  // canonical second operand for gl_add_c
  auto check = [&](int test, u64 got, u64 want) {
    if (gl_canon(got) != want) atomicAdd(bad + test, 1ull);
  };
  check(T_ADD, gl_add(a, b), ref_mod((u128)a + b));
  check(T_ADDC, gl_add_c(a, bc), ref_mod((u128)a + bc));
  check(T_SUB, gl_sub(a, b), ref_mod((u128)a + (u128)GL_P * 2 - b));
  check(T_MUL, gl_mul(a, b), ref_mod((u128)a * b));
  check(T_SQR, gl_sqr(a ^ b), ref_mod((u128)(a ^ b) * (a ^ b)));
  check(T_MULADD, gl_mul_add(a, b, c), ref_mod((u128)a * b + c));
  {
    const u64 x = a + j;
    u64 y = ref_mod((u128)x), q = 1;
    for (int k = 0; k < 7; k++) q = ref_mod((u128)q * y);
    check(T_POW7, gl_pow7(x), q);
  }
  check(T_SHIFT24, gl_mul_2_24(a ^ (b << 1)), ref_mod((u128)(a ^ (b << 1)) << 24));
  check(T_SHIFT48, gl_mul_2_48(a ^ (b << 1)), ref_mod((u128)(a ^ (b << 1)) << 48));
  {
    const u64 x = a ^ (b << 1);
    const u64 x72 = ref_mod((u128)ref_mod((u128)x << 48) << 24);
    check(T_SHIFT72, gl_mul_2_72(x), x72);
  }
  {  // every shift the radix-16 butterfly uses
    const u64 x = a ^ (b >> 1);
    auto sh = [&](int k) { u64 r = ref_mod((u128)x); for (int i = 0; i < k; i++) r = ref_mod((u128)r << 1); return r; };
    check(T_POW2, gl_mul_pow2<12>(x), sh(12));
    check(T_POW2, gl_mul_pow2<24>(x), sh(24));
    check(T_POW2, gl_mul_pow2<36>(x), sh(36));
    check(T_POW2, gl_mul_pow2<48>(x), sh(48));
    check(T_POW2, gl_mul_pow2<60>(x), sh(60));
    check(T_POW2, gl_mul_pow2<72>(x), sh(72));
    check(T_POW2, gl_mul_pow2<84>(x), sh(84));
  }
  // reduce128w over all 8^4 corner-word combinations (first 4096 threads) and over the value grid
  {
    u32 w0, w1, w2, w3;
    if (t < 4096) {
      w0 = corner32(t), w1 = corner32(t >> 3), w2 = corner32(t >> 6), w3 = corner32(t >> 9);
    } else {
      w0 = lo32(a), w1 = hi32(a), w2 = lo32(b), w3 = hi32(b);
    }
    const u128 v = (u128)w0 + ((u128)w1 << 32) + ((u128)w2 << 64) + ((u128)w3 << 96);
    check(T_REDUCE, gl_reduce128w(w0, w1, w2, w3), ref_mod(v));
  }
}

// ---- register-only throughput probes --------------------------------------------------------------------
// S-box shape: 12 independent x^7 chains per thread and trip (what a full round issues)
__global__ void __launch_bounds__(128) k_probe_pow7(u64 *out, int trips) {
  u64 s[12];
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = (u64)threadIdx.x * 0x9E3779B97F4A7C15ull + i + blockIdx.x;
#pragma unroll 1
  for (int t = 0; t < trips; t++) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_pow7(s[i]);
  }
  u64 acc = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) acc ^= s[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
// transform shape: radix-8 butterfly with shift twiddles followed by 7 general twiddle multiplications
__global__ void __launch_bounds__(256) k_probe_dft8(u64 *out, int trips) {
  u64 x[8], w[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    x[i] = (u64)threadIdx.x * 0x9E3779B97F4A7C15ull + i + blockIdx.x;
    w[i] = x[i] * 0xBF58476D1CE4E5B9ull + 1;
  }
#pragma unroll 1
  for (int t = 0; t < trips; t++) {
    gl_dft8(x);
#pragma unroll
    for (int i = 1; i < 8; i++) x[i] = gl_mul(x[i], w[i]);
  }
  u64 acc = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) acc ^= x[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// one butterfly per thread on caller data (host-checked against the DFT by definition in tests/test_gpu_field.py)
__global__ void k_debug_dft(u64 *io, u32 log_points, size_t count) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  if (log_points == 4) {
    u64 x[16];
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = io[16 * t + k];
    gl_dft16(x);
#pragma unroll
    for (int k = 0; k < 16; k++) io[16 * t + k] = gl_canon(x[k]);
  } else {
    u64 x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = io[8 * t + k];
    gl_dft8(x);
#pragma unroll
    for (int k = 0; k < 8; k++) io[8 * t + k] = gl_canon(x[k]);
  }
}

}  // namespace
}  // namespace mp2

using namespace mp2;

extern "C" {

const char *mp2gpu_debug_field_selftest(uint64_t *mismatches_out, size_t ntests) {
  auto fail = [](const std::string &s) -> const char * {
    char *p = (char *)malloc(s.size() + 1);
    if (p) memcpy(p, s.c_str(), s.size() + 1);
    return p;
  };
  if (!mismatches_out || ntests < T_COUNT) return fail("mismatches_out must hold at least 12 counters");
  unsigned long long *bad = nullptr;
  if (cudaMalloc(&bad, sizeof(unsigned long long) * T_COUNT) != cudaSuccess)
    return fail("no usable CUDA device (this library has no CPU fallback)");
  cudaMemset(bad, 0, sizeof(unsigned long long) * T_COUNT);
  const u32 nv = 1024;  // 48 corner values + 976 pseudo-random ones: ~1M pairs
  k_field_selftest<<<(nv * nv + 255) / 256, 256>>>(nv, 0x6D7032ull, bad);
  count_launch();
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(bad);
    return fail(std::string("field selftest kernel: ") + cudaGetErrorString(e));
  }
  std::vector<unsigned long long> h(T_COUNT);
  cudaMemcpy(h.data(), bad, sizeof(unsigned long long) * T_COUNT, cudaMemcpyDeviceToHost);
  cudaFree(bad);
  for (size_t i = 0; i < ntests; i++) mismatches_out[i] = i < T_COUNT ? h[i] : 0;
  return nullptr;
}

const char *mp2gpu_debug_dft(uint64_t *io, uint32_t log_points, size_t count) {
  auto fail = [](const char *s) -> const char * { return strdup(s); };
  if (!io || (log_points != 3 && log_points != 4)) return fail("mp2gpu_debug_dft: io must be non-null and log_points 3 or 4");
  if (!count) return nullptr;
  const size_t elems = count << log_points;
  u64 *d = nullptr;
  if (cudaMalloc(&d, elems * sizeof(u64)) != cudaSuccess) return fail("no usable CUDA device (this library has no CPU fallback)");
  cudaMemcpy(d, io, elems * sizeof(u64), cudaMemcpyHostToDevice);
  k_debug_dft<<<(unsigned)((count + 127) / 128), 128>>>(d, log_points, count);
  count_launch();
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) cudaMemcpy(io, d, elems * sizeof(u64), cudaMemcpyDeviceToHost);
  cudaFree(d);
  return e == cudaSuccess ? nullptr : fail(cudaGetErrorString(e));
}

// thread-level operations per clock per SM: out[0] = x^7 (S-box layer shape), out[1] = radix-8 butterfly
// elements (8 per item, each item also does 7 twiddle multiplications)
const char *mp2gpu_debug_field_probe(double *ops_per_clk_per_sm_out) {
  auto fail = [](const char *s) -> const char * { return strdup(s); };
  if (!ops_per_clk_per_sm_out) return fail("null out");
  int dev = 0, nsm = 0, khz = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return fail("no usable CUDA device (this library has no CPU fallback)");
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  u64 *out = nullptr;
  const int ctas = nsm * 8;
  if (cudaMalloc(&out, sizeof(u64) * ctas * 256) != cudaSuccess) return fail("cudaMalloc failed in field probe");
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const double clk = (double)khz * 1e3;
  for (int which = 0; which < 2; which++) {
    double best = 0;
    const int trips = which == 0 ? 400 : 800;
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0);
      if (which == 0) k_probe_pow7<<<ctas, 128>>>(out, trips);
      else k_probe_dft8<<<ctas, 256>>>(out, trips);
      cudaEventRecord(e1);
      cudaDeviceSynchronize();
      count_launch();
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      const double ops = which == 0 ? (double)ctas * 128 * trips * 12 : (double)ctas * 256 * trips * 8;
      const double rate = ops / (ms * 1e-3) / clk / nsm;
      if (rep > 0 && rate > best) best = rate;
    }
    ops_per_clk_per_sm_out[which] = best;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return nullptr;
}

}  // extern "C"
