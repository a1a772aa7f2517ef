// Batched Goldilocks iNTT and coset low-degree extension, staged through shared memory.
//
// Replaces plonky2_field's PolynomialValues::ifft and PolynomialCoeffs::{lde, coset_fft_with_options}
// as used by PolynomialBatch::from_values / from_coeffs, and
// folds plonky2_util::{transpose, reverse_index_bits_in_place}'s row permutation into the transform:
//
//   * the LDE on the coset 7*<w_N>, N = n*2^r, is computed as 2^r independent size-n transforms
//     (coset k: coefficients scaled by (7*w_N^k)^j) -- the zero-padded top r stages never run;
//   * each size-n transform is a decimation-in-frequency network (natural in, bit-reversed out), and
//     leaf index L = (bitrev_r(k) << log n) | bitrev_n(m) is exactly the position the value
//     P(7*w_N^(k + 2^r m)) has in plonky2's bit-reversed leaf order -- so the network's raw output,
//     written to block bitrev_r(k), IS the leaf order: no separate bit-reversal pass exists;
//   * output is column-major over leaves (column c contiguous), which is what the leaf-hash kernel
//     wants for coalesced loads; the row-major `leaves` copy is produced by that kernel.
//
// n <= 2^14 : one pass, the whole line lives in shared memory (128 KB at 2^14).
// n >  2^14 : two passes (n = n1*n2, four-step), each pass a shared-memory transform on a tile of
//             adjacent lines so that every global access is a >= 128-byte run.
#include "internal.h"
#include "gl.cuh"

namespace mp2 {

static const u32 kMaxSingleLog = 14;  // 2^14 * 8 B = 128 KB of the 227 KB shared memory
static const u32 kTileLog = 13;       // target tile (elements) when several lines share a CTA

struct LdeMap {  // (column c, leaf L) -> offset in the leaf-ordered, column-major, shardable buffer
  u32 ls_log;    // log2(leaves per shard)
  size_t shard_stride, col_stride;
  GL_DEV size_t operator()(size_t L, size_t c) const {
    return (L >> ls_log) * shard_stride + c * col_stride + (L & (((size_t)1 << ls_log) - 1));
  }
};

// W[m] = w_T^m, T = 2^log_t.  Twiddle of stage `stage` (butterfly span 2^stage), index t:
// w_{2^(stage+1)}^t = W[t << (log_t - stage - 1)];  the inverse transform uses W[T - idx].
struct Roots {
  const u64 *W;
  u32 log_t;
  GL_DEV u64 get(size_t idx, bool inverse) const {
    size_t mask = ((size_t)1 << log_t) - 1;
    idx &= mask;
    if (inverse) idx = (((size_t)1 << log_t) - idx) & mask;
    return W[idx];
  }
};

// Decimation-in-frequency network over LINES = 2^lines_log independent lines of S = 2^s points held
// as sm[p*LINES + l]: consecutive threads take consecutive lines of the same butterfly, so shared
// memory accesses are conflict-free and the twiddle is a broadcast.  Leaves X[bitrev_s(p)] at p.
GL_DEV void smem_dif(u64 *sm, u32 s, u32 lines_log, Roots roots, bool inverse) {
  const u32 total = 1u << (s + lines_log - 1);  // butterflies per stage (s >= 1)
  const u32 lmask = (1u << lines_log) - 1;
  for (int stage = (int)s - 1; stage >= 0; stage--) {
    const u32 half = 1u << stage;
    for (u32 b = threadIdx.x; b < total; b += blockDim.x) {
      u32 l = b & lmask, bb = b >> lines_log;
      u32 t = bb & (half - 1);
      u32 i = ((bb >> stage) << (stage + 1)) + t;
      u64 *pi = sm + (((size_t)i << lines_log) + l);
      u64 *pj = pi + ((size_t)half << lines_log);
      u64 u = *pi, v = *pj;
      *pi = gl_add(u, v);
      u64 d = gl_sub(u, v);
      *pj = t ? gl_mul(d, roots.get((size_t)t << (roots.log_t - stage - 1), inverse)) : d;
    }
    __syncthreads();
  }
}

// ---- single pass: lines are columns ------------------------------------------------------------
// grid.x = ceil(ncols / LINES); inverse transform, natural-order output scaled by n^-1
__global__ void k_intt_single(const u64 *__restrict__ in, size_t in_stride, u64 *__restrict__ out,
                              size_t out_stride, u32 ncols, u32 s, u32 lines_log, Roots roots, u64 n_inv) {
  extern __shared__ u64 sm[];
  const u32 S = 1u << s, LINES = 1u << lines_log, c0 = blockIdx.x * LINES;
  for (u32 e = threadIdx.x; e < (S << lines_log); e += blockDim.x) {
    u32 p = e & (S - 1), l = e >> s, c = c0 + l;
    sm[((size_t)p << lines_log) + l] = c < ncols ? in[(size_t)c * in_stride + p] : 0;
  }
  __syncthreads();
  if (s) smem_dif(sm, s, lines_log, roots, true);
  for (u32 e = threadIdx.x; e < (S << lines_log); e += blockDim.x) {
    u32 k = e & (S - 1), l = e >> s, c = c0 + l;
    if (c < ncols)
      out[(size_t)c * out_stride + k] = gl_canon(gl_mul(sm[((size_t)brev_bits(k, s) << lines_log) + l], n_inv));
  }
}

// grid = (ceil(ncols / LINES), 2^r cosets): coset-scaled forward transform, leaf-ordered output
__global__ void k_lde_single(const u64 *__restrict__ coeffs, size_t in_stride, u64 *__restrict__ lde, LdeMap map,
                             u32 ncols, u32 s, u32 lines_log, u32 rate_bits, Roots roots,
                             const u64 *__restrict__ pow7) {
  extern __shared__ u64 sm[];
  const u32 S = 1u << s, LINES = 1u << lines_log, c0 = blockIdx.x * LINES, k = blockIdx.y;
  for (u32 e = threadIdx.x; e < (S << lines_log); e += blockDim.x) {
    u32 p = e & (S - 1), l = e >> s, c = c0 + l;
    u64 v = 0;
    if (c < ncols) {
      v = coeffs[(size_t)c * in_stride + p];
      v = gl_mul(v, gl_mul(pow7[p], roots.get((size_t)k * p, false)));  // (7*w_N^k)^p
    }
    sm[((size_t)p << lines_log) + l] = v;
  }
  __syncthreads();
  if (s) smem_dif(sm, s, lines_log, roots, false);
  const size_t block_base = (size_t)brev_bits(k, rate_bits) << s;
  for (u32 e = threadIdx.x; e < (S << lines_log); e += blockDim.x) {
    u32 p = e & (S - 1), l = e >> s, c = c0 + l;
    if (c < ncols) lde[map(block_base + p, c)] = gl_canon(sm[((size_t)p << lines_log) + l]);
  }
}

// ---- two passes (four-step): n = n1*n2, j = j1*n2 + j2, k = k1 + n1*k2 ---------------------------
struct TwoPass {
  u32 n_log, a, b;   // n1 = 2^a (strided pass 1), n2 = 2^b (contiguous pass 2)
  u32 lines_log;
  u32 rate_bits;
  int inverse;       // 1: iNTT natural -> natural; 0: coset LDE natural -> leaf order
  u64 n_inv;
};

// pass 1: tile = LINES adjacent j2; size-n1 transform over j1 (stride n2); then the four-step
// twiddle rho^(j2*k1).  grid = (n2 / LINES, ncols, cosets)
__global__ void k_pass1(const u64 *__restrict__ in, size_t in_stride, u64 *__restrict__ out, size_t out_stride,
                        LdeMap map, TwoPass tp, Roots roots, const u64 *__restrict__ pow7) {
  extern __shared__ u64 sm[];
  const u32 LINES = 1u << tp.lines_log, S = 1u << tp.a;
  const size_t n2 = (size_t)1 << tp.b;
  const size_t c = blockIdx.y, k = blockIdx.z, q0 = (size_t)blockIdx.x * LINES;
  for (u32 e = threadIdx.x; e < (S << tp.lines_log); e += blockDim.x) {
    u32 l = e & (LINES - 1), p = e >> tp.lines_log;
    size_t j = (size_t)p * n2 + q0 + l;
    u64 v = in[c * in_stride + j];
    if (!tp.inverse) v = gl_mul(v, gl_mul(pow7[j], roots.get(k * j, false)));
    sm[e] = v;  // == sm[p*LINES + l]
  }
  __syncthreads();
  smem_dif(sm, tp.a, tp.lines_log, roots, tp.inverse);
  const size_t block_base = (size_t)brev_bits((u32)k, tp.rate_bits) << tp.n_log;
  for (u32 e = threadIdx.x; e < (S << tp.lines_log); e += blockDim.x) {
    u32 l = e & (LINES - 1), p = e >> tp.lines_log;
    size_t j2 = q0 + l, k1 = brev_bits(p, tp.a);
    u64 w = roots.get((j2 * k1) << (roots.log_t - tp.n_log), tp.inverse);
    u64 v = gl_mul(sm[e], w);
    if (tp.inverse) out[c * out_stride + k1 * n2 + j2] = v;       // row k1 (natural)
    else out[map(block_base + (size_t)p * n2 + j2, c)] = v;       // row bitrev(k1) = p
  }
}

// pass 2: tile = LINES adjacent rows; size-n2 transform along each (contiguous) row.
// grid = (n1 / LINES, ncols, cosets)
__global__ void k_pass2(const u64 *__restrict__ in, size_t in_stride, u64 *__restrict__ out, size_t out_stride,
                        LdeMap map, TwoPass tp, Roots roots) {
  extern __shared__ u64 sm[];
  const u32 LINES = 1u << tp.lines_log, S = 1u << tp.b;
  const size_t n1 = (size_t)1 << tp.a, n2 = (size_t)1 << tp.b;
  const size_t c = blockIdx.y, k = blockIdx.z, r0 = (size_t)blockIdx.x * LINES;
  const size_t block_base = (size_t)brev_bits((u32)k, tp.rate_bits) << tp.n_log;
  for (u32 e = threadIdx.x; e < (S << tp.lines_log); e += blockDim.x) {
    u32 p = e & (S - 1), l = e >> tp.b;
    size_t row = r0 + l;
    sm[((size_t)p << tp.lines_log) + l] =
        tp.inverse ? in[c * in_stride + row * n2 + p] : in[map(block_base + row * n2 + p, c)];
  }
  __syncthreads();
  smem_dif(sm, tp.b, tp.lines_log, roots, tp.inverse);
  if (tp.inverse) {
    for (u32 e = threadIdx.x; e < (S << tp.lines_log); e += blockDim.x) {
      u32 l = e & (LINES - 1), p = e >> tp.lines_log;
      size_t k1 = r0 + l, k2 = brev_bits(p, tp.b);
      out[c * out_stride + k1 + n1 * k2] = gl_canon(gl_mul(sm[e], tp.n_inv));
    }
  } else {
    for (u32 e = threadIdx.x; e < (S << tp.lines_log); e += blockDim.x) {
      u32 p = e & (S - 1), l = e >> tp.b;
      out[map(block_base + (r0 + l) * n2 + p, c)] = gl_canon(sm[((size_t)p << tp.lines_log) + l]);
    }
  }
}

__global__ void k_canonicalize(const u64 *__restrict__ in, size_t in_stride, u64 *__restrict__ out, size_t out_stride,
                               size_t n, size_t total) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  size_t c = t / n, i = t % n;
  out[c * out_stride + i] = gl_canon(in[c * in_stride + i]);
}

// ------------------------------------------------------------------------------------------------
static u32 ceil_log2(size_t x) {
  u32 l = 0;
  while (((size_t)1 << l) < x) l++;
  return l;
}
static u32 threads_for(u32 tile_log) {
  u32 t = tile_log >= 1 ? 1u << (tile_log - 1) : 1;  // one butterfly per thread...
  if (t > 1024) t = 1024;                            // ...up to a full CTA
  if (t < 32) t = 32;
  return t;
}
template <typename K>
static Status allow_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) MP2_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return "";
}

Status ntt_canonicalize(const u64 *in, size_t in_stride, u64 *out, size_t out_stride, size_t ncols, size_t n,
                        cudaStream_t st) {
  size_t total = ncols * n;
  if (!total) return "";
  { ProfScope _p("k_canonicalize", st); k_canonicalize<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, in_stride, out, out_stride, n, total); }
  MP2_LAUNCH_CHECK();
  return "";
}

static Status split_two_pass(u32 n_log, TwoPass *tp) {
  if (n_log > 2 * kMaxSingleLog - 2) return "polynomial degree 2^" + std::to_string(n_log) + " not supported (max 2^26)";
  tp->n_log = n_log;
  tp->b = (n_log + 1) / 2;
  tp->a = n_log - tp->b;
  tp->lines_log = kTileLog - tp->b;  // a <= b <= 13
  return "";
}

Status ntt_intt(const u64 *values, size_t in_stride, u64 *coeffs, size_t out_stride, size_t ncols, u32 n_log,
                cudaStream_t st) {
  if (ncols == 0) return "";
  if (n_log > 32) return "n_log exceeds the field's two-adicity (32)";
  const u64 n_inv = h_inv((u64)1 << n_log);
  Roots roots;
  roots.log_t = n_log;
  MP2_TRY(table_roots(n_log, st, &roots.W));
  LdeMap none = {0, 0, 0};
  if (n_log <= kMaxSingleLog) {
    u32 lines_log = n_log >= kTileLog ? 0 : std::min(kTileLog - n_log, ceil_log2(ncols));
    u32 tile_log = n_log + lines_log;
    size_t smem = sizeof(u64) << tile_log;
    MP2_TRY(allow_smem(k_intt_single, smem));
    unsigned grid = (unsigned)((ncols + ((size_t)1 << lines_log) - 1) >> lines_log);
    { ProfScope _p("k_intt_single", st); k_intt_single<<<grid, threads_for(tile_log), smem, st>>>(values, in_stride, coeffs, out_stride, (u32)ncols,
                                                              n_log, lines_log, roots, n_inv); }
    MP2_LAUNCH_CHECK();
    return "";
  }
  TwoPass tp;
  MP2_TRY(split_two_pass(n_log, &tp));
  tp.rate_bits = 0;
  tp.inverse = 1;
  tp.n_inv = n_inv;
  const size_t n = (size_t)1 << n_log;
  u64 *tmp = nullptr;
  MP2_CUDA(cudaMallocAsync(&tmp, sizeof(u64) * n * ncols, st));
  {
    u32 tile_log = tp.a + tp.lines_log;
    size_t smem = sizeof(u64) << tile_log;
    MP2_TRY(allow_smem(k_pass1, smem));
    dim3 grid((unsigned)(((size_t)1 << tp.b) >> tp.lines_log), (unsigned)ncols, 1);
    { ProfScope _p("k_pass1", st); k_pass1<<<grid, threads_for(tile_log), smem, st>>>(values, in_stride, tmp, n, none, tp, roots, nullptr); }
    MP2_LAUNCH_CHECK();
  }
  {
    u32 lines_log = std::min(tp.lines_log, tp.a);
    TwoPass tp2 = tp;
    tp2.lines_log = lines_log;
    u32 tile_log = tp.b + lines_log;
    size_t smem = sizeof(u64) << tile_log;
    MP2_TRY(allow_smem(k_pass2, smem));
    dim3 grid((unsigned)(((size_t)1 << tp.a) >> lines_log), (unsigned)ncols, 1);
    { ProfScope _p("k_pass2", st); k_pass2<<<grid, threads_for(tile_log), smem, st>>>(tmp, n, coeffs, out_stride, none, tp2, roots); }
    MP2_LAUNCH_CHECK();
  }
  MP2_CUDA(cudaFreeAsync(tmp, st));
  return "";
}

Status ntt_coset_lde(const u64 *coeffs, size_t in_stride, u64 *lde, size_t lde_stride, size_t ncols, u32 n_log,
                     u32 rate_bits, u32 shard_log, size_t shard_stride, cudaStream_t st) {
  if (ncols == 0) return "";
  const u32 N_log = n_log + rate_bits;
  if (N_log > 32) return "n_log + rate_bits exceeds the field's two-adicity (32)";
  if (shard_log > N_log) return "shard_log larger than log2(number of leaves)";
  if (rate_bits > 15) return "rate_bits too large";
  Roots roots;
  roots.log_t = N_log;
  MP2_TRY(table_roots(N_log, st, &roots.W));
  const u64 *pow7 = nullptr;
  MP2_TRY(table_shift_powers(n_log, st, &pow7));
  LdeMap map = {N_log - shard_log, shard_log ? shard_stride : 0, lde_stride};
  const unsigned cosets = 1u << rate_bits;
  if (n_log <= kMaxSingleLog) {
    u32 lines_log = n_log >= kTileLog ? 0 : std::min(kTileLog - n_log, ceil_log2(ncols));
    u32 tile_log = n_log + lines_log;
    size_t smem = sizeof(u64) << tile_log;
    MP2_TRY(allow_smem(k_lde_single, smem));
    dim3 grid((unsigned)((ncols + ((size_t)1 << lines_log) - 1) >> lines_log), cosets, 1);
    { ProfScope _p("k_lde_single", st); k_lde_single<<<grid, threads_for(tile_log), smem, st>>>(coeffs, in_stride, lde, map, (u32)ncols, n_log,
                                                             lines_log, rate_bits, roots, pow7); }
    MP2_LAUNCH_CHECK();
    return "";
  }
  TwoPass tp;
  MP2_TRY(split_two_pass(n_log, &tp));
  tp.rate_bits = rate_bits;
  tp.inverse = 0;
  tp.n_inv = 1;
  {
    u32 tile_log = tp.a + tp.lines_log;
    size_t smem = sizeof(u64) << tile_log;
    MP2_TRY(allow_smem(k_pass1, smem));
    dim3 grid((unsigned)(((size_t)1 << tp.b) >> tp.lines_log), (unsigned)ncols, cosets);
    { ProfScope _p("k_pass1", st); k_pass1<<<grid, threads_for(tile_log), smem, st>>>(coeffs, in_stride, lde, 0, map, tp, roots, pow7); }
    MP2_LAUNCH_CHECK();
  }
  {
    u32 lines_log = std::min(tp.lines_log, tp.a);
    TwoPass tp2 = tp;
    tp2.lines_log = lines_log;
    u32 tile_log = tp.b + lines_log;
    size_t smem = sizeof(u64) << tile_log;
    MP2_TRY(allow_smem(k_pass2, smem));
    dim3 grid((unsigned)(((size_t)1 << tp.a) >> lines_log), (unsigned)ncols, cosets);
    { ProfScope _p("k_pass2", st); k_pass2<<<grid, threads_for(tile_log), smem, st>>>(lde, 0, lde, 0, map, tp2, roots); }
    MP2_LAUNCH_CHECK();
  }
  return "";
}

}  // namespace mp2
