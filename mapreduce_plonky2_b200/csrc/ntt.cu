// Batched Goldilocks iNTT and coset low-degree extension, staged through shared memory.
//
// Replaces plonky2_field's PolynomialValues::ifft and PolynomialCoeffs::{lde, coset_fft_with_options}
// as used by PolynomialBatch::from_values / from_coeffs, and
// folds plonky2_util::{transpose, reverse_index_bits_in_place}'s row permutation into the transform:
//
//   * the LDE on the coset 7*<w_N>, N = n*2^r, is computed as 2^r independent size-n transforms
//     (coset k: coefficients scaled by (7*w_N^k)^j, a precomputed table) -- the zero-padded top r
//     stages never run;
//   * each size-n transform is a decimation-in-frequency network (natural in, bit-reversed out), and
//     leaf index L = (bitrev_r(k) << log n) | bitrev_n(m) is exactly the position the value
//     P(7*w_N^(k + 2^r m)) has in plonky2's bit-reversed leaf order -- so the network's raw output,
//     written to block bitrev_r(k), IS the leaf order: no separate bit-reversal pass exists;
//   * output is column-major over leaves (column c contiguous), which is what the leaf-hash kernel
//     wants for coalesced loads; the row-major `leaves` copy is produced by that kernel;
//   * the inverse transform is the forward one with the output index reversed (coeffs[i] =
//     fft[(n-i)%n] / n, as plonky2's ifft does), so only forward root tables exist.
//
// The network runs radix-8 passes held in registers: a pass is an 8-point DFT whose internal twiddles
// are the 8th roots of unity -- in Goldilocks w_8 = -2^24, w_4 = 2^48 (2^96 = -1), i.e. shifts, no
// multiplier -- followed by ONE general twiddle multiplication per element.  14 stages = 5 shared
// memory round trips and ~4.4 general multiplications per element instead of 14 and 7.
//
// n <= 2^14 : one pass over global memory, the whole line lives in shared memory (128 KB at 2^14).
// n >  2^14 : two passes (n = n1*n2, four-step), each a shared-memory transform on a tile of adjacent
//             lines (4 lines of 2^10 points at n = 2^20: the strided pass moves whole 32-byte sectors).
#include <map>
#include <mutex>
#include <utility>

#include "internal.h"
#ifdef MP2_NTT_MUL_REDUCE_FMA
#define MP2_MUL_REDUCE_FMA 1
#endif
#include "gl.cuh"
#include "dft.cuh"

namespace mp2 {

#ifndef MP2_NTT_MAX_SINGLE_LOG
#define MP2_NTT_MAX_SINGLE_LOG 14
#endif
static const u32 kMaxSingleLog = MP2_NTT_MAX_SINGLE_LOG;  // 2^14 * 8 B = 128 KB of the 227 KB shared memory
// Tile = 2^12 elements (32 KB) and 256 threads: 4 resident CTAs per SM whose load / transform / store
// phases interleave.  2^13-element tiles (2 resident CTAs) were 8 % slower on the 2^20 four-step path.
#ifndef MP2_NTT_TILE_LOG
#define MP2_NTT_TILE_LOG 12
#endif
#ifndef MP2_NTT_THREADS_SHIFT
#define MP2_NTT_THREADS_SHIFT 4  // threads = tile >> shift on the big tiles (2 radix-8 items per thread and pass)
#endif
static const u32 kTileLog = MP2_NTT_TILE_LOG;  // target tile (elements) when several lines share a CTA

struct LdeMap {  // (column c, leaf L) -> address in the leaf-ordered, column-major, shardable buffer
  u32 ls_log;    // log2(leaves per shard)
  size_t shard_stride, col_stride;
  // coset_rot: the launch visits the cosets in the order brev_r((z + coset_rot) mod 2^r), i.e. leaf blocks -- and
  // with them destination shards -- in rotated natural order, so that the ranks of a peer exchange, each passing
  // its own rotation, store to different destinations at any moment instead of all hitting rank 0 first
  u32 coset_rot;
  // peer mode: shard g of the output starts at bases[g] -- a buffer in rank g's HBM, mapped over NVLink
  // (the exchange of the sharded commitment happens in the store of the LDE kernel itself)
  u32 peer;
  u64 *bases[16];
  GL_DEV size_t operator()(size_t L, size_t c) const {
    return (L >> ls_log) * shard_stride + c * col_stride + (L & (((size_t)1 << ls_log) - 1));
  }
  GL_DEV u64 *ptr(u64 *local, size_t L, size_t c) const {
    if (peer) return bases[L >> ls_log] + c * col_stride + (L & (((size_t)1 << ls_log) - 1));
    return local + (*this)(L, c);
  }
};

// Optional shared-memory skew (element a at a + (a >> 4)) and twiddle prefetch.  The short-stride passes
// at the end of the network have 8-way bank conflicts (half of all wavefronts) and the twiddle loads show up
// as long_scoreboard stalls, yet removing either changes the run time by < 4 % on B200 (measured, all four
// combinations): the transform is bound by the alu pipe (62-75 % busy), so both stay off.
// Staging batch: every global->shared staging loop first issues MP2_NTT_BATCH independent global loads per
// thread (data, coset scale, four-step twiddle) and only then consumes them.  ncu's source view of the first
// two-pass kernels showed 33 % (pass 1) / 25 % (pass 2) of all stall samples on the first consumer of those
// loads: one load in flight per thread leaves the DRAM/L2 latency exposed.
#ifndef MP2_NTT_BATCH
#define MP2_NTT_BATCH 4
#endif
#ifndef MP2_NTT_SKEW
#define MP2_NTT_SKEW 0
#endif
#if MP2_NTT_SKEW
GL_DEV size_t sidx(size_t a) { return a + (a >> 4); }
#else
GL_DEV size_t sidx(size_t a) { return a; }
#endif
static inline size_t smem_bytes_for(u32 tile_log) {
  size_t e = (size_t)1 << tile_log;
  return sizeof(u64) * (e + (e >> 4) + 1);
}

// One radix-2^RHO decimation-in-frequency pass over LINES = 2^lines_log lines of 2^s points held as
// sm[p*LINES + l].  The current sub-transform length is L = 2^ell; a work item is the R points
// p = blk*L + j*(L/R) + lo.  After the DFT, the output of frequency r (stored at j = bitrev(r)) is
// multiplied by w_L^(r*lo) = W[r*lo << (s - ell)], W the root table of the line size 2^s.
// All tile indices fit 32 bits (a tile is at most 2^14 elements).
template <int RHO>
GL_DEV void ntt_pass(u64 *sm, u32 ell, u32 s, u32 lines_log, const u64 *__restrict__ W) {
  constexpr int R = 1 << RHO;
  const u32 sub_log = ell - RHO;
  const u32 items = 1u << (s - RHO + lines_log);
  const u32 lmask = (1u << lines_log) - 1, lomask = (1u << sub_log) - 1;
  const u32 jstride = 1u << (sub_log + lines_log);
  for (u32 w = threadIdx.x; w < items; w += blockDim.x) {
    const u32 l = w & lmask, q = w >> lines_log;
    const u32 lo = q & lomask, blk = q >> sub_log;
    const u32 e1 = lo << (s - ell);
    const u32 a0 = (((blk << ell) + lo) << lines_log) + l;
    u64 x[R];
#pragma unroll
    for (int j = 0; j < R; j++) x[j] = sm[sidx(a0 + j * jstride)];
    Dft<RHO>::run(x);
    if (sub_log) {  // lo == 0 for the last pass: all twiddles are 1
#pragma unroll
      for (int j = 1; j < R; j++) x[j] = gl_mul(x[j], __ldg(W + (__brev((u32)j) >> (32 - RHO)) * e1));
    }
#pragma unroll
    for (int j = 0; j < R; j++) sm[sidx(a0 + j * jstride)] = x[j];
  }
  __syncthreads();
}

// Full network on a tile in ceil(s/4) passes: radix 16 where the stage count needs it, radix 8 otherwise
// (10 stages = 4+3+3, 12 = 4+4+4, 14 = 4+4+3+3, 9 = 3+3+3): one shared-memory round trip and one general twiddle
// multiplication per element and PASS, so fewer passes is fewer of both (round 1 ran 3+3+2+2 for 10 stages).
#ifndef MP2_NTT_RADIX16
#define MP2_NTT_RADIX16 1
#endif
GL_DEV void smem_ntt(u64 *sm, u32 s, u32 lines_log, const u64 *__restrict__ W) {
  u32 ell = s;
#if MP2_NTT_RADIX16
  u32 passes = (s + 3) / 4;
  int fours = (int)s - 3 * (int)passes;  // how many of the passes must be radix 16
  while (ell) {
    if (fours > 0 && ell >= 4) {
      ntt_pass<4>(sm, ell, s, lines_log, W);
      ell -= 4;
      fours--;
    } else if (ell >= 3) {
      ntt_pass<3>(sm, ell, s, lines_log, W);
      ell -= 3;
    } else if (ell == 2) {
      ntt_pass<2>(sm, ell, s, lines_log, W);
      ell -= 2;
    } else {
      ntt_pass<1>(sm, ell, s, lines_log, W);
      ell -= 1;
    }
  }
#else
  while (ell) {
    if (ell == 4 || ell == 2) {
      ntt_pass<2>(sm, ell, s, lines_log, W);
      ell -= 2;
    } else if (ell >= 3) {
      ntt_pass<3>(sm, ell, s, lines_log, W);
      ell -= 3;
    } else {
      ntt_pass<1>(sm, ell, s, lines_log, W);
      ell -= 1;
    }
  }
#endif
}

// ---- single pass over global memory: lines are columns ------------------------------------------
// grid.x = ceil(ncols / LINES); forward network + index reversal = inverse transform, scaled by n^-1
__global__ void __launch_bounds__(1024)
k_intt_single(const u64 *__restrict__ in, size_t in_stride, u64 *__restrict__ out, size_t out_stride, u32 ncols,
              u32 s, u32 lines_log, const u64 *__restrict__ W, u64 n_inv) {
  extern __shared__ u64 sm[];
  const u32 S = 1u << s, c0 = blockIdx.x << lines_log;
  const u32 total = S << lines_log, nthr = blockDim.x;
  for (u32 e0 = threadIdx.x; e0 < total; e0 += MP2_NTT_BATCH * nthr) {
    u64 v[MP2_NTT_BATCH];
#pragma unroll
    for (int b = 0; b < MP2_NTT_BATCH; b++) {
      u32 e = e0 + b * nthr, p = e & (S - 1), l = e >> s, c = c0 + l;
      v[b] = (e < total && c < ncols) ? in[(size_t)c * in_stride + p] : 0;
    }
#pragma unroll
    for (int b = 0; b < MP2_NTT_BATCH; b++) {
      u32 e = e0 + b * nthr, p = e & (S - 1), l = e >> s;
      if (e < total) sm[sidx((p << lines_log) + l)] = v[b];
    }
  }
  __syncthreads();
  smem_ntt(sm, s, lines_log, W);
  for (u32 e = threadIdx.x; e < (S << lines_log); e += blockDim.x) {
    u32 i = e & (S - 1), l = e >> s, c = c0 + l;
    u32 k = (S - i) & (S - 1);  // coeffs[i] = fft[(n - i) % n] / n
    if (c < ncols)
      out[(size_t)c * out_stride + i] = gl_canon(gl_mul(sm[sidx((brev_bits(k, s) << lines_log) + l)], n_inv));
  }
}

// grid = (ceil(ncols / LINES), 2^r cosets): coset-scaled forward transform, leaf-ordered output
__global__ void __launch_bounds__(1024)
k_lde_single(const u64 *__restrict__ coeffs, size_t in_stride, u64 *__restrict__ lde, LdeMap map, u32 ncols, u32 s,
             u32 lines_log, u32 rate_bits, const u64 *__restrict__ W, const u64 *__restrict__ scale) {
  extern __shared__ u64 sm[];
  const u32 S = 1u << s, c0 = blockIdx.x << lines_log;
  const u32 k = map.peer ? brev_bits((blockIdx.y + map.coset_rot) & ((1u << rate_bits) - 1), rate_bits) : blockIdx.y;
  const u64 *sc = scale + ((size_t)k << s);  // (7*w_N^k)^j
  const u32 total = S << lines_log, nthr = blockDim.x;
  for (u32 e0 = threadIdx.x; e0 < total; e0 += MP2_NTT_BATCH * nthr) {
    u64 v[MP2_NTT_BATCH], f[MP2_NTT_BATCH];
#pragma unroll
    for (int b = 0; b < MP2_NTT_BATCH; b++) {
      u32 e = e0 + b * nthr, p = e & (S - 1), l = e >> s, c = c0 + l;
      const bool ok = e < total && c < ncols;
      v[b] = ok ? coeffs[(size_t)c * in_stride + p] : 0;
      f[b] = ok ? __ldg(sc + p) : 0;
    }
#pragma unroll
    for (int b = 0; b < MP2_NTT_BATCH; b++) {
      u32 e = e0 + b * nthr, p = e & (S - 1), l = e >> s;
      if (e < total) sm[sidx((p << lines_log) + l)] = gl_mul(v[b], f[b]);
    }
  }
  __syncthreads();
  smem_ntt(sm, s, lines_log, W);
  const size_t block_base = (size_t)brev_bits(k, rate_bits) << s;
  if (lines_log == 0 && ((block_base ^ (block_base + S - 1)) >> map.ls_log) == 0) {
    // one column per CTA whose coset block sits in one shard block: base pointer + 32-bit offset
    u64 *__restrict__ dst = map.ptr(lde, block_base, c0);
    for (u32 p = threadIdx.x; p < S; p += blockDim.x) dst[p] = gl_canon(sm[sidx(p)]);
    return;
  }
  for (u32 e = threadIdx.x; e < (S << lines_log); e += blockDim.x) {
    u32 p = e & (S - 1), l = e >> s, c = c0 + l;
    if (c < ncols) *map.ptr(lde, block_base + p, c) = gl_canon(sm[sidx((p << lines_log) + l)]);
  }
}

// ---- two passes (four-step): n = n1*n2, j = j1*n2 + j2, k = k1 + n1*k2 ---------------------------
struct TwoPass {
  u32 n_log, a, b;   // n1 = 2^a (strided pass 1), n2 = 2^b (contiguous pass 2)
  u32 lines_log;
  u32 rate_bits;
  u32 ncols, tg_log; // pass 1 only: the grid is flattened, see k_pass1 (tg_log = log2 of tiles per group)
  u32 coset0;        // pass 2 only: first coset of this launch (grid.z counts from it)
  int inverse;       // 1: iNTT natural -> natural; 0: coset LDE natural -> leaf order
  u64 n_inv;
};

// pass 1: tile = LINES adjacent j2; size-n1 transform over j1 (stride n2); then the four-step
// twiddle w_n^(j2*k1).  1-D grid of ncols * cosets * (n2 / LINES) CTAs ordered (fast -> slow) 4 adjacent tiles,
// coset, column, tile group:
//   * the 2^r cosets of an input tile are 4 CTAs apart, so the tile comes from HBM once and is re-read from L2
//     (with the coset as the slowest grid dimension pass 1 read the input 8 times: 17.5 GB for 2.1 GB);
//   * 4 adjacent tiles (4 x 32-byte sectors = one 128-byte line of every row) run together, so no fetched
//     sector is left unused (ordering by column first lost that: 21.5 GB at 256 columns);
//   * the coset-scale slices of a tile group (1 MB) stay in L2 while the columns sweep over them.
// Staging note (both passes): a tile's global addresses are base + a 32-bit offset, the base computed once per
// CTA -- whenever the tile lies inside one shard block of the output map, which is every case except more shards
// than cosets.  The first version evaluated the 64-bit LdeMap per element: 63 of the 305 instructions per output
// element in pass 2 were this address arithmetic (profiles/r1f_lde_pass2_opcode_mix.txt).
GL_DEV bool within_shard(const LdeMap &m, size_t L0, size_t len) { return ((L0 ^ (L0 + len - 1)) >> m.ls_log) == 0; }

// MAXT / MINB: launch bounds.  The four-step tiles run 256 threads per CTA.  Measured on the wide batch
// (gpurun_out/r2n_ntt_variants.log): 4 CTAs/SM at 64 registers 45.6 ms per LDE, 3 CTAs at 80 registers 47.6 ms, 2 CTAs at
// 102-114 registers 54.1 ms -- the passes want resident warps, not registers, so the bound stays at 4.
#ifndef MP2_NTT_MINB
#define MP2_NTT_MINB 4
#endif
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
k_pass1(const u64 *__restrict__ in, size_t in_stride, u64 *__restrict__ out, size_t out_stride, LdeMap map,
        TwoPass tp, const u64 *__restrict__ W1, const u64 *__restrict__ Wn, const u64 *__restrict__ scale) {
  extern __shared__ u64 sm[];
  const u32 lines_log = tp.lines_log, LINES = 1u << lines_log, S = 1u << tp.a, b_log = tp.b;
  const u32 t4 = blockIdx.x & ((1u << tp.tg_log) - 1);
  const u32 k = (blockIdx.x >> tp.tg_log) & ((1u << tp.rate_bits) - 1);
  const u32 rest = blockIdx.x >> (tp.tg_log + tp.rate_bits), tg = rest / tp.ncols;
  const u32 c = rest - tg * tp.ncols;
  const u32 q0 = ((tg << tp.tg_log) + t4) << lines_log;
  const u32 total = S << lines_log, nthr = blockDim.x;
  // element e = p*LINES + l of the tile is input j = p*n2 + q0 + l
  const u64 *__restrict__ src = in + (size_t)c * in_stride + q0;
  const u64 *__restrict__ sc = tp.inverse ? nullptr : scale + ((size_t)k << tp.n_log) + q0;
  for (u32 e0 = threadIdx.x; e0 < total; e0 += MP2_NTT_BATCH * nthr) {
    u64 v[MP2_NTT_BATCH], f[MP2_NTT_BATCH];
#pragma unroll
    for (int b = 0; b < MP2_NTT_BATCH; b++) {
      const u32 e = e0 + b * nthr, off = ((e >> lines_log) << b_log) + (e & (LINES - 1));
      v[b] = e < total ? src[off] : 0;
      f[b] = (e < total && !tp.inverse) ? __ldg(sc + off) : 1;
    }
#pragma unroll
    for (int b = 0; b < MP2_NTT_BATCH; b++) {
      const u32 e = e0 + b * nthr;
      if (e < total) sm[sidx(e)] = tp.inverse ? v[b] : gl_mul(v[b], f[b]);
    }
  }
  __syncthreads();
  smem_ntt(sm, tp.a, lines_log, W1);
  const size_t block_base = (size_t)brev_bits(k, tp.rate_bits) << tp.n_log;
  // output: inverse -> row k1 (natural) of the scratch matrix; LDE -> row p = bitrev(k1) of leaf block bitrev_r(k)
  const bool fast = tp.inverse || map.ls_log >= tp.n_log;  // the whole coset block sits in one shard block
  u64 *__restrict__ dst = tp.inverse ? out + (size_t)c * out_stride + q0 : fast ? out + map(block_base + q0, c) : out;
  if (!fast) {  // more shards than cosets: per-element map
    for (u32 e = threadIdx.x; e < total; e += nthr) {
      const u32 l = e & (LINES - 1), p = e >> lines_log;
      out[map(block_base + ((size_t)p << b_log) + q0 + l, c)] = gl_mul(sm[sidx(e)], __ldg(Wn + (q0 + l) * brev_bits(p, tp.a)));
    }
    return;
  }
  for (u32 e0 = threadIdx.x; e0 < total; e0 += MP2_NTT_BATCH * nthr) {
    u64 w[MP2_NTT_BATCH];
    u32 row[MP2_NTT_BATCH];
#pragma unroll
    for (int b = 0; b < MP2_NTT_BATCH; b++) {
      const u32 e = e0 + b * nthr, l = e & (LINES - 1), p = e >> lines_log, k1 = brev_bits(p, tp.a);
      w[b] = e < total ? __ldg(Wn + (q0 + l) * k1) : 0;  // w_n^(j2*k1), j2*k1 < n <= 2^26
      row[b] = ((tp.inverse ? k1 : p) << b_log) + l;
    }
#pragma unroll
    for (int b = 0; b < MP2_NTT_BATCH; b++) {
      const u32 e = e0 + b * nthr;
      if (e < total) dst[row[b]] = gl_mul(sm[sidx(e)], w[b]);
    }
  }
}

// pass 2: tile = LINES adjacent rows; size-n2 transform along each (contiguous) row: the tile is `total`
// consecutive elements on both sides.  grid = (n1 / LINES, ncols, cosets)
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
k_pass2(const u64 *__restrict__ in, size_t in_stride, u64 *__restrict__ out, size_t out_stride, LdeMap map,
        LdeMap out_map, TwoPass tp, const u64 *__restrict__ W2) {
  extern __shared__ u64 sm[];
  const u32 lines_log = tp.lines_log, LINES = 1u << lines_log, S = 1u << tp.b, b_log = tp.b;
  const u32 n = 1u << tp.n_log;
  const u32 c = blockIdx.y, r0 = blockIdx.x << lines_log;
  const u32 k = out_map.peer ? brev_bits((blockIdx.z + out_map.coset_rot) & ((1u << tp.rate_bits) - 1), tp.rate_bits)
                             : blockIdx.z + tp.coset0;
  const size_t block_base = (size_t)brev_bits(k, tp.rate_bits) << tp.n_log;
  const size_t L0 = block_base + ((size_t)r0 << b_log);  // first element of the tile
  const u32 total = S << lines_log, nthr = blockDim.x;
  const bool in_fast = tp.inverse || within_shard(map, L0, total);
  if (in_fast) {  // (uniform per CTA) the tile is `total` consecutive elements
    const u64 *__restrict__ src = tp.inverse ? in + (size_t)c * in_stride + ((size_t)r0 << b_log) : in + map(L0, c);
    for (u32 e0 = threadIdx.x; e0 < total; e0 += MP2_NTT_BATCH * nthr) {
      u64 v[MP2_NTT_BATCH];
#pragma unroll
      for (int b = 0; b < MP2_NTT_BATCH; b++) v[b] = e0 + b * nthr < total ? src[e0 + b * nthr] : 0;
#pragma unroll
      for (int b = 0; b < MP2_NTT_BATCH; b++) {
        const u32 e = e0 + b * nthr;
        if (e < total) sm[sidx(((e & (S - 1)) << lines_log) + (e >> b_log))] = v[b];
      }
    }
  } else {
    for (u32 e = threadIdx.x; e < total; e += nthr) sm[sidx(((e & (S - 1)) << lines_log) + (e >> b_log))] = in[map(L0 + e, c)];
  }
  __syncthreads();
  smem_ntt(sm, tp.b, lines_log, W2);
  if (tp.inverse) {
    u64 *__restrict__ dst = out + (size_t)c * out_stride;
    for (u32 e = threadIdx.x; e < total; e += nthr) {
      const u32 l = e & (LINES - 1), p = e >> lines_log;
      const u32 kk = (r0 + l) + (brev_bits(p, tp.b) << tp.a);  // forward frequency k = k1 + n1*k2
      dst[(n - kk) & (n - 1)] = gl_canon(gl_mul(sm[sidx(e)], tp.n_inv));
    }
  } else {
    if (within_shard(out_map, L0, total)) {
      u64 *__restrict__ dst = out_map.ptr(out, L0, c);
      for (u32 e = threadIdx.x; e < total; e += nthr) dst[e] = gl_canon(sm[sidx(((e & (S - 1)) << lines_log) + (e >> b_log))]);
    } else {
      for (u32 e = threadIdx.x; e < total; e += nthr)
        *out_map.ptr(out, L0 + e, c) = gl_canon(sm[sidx(((e & (S - 1)) << lines_log) + (e >> b_log))]);
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
  u32 t = tile_log >= 3 ? 1u << (tile_log - 3) : 1;  // one radix-8 item per thread and pass...
  if (tile_log >= 12) t = 1u << (tile_log - MP2_NTT_THREADS_SHIFT);  // ...more on the big tiles (<= 1024 threads)
  if (t > 1024) t = 1024;
  if (t < 32) t = 32;
  return t;
}
// The dynamic shared-memory limit is an attribute of the FUNCTION (per device), shared by every host thread: it is
// only ever raised, under a lock.  (Setting it to each launch's own size let one prover thread lower it between
// another thread's set and launch: cudaErrorInvalidValue with several provers per process.)
template <typename K>
static Status allow_smem(K kernel, size_t bytes) {
  if (bytes <= 48 * 1024) return "";
  static std::mutex mu;
  static std::map<std::pair<int, const void *>, size_t> granted;
  int device = 0;
  MP2_CUDA(cudaGetDevice(&device));
  std::lock_guard<std::mutex> lock(mu);
  size_t &cur = granted[std::make_pair(device, (const void *)kernel)];
  if (bytes > cur) {
    MP2_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    cur = bytes;
  }
  return "";
}

template <typename... A>
static Status launch_pass1(dim3 grid, u32 threads, size_t smem, cudaStream_t st, A... args) {
  if (threads <= 256) {
    MP2_TRY(allow_smem(k_pass1<256, MP2_NTT_MINB>, smem));
    { ProfScope _p("k_pass1", st); k_pass1<256, MP2_NTT_MINB><<<grid, threads, smem, st>>>(args...); }
  } else {
    MP2_TRY(allow_smem(k_pass1<1024, 1>, smem));
    { ProfScope _p("k_pass1", st); k_pass1<1024, 1><<<grid, threads, smem, st>>>(args...); }
  }
  MP2_LAUNCH_CHECK();
  return "";
}
template <typename... A>
static Status launch_pass2(dim3 grid, u32 threads, size_t smem, cudaStream_t st, A... args) {
  if (threads <= 256) {
    MP2_TRY(allow_smem(k_pass2<256, MP2_NTT_MINB>, smem));
    { ProfScope _p("k_pass2", st); k_pass2<256, MP2_NTT_MINB><<<grid, threads, smem, st>>>(args...); }
  } else {
    MP2_TRY(allow_smem(k_pass2<1024, 1>, smem));
    { ProfScope _p("k_pass2", st); k_pass2<1024, 1><<<grid, threads, smem, st>>>(args...); }
  }
  MP2_LAUNCH_CHECK();
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
  if (n_log > 26) return "polynomial degree 2^" + std::to_string(n_log) + " not supported (max 2^26)";
  tp->n_log = n_log;
  tp->b = (n_log + 1) / 2;
  tp->a = n_log - tp->b;
  tp->lines_log = tp->b >= kTileLog ? 0 : kTileLog - tp->b;  // a <= b <= 13; one line per CTA once a line fills the tile
  return "";
}

Status ntt_intt(const u64 *values, size_t in_stride, u64 *coeffs, size_t out_stride, size_t ncols, u32 n_log,
                cudaStream_t st) {
  if (ncols == 0) return "";
  if (n_log > 32) return "n_log exceeds the field's two-adicity (32)";
  const u64 n_inv = h_inv((u64)1 << n_log);
  LdeMap none = {};
  if (n_log <= kMaxSingleLog) {
    const u64 *W;
    MP2_TRY(table_roots(n_log, st, &W));
    u32 lines_log = n_log >= kTileLog ? 0 : std::min(kTileLog - n_log, ceil_log2(ncols));
    u32 tile_log = n_log + lines_log;
    size_t smem = smem_bytes_for(tile_log);
    MP2_TRY(allow_smem(k_intt_single, smem));
    unsigned grid = (unsigned)((ncols + ((size_t)1 << lines_log) - 1) >> lines_log);
    { ProfScope _p("k_intt_single", st); k_intt_single<<<grid, threads_for(tile_log), smem, st>>>(values, in_stride, coeffs, out_stride, (u32)ncols, n_log, lines_log, W, n_inv); }
    MP2_LAUNCH_CHECK();
    return "";
  }
  TwoPass tp;
  MP2_TRY(split_two_pass(n_log, &tp));
  if (ncols > 65535) return "batch too wide for one iNTT launch (more than 65535 columns at this degree): split the columns";
  tp.rate_bits = 0;
  tp.inverse = 1;
  tp.n_inv = n_inv;
  const u64 *W1, *W2, *Wn;
  MP2_TRY(table_roots(tp.a, st, &W1));
  MP2_TRY(table_roots(tp.b, st, &W2));
  MP2_TRY(table_roots(n_log, st, &Wn));
  const size_t n = (size_t)1 << n_log;
  DevBuf scratch;
  MP2_TRY(scratch.alloc(n * ncols, st));
  u64 *tmp = scratch.p;
  {
    u32 tile_log = tp.a + tp.lines_log;
    size_t smem = smem_bytes_for(tile_log);
    const size_t ctas = (((size_t)1 << tp.b) >> tp.lines_log) * ncols;
    if (ctas > 0x7fffffffull) return "batch too large for one iNTT launch (columns x tiles > 2^31)";
    tp.ncols = (u32)ncols;
    tp.tg_log = std::min(2u, tp.b - tp.lines_log);
    dim3 grid((unsigned)ctas, 1, 1);
    MP2_TRY(launch_pass1(grid, threads_for(tile_log), smem, st, values, in_stride, tmp, n, none, tp, W1, Wn, nullptr));
  }
  {
    u32 lines_log = std::min(tp.lines_log, tp.a);
    TwoPass tp2 = tp;
    tp2.lines_log = lines_log;
    u32 tile_log = tp.b + lines_log;
    size_t smem = smem_bytes_for(tile_log);
    dim3 grid((unsigned)(((size_t)1 << tp.a) >> lines_log), (unsigned)ncols, 1);
    MP2_TRY(launch_pass2(grid, threads_for(tile_log), smem, st, tmp, n, coeffs, out_stride, none, none, tp2, W2));
  }
  return "";
}

bool ntt_lde_is_two_pass(u32 n_log) { return n_log > kMaxSingleLog; }

Status ntt_coset_lde(const u64 *coeffs, size_t in_stride, u64 *lde, size_t lde_stride, size_t ncols, u32 n_log,
                     u32 rate_bits, u32 shard_log, size_t shard_stride, cudaStream_t st, u64 *const *peer_bases,
                     u64 shift, int phase, u32 coset0, u32 ncosets, u32 first_shard) {
  if (ncols == 0) return "";
  if (phase != LDE_ALL && (peer_bases || !ntt_lde_is_two_pass(n_log))) return "split LDE phases need the local two-pass path";
  if (phase == LDE_PASS2 && (ncosets == 0 || coset0 + ncosets > (1u << rate_bits))) return "bad coset range";
  const u32 N_log = n_log + rate_bits;
  if (N_log > 32) return "n_log + rate_bits exceeds the field's two-adicity (32)";
  if (shard_log > N_log) return "shard_log larger than log2(number of leaves)";
  if (rate_bits > 15) return "rate_bits too large";
  const u64 *scale = nullptr;
  MP2_TRY(table_coset_scale(n_log, rate_bits, shift, st, &scale));
  LdeMap map = {};
  map.ls_log = N_log - shard_log;
  map.shard_stride = shard_log ? shard_stride : 0;
  map.col_stride = lde_stride;
  if (peer_bases) {
    if (shard_log > 4) return "peer exchange supports at most 16 ranks";
    if (first_shard >> shard_log) return "first_shard out of range";
    map.peer = 1;
    // leaf block j belongs to shard j >> (rate_bits - shard_log) (shard_log <= rate_bits is implied by G <= 2^cap
    // only for the usual shapes; with more shards than cosets every block spans several shards and no rotation helps)
    map.coset_rot = shard_log <= rate_bits ? first_shard << (rate_bits - shard_log) : 0;
    for (u32 g = 0; g < (1u << shard_log); g++) map.bases[g] = peer_bases[g];
  }
  const unsigned cosets = 1u << rate_bits;
  if (n_log <= kMaxSingleLog) {
    const u64 *W;
    MP2_TRY(table_roots(n_log, st, &W));
    u32 lines_log = n_log >= kTileLog ? 0 : std::min(kTileLog - n_log, ceil_log2(ncols));
    u32 tile_log = n_log + lines_log;
    size_t smem = smem_bytes_for(tile_log);
    MP2_TRY(allow_smem(k_lde_single, smem));
    dim3 grid((unsigned)((ncols + ((size_t)1 << lines_log) - 1) >> lines_log), cosets, 1);
    { ProfScope _p("k_lde_single", st); k_lde_single<<<grid, threads_for(tile_log), smem, st>>>(coeffs, in_stride, lde, map, (u32)ncols, n_log, lines_log, rate_bits, W, scale); }
    MP2_LAUNCH_CHECK();
    return "";
  }
  TwoPass tp;
  MP2_TRY(split_two_pass(n_log, &tp));
  tp.rate_bits = rate_bits;
  tp.inverse = 0;
  tp.n_inv = 1;
  const u64 *W1, *W2, *Wn;
  MP2_TRY(table_roots(tp.a, st, &W1));
  MP2_TRY(table_roots(tp.b, st, &W2));
  MP2_TRY(table_roots(n_log, st, &Wn));
  // the four-step intermediate stays local: in place in the output buffer, or (peer mode, where the
  // output lives in other ranks' HBM) in a scratch buffer of the same shape
  LdeMap mid_map = map;
  u64 *mid = lde;
  DevBuf scratch;
  if (map.peer) {
    const size_t N = (size_t)1 << N_log;
    mid_map = LdeMap{};
    mid_map.ls_log = N_log;
    mid_map.col_stride = N;
    if (lde) {
      mid = lde;  // caller-provided N * ncols scratch (a per-call multi-GB pool allocation is what stalled the first
                  // step after an idle period: profiles/bench/r2_peer_stall_2gpu.txt)
    } else {
      MP2_TRY(scratch.alloc(N * ncols, st));
      mid = scratch.p;
    }
  }
  tp.coset0 = 0;
  if (ncols > 65535) return "batch too wide for one LDE launch (more than 65535 columns at this degree): split the columns";
  if (phase != LDE_PASS2) {
    u32 tile_log = tp.a + tp.lines_log;
    size_t smem = smem_bytes_for(tile_log);
    const size_t ctas = ((((size_t)1 << tp.b) >> tp.lines_log) << rate_bits) * ncols;
    if (ctas > 0x7fffffffull) return "batch too large for one LDE launch (columns x cosets x tiles > 2^31)";
    tp.ncols = (u32)ncols;
    tp.tg_log = std::min(2u, tp.b - tp.lines_log);
    dim3 grid((unsigned)ctas, 1, 1);
    MP2_TRY(launch_pass1(grid, threads_for(tile_log), smem, st, coeffs, in_stride, mid, 0, mid_map, tp, W1, Wn, scale));
  }
  if (phase != LDE_PASS1) {
    u32 lines_log = std::min(tp.lines_log, tp.a);
    TwoPass tp2 = tp;
    tp2.lines_log = lines_log;
    tp2.coset0 = phase == LDE_PASS2 ? coset0 : 0;
    u32 tile_log = tp.b + lines_log;
    size_t smem = smem_bytes_for(tile_log);
    dim3 grid((unsigned)(((size_t)1 << tp.a) >> lines_log), (unsigned)ncols, phase == LDE_PASS2 ? ncosets : cosets);
    MP2_TRY(launch_pass2(grid, threads_for(tile_log), smem, st, mid, 0, lde, 0, mid_map, map, tp2, W2));
  }
  return "";
}

}  // namespace mp2
