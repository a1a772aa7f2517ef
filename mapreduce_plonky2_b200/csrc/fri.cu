// FRI commit phase on the device: per reduction layer, the Merkle tree of the chunked extension-field
// values, the fold with the challenger's beta, and the re-evaluation on the next coset.
//
// Replaces the loop body of plonky2::fri::prover::fri_committed_trees (plonky2 0.2.2; SURVEY.md 8(a) a10
// and 8(f) rank 2), which every prove() of the reference runs after its three batch commitments
// (recursion-framework/src/circuit_builder.rs:308).  D = 2 (mp2-common/src/lib.rs:36): the extension is
// GF(p^2) = F[X]/(X^2 - 7).  Its 2^k-th roots of unity for k <= 32 are the base field's, so the coset FFT of
// an extension polynomial is the base-field transform applied to both components -- the LDE kernels of
// ntt.cu are reused as they are, with the layer's shift 7^(arity^i).
//
// Extension polynomials live component-major on the device (2 x len); the C ABI converts from / to the
// interleaved [a0, a1] pairs that Vec<QuadraticExtension<GoldilocksField>> is in memory.
#include <cstdlib>
#include <cstring>
#include <exception>
#include <memory>
#include <vector>

#include "../../include/mp2gpu.h"
#include "gl.cuh"
#include "internal.h"

namespace mp2 {

struct Ext {
  u64 a, b;
};
// (a0 + a1 X)(b0 + b1 X) with X^2 = 7
GL_DEV Ext ext_mul(Ext x, Ext y) {
  u64 t = gl_mul(x.b, y.b);
  u64 t7 = gl_add(gl_add(gl_add(t, t), gl_add(t, t)), gl_add(gl_add(t, t), t));  // 7t
  Ext r;
  r.a = gl_add(gl_mul(x.a, y.a), t7);
  r.b = gl_add(gl_mul(x.a, y.b), gl_mul(x.b, y.a));
  return r;
}

// reduce_with_powers over chunks of 2^arity_bits coefficients: Horner from the last term
__global__ void k_fri_fold(const u64 *__restrict__ in, size_t in_stride, u64 *__restrict__ out, size_t out_stride,
                           size_t out_len, u32 arity_bits, u64 beta0, u64 beta1) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= out_len) return;
  const Ext beta = {beta0, beta1};
  Ext acc = {0, 0};
  const size_t base = j << arity_bits;
  for (int t = (1 << arity_bits) - 1; t >= 0; t--) {
    acc = ext_mul(acc, beta);
    acc.a = gl_add(acc.a, in[base + t]);
    acc.b = gl_add(acc.b, in[in_stride + base + t]);
  }
  out[j] = gl_canon(acc.a);
  out[out_stride + j] = gl_canon(acc.b);
}

__global__ void k_fri_interleave(const u64 *__restrict__ vals, size_t stride, u64 *__restrict__ out, size_t len) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < 2 * len) out[e] = vals[(e & 1) * stride + (e >> 1)];
}
__global__ void k_fri_deinterleave(const u64 *__restrict__ in, u64 *__restrict__ out, size_t stride, size_t len) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < 2 * len) out[(e & 1) * stride + (e >> 1)] = gl_canon(in[e]);
}

Status fri_fold(const u64 *coeffs, size_t in_stride, u64 *out, size_t out_stride, size_t out_len, u32 arity_bits,
                u64 beta0, u64 beta1, cudaStream_t st) {
  if (!out_len) return "";
  { ProfScope _p("k_fri_fold", st); k_fri_fold<<<(unsigned)((out_len + 127) / 128), 128, 0, st>>>(coeffs, in_stride, out, out_stride, out_len, arity_bits, beta0, beta1); }
  MP2_LAUNCH_CHECK();
  return "";
}
Status fri_interleave(const u64 *vals, size_t stride, u64 *out, size_t len, cudaStream_t st) {
  if (!len) return "";
  { ProfScope _p("k_fri_interleave", st); k_fri_interleave<<<(unsigned)((2 * len + 255) / 256), 256, 0, st>>>(vals, stride, out, len); }
  MP2_LAUNCH_CHECK();
  return "";
}
Status fri_deinterleave(const u64 *in, u64 *out, size_t stride, size_t len, cudaStream_t st) {
  if (!len) return "";
  { ProfScope _p("k_fri_deinterleave", st); k_fri_deinterleave<<<(unsigned)((2 * len + 255) / 256), 256, 0, st>>>(in, out, stride, len); }
  MP2_LAUNCH_CHECK();
  return "";
}

// ---- prove_openings: the alpha-batched quotient that becomes FRI's input polynomial -------------------------
// Replaces the loop of PolynomialBatch::prove_openings (plonky2 0.2.2 fri/oracle.rs) and the ReducingFactor
// calls it makes (util/reducing.rs: reduce_polys_base, shift_poly), reached from every prove() of the
// reference right after the three batch commitments (recursion-framework/src/circuit_builder.rs:308):
//   final_poly = sum_i alpha^(k_i) (F_i(X) - F_i(z_i)) / (X - z_i),   F_i = sum_j alpha^j f_ij.

// F[m] = sum_j alpha^j f_j[m]: one thread per coefficient, the batch's columns streamed once (coalesced)
__global__ void k_fri_reduce_polys(const u64 *const *__restrict__ polys, const u64 *__restrict__ pw, size_t pw_stride,
                                   u32 count, size_t n, u64 *__restrict__ out, size_t out_stride) {
  size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  u64 a = 0, b = 0;
  for (u32 j = 0; j < count; j++) {
    const u64 v = polys[j][m];
    a = gl_mul_add(v, pw[j], a);
    b = gl_mul_add(v, pw[pw_stride + j], b);
  }
  out[m] = a;
  out[out_stride + m] = b;
}

// divide_by_linear is the recurrence b_k = b_(k+1) z + a_k run from the top coefficient down, quotient
// coefficient k = b_(k+1): an exclusive weighted suffix sum  E[k] = sum_{m>k} x[m] w^(m-k-1)  with w = z.
// One CTA covers kScanT * kScanL elements: Horner over kScanL per thread, a Hillis-Steele suffix scan over the
// thread partials in shared memory (the weights w^(L 2^s) come from the host), Horner again to emit.
// TOTALS = true writes only the CTA's inclusive total; the totals are scanned by the same kernel one level up
// (weight w^(T L)) and come back as `carry`.
constexpr int kScanT = 256, kScanL = 4, kScanSteps = 8;
struct ScanPows {
  u64 a[kScanSteps], b[kScanSteps];
};
template <bool TOTALS>
__global__ void __launch_bounds__(kScanT)
k_suffix_scan(const u64 *__restrict__ x, size_t xs, size_t len, Ext w, Ext wL, ScanPows pw,
              const u64 *__restrict__ carry, size_t cs, u64 *out, size_t os, Ext scale, int fresh) {
  __shared__ u64 va[kScanT], vb[kScanT];
  const int t = threadIdx.x;
  const size_t start = ((size_t)blockIdx.x * kScanT + t) * kScanL;
  Ext xv[kScanL];
#pragma unroll
  for (int i = 0; i < kScanL; i++) {
    const size_t idx = start + i;
    xv[i].a = idx < len ? x[idx] : 0;
    xv[i].b = idx < len ? x[xs + idx] : 0;
  }
  Ext s = {0, 0};
#pragma unroll
  for (int i = kScanL - 1; i >= 0; i--) {
    s = ext_mul(s, w);
    s.a = gl_add(s.a, xv[i].a);
    s.b = gl_add(s.b, xv[i].b);
  }
  Ext k_in = {0, 0};  // suffix sum of everything after this CTA, relative to the CTA's end
  if (!TOTALS && carry) {
    k_in.a = carry[blockIdx.x];
    k_in.b = carry[cs + blockIdx.x];
    if (t == kScanT - 1) {
      Ext p = ext_mul(k_in, wL);
      s.a = gl_add(s.a, p.a);
      s.b = gl_add(s.b, p.b);
    }
  }
  va[t] = s.a;
  vb[t] = s.b;
  __syncthreads();
  for (int sft = 0; sft < kScanSteps; sft++) {
    const int d = 1 << sft;
    const bool has = t + d < kScanT;
    Ext add = {0, 0};
    if (has) add = ext_mul(Ext{va[t + d], vb[t + d]}, Ext{pw.a[sft], pw.b[sft]});
    __syncthreads();
    if (has) {
      va[t] = gl_add(va[t], add.a);
      vb[t] = gl_add(vb[t], add.b);
    }
    __syncthreads();
  }
  if (TOTALS) {
    if (t == 0) {
      out[blockIdx.x] = va[0];
      out[os + blockIdx.x] = vb[0];
    }
    return;
  }
  Ext b = t + 1 < kScanT ? Ext{va[t + 1], vb[t + 1]} : k_in;
#pragma unroll
  for (int i = kScanL - 1; i >= 0; i--) {
    const size_t idx = start + i;
    if (idx < len) {
      Ext e = b;
      if (!fresh) {  // shift_poly then +=
        Ext o = ext_mul(Ext{out[idx], out[os + idx]}, scale);
        e.a = gl_add(e.a, o.a);
        e.b = gl_add(e.b, o.b);
      }
      out[idx] = gl_canon(e.a);
      out[os + idx] = gl_canon(e.b);
    }
    b = ext_mul(b, w);
    b.a = gl_add(b.a, xv[i].a);
    b.b = gl_add(b.b, xv[i].b);
  }
}

// ---- OpeningSet: every polynomial of a batch evaluated at extension points -------------------------------
// Replaces `c.polynomials.par_iter().map(|p| p.to_extension().eval(z))` of plonky2's OpeningSet::new
// (plonk/proof.rs; step 7 of prove(), between the quotient commitment and prove_openings).
// grid = (ncols, npoints); thread t owns the coefficients m = t (mod T): coalesced loads, Horner in
// w = z^T, then z^t from a host-made table and a shared-memory sum.  tab: per point 2 x (T + 1) entries,
// component-major: z^0 .. z^(T-1), z^T.
constexpr int kEvalT = 256;
__global__ void __launch_bounds__(kEvalT)
k_eval_polys(const u64 *__restrict__ coeffs, size_t stride, size_t n, const u64 *__restrict__ tab,
             u64 *__restrict__ out, size_t ncols) {
  __shared__ u64 ra[kEvalT], rb[kEvalT];
  const int t = threadIdx.x;
  const size_t c = blockIdx.x, pt = blockIdx.y;
  const u64 *tb = tab + pt * 2 * (kEvalT + 1);
  const Ext w = {tb[kEvalT], tb[kEvalT + 1 + kEvalT]};
  const u64 *f = coeffs + c * stride;
  Ext acc = {0, 0};
  if ((size_t)t < n) {
    const size_t top = (n - 1 - t) / kEvalT;  // largest i with t + T i < n
    for (size_t i = top + 1; i-- > 0;) {
      acc = ext_mul(acc, w);
      acc.a = gl_add(acc.a, f[t + kEvalT * i]);
    }
    acc = ext_mul(acc, Ext{tb[t], tb[kEvalT + 1 + t]});
  }
  ra[t] = acc.a;
  rb[t] = acc.b;
  __syncthreads();
  for (int d = kEvalT / 2; d > 0; d >>= 1) {
    if (t < d) {
      ra[t] = gl_add(ra[t], ra[t + d]);
      rb[t] = gl_add(rb[t], rb[t + d]);
    }
    __syncthreads();
  }
  if (t == 0) {
    out[2 * (pt * ncols + c)] = gl_canon(ra[0]);
    out[2 * (pt * ncols + c) + 1] = gl_canon(rb[0]);
  }
}

// host-side extension arithmetic (a handful of powers per call; no data-path work)
struct HExt {
  u64 a, b;
};
static u64 h_add(u64 a, u64 b) { return (u64)(((unsigned __int128)a + b) % kP); }
static HExt hx_mul(HExt x, HExt y) {
  return {h_add(h_mul(x.a, y.a), h_mul(7, h_mul(x.b, y.b))), h_add(h_mul(x.a, y.b), h_mul(x.b, y.a))};
}
static HExt hx_pow(HExt x, u64 e) {
  HExt r = {1, 0};
  while (e) {
    if (e & 1) r = hx_mul(r, x);
    x = hx_mul(x, x);
    e >>= 1;
  }
  return r;
}

static Status suffix_scan(const u64 *x, size_t xs, size_t len, HExt w, u64 *out, size_t os, HExt scale, bool fresh,
                          cudaStream_t st) {
  const size_t per = (size_t)kScanT * kScanL, nb = (len + per - 1) / per;
  if (nb > 0x7fffffffull) return "polynomial too long for the quotient scan";
  const HExt wL = hx_pow(w, kScanL);
  ScanPows pw;
  HExt p = wL;
  for (int s = 0; s < kScanSteps; s++) {
    pw.a[s] = p.a;
    pw.b[s] = p.b;
    p = hx_mul(p, p);
  }  // p = w^(T L)
  const Ext dw = {w.a, w.b}, dwL = {wL.a, wL.b}, dscale = {scale.a, scale.b};
  if (nb <= 1) {
    { ProfScope _p("k_suffix_scan", st); k_suffix_scan<false><<<1, kScanT, 0, st>>>(x, xs, len, dw, dwL, pw, nullptr, 0, out, os, dscale, fresh ? 1 : 0); }
    MP2_LAUNCH_CHECK();
    return "";
  }
  DevBuf tot, car;
  MP2_TRY(tot.alloc(2 * nb, st));
  MP2_TRY(car.alloc(2 * nb, st));
  { ProfScope _p("k_suffix_scan", st); k_suffix_scan<true><<<(unsigned)nb, kScanT, 0, st>>>(x, xs, len, dw, dwL, pw, nullptr, 0, tot.p, nb, dscale, 1); }
  MP2_LAUNCH_CHECK();
  MP2_TRY(suffix_scan(tot.p, nb, nb, p, car.p, nb, HExt{0, 0}, true, st));
  { ProfScope _p("k_suffix_scan", st); k_suffix_scan<false><<<(unsigned)nb, kScanT, 0, st>>>(x, xs, len, dw, dwL, pw, car.p, nb, out, os, dscale, fresh ? 1 : 0); }
  MP2_LAUNCH_CHECK();
  return "";
}

// out[pt][c] = polynomial c evaluated at points[pt]; out is a DEVICE buffer of npoints x ncols pairs
Status fri_eval_polys(const u64 *coeffs, size_t stride, size_t ncols, size_t n, const u64 *points_host, size_t npoints,
                      u64 *out, cudaStream_t st) {
  if (!ncols || !npoints || !n) return "";
  if (npoints > 65535) return "too many evaluation points";
  std::vector<u64> tab(npoints * 2 * (kEvalT + 1));
  for (size_t p = 0; p < npoints; p++) {
    const HExt z = {points_host[2 * p] % kP, points_host[2 * p + 1] % kP};
    u64 *ta = tab.data() + p * 2 * (kEvalT + 1), *tb = ta + kEvalT + 1;
    HExt cur = {1, 0};
    for (int t = 0; t <= kEvalT; t++) {
      ta[t] = cur.a;
      tb[t] = cur.b;
      cur = hx_mul(cur, z);
    }
  }
  DevBuf d_tab;
  MP2_TRY(d_tab.alloc(tab.size(), st));
  MP2_CUDA(cudaMemcpyAsync(d_tab.p, tab.data(), sizeof(u64) * tab.size(), cudaMemcpyHostToDevice, st));
  // the table is pageable host memory: the copy above has been staged when the call returns
  if (ncols > 0x7fffffffull) return "too many polynomials";
  if (npoints > 65535) return "too many evaluation points for one launch (more than 65535)";
  dim3 grid((unsigned)ncols, (unsigned)npoints, 1);
  { ProfScope _p("k_eval_polys", st); k_eval_polys<<<grid, kEvalT, 0, st>>>(coeffs, stride, n, d_tab.p, out, ncols); }
  MP2_LAUNCH_CHECK();
  return "";
}
Status fri_reduce_polys_strided(const u64 *const *polys, const u64 *pw, size_t pw_stride, u32 count, size_t n, u64 *out,
                                size_t out_stride, cudaStream_t st) {
  if (!n) return "";
  { ProfScope _p("k_fri_reduce_polys", st); k_fri_reduce_polys<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(polys, pw, pw_stride, count, n, out, out_stride); }
  MP2_LAUNCH_CHECK();
  return "";
}
Status fri_divide_accumulate(const u64 *x, size_t x_stride, size_t len, const u64 z[2], u64 *acc, size_t acc_stride,
                             const u64 scale[2], bool fresh, cudaStream_t st) {
  if (!len) return "";
  return suffix_scan(x, x_stride, len, HExt{z[0] % kP, z[1] % kP}, acc, acc_stride, HExt{scale[0] % kP, scale[1] % kP},
                     fresh, st);
}

}  // namespace mp2

using namespace mp2;

// Device-resident state of one fri_committed_trees loop.
struct mp2gpu_fri {
  int device;
  cudaStream_t owner_stream;  // stream the stream-ordered buffers below were allocated on (and are freed on)
  u32 n_log;       // log2 of the current (non-padded) coefficient count
  u32 rate_bits, cap_height, hash_kind;
  u64 shift;       // coset shift of the current layer: 7^(prod of arities so far)
  u32 last_arity_bits;
  bool committed;  // commit_layer was called for the current polynomial
  u64 *coeffs;     // 2 x n, component-major
  struct Layer {
    size_t nleaves, leaf_len, ndigests, ncap;
    u64 *leaves, *digests, *cap;
  };
  std::vector<Layer> layers;
};

namespace {
const char *dup_c(const Status &s) {
  if (s.empty()) return nullptr;
  char *p = (char *)malloc(s.size() + 1);
  if (p) memcpy(p, s.c_str(), s.size() + 1);
  return p;
}
// Every entry point runs inside a DeviceScope for the handle's device (the caller's device binding -- the
// library's and CUDA's -- is restored on return) and on the calling thread's private stream for that device,
// exactly like the batch entry points of api.cu.
#define FRI_ON_DEVICE(f, st)                  \
  if (!(f)) return "null fri handle";         \
  DeviceScope _scope((f)->device);            \
  cudaStream_t st;                            \
  MP2_TRY(ctx_stream(&st))
// stream-ordered allocation that the handle keeps (released by mp2gpu_fri_free / the next fold)
Status fri_alloc(u64 **p, size_t elems, cudaStream_t st) {
  return pool_alloc(p, sizeof(u64) * (elems ? elems : 1), st);
}
template <typename F>
const char *guard(F f) {
  try {
    return dup_c(f());
  } catch (const std::exception &e) {
    return dup_c(std::string("exception: ") + e.what());
  } catch (...) {
    return dup_c("unknown exception");
  }
}
struct FriDeleter {
  void operator()(mp2gpu_fri *f) const { mp2gpu_fri_free(f); }
};
typedef std::unique_ptr<mp2gpu_fri, FriDeleter> FriPtr;
// must be called inside a DeviceScope for `device`; st = the calling thread's stream on it
Status new_fri(u32 n_log, u32 rate_bits, u32 cap_height, u32 hash_kind, int device, FriPtr *out, cudaStream_t st) {
  if (hash_kind > 1) return "unknown hash_kind " + std::to_string(hash_kind);
  if (n_log + rate_bits > 32) return "degree_log + rate_bits exceeds two-adicity 32";
  FriPtr f(new mp2gpu_fri());
  f->device = device;
  f->owner_stream = st;
  f->n_log = n_log;
  f->rate_bits = rate_bits;
  f->cap_height = cap_height;
  f->hash_kind = hash_kind;
  f->shift = kCosetShift;
  f->last_arity_bits = 0;
  f->committed = false;
  f->coeffs = nullptr;
  MP2_TRY(fri_alloc(&f->coeffs, 2 * ((size_t)1 << n_log), st));
  *out = std::move(f);
  return "";
}
}  // namespace

extern "C" {

const char *mp2gpu_fri_begin(const uint64_t *coeffs_ext, uint32_t n_log, uint32_t rate_bits, uint32_t cap_height,
                             uint32_t hash_kind, mp2gpu_fri **out) {
  return guard([&]() -> Status {
    if (!coeffs_ext || !out) return "null coeffs / out";
    FriPtr f;
    DeviceScope scope(ctx_device());  // the device this thread chose with mp2gpu_init
    cudaStream_t st;
    MP2_TRY(ctx_stream(&st));
    MP2_TRY(new_fri(n_log, rate_bits, cap_height, hash_kind, ctx_device(), &f, st));
    const size_t n = (size_t)1 << n_log;
    DevBuf tmp;
    MP2_TRY(tmp.alloc(2 * n, st));
    MP2_CUDA(cudaMemcpyAsync(tmp.p, coeffs_ext, sizeof(u64) * 2 * n, cudaMemcpyHostToDevice, st));
    MP2_TRY(fri_deinterleave(tmp.p, f->coeffs, n, n, st));
    MP2_CUDA(cudaStreamSynchronize(st));
    *out = f.release();
    return "";
  });
}

const char *mp2gpu_batch_eval(const mp2gpu_batch *b, const uint64_t *points, size_t npoints, uint64_t *out) {
  return guard([&]() -> Status {
    if (!b) return "null batch handle";
    if (npoints && (!points || !out)) return "null points / out";
    if (!npoints) return "";
    DeviceScope scope(b->device);
    cudaStream_t st;
    MP2_TRY(ctx_stream(&st));
    DevBuf d_out;
    MP2_TRY(d_out.alloc(2 * npoints * b->ncols, st));
    MP2_TRY(fri_eval_polys(b->coeffs, (size_t)1 << b->n_log, b->ncols, (size_t)1 << b->n_log, (const u64 *)points, npoints,
                           d_out.p, st));
    MP2_CUDA(cudaMemcpyAsync(out, d_out.p, sizeof(u64) * 2 * npoints * b->ncols, cudaMemcpyDeviceToHost, st));
    MP2_CUDA(cudaStreamSynchronize(st));
    return "";
  });
}

const char *mp2gpu_fri_begin_openings(const mp2gpu_batch *const *oracles, size_t noracles, const uint64_t *points,
                                      const uint32_t *batch_sizes, size_t nbatches, const uint32_t *oracle_index,
                                      const uint32_t *polynomial_index, const uint64_t alpha[2], uint32_t cap_height,
                                      uint32_t hash_kind, uint64_t *final_poly_out, mp2gpu_fri **out) {
  return guard([&]() -> Status {
    if (!oracles || !noracles || !points || !batch_sizes || !nbatches || !oracle_index || !polynomial_index || !alpha || !out)
      return "prove_openings: null / empty argument";
    for (size_t o = 0; o < noracles; o++) {
      if (!oracles[o]) return "prove_openings: null oracle handle";
      if (oracles[o]->n_log != oracles[0]->n_log || oracles[o]->rate_bits != oracles[0]->rate_bits ||
          oracles[o]->device != oracles[0]->device)
        return "prove_openings: the oracles must share degree, rate_bits and device";
    }
    const u32 n_log = oracles[0]->n_log;
    const size_t n = (size_t)1 << n_log;
    size_t total = 0, max_count = 0;
    for (size_t i = 0; i < nbatches; i++) {
      if (!batch_sizes[i]) return "prove_openings: empty batch";
      total += batch_sizes[i];
      if (batch_sizes[i] > max_count) max_count = batch_sizes[i];
    }
    std::vector<const u64 *> ptrs(total);
    for (size_t j = 0; j < total; j++) {
      if (oracle_index[j] >= noracles) return "prove_openings: oracle_index " + std::to_string(oracle_index[j]) + " out of range";
      const mp2gpu_batch *b = oracles[oracle_index[j]];
      if (polynomial_index[j] >= b->ncols)
        return "prove_openings: polynomial_index " + std::to_string(polynomial_index[j]) + " out of range (oracle has " +
               std::to_string(b->ncols) + " polynomials)";
      ptrs[j] = b->coeffs + (size_t)polynomial_index[j] * n;
    }
    // alpha^j, j < the largest batch: base.powers() restarts for every batch
    const HExt al = {alpha[0] % kP, alpha[1] % kP};
    std::vector<u64> pw(2 * max_count);
    HExt cur = {1, 0};
    for (size_t j = 0; j < max_count; j++) {
      pw[j] = cur.a;
      pw[max_count + j] = cur.b;
      cur = hx_mul(cur, al);
    }
    FriPtr f;
    DeviceScope scope(oracles[0]->device);
    cudaStream_t st;
    MP2_TRY(ctx_stream(&st));
    MP2_TRY(new_fri(n_log, oracles[0]->rate_bits, cap_height, hash_kind, oracles[0]->device, &f, st));
    DevBuf d_ptrs, d_pw, comp;
    MP2_TRY(d_ptrs.alloc(total, st));
    MP2_TRY(d_pw.alloc(2 * max_count, st));
    MP2_TRY(comp.alloc(2 * n, st));
    MP2_CUDA(cudaMemcpyAsync(d_ptrs.p, ptrs.data(), sizeof(u64 *) * total, cudaMemcpyHostToDevice, st));
    MP2_CUDA(cudaMemcpyAsync(d_pw.p, pw.data(), sizeof(u64) * 2 * max_count, cudaMemcpyHostToDevice, st));
    size_t at = 0;
    for (size_t i = 0; i < nbatches; i++) {
      const u32 count = batch_sizes[i];
      // the powers table is component-major over max_count entries; a batch of `count` reads a prefix of each half
      MP2_TRY(fri_reduce_polys_strided((const u64 *const *)d_ptrs.p + at, d_pw.p, max_count, count, n, comp.p, n, st));
      const HExt shift = hx_pow(al, count);  // ReducingFactor::shift_poly: alpha^count of THIS batch
      const u64 sc[2] = {shift.a, shift.b};
      MP2_TRY(fri_divide_accumulate(comp.p, n, n, (const u64 *)points + 2 * i, f->coeffs, n, sc, i == 0, st));
      at += count;
    }
    if (final_poly_out) {
      DevBuf tmp;
      MP2_TRY(tmp.alloc(2 * n, st));
      MP2_TRY(fri_interleave(f->coeffs, n, tmp.p, n, st));
      MP2_CUDA(cudaMemcpyAsync(final_poly_out, tmp.p, sizeof(u64) * 2 * n, cudaMemcpyDeviceToHost, st));
    }
    MP2_CUDA(cudaStreamSynchronize(st));
    *out = f.release();
    return "";
  });
}

const char *mp2gpu_fri_commit_layer(mp2gpu_fri *f, uint32_t arity_bits, uint64_t *cap_out) {
  return guard([&]() -> Status {
    FRI_ON_DEVICE(f, st);
    if (!cap_out) return "null cap_out";
    if (f->committed) return "fri_commit_layer called twice without a fold in between";
    const u32 N_log = f->n_log + f->rate_bits;
    if (arity_bits == 0 || arity_bits > N_log) return "bad arity_bits";
    if (f->n_log < arity_bits) return "polynomial shorter than the arity";
    const size_t n = (size_t)1 << f->n_log, N = (size_t)1 << N_log;
    mp2gpu_fri::Layer L;
    L.nleaves = N >> arity_bits;
    L.leaf_len = (size_t)2 << arity_bits;
    const u32 leaves_log = N_log - arity_bits;
    const u32 cap_h = f->cap_height;
    if (cap_h > leaves_log)
      return "MerkleTree::new: cap_height=" + std::to_string(cap_h) + " should be at most log2(leaves.len())=" + std::to_string(leaves_log);
    L.ncap = (size_t)1 << cap_h;
    L.ndigests = 2 * (L.nleaves - L.ncap);
    L.leaves = L.digests = L.cap = nullptr;
    Status built = [&]() -> Status {
      DevBuf vals;
      MP2_TRY(vals.alloc(2 * N, st));
      // values on the coset shift*<w_N>, leaf (= bit-reversed) order, both components
      MP2_TRY(ntt_coset_lde(f->coeffs, n, vals.p, N, 2, f->n_log, f->rate_bits, 0, 0, st, nullptr, f->shift));
      MP2_TRY(fri_alloc(&L.leaves, 2 * N, st));
      MP2_TRY(fri_alloc(&L.digests, 4 * L.ndigests, st));
      MP2_TRY(fri_alloc(&L.cap, 4 * L.ncap, st));
      MP2_TRY(fri_interleave(vals.p, N, L.leaves, N, st));
      MP2_TRY(merkle_rowmajor(L.leaves, L.nleaves, L.leaf_len, cap_h, f->hash_kind, L.digests, L.cap, st));
      MP2_CUDA(cudaMemcpyAsync(cap_out, L.cap, sizeof(u64) * 4 * L.ncap, cudaMemcpyDeviceToHost, st));
      MP2_CUDA(cudaStreamSynchronize(st));
      return "";
    }();
    if (!built.empty()) {
      for (u64 *p : {L.leaves, L.digests, L.cap})
        if (p) pool_free(p, st);
      return built;
    }
    f->layers.push_back(L);
    f->last_arity_bits = arity_bits;
    f->committed = true;
    return "";
  });
}

const char *mp2gpu_fri_fold(mp2gpu_fri *f, const uint64_t beta[2]) {
  return guard([&]() -> Status {
    FRI_ON_DEVICE(f, st);
    if (!beta) return "null beta";
    if (!f->committed) return "fri_fold without a committed layer";
    const u32 ab = f->last_arity_bits;
    const size_t n = (size_t)1 << f->n_log, n_out = n >> ab;
    u64 *next = nullptr;
    MP2_TRY(fri_alloc(&next, 2 * n_out, st));
    Status folded = fri_fold(f->coeffs, n, next, n_out, n_out, ab, beta[0] % kP, beta[1] % kP, st);
    if (!folded.empty()) {
      pool_free(next, st);
      return folded;
    }
    pool_free(f->coeffs, st);  // stream order: after the fold that read it
    MP2_CUDA(cudaStreamSynchronize(st));      // the handle may be used from another thread (another stream) next
    f->coeffs = next;
    f->n_log -= ab;
    f->shift = h_pow(f->shift, (u64)1 << ab);
    f->committed = false;
    return "";
  });
}

const char *mp2gpu_fri_fetch_layer(const mp2gpu_fri *f, uint32_t layer, uint64_t *leaves_out, uint64_t *digests_out,
                                   uint64_t *cap_out) {
  return guard([&]() -> Status {
    FRI_ON_DEVICE(f, st);
    if (layer >= f->layers.size()) return "no such FRI layer";
    const mp2gpu_fri::Layer &L = f->layers[layer];
    if (leaves_out) MP2_CUDA(cudaMemcpyAsync(leaves_out, L.leaves, sizeof(u64) * L.nleaves * L.leaf_len, cudaMemcpyDeviceToHost, st));
    if (digests_out && L.ndigests) MP2_CUDA(cudaMemcpyAsync(digests_out, L.digests, sizeof(u64) * 4 * L.ndigests, cudaMemcpyDeviceToHost, st));
    if (cap_out) MP2_CUDA(cudaMemcpyAsync(cap_out, L.cap, sizeof(u64) * 4 * L.ncap, cudaMemcpyDeviceToHost, st));
    MP2_CUDA(cudaStreamSynchronize(st));
    return "";
  });
}

const char *mp2gpu_fri_open_layer(const mp2gpu_fri *f, uint32_t layer, const uint64_t *leaf_idx, size_t count,
                                  uint64_t *leaves_out, uint64_t *siblings_out) {
  return guard([&]() -> Status {
    FRI_ON_DEVICE(f, st);
    if (layer >= f->layers.size()) return "no such FRI layer";
    if (count && !leaf_idx) return "null leaf_idx";
    const mp2gpu_fri::Layer &L = f->layers[layer];
    u32 cap_h = 0;
    while (((size_t)1 << cap_h) < L.ncap) cap_h++;
    return merkle_open(L.leaves, nullptr, 0, L.leaf_len, L.digests, L.nleaves, cap_h, (const u64 *)leaf_idx, count,
                       (u64 *)leaves_out, (u64 *)siblings_out, st);
  });
}

const char *mp2gpu_fri_layer_shape(const mp2gpu_fri *f, uint32_t layer, size_t *nleaves, size_t *leaf_len,
                                   size_t *ndigests, size_t *ncap) {
  return guard([&]() -> Status {
    if (!f) return "null fri handle";
    if (layer >= f->layers.size()) return "no such FRI layer";
    const mp2gpu_fri::Layer &L = f->layers[layer];
    if (nleaves) *nleaves = L.nleaves;
    if (leaf_len) *leaf_len = L.leaf_len;
    if (ndigests) *ndigests = L.ndigests;
    if (ncap) *ncap = L.ncap;
    return "";
  });
}

const char *mp2gpu_fri_finish(mp2gpu_fri *f, uint64_t *final_coeffs_out, size_t *len_out) {
  return guard([&]() -> Status {
    FRI_ON_DEVICE(f, st);
    const size_t n = (size_t)1 << f->n_log;
    if (len_out) *len_out = n;
    if (final_coeffs_out) {
      DevBuf tmp;
      MP2_TRY(tmp.alloc(2 * n, st));
      MP2_TRY(fri_interleave(f->coeffs, n, tmp.p, n, st));
      MP2_CUDA(cudaMemcpyAsync(final_coeffs_out, tmp.p, sizeof(u64) * 2 * n, cudaMemcpyDeviceToHost, st));
      MP2_CUDA(cudaStreamSynchronize(st));
    }
    return "";
  });
}

void mp2gpu_fri_free(mp2gpu_fri *f) {
  if (!f) return;
  // every entry point that touches the handle synchronises or orders its work on the owner's stream; the
  // buffers go back to the stream-ordered pool on the stream they came from (see mp2gpu_batch_free)
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(f->device);
  if (f->coeffs) pool_free(f->coeffs, f->owner_stream);
  for (auto &L : f->layers)
    for (u64 *p : {L.leaves, L.digests, L.cap})
      if (p) pool_free(p, f->owner_stream);
  cudaSetDevice(prev);
  delete f;
}

}  // extern "C"
