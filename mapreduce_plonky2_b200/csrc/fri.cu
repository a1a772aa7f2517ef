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

}  // namespace mp2

using namespace mp2;

// Device-resident state of one fri_committed_trees loop.
struct mp2gpu_fri {
  int device;
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
Status use_device(const mp2gpu_fri *f, cudaStream_t *st) {
  if (!f) return "null fri handle";
  MP2_CUDA(cudaSetDevice(f->device));
  *st = cudaStreamPerThread;
  return "";
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
}  // namespace

extern "C" {

const char *mp2gpu_fri_begin(const uint64_t *coeffs_ext, uint32_t n_log, uint32_t rate_bits, uint32_t cap_height,
                             uint32_t hash_kind, mp2gpu_fri **out) {
  return guard([&]() -> Status {
    if (!coeffs_ext || !out) return "null coeffs / out";
    if (hash_kind > 1) return "unknown hash_kind " + std::to_string(hash_kind);
    if (n_log + rate_bits > 32) return "degree_log + rate_bits exceeds two-adicity 32";
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return std::string("no usable CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(e);
    cudaStream_t st = cudaStreamPerThread;
    const size_t n = (size_t)1 << n_log;
    mp2gpu_fri *f = new mp2gpu_fri();
    f->device = dev;
    f->n_log = n_log;
    f->rate_bits = rate_bits;
    f->cap_height = cap_height;
    f->hash_kind = hash_kind;
    f->shift = kCosetShift;
    f->last_arity_bits = 0;
    f->committed = false;
    f->coeffs = nullptr;
    u64 *tmp = nullptr;
    MP2_CUDA(cudaMalloc(&f->coeffs, sizeof(u64) * 2 * n));
    MP2_CUDA(cudaMallocAsync(&tmp, sizeof(u64) * 2 * n, st));
    MP2_CUDA(cudaMemcpyAsync(tmp, coeffs_ext, sizeof(u64) * 2 * n, cudaMemcpyHostToDevice, st));
    MP2_TRY(fri_deinterleave(tmp, f->coeffs, n, n, st));
    MP2_CUDA(cudaFreeAsync(tmp, st));
    MP2_CUDA(cudaStreamSynchronize(st));
    *out = f;
    return "";
  });
}

const char *mp2gpu_fri_commit_layer(mp2gpu_fri *f, uint32_t arity_bits, uint64_t *cap_out) {
  return guard([&]() -> Status {
    cudaStream_t st;
    MP2_TRY(use_device(f, &st));
    if (!cap_out) return "null cap_out";
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
    u64 *vals = nullptr;
    MP2_CUDA(cudaMallocAsync(&vals, sizeof(u64) * 2 * N, st));
    // values on the coset shift*<w_N>, leaf (= bit-reversed) order, both components
    MP2_TRY(ntt_coset_lde(f->coeffs, n, vals, N, 2, f->n_log, f->rate_bits, 0, 0, st, nullptr, f->shift));
    MP2_CUDA(cudaMalloc(&L.leaves, sizeof(u64) * 2 * N));
    MP2_CUDA(cudaMalloc(&L.digests, sizeof(u64) * 4 * (L.ndigests ? L.ndigests : 1)));
    MP2_CUDA(cudaMalloc(&L.cap, sizeof(u64) * 4 * L.ncap));
    MP2_TRY(fri_interleave(vals, N, L.leaves, N, st));
    MP2_CUDA(cudaFreeAsync(vals, st));
    MP2_TRY(merkle_rowmajor(L.leaves, L.nleaves, L.leaf_len, cap_h, f->hash_kind, L.digests, L.cap, st));
    MP2_CUDA(cudaMemcpyAsync(cap_out, L.cap, sizeof(u64) * 4 * L.ncap, cudaMemcpyDeviceToHost, st));
    MP2_CUDA(cudaStreamSynchronize(st));
    f->layers.push_back(L);
    f->last_arity_bits = arity_bits;
    f->committed = true;
    return "";
  });
}

const char *mp2gpu_fri_fold(mp2gpu_fri *f, const uint64_t beta[2]) {
  return guard([&]() -> Status {
    cudaStream_t st;
    MP2_TRY(use_device(f, &st));
    if (!beta) return "null beta";
    if (!f->committed) return "fri_fold without a committed layer";
    const u32 ab = f->last_arity_bits;
    const size_t n = (size_t)1 << f->n_log, n_out = n >> ab;
    u64 *next = nullptr;
    MP2_CUDA(cudaMalloc(&next, sizeof(u64) * 2 * n_out));
    MP2_TRY(fri_fold(f->coeffs, n, next, n_out, n_out, ab, beta[0] % kP, beta[1] % kP, st));
    MP2_CUDA(cudaStreamSynchronize(st));
    MP2_CUDA(cudaFree(f->coeffs));
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
    cudaStream_t st;
    MP2_TRY(use_device(f, &st));
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
    cudaStream_t st;
    MP2_TRY(use_device(f, &st));
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
    cudaStream_t st;
    MP2_TRY(use_device(f, &st));
    const size_t n = (size_t)1 << f->n_log;
    if (len_out) *len_out = n;
    if (final_coeffs_out) {
      u64 *tmp = nullptr;
      MP2_CUDA(cudaMallocAsync(&tmp, sizeof(u64) * 2 * n, st));
      MP2_TRY(fri_interleave(f->coeffs, n, tmp, n, st));
      MP2_CUDA(cudaMemcpyAsync(final_coeffs_out, tmp, sizeof(u64) * 2 * n, cudaMemcpyDeviceToHost, st));
      MP2_CUDA(cudaFreeAsync(tmp, st));
      MP2_CUDA(cudaStreamSynchronize(st));
    }
    return "";
  });
}

void mp2gpu_fri_free(mp2gpu_fri *f) {
  if (!f) return;
  cudaSetDevice(f->device);
  if (f->coeffs) cudaFree(f->coeffs);
  for (auto &L : f->layers) {
    cudaFree(L.leaves);
    cudaFree(L.digests);
    cudaFree(L.cap);
  }
  delete f;
}

}  // extern "C"
