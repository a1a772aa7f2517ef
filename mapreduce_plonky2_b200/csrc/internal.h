// Internal C++ interface between the translation units of libmp2gpu.so (not installed).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <string>

typedef unsigned long long u64;
typedef unsigned int u32;

namespace mp2 {

// "" = success; anything else is the message handed back across the C ABI.
typedef std::string Status;

extern std::atomic<uint64_t> g_launches;
inline void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// Optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline leg).
// Disabled by default: a ProfScope is then two predictable branches.
extern std::atomic<int> g_profile_on;
void prof_record(const char *name, cudaStream_t st, bool begin);
struct ProfScope {
  const char *name;
  cudaStream_t st;
  bool on;
  ProfScope(const char *n, cudaStream_t s) : name(n), st(s), on(g_profile_on.load(std::memory_order_relaxed) != 0) {
    if (on) prof_record(name, st, true);
  }
  ~ProfScope() {
    if (on) prof_record(name, st, false);
  }
};

#define MP2_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return std::string(#expr) + ": " + cudaGetErrorName(_e) + ": " + cudaGetErrorString(_e); \
  } while (0)
#define MP2_TRY(expr)            \
  do {                           \
    mp2::Status _s = (expr);     \
    if (!_s.empty()) return _s;  \
  } while (0)
#define MP2_LAUNCH_CHECK()     \
  do {                         \
    mp2::count_launch();       \
    MP2_CUDA(cudaGetLastError()); \
  } while (0)

// ---- per-thread device binding (api.cu) ----
// Binds the calling thread to the device chosen with mp2gpu_init (default 0) and returns that thread's private
// compute stream on it (optionally also its device->host and host->device copy streams).
Status ctx_stream(cudaStream_t *out, cudaStream_t *copy_out = nullptr, cudaStream_t *up_out = nullptr);
// the device the calling thread is bound to
int ctx_device();
// Runs the rest of the scope on the device a handle lives on, then gives the thread its own device back
// (both the library's binding and CUDA's current device).
struct DeviceScope {
  int saved;
  explicit DeviceScope(int device);
  ~DeviceScope();
  DeviceScope(const DeviceScope &) = delete;
  DeviceScope &operator=(const DeviceScope &) = delete;
};

// ---- host-buffer plumbing shared by api.cu and sharded.cu ----
// Column-wise copies merged over runs of columns that are adjacent in host memory
Status copy_columns_h2d(u64 *dev, const uint64_t *const *cols, size_t ncols, size_t n, cudaStream_t st);
Status copy_columns_d2h(uint64_t *const *cols, const u64 *dev, size_t ncols, size_t n, cudaStream_t st);
Status check_commit_args(size_t ncols, u32 n_log, u32 rate_bits, u32 cap_height, u32 hash_kind);

// Device blocks for scratch and handles.  Behind these two calls sits a small per-thread, per-device cache in front
// of the stream-ordered pool (api.cu): a block freed on the calling thread's own compute stream is kept (exact size)
// and handed to that thread's next request of the same size -- a prover thread repeats the same batch shapes, and
// going through cudaMallocAsync / cudaFreeAsync every time let the pool re-grow under several concurrent provers
// (100 ms-scale stalls, tools/prover_trace_var.py).  Blocks above 1 GiB, a cache above MP2GPU_THREAD_CACHE_MB
// (default 3072), foreign streams and foreign threads go straight to the pool.  mp2gpu_trim() empties it.
Status pool_alloc(u64 **p, size_t bytes, cudaStream_t st);
void pool_free(void *p, cudaStream_t st);

// Stream-ordered scratch, returned on every exit path
struct DevBuf {
  u64 *p = nullptr;
  cudaStream_t st = nullptr;
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  Status alloc(size_t elems, cudaStream_t s) {
    st = s;
    if (elems == 0) elems = 1;
    return pool_alloc(&p, elems * sizeof(u64), s);
  }
  u64 *release() {
    u64 *r = p;
    p = nullptr;
    return r;
  }
  ~DevBuf() {
    if (p) pool_free(p, st);
  }
};

// ---- host-side Goldilocks helpers (table seeds only; no data-path work happens on the host) ----
static const u64 kP = 0xFFFFFFFF00000001ULL;
inline u64 h_mul(u64 a, u64 b) { return (u64)(((unsigned __int128)a * b) % kP); }
inline u64 h_pow(u64 a, u64 e) {
  u64 r = 1;
  a %= kP;
  while (e) {
    if (e & 1) r = h_mul(r, a);
    a = h_mul(a, a);
    e >>= 1;
  }
  return r;
}
inline u64 h_inv(u64 a) { return h_pow(a, kP - 2); }
// plonky2 primitive_root_of_unity(k) = POWER_OF_TWO_GENERATOR^(2^(32-k)), generator 7^((p-1)/2^32)
inline u64 h_root_of_unity(u32 log_n) {
  u64 w = h_pow(7, (kP - 1) >> 32);
  for (u32 i = log_n; i < 32; i++) w = h_mul(w, w);
  return w;
}

// ---- device tables, cached per device (tables.cu) ----
// W[m] = w_T^m for m < T = 2^log_t (T >= 1)
Status table_roots(u32 log_t, cudaStream_t st, const u64 **out);
// scale[(k << log_n) + j] = (shift * w_N^k)^j for k < 2^rate_bits, j < 2^log_n, N = 2^(log_n + rate_bits):
// the coset pre-scaling of PolynomialCoeffs::coset_fft (shift = MULTIPLICATIVE_GROUP_GENERATOR = 7 for the
// polynomial batches; 7^(arity^i) for FRI layer i)
static const u64 kCosetShift = 7;
Status table_coset_scale(u32 log_n, u32 rate_bits, u64 shift, cudaStream_t st, const u64 **out);
void table_cache_clear();  // frees the current device's cached tables (mp2gpu_trim)

// ---- transforms (ntt.cu) ----
Status ntt_intt(const u64 *values, size_t in_stride, u64 *coeffs, size_t out_stride, size_t ncols,
                u32 n_log, cudaStream_t st);
// peer_bases (optional, host array of 2^shard_log device pointers): shard g is written to peer_bases[g]
// (column c at + c*lde_stride) instead of lde + g*shard_stride -- the exchange fused into the store.
// first_shard: the destination the launch stores to first (the others follow in rotated order); ranks pass
// their own index so that they do not all target the same peer at the same time.
// phase (two-pass sizes only, local output): LDE_PASS1 runs the strided pass of every coset and leaves the
// four-step intermediate in `lde`; LDE_PASS2 finishes cosets [coset0, coset0 + ncosets), i.e. the leaf blocks
// bitrev_r(k) -- the host entry point interleaves those with the hashing and the copy-out of each block.
enum { LDE_ALL = 0, LDE_PASS1 = 1, LDE_PASS2 = 2 };
bool ntt_lde_is_two_pass(u32 n_log);
Status ntt_coset_lde(const u64 *coeffs, size_t in_stride, u64 *lde, size_t lde_stride, size_t ncols,
                     u32 n_log, u32 rate_bits, u32 shard_log, size_t shard_stride, cudaStream_t st,
                     u64 *const *peer_bases = nullptr, u64 shift = kCosetShift, int phase = LDE_ALL,
                     u32 coset0 = 0, u32 ncosets = 0, u32 first_shard = 0);

// ---- FRI commit phase (fri.cu) ----
// coeffs' = chunks(2^arity_bits) reduced with powers of beta; ext polys are component-major (2 x len)
Status fri_fold(const u64 *coeffs, size_t in_stride, u64 *out, size_t out_stride, size_t out_len, u32 arity_bits,
                u64 beta0, u64 beta1, cudaStream_t st);
// out[2*pos + comp] = vals[comp*stride + pos]: leaf-ordered ext values -> flattened FRI leaves
Status fri_interleave(const u64 *vals, size_t stride, u64 *out, size_t len, cudaStream_t st);
// inverse: interleaved host layout -> component-major (canonicalised)
Status fri_deinterleave(const u64 *in, u64 *out, size_t stride, size_t len, cudaStream_t st);
Status ntt_canonicalize(const u64 *in, size_t in_stride, u64 *out, size_t out_stride, size_t ncols,
                        size_t n, cudaStream_t st);

// ---- hashing / Merkle (merkle.cu) ----
Status merkle_colmajor(const u64 *lde, size_t lde_stride, size_t ncols, size_t nleaves,
                       u32 cap_height, u32 hash_kind, u64 *leaves_out, u64 *digests, u64 *cap,
                       cudaStream_t st);
// the two halves of merkle_colmajor, for callers that pipeline leaf ranges against copies
Status merkle_colmajor_leaves(const u64 *lde, size_t lde_stride, size_t ncols, size_t nleaves, u32 cap_height,
                              u32 hash_kind, size_t leaf_begin, size_t leaf_end, u64 *leaves_out, u64 *digests,
                              u64 *cap, cudaStream_t st);
Status merkle_levels(size_t nleaves, u32 cap_height, u32 hash_kind, u64 *digests, u64 *cap, cudaStream_t st);
Status merkle_levels_subtrees(size_t nleaves, u32 cap_height, u32 hash_kind, u64 *digests, u64 *cap, size_t sub0,
                              size_t nsub, cudaStream_t st);
Status merkle_rowmajor(const u64 *leaves, size_t nleaves, size_t leaf_len, u32 cap_height,
                       u32 hash_kind, u64 *digests, u64 *cap, cudaStream_t st);
// flat: concatenated leaves, offsets: nleaves+1 prefix sums (device)
Status merkle_ragged(const u64 *flat, const u64 *offsets, size_t nleaves, u32 cap_height,
                     u32 hash_kind, u64 *digests, u64 *cap, cudaStream_t st);
Status hash_no_pad_batch(const u64 *inputs, size_t count, size_t input_len, u32 hash_kind, u64 *out,
                         cudaStream_t st);
Status two_to_one_batch(const u64 *a, const u64 *b, size_t count, u32 hash_kind, u64 *out,
                        cudaStream_t st);
Status permute_batch(u64 *states, size_t count, u32 hash_kind, cudaStream_t st);
// smallest c in [start, start + count) with clz(canon(permute(state | state[pos] = c)[7])) >= min_lz, else ~0
Status pow_search(const u64 *state12_host, u32 pos, u32 min_lz, u32 hash_kind, u64 start, u64 count, u64 *found,
                  cudaStream_t st);
Status gather_rows(const u64 *leaves_rowmajor, const u64 *lde_colmajor, size_t lde_stride,
                   size_t ncols, const u64 *row_idx, size_t nrows, u64 *out, cudaStream_t st);

// Query-time openings (MerkleTree::get + MerkleTree::prove for a list of leaf indices): rows and sibling
// digests are gathered on the device and copied back in two transfers.  idx / outputs are HOST pointers.
Status merkle_open(const u64 *leaves_rowmajor, const u64 *lde_colmajor, size_t lde_stride, size_t leaf_len,
                   const u64 *digests, size_t nleaves, u32 cap_height, const u64 *idx_host, size_t count,
                   u64 *rows_out_host, u64 *siblings_out_host, cudaStream_t st);

// ---- prove_openings (fri.cu) ----
// out[m] (+ out[stride + m]) = sum_j pw[j] * polys[j][m]: ReducingFactor::reduce_polys_base.  polys / pw are
// device arrays (count pointers; powers component-major, the two halves pw_stride apart)
Status fri_reduce_polys_strided(const u64 *const *polys, const u64 *pw, size_t pw_stride, u32 count, size_t n, u64 *out,
                                size_t out_stride, cudaStream_t st);
// out[pt * ncols + c] (pairs, device) = polynomial c evaluated at the extension point points_host[pt]
Status fri_eval_polys(const u64 *coeffs, size_t stride, size_t ncols, size_t n, const u64 *points_host, size_t npoints,
                      u64 *out, cudaStream_t st);
// acc[k] = acc[k] * scale + sum_{m > k} x[m] * z^(m-k-1)   (divide_by_linear + push(0), then shift_poly / +=);
// `fresh` skips the read of acc.  x, acc: component-major extension polynomials of length len.
Status fri_divide_accumulate(const u64 *x, size_t x_stride, size_t len, const u64 z[2], u64 *acc, size_t acc_stride,
                             const u64 scale[2], bool fresh, cudaStream_t st);

inline int log2_exact(size_t n) {
  if (n == 0 || (n & (n - 1))) return -1;
  int l = 0;
  while (((size_t)1 << l) < n) l++;
  return l;
}

}  // namespace mp2

struct mp2gpu_batch;
namespace mp2 {
// PolynomialBatch::from_values / from_coeffs with host buffers on the calling thread's device (api.cu)
Status commit_host(const uint64_t *const *cols, size_t ncols, u32 n_log, u32 rate_bits, u32 cap_height, u32 hash_kind,
                   int from_coeffs, uint64_t *const *coeffs_out, uint64_t *leaves_out, uint64_t *digests_out,
                   uint64_t *cap_out, mp2gpu_batch **handle_out);
}  // namespace mp2

// Device-resident PolynomialBatch behind the C ABI's opaque handle (api.cu creates it, fri.cu reads coeffs).
struct mp2gpu_batch {
  int device;
  size_t ncols;
  u32 n_log, rate_bits, cap_height, hash_kind;
  u64 *coeffs, *lde, *leaves, *digests, *cap;  // device; coeffs: ncols x n, lde: ncols x N, column-major
  cudaStream_t owner_stream;  // the stream the buffers were allocated on (stream-ordered pool): they are freed on it
};
