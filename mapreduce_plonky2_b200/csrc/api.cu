// extern "C" surface of libmp2gpu.so -- see include/mp2gpu.h for the contract of every symbol and
// the reference interface it replaces.  Error convention copied from the reference's only FFI
// (gnark-utils/src/lib.rs:17-52, gnark-utils/src/utils.rs:9-20): NULL = ok, else a malloc'ed string.
#include <cstdlib>
#include <cstring>
#include <exception>
#include <map>
#include <unordered_map>
#include <vector>

#include "../../include/mp2gpu.h"
#include "internal.h"

namespace mp2 {
std::atomic<uint64_t> g_launches{0};

// Per calling thread: the device chosen with mp2gpu_init and one (compute, copy) stream pair per device the
// thread has touched -- a prover thread that also reads a batch living on another device keeps both pairs.
struct DevStreams {
  cudaStream_t stream = nullptr;       // compute
  cudaStream_t copy_stream = nullptr;  // device->host copies overlapped with compute
  cudaStream_t up_stream = nullptr;    // host->device uploads overlapped with compute
  // block cache of this thread on this device (see pool_alloc in internal.h)
  std::unordered_map<void *, size_t> live;   // blocks this thread allocated on `stream`: address -> bytes
  std::multimap<size_t, void *> idle;        // freed on `stream`, ready for the next request of the same size
  size_t idle_bytes = 0;
};
struct ThreadCtx {
  int device = 0;
  std::vector<DevStreams> per_device;
  ~ThreadCtx() {  // thread exit: hand the idle blocks back (errors are ignored: the context may be gone already)
    for (size_t d = 0; d < per_device.size(); d++)
      for (auto &kv : per_device[d].idle) {
        if (cudaSetDevice((int)d) == cudaSuccess) cudaFreeAsync(kv.second, per_device[d].stream);
      }
  }
};
static thread_local ThreadCtx t_ctx;

// Binds the calling thread to its device and returns its private stream (and, optionally, its copy stream).
Status ctx_stream(cudaStream_t *out, cudaStream_t *copy_out, cudaStream_t *up_out) {
  ThreadCtx &c = t_ctx;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return std::string("no usable CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(e);
  if (c.device < 0 || c.device >= ndev) return "device " + std::to_string(c.device) + " out of range";
  MP2_CUDA(cudaSetDevice(c.device));
  if (c.per_device.size() < (size_t)ndev) c.per_device.resize(ndev);
  DevStreams &d = c.per_device[c.device];
  if (!d.stream) {
    cudaStream_t s1 = nullptr, s2 = nullptr, s3 = nullptr;
    MP2_CUDA(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    if (cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&s3, cudaStreamNonBlocking) != cudaSuccess) {
      cudaStreamDestroy(s1);
      if (s2) cudaStreamDestroy(s2);
      return "cudaStreamCreateWithFlags failed for the copy streams";
    }
    d.stream = s1;
    d.copy_stream = s2;
    d.up_stream = s3;
    // keep freed blocks in the stream-ordered pool: commitments reuse the same sizes over and over
    cudaMemPool_t pool;
    MP2_CUDA(cudaDeviceGetDefaultMemPool(&pool, c.device));
    uint64_t keep = UINT64_MAX;
    MP2_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  *out = d.stream;
  if (copy_out) *copy_out = d.copy_stream;
  if (up_out) *up_out = d.up_stream;
  return "";
}

namespace {
DevStreams *own_streams(cudaStream_t st) {  // the calling thread's stream set if `st` is its compute stream on the current device
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || (size_t)dev >= t_ctx.per_device.size()) return nullptr;
  DevStreams &d = t_ctx.per_device[dev];
  return (d.stream && d.stream == st) ? &d : nullptr;
}
size_t cache_limit_bytes() {
  static const size_t v = [] {
    const char *e = getenv("MP2GPU_THREAD_CACHE_MB");
    return (size_t)(e && *e ? atoll(e) : 3072) << 20;
  }();
  return v;
}
const size_t kMaxCachedBlock = (size_t)1 << 30;
}  // namespace

Status pool_alloc(u64 **p, size_t bytes, cudaStream_t st) {
  if (bytes == 0) bytes = sizeof(u64);
  DevStreams *d = own_streams(st);
  if (d) {
    auto it = d->idle.find(bytes);
    if (it != d->idle.end()) {
      *p = (u64 *)it->second;
      d->idle.erase(it);
      d->idle_bytes -= bytes;
      d->live[*p] = bytes;
      return "";
    }
  }
  MP2_CUDA(cudaMallocAsync((void **)p, bytes, st));
  if (d && bytes <= kMaxCachedBlock) d->live[*p] = bytes;
  return "";
}
void pool_free(void *p, cudaStream_t st) {
  if (!p) return;
  DevStreams *d = own_streams(st);
  if (d) {
    auto it = d->live.find(p);
    if (it != d->live.end()) {
      const size_t bytes = it->second;
      d->live.erase(it);
      if (bytes <= kMaxCachedBlock && d->idle_bytes + bytes <= cache_limit_bytes()) {
        d->idle.emplace(bytes, p);  // same thread, same stream: the next user is ordered after every pending use
        d->idle_bytes += bytes;
        return;
      }
    }
  }
  cudaFreeAsync(p, st);
}

// Runs the rest of the scope on the device a handle lives on, then gives the thread its own device back.
DeviceScope::DeviceScope(int device) : saved(t_ctx.device) { t_ctx.device = device; }
DeviceScope::~DeviceScope() {
  t_ctx.device = saved;
  cudaSetDevice(saved);
}
int ctx_device() { return t_ctx.device; }


namespace {
Status pick_stream(void *user, cudaStream_t *out) {
  cudaStream_t own;
  MP2_TRY(ctx_stream(&own));  // also validates the device
  *out = user == MP2GPU_STREAM_THREAD ? own : (cudaStream_t)user;  // NULL = CUDA's legacy default stream
  return "";
}

const char *to_c(const Status &s) {
  if (s.empty()) return nullptr;
  char *p = (char *)malloc(s.size() + 1);
  if (p) memcpy(p, s.c_str(), s.size() + 1);
  return p;
}

}  // namespace

// Column-wise host<->device copies, merged over runs of columns that are adjacent in host memory (one
// transfer for a contiguous matrix; one per column for separately allocated Vecs -- each cudaMemcpyAsync
// costs a few microseconds of launch time, which at 135 columns was a fifth of a config-1 call).
Status copy_columns_h2d(u64 *dev, const uint64_t *const *cols, size_t ncols, size_t n, cudaStream_t st) {
  for (size_t c = 0; c < ncols;) {
    if (!cols[c]) return "null column pointer";
    size_t run = 1;
    while (c + run < ncols && cols[c + run] == cols[c] + run * n) run++;
    MP2_CUDA(cudaMemcpyAsync(dev + c * n, cols[c], run * n * sizeof(u64), cudaMemcpyHostToDevice, st));
    c += run;
  }
  return "";
}
Status copy_columns_d2h(uint64_t *const *cols, const u64 *dev, size_t ncols, size_t n, cudaStream_t st) {
  for (size_t c = 0; c < ncols;) {
    if (!cols[c]) {
      c++;
      continue;
    }
    size_t run = 1;
    while (c + run < ncols && cols[c + run] == cols[c] + run * n) run++;
    MP2_CUDA(cudaMemcpyAsync(cols[c], dev + c * n, run * n * sizeof(u64), cudaMemcpyDeviceToHost, st));
    c += run;
  }
  return "";
}

Status check_commit_args(size_t ncols, u32 n_log, u32 rate_bits, u32 cap_height, u32 hash_kind) {
  if (ncols == 0) return "PolynomialBatch: no polynomials";
  if (n_log + rate_bits > 32) return "PolynomialBatch: degree_log + rate_bits exceeds two-adicity 32";
  if (cap_height > n_log + rate_bits)
    return "MerkleTree::new: cap_height=" + std::to_string(cap_height) + " should be at most log2(leaves.len())=" +
           std::to_string(n_log + rate_bits);
  if (hash_kind > 1) return "unknown hash_kind " + std::to_string(hash_kind);
  return "";
}

namespace {
Status dev_commit(const u64 *cols, size_t ncols, u32 n_log, u32 rate_bits, u32 cap_height, u32 hash_kind,
                  int from_coeffs, u64 *coeffs, u64 *lde, u64 *leaves, u64 *digests, u64 *cap, cudaStream_t st) {
  MP2_TRY(check_commit_args(ncols, n_log, rate_bits, cap_height, hash_kind));
  const size_t n = (size_t)1 << n_log, N = n << rate_bits;
  if (from_coeffs) MP2_TRY(ntt_canonicalize(cols, n, coeffs, n, ncols, n, st));
  else MP2_TRY(ntt_intt(cols, n, coeffs, n, ncols, n_log, st));
  MP2_TRY(ntt_coset_lde(coeffs, n, lde, N, ncols, n_log, rate_bits, 0, 0, st));
  MP2_TRY(merkle_colmajor(lde, N, ncols, N, cap_height, hash_kind, leaves, digests, cap, st));
  return "";
}

}  // namespace
}  // namespace mp2

namespace mp2 {
Status commit_host(const uint64_t *const *cols, size_t ncols, u32 n_log, u32 rate_bits, u32 cap_height,
                   u32 hash_kind, int from_coeffs, uint64_t *const *coeffs_out, uint64_t *leaves_out,
                   uint64_t *digests_out, uint64_t *cap_out, mp2gpu_batch **handle_out) {
  MP2_TRY(check_commit_args(ncols, n_log, rate_bits, cap_height, hash_kind));
  if (!cols || !cap_out) return "null cols / cap_out";
  cudaStream_t st, cp, up;
  MP2_TRY(ctx_stream(&st, &cp, &up));
  const size_t n = (size_t)1 << n_log, N = n << rate_bits, ncap = (size_t)1 << cap_height;
  const size_t ndig = 2 * (N - ncap);
  const bool want_rows = leaves_out != nullptr || handle_out != nullptr;
  DevBuf d_in, d_coeffs, d_lde, d_leaves, d_dig, d_cap;
  // declared after the buffers, so it runs before they are returned to the pool: on an early error return the
  // copy stream may still be reading them
  struct CopyDrain {
    cudaStream_t a, b;
    ~CopyDrain() {
      cudaStreamSynchronize(a);
      cudaStreamSynchronize(b);
    }
  } copy_drain{cp, up};
  MP2_TRY(d_in.alloc(ncols * n, st));
  MP2_TRY(d_coeffs.alloc(ncols * n, st));
  MP2_TRY(d_lde.alloc(ncols * N, st));
  if (want_rows) MP2_TRY(d_leaves.alloc(ncols * N, st));
  MP2_TRY(d_dig.alloc(ndig * 4, st));
  MP2_TRY(d_cap.alloc(ncap * 4, st));
  // The copy stream trails the compute stream: coefficients go back while the LDE runs, each block of
  // leaf rows goes back while the next block is hashed.  PCIe, not HBM, bounds this entry point
  // (SURVEY.md section 7), so hiding the copies behind the hashing is worth ~2x end to end.
  cudaEvent_t ev;
  MP2_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  struct EvGuard {
    cudaEvent_t e;
    ~EvGuard() { cudaEventDestroy(e); }
  } ev_guard{ev};
  // The transforms are per column, so big batches go through in column blocks: the upload of block k+1 overlaps
  // the iNTT + LDE of block k, and block k's coefficients travel back while block k+1 is transformed.
  static const size_t kBlocks = [] {
    const char *e = getenv("MP2_COMMIT_BLOCKS");
    return (size_t)(e && *e ? atoi(e) : 16);
  }();
  const size_t nblocks = (ncols >= 2 * kBlocks && ncols * n >= ((size_t)1 << 24)) ? kBlocks : 1;
  // Leaf rows go back to the host in blocks while the next block is hashed.  When the LDE is the two-pass kind and
  // a block is one coset (leaf block j = coset bitrev_r(j)), the contiguous pass of that coset is run just before
  // its block is hashed, so the first rows start travelling one pass earlier.
  const size_t cosets = (size_t)1 << rate_bits;
  const size_t nchunks = !(leaves_out && N >= ((size_t)1 << 16)) ? 1 : (cosets >= 4 && cosets <= 16) ? cosets : 8;
  const bool split_lde = nchunks == cosets && nchunks > 1 && ntt_lde_is_two_pass(n_log);
  if (nblocks > 1) {  // the buffers were allocated in st's order: the upload stream may touch them only after that
    MP2_CUDA(cudaEventRecord(ev, st));
    MP2_CUDA(cudaStreamWaitEvent(up, ev, 0));
  }
  for (size_t k = 0; k < nblocks; k++) {
    const size_t c0 = k * ncols / nblocks, c1 = (k + 1) * ncols / nblocks, cnt = c1 - c0;
    if (!cnt) continue;
    u64 *in_k = d_in.p + c0 * n, *co_k = d_coeffs.p + c0 * n;
    MP2_TRY(copy_columns_h2d(in_k, cols + c0, cnt, n, nblocks > 1 ? up : st));
    if (nblocks > 1) {
      MP2_CUDA(cudaEventRecord(ev, up));
      MP2_CUDA(cudaStreamWaitEvent(st, ev, 0));
    }
    if (from_coeffs) MP2_TRY(ntt_canonicalize(in_k, n, co_k, n, cnt, n, st));
    else MP2_TRY(ntt_intt(in_k, n, co_k, n, cnt, n_log, st));
    if (coeffs_out) {
      MP2_CUDA(cudaEventRecord(ev, st));
      MP2_CUDA(cudaStreamWaitEvent(cp, ev, 0));
      MP2_TRY(copy_columns_d2h(coeffs_out + c0, co_k, cnt, n, cp));
    }
    MP2_TRY(ntt_coset_lde(co_k, n, d_lde.p + c0 * N, N, cnt, n_log, rate_bits, 0, 0, st, nullptr, kCosetShift,
                          split_lde ? LDE_PASS1 : LDE_ALL));
  }
  for (size_t j = 0; j < nchunks; j++) {
    const size_t lb = j * (N / nchunks), le = (j + 1) * (N / nchunks);
    if (split_lde) {
      u32 k = 0;  // bitrev_r(j)
      for (u32 bit = 0; bit < rate_bits; bit++) k |= ((j >> bit) & 1) << (rate_bits - 1 - bit);
      MP2_TRY(ntt_coset_lde(d_coeffs.p, n, d_lde.p, N, ncols, n_log, rate_bits, 0, 0, st, nullptr, kCosetShift, LDE_PASS2, k, 1));
    }
    MP2_TRY(merkle_colmajor_leaves(d_lde.p, N, ncols, N, cap_height, hash_kind, lb, le, d_leaves.p, d_dig.p, d_cap.p, st));
    if (leaves_out) {
      MP2_CUDA(cudaEventRecord(ev, st));
      MP2_CUDA(cudaStreamWaitEvent(cp, ev, 0));
      MP2_CUDA(cudaMemcpyAsync(leaves_out + lb * ncols, d_leaves.p + lb * ncols, (le - lb) * ncols * sizeof(u64),
                               cudaMemcpyDeviceToHost, cp));
    }
  }
  // Levels and the digest copy: on big trees the cap subtrees are finished in 4 groups, and a group's digest chunk
  // (contiguous in plonky2's layout) travels to the host while the next group's levels are built -- otherwise the
  // 0.5 GB of digests of the wide batch is 10 ms of exposed PCIe time at the end of the call.
  const size_t groups = (digests_out && ndig && ncap >= 4 && N >= ((size_t)1 << 20)) ? 4 : 1;
  if (groups == 1) {
    MP2_TRY(merkle_levels(N, cap_height, hash_kind, d_dig.p, d_cap.p, st));
    if (digests_out && ndig)
      MP2_CUDA(cudaMemcpyAsync(digests_out, d_dig.p, ndig * 4 * sizeof(u64), cudaMemcpyDeviceToHost, st));
  } else {
    const size_t nsub = ncap / groups, chunk = ndig / groups * 4;  // u64 per group
    for (size_t gq = 0; gq < groups; gq++) {
      MP2_TRY(merkle_levels_subtrees(N, cap_height, hash_kind, d_dig.p, d_cap.p, gq * nsub, nsub, st));
      MP2_CUDA(cudaEventRecord(ev, st));
      MP2_CUDA(cudaStreamWaitEvent(cp, ev, 0));
      MP2_CUDA(cudaMemcpyAsync(digests_out + gq * chunk, d_dig.p + gq * chunk, chunk * sizeof(u64), cudaMemcpyDeviceToHost, cp));
    }
  }
  MP2_CUDA(cudaMemcpyAsync(cap_out, d_cap.p, ncap * 4 * sizeof(u64), cudaMemcpyDeviceToHost, st));
  MP2_CUDA(cudaStreamSynchronize(st));
  MP2_CUDA(cudaStreamSynchronize(cp));
  if (handle_out) {
    int device = 0;
    MP2_CUDA(cudaGetDevice(&device));
    mp2gpu_batch *b = new mp2gpu_batch();
    b->device = device;
    b->ncols = ncols;
    b->n_log = n_log;
    b->rate_bits = rate_bits;
    b->cap_height = cap_height;
    b->hash_kind = hash_kind;
    b->coeffs = d_coeffs.release();
    // the row-major leaves serve every reader of the handle; the column-major copy is kept as well when it is small
    // (proof-sized batches: the quotient kernel reads it coalesced), dropped for big ones
    b->lde = ncols * N * sizeof(u64) <= ((size_t)1 << 29) ? d_lde.release() : nullptr;
    b->leaves = d_leaves.release();
    b->digests = d_dig.release();
    b->cap = d_cap.release();
    b->owner_stream = st;
    *handle_out = b;
  }
  return "";
}

}  // namespace mp2

namespace mp2 {
// from_values / from_coeffs of columns that are already on the device (quotient.cu): whole-batch, no pipelining
// (these batches are a few columns wide); requested outputs are copied to the host, the rest stays behind the handle.
Status commit_device_columns(const u64 *d_cols, size_t ncols, u32 n_log, u32 rate_bits, u32 cap_height, u32 hash_kind,
                             int from_coeffs, uint64_t *const *coeffs_out, uint64_t *leaves_out, uint64_t *digests_out,
                             uint64_t *cap_out, mp2gpu_batch **handle_out, cudaStream_t st) {
  MP2_TRY(check_commit_args(ncols, n_log, rate_bits, cap_height, hash_kind));
  if (!cap_out) return "null cap_out";
  const size_t n = (size_t)1 << n_log, N = n << rate_bits, ncap = (size_t)1 << cap_height, ndig = 2 * (N - ncap);
  const bool want_rows = leaves_out != nullptr || handle_out != nullptr;
  DevBuf d_coeffs, d_lde, d_leaves, d_dig, d_cap;
  MP2_TRY(d_coeffs.alloc(ncols * n, st));
  MP2_TRY(d_lde.alloc(ncols * N, st));
  if (want_rows) MP2_TRY(d_leaves.alloc(ncols * N, st));
  MP2_TRY(d_dig.alloc(ndig * 4, st));
  MP2_TRY(d_cap.alloc(ncap * 4, st));
  MP2_TRY(dev_commit(d_cols, ncols, n_log, rate_bits, cap_height, hash_kind, from_coeffs, d_coeffs.p, d_lde.p, d_leaves.p,
                     d_dig.p, d_cap.p, st));
  if (coeffs_out) MP2_TRY(copy_columns_d2h(coeffs_out, d_coeffs.p, ncols, n, st));
  if (leaves_out) MP2_CUDA(cudaMemcpyAsync(leaves_out, d_leaves.p, ncols * N * sizeof(u64), cudaMemcpyDeviceToHost, st));
  if (digests_out && ndig) MP2_CUDA(cudaMemcpyAsync(digests_out, d_dig.p, ndig * 4 * sizeof(u64), cudaMemcpyDeviceToHost, st));
  MP2_CUDA(cudaMemcpyAsync(cap_out, d_cap.p, ncap * 4 * sizeof(u64), cudaMemcpyDeviceToHost, st));
  MP2_CUDA(cudaStreamSynchronize(st));
  if (handle_out) {
    int device = 0;
    MP2_CUDA(cudaGetDevice(&device));
    mp2gpu_batch *b = new mp2gpu_batch();
    b->device = device;
    b->ncols = ncols;
    b->n_log = n_log;
    b->rate_bits = rate_bits;
    b->cap_height = cap_height;
    b->hash_kind = hash_kind;
    b->coeffs = d_coeffs.release();
    b->lde = ncols * N * sizeof(u64) <= ((size_t)1 << 29) ? d_lde.release() : nullptr;
    b->leaves = d_leaves.release();
    b->digests = d_dig.release();
    b->cap = d_cap.release();
    b->owner_stream = st;
    *handle_out = b;
  }
  return "";
}
}  // namespace mp2

using namespace mp2;

namespace {

Status merkle_prove_indices(size_t nleaves, u32 cap_height, size_t leaf_index, std::vector<size_t> *idx) {
  int lg = log2_exact(nleaves);
  if (lg < 0) return "MerkleTree::prove: number of leaves is not a power of two";
  if ((int)cap_height > lg) return "MerkleTree::prove: cap_height > log2(leaves.len())";
  if (leaf_index >= nleaves) return "MerkleTree::prove: leaf_index out of range";
  u32 h = (u32)lg - cap_height;
  size_t per = 2 * (((size_t)1 << h) - 1);
  size_t base = (leaf_index >> h) * per;
  size_t pair_index = leaf_index & (((size_t)1 << h) - 1);
  for (u32 i = 0; i < h; i++) {
    size_t parity = pair_index & 1;
    pair_index >>= 1;
    size_t q = (pair_index << (i + 1)) + ((size_t)1 << i) - 1;
    idx->push_back(base + 2 * q + (1 - parity));
  }
  return "";
}

template <typename F>
const char *guarded(F f) {
  try {
    return to_c(f());
  } catch (const std::exception &e) {
    return to_c(std::string("exception: ") + e.what());
  } catch (...) {
    return to_c("unknown exception");
  }
}

}  // namespace

extern "C" {

const char *mp2gpu_version(void) { return "0.1.0 (sm_100a)"; }

void mp2gpu_free_string(const char *s) { free((void *)s); }

const char *mp2gpu_device_count(int *count_out) {
  return guarded([&]() -> Status {
    if (!count_out) return "null count_out";
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
      *count_out = 0;
      return std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e);
    }
    *count_out = n;
    return "";
  });
}

const char *mp2gpu_init(int device) {
  return guarded([&]() -> Status {
    if (device < 0) return "negative device index";
    t_ctx.device = device;
    cudaStream_t st;
    MP2_TRY(ctx_stream(&st));
    cudaDeviceProp prop;
    MP2_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
      return "device " + std::to_string(device) + " (" + prop.name + ", sm_" + std::to_string(prop.major) +
             std::to_string(prop.minor) + ") is not a Blackwell sm_100 part; libmp2gpu is built for sm_100a only";
    return "";
  });
}

const char *mp2gpu_trim(void) {
  return guarded([&]() -> Status {
    cudaStream_t st;
    MP2_TRY(ctx_stream(&st));
    DevStreams &d = t_ctx.per_device[t_ctx.device];
    for (auto &kv : d.idle) cudaFreeAsync(kv.second, st);
    d.idle.clear();
    d.idle_bytes = 0;
    MP2_CUDA(cudaStreamSynchronize(st));
    table_cache_clear();
    cudaMemPool_t pool;
    MP2_CUDA(cudaDeviceGetDefaultMemPool(&pool, t_ctx.device));
    MP2_CUDA(cudaMemPoolTrimTo(pool, 0));
    return "";
  });
}

const char *mp2gpu_host_alloc(void **ptr_out, size_t bytes) {
  return guarded([&]() -> Status {
    if (!ptr_out) return "null ptr_out";
    cudaStream_t st;
    MP2_TRY(ctx_stream(&st));
    MP2_CUDA(cudaHostAlloc(ptr_out, bytes ? bytes : 1, cudaHostAllocPortable));  // pinned for every device of the process (sharded entry point)
    return "";
  });
}
const char *mp2gpu_host_free(void *ptr) {
  return guarded([&]() -> Status {
    if (ptr) MP2_CUDA(cudaFreeHost(ptr));
    return "";
  });
}

const char *mp2gpu_commit_from_values(const uint64_t *const *cols, size_t ncols, uint32_t n_log, uint32_t rate_bits,
                                      uint32_t cap_height, uint32_t hash_kind, uint64_t *const *coeffs_out,
                                      uint64_t *leaves_out, uint64_t *digests_out, uint64_t *cap_out,
                                      mp2gpu_batch **handle_out) {
  return guarded([&]() {
    return commit_host(cols, ncols, n_log, rate_bits, cap_height, hash_kind, 0, coeffs_out, leaves_out, digests_out,
                       cap_out, handle_out);
  });
}
const char *mp2gpu_commit_from_coeffs(const uint64_t *const *cols, size_t ncols, uint32_t n_log, uint32_t rate_bits,
                                      uint32_t cap_height, uint32_t hash_kind, uint64_t *const *coeffs_out,
                                      uint64_t *leaves_out, uint64_t *digests_out, uint64_t *cap_out,
                                      mp2gpu_batch **handle_out) {
  return guarded([&]() {
    return commit_host(cols, ncols, n_log, rate_bits, cap_height, hash_kind, 1, coeffs_out, leaves_out, digests_out,
                       cap_out, handle_out);
  });
}

const char *mp2gpu_merkle_new(const uint64_t *leaves, size_t nleaves, size_t leaf_len, uint32_t cap_height,
                              uint32_t hash_kind, uint64_t *digests_out, uint64_t *cap_out) {
  return guarded([&]() -> Status {
    int lg = log2_exact(nleaves);
    if (lg < 0) return "MerkleTree::new: number of leaves (" + std::to_string(nleaves) + ") is not a power of two";
    if ((int)cap_height > lg)
      return "MerkleTree::new: cap_height=" + std::to_string(cap_height) +
             " should be at most log2(leaves.len())=" + std::to_string(lg);
    if (!cap_out || (!leaves && leaf_len)) return "null leaves / cap_out";
    cudaStream_t st;
    MP2_TRY(ctx_stream(&st));
    size_t ncap = (size_t)1 << cap_height, ndig = 2 * (nleaves - ncap);
    DevBuf d_lv, d_dig, d_cap;
    MP2_TRY(d_lv.alloc(nleaves * leaf_len, st));
    MP2_TRY(d_dig.alloc(ndig * 4, st));
    MP2_TRY(d_cap.alloc(ncap * 4, st));
    if (leaf_len)
      MP2_CUDA(cudaMemcpyAsync(d_lv.p, leaves, nleaves * leaf_len * sizeof(u64), cudaMemcpyHostToDevice, st));
    MP2_TRY(merkle_rowmajor(d_lv.p, nleaves, leaf_len, cap_height, hash_kind, d_dig.p, d_cap.p, st));
    MP2_CUDA(cudaMemcpyAsync(cap_out, d_cap.p, ncap * 4 * sizeof(u64), cudaMemcpyDeviceToHost, st));
    if (digests_out && ndig)
      MP2_CUDA(cudaMemcpyAsync(digests_out, d_dig.p, ndig * 4 * sizeof(u64), cudaMemcpyDeviceToHost, st));
    MP2_CUDA(cudaStreamSynchronize(st));
    return "";
  });
}

const char *mp2gpu_merkle_new_ragged(const uint64_t *const *leaves, const size_t *leaf_lens, size_t nleaves,
                                     uint32_t cap_height, uint32_t hash_kind, uint64_t *digests_out,
                                     uint64_t *cap_out) {
  return guarded([&]() -> Status {
    int lg = log2_exact(nleaves);
    if (lg < 0) return "MerkleTree::new: number of leaves (" + std::to_string(nleaves) + ") is not a power of two";
    if ((int)cap_height > lg)
      return "MerkleTree::new: cap_height=" + std::to_string(cap_height) +
             " should be at most log2(leaves.len())=" + std::to_string(lg);
    if (!cap_out || !leaves || !leaf_lens) return "null leaves / leaf_lens / cap_out";
    cudaStream_t st;
    MP2_TRY(ctx_stream(&st));
    std::vector<u64> off(nleaves + 1, 0);
    for (size_t i = 0; i < nleaves; i++) off[i + 1] = off[i] + leaf_lens[i];
    std::vector<u64> flat(off[nleaves] ? off[nleaves] : 1);
    for (size_t i = 0; i < nleaves; i++)
      if (leaf_lens[i]) memcpy(flat.data() + off[i], leaves[i], leaf_lens[i] * sizeof(u64));
    size_t ncap = (size_t)1 << cap_height, ndig = 2 * (nleaves - ncap);
    DevBuf d_flat, d_off, d_dig, d_cap;
    MP2_TRY(d_flat.alloc(flat.size(), st));
    MP2_TRY(d_off.alloc(off.size(), st));
    MP2_TRY(d_dig.alloc(ndig * 4, st));
    MP2_TRY(d_cap.alloc(ncap * 4, st));
    MP2_CUDA(cudaMemcpyAsync(d_flat.p, flat.data(), flat.size() * sizeof(u64), cudaMemcpyHostToDevice, st));
    MP2_CUDA(cudaMemcpyAsync(d_off.p, off.data(), off.size() * sizeof(u64), cudaMemcpyHostToDevice, st));
    MP2_TRY(merkle_ragged(d_flat.p, d_off.p, nleaves, cap_height, hash_kind, d_dig.p, d_cap.p, st));
    MP2_CUDA(cudaMemcpyAsync(cap_out, d_cap.p, ncap * 4 * sizeof(u64), cudaMemcpyDeviceToHost, st));
    if (digests_out && ndig)
      MP2_CUDA(cudaMemcpyAsync(digests_out, d_dig.p, ndig * 4 * sizeof(u64), cudaMemcpyDeviceToHost, st));
    MP2_CUDA(cudaStreamSynchronize(st));
    return "";
  });
}

const char *mp2gpu_merkle_prove(const uint64_t *digests, size_t nleaves, uint32_t cap_height, size_t leaf_index,
                                uint64_t *siblings_out, size_t *nsiblings_out) {
  return guarded([&]() -> Status {
    std::vector<size_t> idx;
    MP2_TRY(merkle_prove_indices(nleaves, cap_height, leaf_index, &idx));
    if (!idx.empty() && (!digests || !siblings_out)) return "null digests / siblings_out";
    for (size_t i = 0; i < idx.size(); i++) memcpy(siblings_out + 4 * i, digests + 4 * idx[i], 32);
    if (nsiblings_out) *nsiblings_out = idx.size();
    return "";
  });
}

const char *mp2gpu_hash_no_pad_batch(const uint64_t *inputs, size_t count, size_t input_len, uint32_t hash_kind,
                                     uint64_t *out) {
  return guarded([&]() -> Status {
    if (count == 0) return "";
    if (!out || (!inputs && input_len)) return "null inputs / out";
    cudaStream_t st;
    MP2_TRY(ctx_stream(&st));
    DevBuf d_in, d_out;
    MP2_TRY(d_in.alloc(count * input_len, st));
    MP2_TRY(d_out.alloc(count * 4, st));
    if (input_len)
      MP2_CUDA(cudaMemcpyAsync(d_in.p, inputs, count * input_len * sizeof(u64), cudaMemcpyHostToDevice, st));
    MP2_TRY(hash_no_pad_batch(d_in.p, count, input_len, hash_kind, d_out.p, st));
    MP2_CUDA(cudaMemcpyAsync(out, d_out.p, count * 4 * sizeof(u64), cudaMemcpyDeviceToHost, st));
    MP2_CUDA(cudaStreamSynchronize(st));
    return "";
  });
}

const char *mp2gpu_two_to_one_batch(const uint64_t *a, const uint64_t *b, size_t count, uint32_t hash_kind,
                                    uint64_t *out) {
  return guarded([&]() -> Status {
    if (count == 0) return "";
    if (!a || !b || !out) return "null a / b / out";
    cudaStream_t st;
    MP2_TRY(ctx_stream(&st));
    DevBuf d_a, d_b, d_out;
    MP2_TRY(d_a.alloc(count * 4, st));
    MP2_TRY(d_b.alloc(count * 4, st));
    MP2_TRY(d_out.alloc(count * 4, st));
    MP2_CUDA(cudaMemcpyAsync(d_a.p, a, count * 32, cudaMemcpyHostToDevice, st));
    MP2_CUDA(cudaMemcpyAsync(d_b.p, b, count * 32, cudaMemcpyHostToDevice, st));
    MP2_TRY(two_to_one_batch(d_a.p, d_b.p, count, hash_kind, d_out.p, st));
    MP2_CUDA(cudaMemcpyAsync(out, d_out.p, count * 32, cudaMemcpyDeviceToHost, st));
    MP2_CUDA(cudaStreamSynchronize(st));
    return "";
  });
}

const char *mp2gpu_permute_batch(uint64_t *states, size_t count, uint32_t hash_kind) {
  return guarded([&]() -> Status {
    if (count == 0) return "";
    if (!states) return "null states";
    cudaStream_t st;
    MP2_TRY(ctx_stream(&st));
    DevBuf d;
    MP2_TRY(d.alloc(count * 12, st));
    MP2_CUDA(cudaMemcpyAsync(d.p, states, count * 96, cudaMemcpyHostToDevice, st));
    MP2_TRY(permute_batch(d.p, count, hash_kind, st));
    MP2_CUDA(cudaMemcpyAsync(states, d.p, count * 96, cudaMemcpyDeviceToHost, st));
    MP2_CUDA(cudaStreamSynchronize(st));
    return "";
  });
}

const char *mp2gpu_fri_proof_of_work(const uint64_t *duplex_state, uint32_t witness_pos, uint32_t min_leading_zeros,
                                     uint32_t hash_kind, uint64_t *witness_out) {
  return guarded([&]() -> Status {
    if (!duplex_state || !witness_out) return "null state / witness_out";
    cudaStream_t st;
    MP2_TRY(ctx_stream(&st));
    if (min_leading_zeros > 40) return "proof_of_work_bits too large";
    // The smallest witness is geometric with mean 2^min_lz, and ranges are scanned in order (the atomicMin inside a
    // range keeps the "smallest witness" rule), so the first range is sized at twice the mean -- 86 % of the grinds
    // end there -- and later ranges double up to 2^22.  (A fixed 2^20 range hashed 16x more candidates than needed
    // at the reference's 16 bits: 0.95 ms of a 2.5 ms fri_proof.)  Never below ~1.4 waves of the hash kernel.
    u64 batch = (u64)1 << (min_leading_zeros + 1 < 17 ? 17 : min_leading_zeros + 1 > 22 ? 22 : min_leading_zeros + 1);
    for (u64 start = 0; start < kP; start += batch, batch = batch < ((u64)1 << 22) ? batch << 1 : batch) {
      u64 found = ~(u64)0;
      u64 count = kP - start < batch ? kP - start : batch;
      MP2_TRY(pow_search((const u64 *)duplex_state, witness_pos, min_leading_zeros, hash_kind, start, count, &found, st));
      if (found != ~(u64)0) {
        *witness_out = found;
        return "";
      }
    }
    return "Proof of work failed. This is highly unlikely!";
  });
}

// ---- handle ---------------------------------------------------------------------------------
const char *mp2gpu_batch_shape(const mp2gpu_batch *b, size_t *ncols, uint32_t *n_log, uint32_t *rate_bits,
                               uint32_t *cap_height, uint32_t *hash_kind) {
  return guarded([&]() -> Status {
    if (!b) return "null batch handle";
    if (ncols) *ncols = b->ncols;
    if (n_log) *n_log = b->n_log;
    if (rate_bits) *rate_bits = b->rate_bits;
    if (cap_height) *cap_height = b->cap_height;
    if (hash_kind) *hash_kind = b->hash_kind;
    return "";
  });
}

const char *mp2gpu_batch_fetch_rows(const mp2gpu_batch *b, const uint64_t *row_idx, size_t nrows, uint64_t *out) {
  return guarded([&]() -> Status {
    if (!b) return "null batch handle";
    if (nrows == 0) return "";
    if (!row_idx || !out) return "null row_idx / out";
    const size_t N = ((size_t)1 << b->n_log) << b->rate_bits;
    for (size_t i = 0; i < nrows; i++)
      if (row_idx[i] >= N) return "get_lde_values: row index out of range";
    DeviceScope on_batch_device(b->device);
    cudaStream_t st;
    MP2_TRY(ctx_stream(&st));
    DevBuf d_idx, d_out;
    MP2_TRY(d_idx.alloc(nrows, st));
    MP2_TRY(d_out.alloc(nrows * b->ncols, st));
    MP2_CUDA(cudaMemcpyAsync(d_idx.p, row_idx, nrows * sizeof(u64), cudaMemcpyHostToDevice, st));
    MP2_TRY(gather_rows(b->leaves, b->lde, N, b->ncols, d_idx.p, nrows, d_out.p, st));
    MP2_CUDA(cudaMemcpyAsync(out, d_out.p, nrows * b->ncols * sizeof(u64), cudaMemcpyDeviceToHost, st));
    MP2_CUDA(cudaStreamSynchronize(st));
    return "";
  });
}

const char *mp2gpu_batch_prove(const mp2gpu_batch *b, size_t leaf_index, uint64_t *siblings_out,
                               size_t *nsiblings_out) {
  return guarded([&]() -> Status {
    if (!b) return "null batch handle";
    const size_t N = ((size_t)1 << b->n_log) << b->rate_bits;
    std::vector<size_t> idx;
    MP2_TRY(merkle_prove_indices(N, b->cap_height, leaf_index, &idx));
    if (!idx.empty() && !siblings_out) return "null siblings_out";
    DeviceScope on_batch_device(b->device);
    cudaStream_t st;
    MP2_TRY(ctx_stream(&st));
    // one device gather + one copy for the whole path (round 1 issued one 32-byte copy per sibling)
    const u64 leaf = (u64)leaf_index;
    if (!idx.empty())
      MP2_TRY(merkle_open(b->leaves, b->lde, N, b->ncols, b->digests, N, b->cap_height, &leaf, 1, nullptr,
                          (u64 *)siblings_out, st));
    if (nsiblings_out) *nsiblings_out = idx.size();
    return "";
  });
}

const char *mp2gpu_batch_open(const mp2gpu_batch *b, const uint64_t *leaf_idx, size_t count, uint64_t *rows_out,
                              uint64_t *siblings_out) {
  return guarded([&]() -> Status {
    if (!b) return "null batch handle";
    if (count && !leaf_idx) return "null leaf_idx";
    DeviceScope on_batch_device(b->device);
    cudaStream_t st;
    MP2_TRY(ctx_stream(&st));
    const size_t N = ((size_t)1 << b->n_log) << b->rate_bits;
    return merkle_open(b->leaves, b->lde, N, b->ncols, b->digests, N, b->cap_height, (const u64 *)leaf_idx, count,
                       (u64 *)rows_out, (u64 *)siblings_out, st);
  });
}

const char *mp2gpu_batch_fetch(const mp2gpu_batch *b, uint64_t *const *coeffs_out, uint64_t *leaves_out,
                               uint64_t *digests_out, uint64_t *cap_out) {
  return guarded([&]() -> Status {
    if (!b) return "null batch handle";
    DeviceScope on_batch_device(b->device);
    cudaStream_t st;
    MP2_TRY(ctx_stream(&st));
    const size_t n = (size_t)1 << b->n_log, N = n << b->rate_bits, ncap = (size_t)1 << b->cap_height;
    if (coeffs_out) MP2_TRY(copy_columns_d2h(coeffs_out, b->coeffs, b->ncols, n, st));
    if (leaves_out) {
      if (!b->leaves) return "batch holds no row-major leaves";
      MP2_CUDA(cudaMemcpyAsync(leaves_out, b->leaves, N * b->ncols * sizeof(u64), cudaMemcpyDeviceToHost, st));
    }
    if (digests_out && N > ncap)
      MP2_CUDA(cudaMemcpyAsync(digests_out, b->digests, 2 * (N - ncap) * 32, cudaMemcpyDeviceToHost, st));
    if (cap_out) MP2_CUDA(cudaMemcpyAsync(cap_out, b->cap, ncap * 32, cudaMemcpyDeviceToHost, st));
    MP2_CUDA(cudaStreamSynchronize(st));
    return "";
  });
}

void mp2gpu_batch_free(mp2gpu_batch *b) {
  if (!b) return;
  // Every call that reads a handle synchronises before it returns, so nothing is in flight on these buffers.
  // They are freed on the stream they were allocated on: the pool then hands the same blocks to that thread's
  // next commitment without growing (freeing 17 GB on another stream, or with cudaFree, made the next
  // allocation map fresh pages: 350 -> 530..960 ms per wide-batch call).
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(b->device);
  for (u64 *p : {b->coeffs, b->lde, b->leaves, b->digests, b->cap})
    if (p) pool_free(p, b->owner_stream);  // the owner thread gets them back into its cache, anyone else frees to the pool
  cudaSetDevice(prev);
  delete b;
}

// ---- device-pointer stages --------------------------------------------------------------------
const char *mp2gpu_dev_intt(const uint64_t *values, size_t in_stride, uint64_t *coeffs, size_t out_stride,
                            size_t ncols, uint32_t n_log, void *stream) {
  return guarded([&]() -> Status {
    cudaStream_t st;
    MP2_TRY(pick_stream(stream, &st));
    return ntt_intt((const u64 *)values, in_stride, (u64 *)coeffs, out_stride, ncols, n_log, st);
  });
}

const char *mp2gpu_dev_coset_lde(const uint64_t *coeffs, size_t in_stride, uint64_t *lde, size_t lde_stride,
                                 size_t ncols, uint32_t n_log, uint32_t rate_bits, uint32_t shard_log,
                                 size_t shard_stride, void *stream) {
  return guarded([&]() -> Status {
    cudaStream_t st;
    MP2_TRY(pick_stream(stream, &st));
    return ntt_coset_lde((const u64 *)coeffs, in_stride, (u64 *)lde, lde_stride, ncols, n_log, rate_bits, shard_log,
                         shard_stride, st);
  });
}

const char *mp2gpu_dev_coset_lde_peer(const uint64_t *coeffs, size_t in_stride, uint64_t *const *shard_bases,
                                      size_t lde_stride, size_t ncols, uint32_t n_log, uint32_t rate_bits,
                                      uint32_t shard_log, uint32_t first_shard, uint64_t *scratch, void *stream) {
  return guarded([&]() -> Status {
    cudaStream_t st;
    MP2_TRY(pick_stream(stream, &st));
    if (!shard_bases) return "null shard_bases";
    return ntt_coset_lde((const u64 *)coeffs, in_stride, (u64 *)scratch, lde_stride, ncols, n_log, rate_bits, shard_log, 0, st,
                         (u64 *const *)shard_bases, kCosetShift, LDE_ALL, 0, 0, first_shard);
  });
}

const char *mp2gpu_dev_merkle_colmajor(const uint64_t *lde, size_t lde_stride, size_t ncols, size_t nleaves,
                                       uint32_t cap_height, uint32_t hash_kind, uint64_t *leaves_out,
                                       uint64_t *digests_out, uint64_t *cap_out, void *stream) {
  return guarded([&]() -> Status {
    cudaStream_t st;
    MP2_TRY(pick_stream(stream, &st));
    return merkle_colmajor((const u64 *)lde, lde_stride, ncols, nleaves, cap_height, hash_kind, (u64 *)leaves_out,
                           (u64 *)digests_out, (u64 *)cap_out, st);
  });
}

const char *mp2gpu_dev_merkle_colmajor_leaves(const uint64_t *lde, size_t lde_stride, size_t ncols, size_t nleaves,
                                              uint32_t cap_height, uint32_t hash_kind, size_t leaf_begin,
                                              size_t leaf_end, uint64_t *leaves_out, uint64_t *digests_out,
                                              uint64_t *cap_out, void *stream) {
  return guarded([&]() -> Status {
    cudaStream_t st;
    MP2_TRY(pick_stream(stream, &st));
    return merkle_colmajor_leaves((const u64 *)lde, lde_stride, ncols, nleaves, cap_height, hash_kind, leaf_begin,
                                  leaf_end, (u64 *)leaves_out, (u64 *)digests_out, (u64 *)cap_out, st);
  });
}

const char *mp2gpu_dev_merkle_levels(size_t nleaves, uint32_t cap_height, uint32_t hash_kind, uint64_t *digests,
                                     uint64_t *cap, void *stream) {
  return guarded([&]() -> Status {
    cudaStream_t st;
    MP2_TRY(pick_stream(stream, &st));
    return merkle_levels(nleaves, cap_height, hash_kind, (u64 *)digests, (u64 *)cap, st);
  });
}

const char *mp2gpu_dev_merkle_rowmajor(const uint64_t *leaves, size_t nleaves, size_t leaf_len, uint32_t cap_height,
                                       uint32_t hash_kind, uint64_t *digests_out, uint64_t *cap_out, void *stream) {
  return guarded([&]() -> Status {
    cudaStream_t st;
    MP2_TRY(pick_stream(stream, &st));
    return merkle_rowmajor((const u64 *)leaves, nleaves, leaf_len, cap_height, hash_kind, (u64 *)digests_out,
                           (u64 *)cap_out, st);
  });
}

const char *mp2gpu_dev_commit(const uint64_t *cols_dev, size_t ncols, uint32_t n_log, uint32_t rate_bits,
                              uint32_t cap_height, uint32_t hash_kind, int from_coeffs, uint64_t *coeffs_dev,
                              uint64_t *lde_dev, uint64_t *leaves_dev, uint64_t *digests_dev, uint64_t *cap_dev,
                              void *stream) {
  return guarded([&]() -> Status {
    cudaStream_t st;
    MP2_TRY(pick_stream(stream, &st));
    if (!cols_dev || !coeffs_dev || !lde_dev || !cap_dev) return "null device buffer";
    if (!digests_dev && cap_height < n_log + rate_bits) return "null digests buffer";
    return dev_commit((const u64 *)cols_dev, ncols, n_log, rate_bits, cap_height, hash_kind, from_coeffs,
                      (u64 *)coeffs_dev, (u64 *)lde_dev, (u64 *)leaves_dev, (u64 *)digests_dev, (u64 *)cap_dev, st);
  });
}

const char *mp2gpu_dev_canonicalize(const uint64_t *in, uint64_t *out, size_t count, void *stream) {
  return guarded([&]() -> Status {
    cudaStream_t st;
    MP2_TRY(pick_stream(stream, &st));
    return ntt_canonicalize((const u64 *)in, count, (u64 *)out, count, 1, count, st);
  });
}

const char *mp2gpu_sync(void *stream) {
  return guarded([&]() -> Status {
    cudaStream_t st;
    MP2_TRY(pick_stream(stream, &st));
    MP2_CUDA(cudaStreamSynchronize(st));
    return "";
  });
}

uint64_t mp2gpu_launch_count(void) { return g_launches.load(); }

}  // extern "C"
