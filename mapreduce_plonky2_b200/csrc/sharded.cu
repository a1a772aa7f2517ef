// One wide PolynomialBatch over the G GPUs of ONE process, behind the C ABI:
//
//   columns sharded -> per-device upload + iNTT + coset LDE whose stores land directly in the HBM of the device
//   that owns the leaf (peer stores over NVLink/NVSwitch: the column-shard -> row-shard exchange IS the LDE
//   kernel's store) -> barrier -> per-device leaf hashing of its row shard, block by block with the host copy of
//   each block overlapped -> the device's 2^cap/G subtrees -> its contiguous slice of digests / cap.
//
// Device g of G owns columns [g*c/G, (g+1)*c/G) before the exchange and leaves [g*N/G, (g+1)*N/G) after it; a
// rank's digests and cap entries are contiguous slices of plonky2's arrays (MerkleTree::new splits `digests` into
// 2^cap_height per-subtree chunks), so the host outputs are written in place with no reassembly.  No NCCL: inside
// one process the exchange needs only peer access, and the 512-byte cap travels through the host buffer it is
// returned in.  (The one-process-per-GPU form of the same path is sharded.py, where NCCL / symmetric memory
// provide the addressing across processes.)
//
// Replaces PolynomialBatch::from_values / from_coeffs for the wide-batch configuration (BASELINE configs[2]);
// the prover process that owns the box is the reference's deployment unit (mp2-v1/src/api.rs:154-165).
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/mp2gpu.h"
#include "internal.h"

using namespace mp2;

struct mp2gpu_comm {
  struct Rank {
    int device = 0;
    cudaStream_t st = nullptr, cp = nullptr, up = nullptr;
    u64 *recv = nullptr;  // G x c_loc x n_loc: block s = rank s's columns restricted to MY leaves (peer-writable)
    size_t recv_elems = 0;
    u64 *mid = nullptr;   // c_loc x N: four-step intermediate of the LDE (persistent: no multi-GB pool traffic per call)
    size_t mid_elems = 0;
  };
  std::vector<Rank> ranks;
  std::mutex call_mu;  // one sharded commitment at a time per communicator
  // reusable barrier over the G worker threads of a call
  std::mutex bar_mu;
  std::condition_variable bar_cv;
  int bar_count = 0, bar_gen = 0;
  bool failed = false;
  std::string error;
};

namespace {

const char *to_c(const Status &s) {
  if (s.empty()) return nullptr;
  char *p = (char *)malloc(s.size() + 1);
  if (p) memcpy(p, s.c_str(), s.size() + 1);
  return p;
}

void comm_barrier(mp2gpu_comm *c) {
  std::unique_lock<std::mutex> lk(c->bar_mu);
  const int gen = c->bar_gen;
  if (++c->bar_count == (int)c->ranks.size()) {
    c->bar_count = 0;
    c->bar_gen++;
    c->bar_cv.notify_all();
  } else {
    c->bar_cv.wait(lk, [&] { return c->bar_gen != gen; });
  }
}
void comm_fail(mp2gpu_comm *c, int g, const Status &s) {
  std::lock_guard<std::mutex> lk(c->bar_mu);
  if (!c->failed) {
    c->failed = true;
    c->error = "device " + std::to_string(c->ranks[g].device) + " (rank " + std::to_string(g) + "): " + s;
  }
}
bool comm_failed(mp2gpu_comm *c) {
  std::lock_guard<std::mutex> lk(c->bar_mu);
  return c->failed;
}

struct ShardedArgs {
  const uint64_t *const *cols;
  size_t ncols;
  u32 n_log, rate_bits, cap_height, hash_kind;
  int from_coeffs;
  uint64_t *const *coeffs_out;
  uint64_t *leaves_out, *digests_out, *cap_out;
};

// stage A: buffers, upload, iNTT (coefficients start travelling back)
struct RankState {
  DevBuf in, coeffs, leaves, dig, cap;
  cudaEvent_t ev = nullptr;
  ~RankState() {
    if (ev) cudaEventDestroy(ev);
  }
};

Status rank_stage_a(mp2gpu_comm *c, int g, const ShardedArgs &a, RankState &rs) {
  mp2gpu_comm::Rank &r = c->ranks[g];
  const size_t G = c->ranks.size();
  MP2_CUDA(cudaSetDevice(r.device));
  const size_t n = (size_t)1 << a.n_log, N = n << a.rate_bits;
  const size_t c_loc = a.ncols / G, n_loc = N / G;
  u32 glog = 0;
  while (((size_t)1 << glog) < G) glog++;
  const size_t ncap_loc = ((size_t)1 << a.cap_height) >> glog, ndig_loc = 2 * (n_loc - ncap_loc);
  MP2_CUDA(cudaEventCreateWithFlags(&rs.ev, cudaEventDisableTiming));
  const size_t need = G * c_loc * n_loc;
  if (r.recv_elems < need) {  // peer-visible: plain cudaMalloc (pool memory would need cudaMemPoolSetAccess)
    MP2_CUDA(cudaStreamSynchronize(r.st));
    if (r.recv) MP2_CUDA(cudaFree(r.recv));
    r.recv = nullptr;
    r.recv_elems = 0;
    MP2_CUDA(cudaMalloc(&r.recv, need * sizeof(u64)));
    r.recv_elems = need;
  }
  if (ntt_lde_is_two_pass(a.n_log) && r.mid_elems < c_loc * N) {
    MP2_CUDA(cudaStreamSynchronize(r.st));
    if (r.mid) MP2_CUDA(cudaFree(r.mid));
    r.mid = nullptr;
    r.mid_elems = 0;
    MP2_CUDA(cudaMalloc(&r.mid, c_loc * N * sizeof(u64)));
    r.mid_elems = c_loc * N;
  }
  MP2_TRY(rs.in.alloc(c_loc * n, r.st));
  MP2_TRY(rs.coeffs.alloc(c_loc * n, r.st));
  if (a.leaves_out) MP2_TRY(rs.leaves.alloc(n_loc * a.ncols, r.st));
  MP2_TRY(rs.dig.alloc(ndig_loc * 4, r.st));
  MP2_TRY(rs.cap.alloc(ncap_loc * 4, r.st));
  // upload in column blocks on the upload stream; the iNTT of block k overlaps the upload of block k+1 and the
  // coefficients of block k go back on the copy stream (same scheme as the single-GPU entry point)
  const size_t nblocks = (c_loc >= 8 && c_loc * n >= ((size_t)1 << 23)) ? 4 : 1;
  MP2_CUDA(cudaEventRecord(rs.ev, r.st));
  MP2_CUDA(cudaStreamWaitEvent(r.up, rs.ev, 0));
  for (size_t k = 0; k < nblocks; k++) {
    const size_t c0 = k * c_loc / nblocks, c1 = (k + 1) * c_loc / nblocks, cnt = c1 - c0;
    if (!cnt) continue;
    u64 *in_k = rs.in.p + c0 * n, *co_k = rs.coeffs.p + c0 * n;
    MP2_TRY(copy_columns_h2d(in_k, a.cols + g * c_loc + c0, cnt, n, r.up));
    MP2_CUDA(cudaEventRecord(rs.ev, r.up));
    MP2_CUDA(cudaStreamWaitEvent(r.st, rs.ev, 0));
    if (a.from_coeffs) MP2_TRY(ntt_canonicalize(in_k, n, co_k, n, cnt, n, r.st));
    else MP2_TRY(ntt_intt(in_k, n, co_k, n, cnt, a.n_log, r.st));
    if (a.coeffs_out) {
      MP2_CUDA(cudaEventRecord(rs.ev, r.st));
      MP2_CUDA(cudaStreamWaitEvent(r.cp, rs.ev, 0));
      MP2_TRY(copy_columns_d2h(a.coeffs_out + g * c_loc + c0, co_k, cnt, n, r.cp));
    }
  }
  return "";
}

// stage B: the LDE with the exchange fused into its stores; returns once this rank's stores have landed
Status rank_stage_b(mp2gpu_comm *c, int g, const ShardedArgs &a, RankState &rs) {
  mp2gpu_comm::Rank &r = c->ranks[g];
  const size_t G = c->ranks.size();
  MP2_CUDA(cudaSetDevice(r.device));
  const size_t n = (size_t)1 << a.n_log, N = n << a.rate_bits;
  const size_t c_loc = a.ncols / G, n_loc = N / G;
  u32 glog = 0;
  while (((size_t)1 << glog) < G) glog++;
  std::vector<u64 *> bases(G);
  for (size_t s = 0; s < G; s++) bases[s] = c->ranks[s].recv + (size_t)g * c_loc * n_loc;  // MY block in rank s's buffer
  // rank g stores to rank g first, then g+1, ...: at any moment the ranks target different peers
  MP2_TRY(ntt_coset_lde(rs.coeffs.p, n, ntt_lde_is_two_pass(a.n_log) ? r.mid : nullptr, n_loc, c_loc, a.n_log, a.rate_bits, glog, 0, r.st, bases.data(),
                        kCosetShift, LDE_ALL, 0, 0, (u32)g));
  MP2_CUDA(cudaStreamSynchronize(r.st));
  return "";
}

// stage C: my leaves, my subtrees, my slices of the host outputs
Status rank_stage_c(mp2gpu_comm *c, int g, const ShardedArgs &a, RankState &rs) {
  mp2gpu_comm::Rank &r = c->ranks[g];
  const size_t G = c->ranks.size();
  MP2_CUDA(cudaSetDevice(r.device));
  const size_t n = (size_t)1 << a.n_log, N = n << a.rate_bits;
  const size_t n_loc = N / G;
  u32 glog = 0;
  while (((size_t)1 << glog) < G) glog++;
  const u32 cap_loc_h = a.cap_height - glog;
  const size_t ncap_loc = (size_t)1 << cap_loc_h, ndig_loc = 2 * (n_loc - ncap_loc);
  const size_t nchunks = (a.leaves_out && n_loc >= ((size_t)1 << 16)) ? 8 : 1;
  for (size_t j = 0; j < nchunks; j++) {
    const size_t lb = j * (n_loc / nchunks), le = (j + 1) * (n_loc / nchunks);
    MP2_TRY(merkle_colmajor_leaves(r.recv, n_loc, a.ncols, n_loc, cap_loc_h, a.hash_kind, lb, le, rs.leaves.p, rs.dig.p,
                                   rs.cap.p, r.st));
    if (a.leaves_out) {
      MP2_CUDA(cudaEventRecord(rs.ev, r.st));
      MP2_CUDA(cudaStreamWaitEvent(r.cp, rs.ev, 0));
      MP2_CUDA(cudaMemcpyAsync(a.leaves_out + ((size_t)g * n_loc + lb) * a.ncols, rs.leaves.p + lb * a.ncols,
                               (le - lb) * a.ncols * sizeof(u64), cudaMemcpyDeviceToHost, r.cp));
    }
  }
  MP2_TRY(merkle_levels(n_loc, cap_loc_h, a.hash_kind, rs.dig.p, rs.cap.p, r.st));
  MP2_CUDA(cudaMemcpyAsync(a.cap_out + (size_t)g * ncap_loc * 4, rs.cap.p, ncap_loc * 4 * sizeof(u64),
                           cudaMemcpyDeviceToHost, r.st));
  if (a.digests_out && ndig_loc)
    MP2_CUDA(cudaMemcpyAsync(a.digests_out + (size_t)g * ndig_loc * 4, rs.dig.p, ndig_loc * 4 * sizeof(u64),
                             cudaMemcpyDeviceToHost, r.st));
  MP2_CUDA(cudaStreamSynchronize(r.st));
  MP2_CUDA(cudaStreamSynchronize(r.cp));
  return "";
}

void rank_main(mp2gpu_comm *c, int g, const ShardedArgs *a) {
  RankState rs;
  mp2gpu_comm::Rank &r = c->ranks[g];
  // Every stage is followed by a barrier that ALL ranks reach, failed or not; after it they all see the same
  // `failed` flag and leave together, so an error on one device can never strand the others at a barrier.
  Status s = rank_stage_a(c, g, *a, rs);
  if (!s.empty()) comm_fail(c, g, s);
  comm_barrier(c);  // every receive buffer exists (and nobody still reads the previous call's)
  if (!comm_failed(c)) {
    s = rank_stage_b(c, g, *a, rs);
    if (!s.empty()) comm_fail(c, g, s);
  }
  comm_barrier(c);  // every rank's stores into my receive buffer have landed
  if (!comm_failed(c)) {
    s = rank_stage_c(c, g, *a, rs);
    if (!s.empty()) comm_fail(c, g, s);
  }
  // drain before the stream-ordered buffers of `rs` go back to the pool
  cudaSetDevice(r.device);
  cudaStreamSynchronize(r.up);
  cudaStreamSynchronize(r.st);
  cudaStreamSynchronize(r.cp);
  comm_barrier(c);  // nobody frees while a peer could still be storing
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

const char *mp2gpu_comm_init(int ndev, const int *devs, mp2gpu_comm **comm_out) {
  return guarded([&]() -> Status {
    if (!comm_out) return "null comm_out";
    *comm_out = nullptr;
    if (ndev <= 0 || (ndev & (ndev - 1))) return "mp2gpu_comm_init: the number of devices must be a power of two";
    if (ndev > 16) return "mp2gpu_comm_init: at most 16 devices (one per cap subtree at cap_height 4)";
    int have = 0;
    cudaError_t e = cudaGetDeviceCount(&have);
    if (e != cudaSuccess || have == 0)
      return std::string("no usable CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(e);
    int prev = 0;
    cudaGetDevice(&prev);
    std::unique_ptr<mp2gpu_comm, void (*)(mp2gpu_comm *)> c(new mp2gpu_comm(), mp2gpu_comm_free);
    c->ranks.resize(ndev);
    for (int g = 0; g < ndev; g++) {
      const int d = devs ? devs[g] : g;
      if (d < 0 || d >= have) return "mp2gpu_comm_init: device " + std::to_string(d) + " out of range";
      for (int h = 0; h < g; h++)
        if (c->ranks[h].device == d) return "mp2gpu_comm_init: device " + std::to_string(d) + " listed twice";
      c->ranks[g].device = d;
    }
    for (int g = 0; g < ndev; g++) {
      mp2gpu_comm::Rank &r = c->ranks[g];
      MP2_CUDA(cudaSetDevice(r.device));
      cudaDeviceProp prop;
      MP2_CUDA(cudaGetDeviceProperties(&prop, r.device));
      if (prop.major != 10) return "device " + std::to_string(r.device) + " is not a Blackwell sm_100 part";
      for (int h = 0; h < ndev; h++) {
        if (h == g) continue;
        int can = 0;
        MP2_CUDA(cudaDeviceCanAccessPeer(&can, r.device, c->ranks[h].device));
        if (!can)
          return "devices " + std::to_string(r.device) + " and " + std::to_string(c->ranks[h].device) +
                 " cannot access each other's memory (no NVLink/PCIe peer path)";
        cudaError_t pe = cudaDeviceEnablePeerAccess(c->ranks[h].device, 0);
        if (pe == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (pe != cudaSuccess) return std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(pe);
      }
      MP2_CUDA(cudaStreamCreateWithFlags(&r.st, cudaStreamNonBlocking));
      MP2_CUDA(cudaStreamCreateWithFlags(&r.cp, cudaStreamNonBlocking));
      MP2_CUDA(cudaStreamCreateWithFlags(&r.up, cudaStreamNonBlocking));
      cudaMemPool_t pool;
      MP2_CUDA(cudaDeviceGetDefaultMemPool(&pool, r.device));
      uint64_t keep = UINT64_MAX;
      MP2_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    cudaSetDevice(prev);
    *comm_out = c.release();
    return "";
  });
}

void mp2gpu_comm_free(mp2gpu_comm *c) {
  if (!c) return;
  int prev = 0;
  cudaGetDevice(&prev);
  for (auto &r : c->ranks) {
    if (!r.st && !r.recv) continue;
    cudaSetDevice(r.device);
    if (r.st) cudaStreamSynchronize(r.st);
    if (r.recv) cudaFree(r.recv);
    if (r.mid) cudaFree(r.mid);
    for (cudaStream_t s : {r.st, r.cp, r.up})
      if (s) cudaStreamDestroy(s);
  }
  cudaSetDevice(prev);
  delete c;
}

const char *mp2gpu_commit_from_values_sharded(mp2gpu_comm *comm, const uint64_t *const *cols, size_t ncols,
                                              uint32_t n_log, uint32_t rate_bits, uint32_t cap_height,
                                              uint32_t hash_kind, int from_coeffs, uint64_t *const *coeffs_out,
                                              uint64_t *leaves_out, uint64_t *digests_out, uint64_t *cap_out) {
  return guarded([&]() -> Status {
    if (!comm) return "null communicator";
    MP2_TRY(check_commit_args(ncols, n_log, rate_bits, cap_height, hash_kind));
    if (!cols || !cap_out) return "null cols / cap_out";
    const size_t G = comm->ranks.size();
    u32 glog = 0;
    while (((size_t)1 << glog) < G) glog++;
    if (ncols % G) return "columns must split evenly: " + std::to_string(ncols) + " over " + std::to_string(G) + " devices";
    if (glog > cap_height)
      return "world size " + std::to_string(G) + " exceeds the number of cap subtrees 2^" + std::to_string(cap_height);
    for (size_t c = 0; c < ncols; c++)
      if (!cols[c]) return "null column pointer";
    std::lock_guard<std::mutex> one_call(comm->call_mu);
    if (G == 1) {  // nothing to exchange: the single-GPU entry point on that device
      DeviceScope scope(comm->ranks[0].device);
      return commit_host(cols, ncols, n_log, rate_bits, cap_height, hash_kind, from_coeffs, coeffs_out, leaves_out,
                         digests_out, cap_out, nullptr);
    }
    int prev = 0;
    cudaGetDevice(&prev);
    ShardedArgs a{cols, ncols, n_log, rate_bits, cap_height, hash_kind, from_coeffs, coeffs_out, leaves_out, digests_out, cap_out};
    comm->failed = false;
    comm->error.clear();
    std::vector<std::thread> workers;
    for (size_t g = 0; g < G; g++) workers.emplace_back(rank_main, comm, (int)g, &a);
    for (auto &w : workers) w.join();
    cudaSetDevice(prev);
    return comm->failed ? comm->error : Status("");
  });
}

}  // extern "C"
