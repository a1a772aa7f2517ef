// Device-resident twiddle tables, built once per (device, size) and cached for the process lifetime.
// They stand in for plonky2's FftRootTable (the `fft_root_table` argument of
// PolynomialBatch::from_values, which the GPU path ignores -- SURVEY.md 8(b) "Threading").
#include <map>
#include <mutex>
#include <vector>

#include "gl.cuh"
#include "internal.h"

namespace mp2 {

__global__ void k_fill_powers(u64 *out, u64 base, size_t count) {
  size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m < count) out[m] = gl_canon(gl_pow(base, m));
}

// out[(k << log_n) + j] = bases[k]^j
__global__ void k_fill_coset_scale(u64 *out, const u64 *bases, u32 log_n, size_t count) {
  size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m < count) out[m] = gl_canon(gl_pow(bases[m >> log_n], m & (((size_t)1 << log_n) - 1)));
}

namespace {
struct Key {
  int device;
  int kind;  // 0 roots, 2 + rate_bits coset scale
  u32 log;
  u64 shift;  // coset shift of a scale table (0 for root tables)
  bool operator<(const Key &o) const {
    if (device != o.device) return device < o.device;
    if (kind != o.kind) return kind < o.kind;
    if (log != o.log) return log < o.log;
    return shift < o.shift;
  }
};
std::mutex g_mu;
std::map<Key, u64 *> g_tables;

Status get_table(int kind, u32 log, u64 base, cudaStream_t st, const u64 **out) {
  int dev = 0;
  MP2_CUDA(cudaGetDevice(&dev));
  Key key = {dev, kind, log, 0};
  std::lock_guard<std::mutex> lock(g_mu);
  auto it = g_tables.find(key);
  if (it != g_tables.end()) {
    *out = it->second;
    return "";
  }
  size_t count = (size_t)1 << log;
  u64 *d = nullptr;
  MP2_CUDA(cudaMalloc(&d, sizeof(u64) * count));
  Status built = [&]() -> Status {
    k_fill_powers<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(d, base, count);
    MP2_LAUNCH_CHECK();
    // other threads (other streams) may pick the table up from the cache right away
    MP2_CUDA(cudaStreamSynchronize(st));
    return "";
  }();
  if (!built.empty()) {
    cudaFree(d);
    return built;
  }
  g_tables[key] = d;
  *out = d;
  return "";
}
}  // namespace

Status table_roots(u32 log_t, cudaStream_t st, const u64 **out) {
  return get_table(0, log_t, h_root_of_unity(log_t), st, out);
}
Status table_coset_scale(u32 log_n, u32 rate_bits, u64 shift, cudaStream_t st, const u64 **out) {
  int dev = 0;
  MP2_CUDA(cudaGetDevice(&dev));
  Key key = {dev, 2 + (int)rate_bits, log_n, shift};
  std::lock_guard<std::mutex> lock(g_mu);
  auto it = g_tables.find(key);
  if (it != g_tables.end()) {
    *out = it->second;
    return "";
  }
  const size_t cosets = (size_t)1 << rate_bits, count = cosets << log_n;
  std::vector<u64> bases(cosets);
  const u64 wN = h_root_of_unity(log_n + rate_bits);
  u64 wk = 1;
  for (size_t k = 0; k < cosets; k++) {
    bases[k] = h_mul(shift, wk);
    wk = h_mul(wk, wN);
  }
  u64 *d = nullptr, *d_bases = nullptr;
  MP2_CUDA(cudaMalloc(&d, sizeof(u64) * count));
  Status built = [&]() -> Status {
    MP2_CUDA(cudaMalloc(&d_bases, sizeof(u64) * cosets));
    MP2_CUDA(cudaMemcpyAsync(d_bases, bases.data(), sizeof(u64) * cosets, cudaMemcpyHostToDevice, st));
    k_fill_coset_scale<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(d, d_bases, log_n, count);
    MP2_LAUNCH_CHECK();
    MP2_CUDA(cudaStreamSynchronize(st));
    return "";
  }();
  if (d_bases) cudaFree(d_bases);
  if (!built.empty()) {
    cudaFree(d);
    return built;
  }
  g_tables[key] = d;
  *out = d;
  return "";
}

// mp2gpu_trim: drop the cached tables of the current device (rebuilt on demand).  Callers must not have transforms in
// flight on other threads.
void table_cache_clear() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return;
  std::lock_guard<std::mutex> lock(g_mu);
  for (auto it = g_tables.begin(); it != g_tables.end();) {
    if (it->first.device == dev) {
      cudaFree(it->second);
      it = g_tables.erase(it);
    } else {
      ++it;
    }
  }
}

}  // namespace mp2
