// Measurement hooks of libmp2gpu.so: per-kernel CUDA-event timing (bench.py's roofline leg) and a
// live integer-pipe peak probe (the Poseidon roofline denominator).
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/mp2gpu.h"
#include "internal.h"

namespace mp2 {
std::atomic<int> g_profile_on{0};

namespace {
struct Pending {
  const char *name;
  cudaEvent_t e0, e1;
};
std::mutex g_prof_mu;
std::vector<Pending> g_pending;
std::vector<cudaEvent_t> g_free_events;

cudaEvent_t get_event() {
  if (!g_free_events.empty()) {
    cudaEvent_t e = g_free_events.back();
    g_free_events.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

void prof_record(const char *name, cudaStream_t st, bool begin) {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  if (begin) {
    Pending p = {name, get_event(), get_event()};
    cudaEventRecord(p.e0, st);
    g_pending.push_back(p);
  } else {
    for (size_t i = g_pending.size(); i-- > 0;)
      if (g_pending[i].name == name) {
        cudaEventRecord(g_pending[i].e1, st);
        break;
      }
  }
}

// ---- integer-pipe probe: dependent-free chains of 32-bit IMAD, one CTA of 1024 threads per SM -----
__global__ void __launch_bounds__(1024) k_imad_probe(unsigned *out, unsigned long long *cycles, unsigned seed, int iters) {
  unsigned x[8];
  unsigned b = threadIdx.x * 2654435761u + seed;
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = b + i * 77u;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(x[(i + 1) % 8]), "r"(b));
  }
  long long t1 = clock64();
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

}  // namespace mp2

using namespace mp2;

static const char *dup_err(const std::string &s) {
  if (s.empty()) return nullptr;
  char *p = (char *)malloc(s.size() + 1);
  if (p) memcpy(p, s.c_str(), s.size() + 1);
  return p;
}

extern "C" {

const char *mp2gpu_profile_enable(int on) {
  g_profile_on.store(on ? 1 : 0);
  return nullptr;
}

// Drains the recorded (kernel, start, stop) event pairs: synchronises the device, then writes one
// line per kernel name, "name launches total_ms\n", into buf (NUL terminated, truncated to buf_len).
const char *mp2gpu_profile_report(char *buf, size_t buf_len) {
  if (!buf || buf_len == 0) return dup_err("null report buffer");
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return dup_err(std::string("cudaDeviceSynchronize: ") + cudaGetErrorString(e));
  std::map<std::string, std::pair<long, double>> acc;
  {
    std::lock_guard<std::mutex> lock(g_prof_mu);
    for (auto &p : g_pending) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) {
        auto &a = acc[p.name];
        a.first += 1;
        a.second += ms;
      }
      g_free_events.push_back(p.e0);
      g_free_events.push_back(p.e1);
    }
    g_pending.clear();
  }
  std::string out;
  for (auto &kv : acc) out += kv.first + " " + std::to_string(kv.second.first) + " " + std::to_string(kv.second.second) + "\n";
  size_t n = out.size() < buf_len - 1 ? out.size() : buf_len - 1;
  memcpy(buf, out.data(), n);
  buf[n] = 0;
  return nullptr;
}

// Measured 32-bit IMAD issue rate of the current device: thread-instructions per clock per SM
// (clock64 deltas with one full CTA per SM), the SM clock held during the probe (cycles / CUDA-event
// time) and the product, in T IMAD/s.
const char *mp2gpu_debug_int_pipe_peak(double *imad_per_clk_per_sm, double *sm_clock_mhz, double *t_imad_per_s) {
  int dev = 0, nsm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return dup_err("no usable CUDA device");
  const int iters = 1 << 15, threads = 1024;
  unsigned *out = nullptr;
  unsigned long long *cyc = nullptr;
  if (cudaMalloc(&out, sizeof(unsigned) * nsm * threads) != cudaSuccess || cudaMalloc(&cyc, sizeof(unsigned long long) * nsm) != cudaSuccess)
    return dup_err("cudaMalloc failed in int-pipe probe");
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best_rate = 0, best_clock = 0, best_t = 0;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    k_imad_probe<<<nsm, threads>>>(out, cyc, rep, iters);
    cudaEventRecord(e1);
    if (cudaDeviceSynchronize() != cudaSuccess) break;
    count_launch();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<unsigned long long> h(nsm);
    cudaMemcpy(h.data(), cyc, sizeof(unsigned long long) * nsm, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < nsm; i++) avg += (double)h[i];
    avg /= nsm;
    double instr_sm = (double)threads * iters * 8;
    double rate = instr_sm / avg, clock_mhz = avg / (ms * 1e-3) / 1e6, t = instr_sm * nsm / (ms * 1e-3) / 1e12;
    if (rep > 0 && t > best_t) {
      best_rate = rate;
      best_clock = clock_mhz;
      best_t = t;
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  cudaFree(cyc);
  if (imad_per_clk_per_sm) *imad_per_clk_per_sm = best_rate;
  if (sm_clock_mhz) *sm_clock_mhz = best_clock;
  if (t_imad_per_s) *t_imad_per_s = best_t;
  return nullptr;
}

}  // extern "C"
