// Can the FP64 pipe run beside the two integer pipes on B200?  Self-dependent chains only (each chain reads
// itself and two loop-invariant registers), no asm volatile: ptxas is free to interleave the three kinds.
// ND/NI/NL = number of DFMA / IMAD / LOP3 chains per thread.  Reported: thread-instructions per clock per SM
// per kind, from clock64() on one 1024-thread CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_mix tools/pipe_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
#define ITERS 8192

template <int ND, int NI, int NL, int NA>
__global__ void __launch_bounds__(1024) k(u64 *out, u64 *cycles, u32 seed, double dc, double de, u32 b, u32 c) {
  double d[ND > 0 ? ND : 1];
  u32 x[NI > 0 ? NI : 1], y[NL > 0 ? NL : 1], z[NA > 0 ? NA : 1];
#pragma unroll
  for (int i = 0; i < ND; i++) d[i] = 1.0 + 1e-9 * (double)(threadIdx.x + i + seed);
#pragma unroll
  for (int i = 0; i < NI; i++) x[i] = threadIdx.x * 2654435761u + i + seed;
#pragma unroll
  for (int i = 0; i < NL; i++) y[i] = threadIdx.x * 40503u + i + seed;
#pragma unroll
  for (int i = 0; i < NA; i++) z[i] = threadIdx.x * 9973u + i + seed;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int rep = 0; rep < 4; rep++) {
#pragma unroll
      for (int i = 0; i < ND; i++) asm("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dc), "d"(de));
#pragma unroll
      for (int i = 0; i < NI; i++) asm("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(c));
#pragma unroll
      for (int i = 0; i < NL; i++) asm("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(b), "r"(c));
#pragma unroll
      for (int i = 0; i < NA; i++) asm("add.f64 %0, %0, %1;" : "+d"(d[i % (ND > 0 ? ND : 1)]) : "d"(de));
    }
  }
  long long t1 = clock64();
  u64 s = 0;
#pragma unroll
  for (int i = 0; i < ND; i++) s += (u64)__double_as_longlong(d[i]);
#pragma unroll
  for (int i = 0; i < NI; i++) s += x[i];
#pragma unroll
  for (int i = 0; i < NL; i++) s += y[i];
#pragma unroll
  for (int i = 0; i < NA; i++) s += z[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = (u64)(t1 - t0);
}

template <int ND, int NI, int NL>
void run(int nsm) {
  u64 *out, *cyc;
  cudaMalloc(&out, sizeof(u64) * nsm * 1024);
  cudaMalloc(&cyc, sizeof(u64) * nsm);
  for (int r = 0; r < 2; r++) k<ND, NI, NL, 0><<<nsm, 1024>>>(out, cyc, r, 1.0000001, 1e-9, 0x9e3779b9u, 12345u);
  cudaDeviceSynchronize();
  u64 *h = new u64[nsm];
  cudaMemcpy(h, cyc, sizeof(u64) * nsm, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < nsm; i++) avg += h[i];
  avg /= nsm;
  const double per = 1024.0 * ITERS * 4 / avg;
  printf("DFMA x%-2d IMAD x%-2d LOP3 x%-2d : DFMA %6.1f  IMAD %6.1f  LOP3 %6.1f  total %6.1f thread-instr/clk/SM\n", ND, NI, NL,
         per * ND, per * NI, per * NL, per * (ND + NI + NL));
  cudaFree(out); cudaFree(cyc); delete[] h;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int nsm = p.multiProcessorCount;
  printf("device %s, %d SMs\n", p.name, nsm);
  run<8, 0, 0>(nsm);
  run<0, 8, 0>(nsm);
  run<0, 0, 8>(nsm);
  run<0, 8, 8>(nsm);
  run<8, 8, 0>(nsm);
  run<8, 0, 8>(nsm);
  run<8, 8, 8>(nsm);
  run<4, 8, 8>(nsm);
  run<2, 8, 8>(nsm);
  run<4, 6, 6>(nsm);
  run<8, 4, 4>(nsm);
  return 0;
}
