// Integer-pipe issue-rate microbenchmark for the Poseidon roofline denominator (SURVEY.md 8(d),
// Appendix C.4): measures thread-instructions per clock per SM for the instruction kinds the
// Goldilocks arithmetic is made of.  Cycles come from clock64() inside the kernel, so the result
// is independent of the SM clock held during the run; wall time (CUDA events) gives the clock.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o intpipe_peak tools/intpipe_peak.cu && ./intpipe_peak
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
typedef unsigned int u32;
#define CHAINS 8
#define ITERS 16384

template <int MODE>
__global__ void __launch_bounds__(1024) k(u64 *out, u64 *cycles, u32 seed) {
  u64 acc[CHAINS];
  u32 a = threadIdx.x * 2654435761u + seed, b = a ^ 0x9e3779b9u;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) acc[i] = (u64)a * (i + 3) + b;
  u32 x[CHAINS], y[CHAINS];
  double dd[CHAINS];
  const double dc = 1.0000001;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) dd[i] = 1.0 + 1e-9 * (double)(threadIdx.x + i);
#pragma unroll
  for (int i = 0; i < CHAINS; i++) { x[i] = a + i; y[i] = b - i; }
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
      const int j = (i + 1) % CHAINS;  // operands come from a neighbouring chain: nothing is loop invariant
      if (MODE == 0) {  // IMAD.WIDE.U32
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"((u32)acc[i]), "r"(b));
      } else if (MODE == 13) {  // IMAD.WIDE without accumulate (mul.wide.u32)
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(acc[i]) : "r"((u32)(acc[i] >> 7)), "r"(b));
      } else if (MODE == 14) {  // IMAD.WIDE, multiplicands from 32-bit chains (independent of the accumulator)
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(x[i]), "r"(y[j]));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(x[j]), "r"(b));
      } else if (MODE == 1) {  // IMAD (32-bit lo)
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(x[j]), "r"(b));
      } else if (MODE == 5) {  // IMAD.HI
        asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(x[j]), "r"(b));
      } else if (MODE == 2) {  // LOP3 + SHF (alu pipe only)
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(x[j]), "r"(y[i]));
        asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(y[i]) : "r"(x[i]));
      } else if (MODE == 7) {  // SHF only
        asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(x[i]) : "r"(x[j]));
      } else if (MODE == 11) {  // LOP3 only
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(x[j]), "r"(y[i]));
      } else if (MODE == 4) {  // 64-bit add with carry, data dependent
        asm volatile("{add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;}" : "+r"(x[i]), "+r"(y[i]) : "r"(x[j]), "r"(y[j]));
      } else if (MODE == 3) {  // 1 IMAD.WIDE : 1 LOP3
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(x[j]), "r"(b));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"((u32)acc[i]), "r"(y[i]));
      } else if (MODE == 6) {  // 2 IMAD.WIDE : 1 LOP3
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(x[j]), "r"(b));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(y[j]), "r"(a));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"((u32)acc[i]), "r"(y[i]));
      } else if (MODE == 9) {  // 1 IMAD.WIDE : 2 alu
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(x[j]), "r"(b));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"((u32)acc[i]), "r"(y[i]));
        asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(y[i]) : "r"(x[i]));
      } else if (MODE == 10) {  // 1 IMAD.WIDE : 3 alu
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(x[j]), "r"(b));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"((u32)acc[i]), "r"(y[i]));
        asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(y[i]) : "r"(x[i]));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(y[i]) : "r"(x[j]), "r"(a));
      } else if (MODE == 8) {  // 1 IMAD (lo) : 1 LOP3
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(x[j]), "r"(b));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(x[i]), "r"(y[j]));
      } else if (MODE == 20) {  // DFMA
        asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dd[i]) : "d"(dd[j]), "d"(dc));
      } else if (MODE == 21) {  // DADD
        asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(dd[i]) : "d"(dd[j]));
      } else if (MODE == 22) {  // 1 DFMA : 1 LOP3 : 1 IMAD  (three pipes)
        asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dd[i]) : "d"(dd[j]), "d"(dc));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(x[j]), "r"(y[i]));
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(y[i]) : "r"(y[j]), "r"(b));
      } else if (MODE == 23) {  // 1 DFMA : 2 LOP3
        asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dd[i]) : "d"(dd[j]), "d"(dc));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(x[j]), "r"(y[i]));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(y[i]) : "r"(y[j]), "r"(b));
      } else if (MODE == 12) {  // IADD3 three-operand adds
        asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(x[i]) : "r"(x[j]), "r"(y[i]));
      }
    }
  }
  long long t1 = clock64();
  u64 s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) s += acc[i] + x[i] + y[i] + (u64)__double_as_longlong(dd[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = (u64)(t1 - t0);
}

template <int MODE>
void run(const char *name, double instr_per_chain_iter, int nsm) {
  int blocks = nsm, threads = 1024;  // exactly one CTA per SM: clock64() deltas are per-SM busy cycles
  u64 *out, *cyc;
  cudaMalloc(&out, sizeof(u64) * blocks * threads);
  cudaMalloc(&cyc, sizeof(u64) * blocks);
  k<MODE><<<blocks, threads>>>(out, cyc, 1);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, cyc, 2);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  u64 *h = new u64[blocks];
  cudaMemcpy(h, cyc, sizeof(u64) * blocks, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
  double instr_sm = 1.0 * threads * (double)ITERS * CHAINS * instr_per_chain_iter;
  double total = (double)blocks * threads * ITERS * CHAINS * instr_per_chain_iter;
  printf("%-44s %7.1f thread-instr/clk/SM   %.2f T instr/s  (%.3f ms, implied SM clock %.0f MHz)\n", name,
         instr_sm / avg, total / (ms * 1e-3) / 1e12, ms, avg / (ms * 1e-3) / 1e6);
  cudaFree(out); cudaFree(cyc); delete[] h;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int nsm = p.multiProcessorCount;
  printf("device %s, %d SMs\n", p.name, nsm);
  run<20>("DFMA", 1, nsm);
  run<21>("DADD", 1, nsm);
  run<22>("1 DFMA : 1 LOP3 : 1 IMAD", 3, nsm);
  run<23>("1 DFMA : 2 LOP3", 3, nsm);
  run<0>("IMAD.WIDE.U32 (acc-dependent)", 1, nsm);
  run<13>("mul.wide.u32", 1, nsm);
  run<14>("IMAD.WIDE (indep. multiplicands) + LOP3", 2, nsm);
  run<1>("IMAD (lo)", 1, nsm);
  run<5>("IMAD.HI", 1, nsm);
  run<11>("LOP3", 1, nsm);
  run<7>("SHF", 1, nsm);
  run<2>("LOP3 + SHF", 2, nsm);
  run<12>("IADD3 (3-operand)", 1, nsm);
  run<4>("64-bit add (add.cc + addc)", 2, nsm);
  run<8>("1 IMAD : 1 LOP3", 2, nsm);
  run<3>("1 IMAD.WIDE : 1 LOP3", 2, nsm);
  run<6>("2 IMAD.WIDE : 1 LOP3", 3, nsm);
  run<9>("1 IMAD.WIDE : 2 alu", 3, nsm);
  run<10>("1 IMAD.WIDE : 3 alu", 4, nsm);
  return 0;
}
