// The permutation argument's running products on the device: plonky2 0.2.2
// `wires_permutation_partial_products_and_zs` / `all_wires_permutation_partial_products` (plonk/prover.rs) followed by
// the prover's second commitment `PolynomialBatch::from_values(zs_partial_products, ...)` -- step 3 of every
// `circuit_data.prove(pw)` of the reference (recursion-framework/src/circuit_builder.rs:308,
// .../universal_verifier_gadget/wrap_circuit.rs:143), the values that FEED the hot path's second call.  With it the
// whole prove() runs from the wire values on without a host round trip of field data.
//
// Inputs are the two device-resident batches (constants+sigmas, wires): the values on the subgroup H are the forward
// NTT of the coefficient columns they keep in HBM (the routed wires and the sigmas only).
//   k_pp_chunks : one thread per (row, challenge): the R quotients (w + beta k_j x + gamma) / (w + beta sigma_j + gamma)
//                 as products over chunks of 2^qb -- numerators and denominators multiplied separately, ONE inversion
//                 per thread for all chunk denominators (Montgomery's trick); exact field arithmetic, so the values
//                 equal plonky2's product of per-wire quotients
//   k_pp_scan   : Z(g^i) = product of the row totals before row i (one CTA per challenge: serial runs per thread,
//                 shared-memory scan of the run products)
//   k_pp_write  : partial products Z(x) * chunk_0 * ... * chunk_k and Z itself, columns [Z_0.., pp(ch 0).., pp(ch 1)..]
// A zero denominator (plonky2's batch_multiplicative_inverse panics on it) is reported as an error.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/mp2gpu.h"
#include "internal.h"
#include "gl.cuh"

namespace mp2 {

Status commit_device_columns(const u64 *d_cols, size_t ncols, u32 n_log, u32 rate_bits, u32 cap_height, u32 hash_kind,
                             int from_coeffs, uint64_t *const *coeffs_out, uint64_t *leaves_out, uint64_t *digests_out,
                             uint64_t *cap_out, mp2gpu_batch **handle_out, cudaStream_t st);

namespace {

constexpr u32 kMaxChunks = 32, kMaxChallenges = 4;

struct PParams {
  u32 n_log, R, md, nchunks, nch;
  const u64 *wv, *sg;   // forward NTTs, leaf (bit-reversed) order: column j at + j * n
  const u64 *roots;     // w_n^m
  const u64 *k_is;      // 7^j
  u64 betas[kMaxChallenges], gammas[kMaxChallenges];
  u64 *cp;              // (nch, nchunks, n) chunk products; natural row order
  u64 *tot;             // (nch, n) row totals, then in place the exclusive prefix products = Z
  u32 *err;
};

GL_DEV u64 gl_inv(u64 a) { return gl_pow(a, GL_P - 2); }

__global__ void __launch_bounds__(128) k_pp_chunks(const __grid_constant__ PParams P) {
  const size_t n = (size_t)1 << P.n_log;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * P.nch) return;
  const u32 c = (u32)(t >> P.n_log);
  const u32 i = (u32)(t & (n - 1));
  const u32 leaf = brev_bits(i, P.n_log);
  const u64 beta = P.betas[c], gamma = P.gammas[c];
  const u64 bx = gl_mul(beta, P.roots[i]);
  u64 num[kMaxChunks], den[kMaxChunks];
  for (u32 k = 0; k < P.nchunks; k++) {
    u64 pn = 1, pd = 1;
    const u32 j1 = min(P.R, (k + 1) * P.md);
    for (u32 j = k * P.md; j < j1; j++) {
      const u64 w = gl_add(P.wv[(size_t)j * n + leaf], gamma);
      pn = gl_mul(pn, gl_mul_add(bx, P.k_is[j], w));
      pd = gl_mul(pd, gl_mul_add(beta, P.sg[(size_t)j * n + leaf], w));
    }
    num[k] = pn;
    den[k] = pd;
  }
  // one inversion for all chunk denominators: prefix products, invert the total, walk back
  u64 pre[kMaxChunks];
  u64 acc = 1;
  for (u32 k = 0; k < P.nchunks; k++) {
    pre[k] = acc;
    acc = gl_mul(acc, den[k]);
  }
  if (gl_canon(acc) == 0) atomicOr(P.err, 1u);
  u64 inv = gl_inv(acc), total = 1;
  for (u32 k = P.nchunks; k-- > 0;) {
    const u64 dinv = gl_mul(inv, pre[k]);
    inv = gl_mul(inv, den[k]);
    num[k] = gl_mul(num[k], dinv);
  }
  for (u32 k = 0; k < P.nchunks; k++) {
    P.cp[((size_t)c * P.nchunks + k) * n + i] = num[k];
    total = gl_mul(total, num[k]);
  }
  P.tot[(size_t)c * n + i] = total;
}

// exclusive prefix product of tot[c][*], in place; blockDim.x = 1024, one CTA per challenge
__global__ void __launch_bounds__(1024) k_pp_scan(u64 *tot, u32 n_log) {
  __shared__ u64 run[1024];
  const size_t n = (size_t)1 << n_log;
  u64 *v = tot + (size_t)blockIdx.x * n;
  const size_t per = (n + 1023) >> 10;
  const size_t b = threadIdx.x * per, e = min(n, b + per);
  u64 acc = 1;
  for (size_t i = b; i < e; i++) acc = gl_mul(acc, v[i]);
  run[threadIdx.x] = acc;
  __syncthreads();
  for (u32 d = 1; d < 1024; d <<= 1) {  // inclusive scan of the run products
    const u64 o = threadIdx.x >= d ? run[threadIdx.x - d] : 1;
    __syncthreads();
    run[threadIdx.x] = gl_mul(run[threadIdx.x], o);
    __syncthreads();
  }
  acc = threadIdx.x ? run[threadIdx.x - 1] : 1;
  for (size_t i = b; i < e; i++) {
    const u64 x = v[i];
    v[i] = acc;
    acc = gl_mul(acc, x);
  }
}

// out: columns of n values, [Z_0..Z_{nch-1}, pp(ch 0) (npp columns), pp(ch 1), ...], canonical
__global__ void __launch_bounds__(256) k_pp_write(const __grid_constant__ PParams P, u64 *__restrict__ out) {
  const size_t n = (size_t)1 << P.n_log;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * P.nch) return;
  const u32 c = (u32)(t >> P.n_log);
  const size_t i = t & (n - 1);
  const u32 npp = P.nchunks - 1;
  u64 acc = P.tot[(size_t)c * n + i];
  out[(size_t)c * n + i] = gl_canon(acc);
  for (u32 k = 0; k < npp; k++) {
    acc = gl_mul(acc, P.cp[((size_t)c * P.nchunks + k) * n + i]);
    out[((size_t)P.nch + (size_t)c * npp + k) * n + i] = gl_canon(acc);
  }
}

}  // namespace

// d_out: nch * (1 + npp) columns of n values on the device (natural row order), on stream st
Status partial_products_and_zs_dev(const mp2gpu_circuit *ci, const mp2gpu_batch *bcs, const mp2gpu_batch *bwi,
                                   const uint64_t *betas, const uint64_t *gammas, u64 *d_out, cudaStream_t st) {
  const u32 n_log = ci->degree_bits, qb = ci->quotient_degree_bits, nch = ci->num_challenges, R = ci->num_routed_wires;
  if (nch == 0 || nch > kMaxChallenges) return "partial_products: num_challenges must be 1.." + std::to_string(kMaxChallenges);
  if (qb == 0 || qb > 4) return "partial_products: quotient_degree_bits must be 1..4";
  if (R == 0 || R > ci->num_wires) return "partial_products: num_routed_wires out of range";
  const u32 md = 1u << qb, nchunks = (R + md - 1) / md;
  if (nchunks > kMaxChunks) return "partial_products: more than " + std::to_string(kMaxChunks) + " chunks of routed wires";
  if (bcs->n_log != n_log || bwi->n_log != n_log) return "partial_products: batch degree differs from the circuit's degree_bits";
  if (bcs->device != bwi->device) return "partial_products: the batches live on different devices";
  if (bcs->ncols != ci->num_constants + R) return "partial_products: constants_sigmas batch must hold num_constants + num_routed_wires columns";
  if (bwi->ncols != ci->num_wires) return "partial_products: wires batch must hold num_wires columns";
  if (!bcs->coeffs || !bwi->coeffs) return "partial_products: batches hold no coefficients";
  const size_t n = (size_t)1 << n_log;
  DevBuf d_vals, d_cp, d_tot, d_tab;
  MP2_TRY(d_vals.alloc(2 * (size_t)R * n, st));
  MP2_TRY(d_cp.alloc((size_t)nch * nchunks * n, st));
  MP2_TRY(d_tot.alloc((size_t)nch * n + 1, st));
  MP2_TRY(d_tab.alloc(R, st));
  // values on H (leaf order): the rate-0 "coset" transform with shift 1
  MP2_TRY(ntt_coset_lde(bwi->coeffs, n, d_vals.p, n, R, n_log, 0, 0, 0, st, nullptr, 1));
  MP2_TRY(ntt_coset_lde(bcs->coeffs + (size_t)ci->num_constants * n, n, d_vals.p + (size_t)R * n, n, R, n_log, 0, 0, 0, st,
                        nullptr, 1));
  std::vector<u64> k_is(R);
  u64 k = 1;
  for (u32 j = 0; j < R; j++) {  // get_unique_coset_shifts: powers of the multiplicative generator
    k_is[j] = k;
    k = h_mul(k, kCosetShift);
  }
  MP2_CUDA(cudaMemcpyAsync(d_tab.p, k_is.data(), R * sizeof(u64), cudaMemcpyHostToDevice, st));
  u32 *d_err = (u32 *)(d_tot.p + (size_t)nch * n);
  MP2_CUDA(cudaMemsetAsync(d_err, 0, sizeof(u64), st));
  PParams P = {};
  P.n_log = n_log; P.R = R; P.md = md; P.nchunks = nchunks; P.nch = nch;
  P.wv = d_vals.p; P.sg = d_vals.p + (size_t)R * n;
  MP2_TRY(table_roots(n_log, st, &P.roots));
  P.k_is = d_tab.p;
  for (u32 c = 0; c < nch; c++) { P.betas[c] = betas[c] % kP; P.gammas[c] = gammas[c] % kP; }
  P.cp = d_cp.p; P.tot = d_tot.p; P.err = d_err;
  const size_t threads = n * nch;
  { ProfScope _p("k_pp_chunks", st); k_pp_chunks<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(P); }
  MP2_LAUNCH_CHECK();
  { ProfScope _p("k_pp_scan", st); k_pp_scan<<<nch, 1024, 0, st>>>(d_tot.p, n_log); }
  MP2_LAUNCH_CHECK();
  { ProfScope _p("k_pp_write", st); k_pp_write<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(P, d_out); }
  MP2_LAUNCH_CHECK();
  u32 err = 0;
  MP2_CUDA(cudaMemcpyAsync(&err, d_err, sizeof(u32), cudaMemcpyDeviceToHost, st));
  MP2_CUDA(cudaStreamSynchronize(st));  // also keeps k_is alive past its upload
  if (err) return "partial_products: zero denominator w + beta*sigma + gamma (plonky2's batch inversion panics here)";
  return "";
}

Status partial_products_and_zs(const mp2gpu_circuit *ci, const mp2gpu_batch *bcs, const mp2gpu_batch *bwi,
                               const uint64_t *betas, const uint64_t *gammas, u32 rate_bits, u32 cap_height, u32 hash_kind,
                               uint64_t *const *values_out, uint64_t *leaves_out, uint64_t *digests_out, uint64_t *cap_out,
                               mp2gpu_batch **handle_out) {
  if (!ci || !bcs || !bwi || !betas || !gammas || !cap_out) return "partial_products: null argument";
  DeviceScope scope(bcs->device);
  cudaStream_t st;
  MP2_TRY(ctx_stream(&st));
  const u32 md = 1u << ci->quotient_degree_bits;
  if (ci->quotient_degree_bits > 4 || ci->num_routed_wires == 0) return "partial_products: bad circuit descriptor";
  const size_t ncols = (size_t)ci->num_challenges * ((ci->num_routed_wires + md - 1) / md);
  const size_t n = (size_t)1 << ci->degree_bits;
  DevBuf d_out;
  MP2_TRY(d_out.alloc(ncols * n, st));
  MP2_TRY(partial_products_and_zs_dev(ci, bcs, bwi, betas, gammas, d_out.p, st));
  if (values_out) MP2_TRY(copy_columns_d2h(values_out, d_out.p, ncols, n, st));
  return commit_device_columns(d_out.p, ncols, ci->degree_bits, rate_bits, cap_height, hash_kind, 0, nullptr, leaves_out,
                               digests_out, cap_out, handle_out, st);
}

}  // namespace mp2

extern "C" const char *mp2gpu_partial_products_and_zs(const mp2gpu_circuit *circuit, const mp2gpu_batch *constants_sigmas,
                                                      const mp2gpu_batch *wires, const uint64_t *betas,
                                                      const uint64_t *gammas, uint32_t rate_bits, uint32_t cap_height,
                                                      uint32_t hash_kind, uint64_t *const *values_out, uint64_t *leaves_out,
                                                      uint64_t *digests_out, uint64_t *cap_out,
                                                      mp2gpu_batch **zs_partial_products_batch_out) {
  mp2::Status s;
  try {
    s = mp2::partial_products_and_zs(circuit, constants_sigmas, wires, betas, gammas, rate_bits, cap_height, hash_kind,
                                     values_out, leaves_out, digests_out, cap_out, zs_partial_products_batch_out);
  } catch (const std::exception &e) {
    s = std::string("exception: ") + e.what();
  }
  if (s.empty()) return nullptr;
  char *m = (char *)malloc(s.size() + 1);
  if (m) memcpy(m, s.c_str(), s.size() + 1);
  return m;
}
