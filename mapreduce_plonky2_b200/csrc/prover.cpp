// plonky2 0.2.2 `prove_with_partition_witness` (plonk/prover.rs) from the point where the witness is known, as ONE
// native call: the orchestration that the reference runs inside every `circuit_data.prove(pw)`
// (recursion-framework/src/circuit_builder.rs:308, .../universal_verifier_gadget/wrap_circuit.rs:143), with every
// data-path step on the device and the Fiat-Shamir transcript on the host (it is host code in the reference too):
//
//   wires commitment -> challenger(circuit digest, public-inputs hash, wires cap) -> betas, gammas
//   -> Z / partial products + their commitment (permutation.cu) -> alphas
//   -> quotient polynomials + their commitment (quotient.cu) -> zeta
//   -> OpeningSet at zeta and g*zeta (mp2gpu_batch_eval) -> prove_openings: alpha-batched quotient, FRI commit
//      phase, proof-of-work grind, query rounds (fri.cu, merkle.cu)
//   -> bincode(ProofWithPublicInputs) (mp2-common/src/proof.rs:86-90 `serialize_proof`; layout as in wire.py)
//
// The Python mirror of the same sequence is prover.py (tests compare the two byte for byte); this file exists so
// that a prover thread makes one FFI call per proof instead of ~40 and holds no interpreter lock in between.
// Host code only: it calls the library's own C ABI entry points; no field arithmetic beyond g*zeta and the transcript.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mp2gpu.h"
#include "internal.h"

namespace mp2 {
Status partial_products_and_zs(const mp2gpu_circuit *ci, const mp2gpu_batch *bcs, const mp2gpu_batch *bwi,
                               const uint64_t *betas, const uint64_t *gammas, u32 rate_bits, u32 cap_height, u32 hash_kind,
                               uint64_t *const *values_out, uint64_t *leaves_out, uint64_t *digests_out, uint64_t *cap_out,
                               mp2gpu_batch **handle_out);
Status quotient_polys(const mp2gpu_circuit *ci, const mp2gpu_batch *bcs, const mp2gpu_batch *bwi, const mp2gpu_batch *bzp,
                      const uint64_t *betas, const uint64_t *gammas, const uint64_t *alphas, const uint64_t *pi_hash,
                      u32 rate_bits, u32 cap_height, u32 hash_kind, uint64_t *const *chunks_out, uint64_t *leaves_out,
                      uint64_t *digests_out, uint64_t *cap_out, mp2gpu_batch **handle_out);

namespace {

// error string of a C ABI call -> Status (and release it)
Status own(const char *e) {
  if (!e) return "";
  std::string s(e);
  free((void *)e);
  return s;
}

// `Challenger<F, H>` (plonky2 iop/challenger.rs): overwrite-mode duplex sponge, challenges popped from the END of
// the squeezed rate
struct Challenger {
  uint64_t state[12] = {0}, in[8] = {0}, out[8] = {0};
  u32 in_len = 0, out_len = 0, kind;
  Status err;
  explicit Challenger(u32 k) : kind(k) {}
  void duplexing() {
    for (u32 i = 0; i < in_len; i++) state[i] = in[i];
    in_len = 0;
    Status s = own(mp2gpu_transcript_permute(state, kind));
    if (!s.empty() && err.empty()) err = s;
    for (u32 i = 0; i < 8; i++) out[i] = state[i];
    out_len = 8;
  }
  void observe_element(uint64_t x) {
    out_len = 0;
    in[in_len++] = x % kP;
    if (in_len == 8) duplexing();
  }
  void observe_elements(const uint64_t *xs, size_t n) {
    for (size_t i = 0; i < n; i++) observe_element(xs[i]);
  }
  uint64_t get_challenge() {
    if (in_len || !out_len) duplexing();
    return out[--out_len];
  }
  void get_n_challenges(uint64_t *dst, u32 n) {
    for (u32 i = 0; i < n; i++) dst[i] = get_challenge();
  }
};

struct Writer {  // bincode 1.x default options: little-endian fixed-width integers, Vec = uint64_t length + elements
  std::vector<uint8_t> b;
  void u64le(uint64_t v) {
    for (int i = 0; i < 8; i++) b.push_back((uint8_t)(v >> (8 * i)));
  }
  void felts(const uint64_t *x, size_t n) {
    for (size_t i = 0; i < n; i++) u64le(x[i] >= kP ? x[i] - kP : x[i]);
  }
  void vec(const uint64_t *x, size_t count, size_t width) {  // Vec of `count` items of `width` field elements
    u64le(count);
    felts(x, count * width);
  }
};

struct Handles {  // everything prove() creates on the device, released on every exit path
  mp2gpu_batch *wires = nullptr, *zs = nullptr, *quotient = nullptr;
  mp2gpu_fri *fri = nullptr;
  ~Handles() {
    if (fri) mp2gpu_fri_free(fri);
    for (mp2gpu_batch *b : {wires, zs, quotient})
      if (b) mp2gpu_batch_free(b);
  }
};

inline u32 clz64(uint64_t x) { return x ? (u32)__builtin_clzll(x) : 64u; }

Status prove(const mp2gpu_circuit *ci, const mp2gpu_prove_config *cf, const mp2gpu_batch *bcs, const uint64_t *digest,
             const uint64_t *const *wires_values, const uint64_t *public_inputs, size_t npi, const uint64_t *pi_hash,
             std::vector<uint8_t> *bytes) {
  if (!ci || !cf || !bcs || !digest || !wires_values || !pi_hash || (npi && !public_inputs)) return "prove: null argument";
  const u32 nch = ci->num_challenges, n_log = ci->degree_bits, R = ci->num_routed_wires, nc = ci->num_constants;
  if (nch == 0 || nch > 4) return "prove: num_challenges must be 1..4";
  if (cf->num_reductions && !cf->reduction_arity_bits) return "prove: null reduction_arity_bits";
  if (bcs->n_log != n_log) return "prove: constants_sigmas batch degree differs from the circuit's degree_bits";
  if (bcs->rate_bits != cf->rate_bits) return "prove: constants_sigmas batch was committed at another rate";
  const u32 kind = cf->hash_kind, rate_bits = cf->rate_bits, cap_height = cf->cap_height;
  const u32 lde_bits = n_log + rate_bits;
  if (cap_height > lde_bits) return "prove: cap_height > log2(LDE size)";
  u32 red = 0;
  for (u32 i = 0; i < cf->num_reductions; i++) red += cf->reduction_arity_bits[i];
  if (red > n_log) return "prove: FRI reductions exceed the degree";
  const size_t ncap = (size_t)1 << cap_height;
  const u32 md = 1u << ci->quotient_degree_bits, npp = (R + md - 1) / md - 1;
  DeviceScope scope(bcs->device);  // the whole proof runs where the circuit's constants/sigmas batch lives
  Handles h;
  std::vector<uint64_t> cap_w(ncap * 4), cap_z(ncap * 4), cap_q(ncap * 4);

  // 1. wires commitment
  MP2_TRY(commit_host(wires_values, ci->num_wires, n_log, rate_bits, cap_height, kind, 0, nullptr, nullptr, nullptr,
                      cap_w.data(), &h.wires));
  Challenger ch(kind);
  ch.observe_elements(digest, 4);
  ch.observe_elements(pi_hash, 4);
  ch.observe_elements(cap_w.data(), cap_w.size());
  uint64_t betas[4], gammas[4], alphas[4];
  ch.get_n_challenges(betas, nch);
  ch.get_n_challenges(gammas, nch);
  // 2. permutation argument
  MP2_TRY(partial_products_and_zs(ci, bcs, h.wires, betas, gammas, rate_bits, cap_height, kind, nullptr, nullptr, nullptr,
                                  cap_z.data(), &h.zs));
  ch.observe_elements(cap_z.data(), cap_z.size());
  ch.get_n_challenges(alphas, nch);
  // 3. quotient
  MP2_TRY(quotient_polys(ci, bcs, h.wires, h.zs, betas, gammas, alphas, pi_hash, rate_bits, cap_height, kind, nullptr, nullptr,
                         nullptr, cap_q.data(), &h.quotient));
  ch.observe_elements(cap_q.data(), cap_q.size());
  uint64_t zeta[2];
  ch.get_n_challenges(zeta, 2);
  // 4. openings at zeta and g * zeta
  const uint64_t g = h_root_of_unity(n_log);
  const uint64_t points[4] = {zeta[0], zeta[1], h_mul(zeta[0] % kP, g), h_mul(zeta[1] % kP, g)};
  const mp2gpu_batch *oracles[4] = {bcs, h.wires, h.zs, h.quotient};
  const size_t cols[4] = {bcs->ncols, h.wires->ncols, h.zs->ncols, h.quotient->ncols};
  const size_t total_cols = cols[0] + cols[1] + cols[2] + cols[3];
  std::vector<uint64_t> at_zeta(total_cols * 2), zs_next((size_t)nch * 2);
  {
    size_t off = 0;
    for (int o = 0; o < 4; o++) {
      const size_t np = o == 2 ? 2 : 1;  // only the Zs are opened at g * zeta as well
      std::vector<uint64_t> ev(np * cols[o] * 2);
      MP2_TRY(own(mp2gpu_batch_eval(oracles[o], points, np, ev.data())));
      memcpy(at_zeta.data() + off * 2, ev.data(), cols[o] * 2 * sizeof(uint64_t));
      if (o == 2) memcpy(zs_next.data(), ev.data() + cols[o] * 2, (size_t)nch * 2 * sizeof(uint64_t));
      off += cols[o];
    }
  }
  ch.observe_elements(at_zeta.data(), at_zeta.size());   // challenger.observe_openings(&openings.to_fri_openings())
  ch.observe_elements(zs_next.data(), zs_next.size());
  // 5. prove_openings
  uint64_t alpha[2];
  ch.get_n_challenges(alpha, 2);
  std::vector<u32> oi, pi;
  for (u32 o = 0; o < 4; o++)
    for (size_t p = 0; p < cols[o]; p++) {
      oi.push_back(o);
      pi.push_back((u32)p);
    }
  for (u32 p = 0; p < nch; p++) {
    oi.push_back(2);
    pi.push_back(p);
  }
  const uint32_t batch_sizes[2] = {(uint32_t)total_cols, nch};
  MP2_TRY(own(mp2gpu_fri_begin_openings(oracles, 4, points, batch_sizes, 2, oi.data(), pi.data(),
                                        alpha, cap_height, kind, nullptr, &h.fri)));
  std::vector<std::vector<uint64_t>> layer_caps;
  for (u32 i = 0; i < cf->num_reductions; i++) {
    std::vector<uint64_t> cap(ncap * 4);
    MP2_TRY(own(mp2gpu_fri_commit_layer(h.fri, cf->reduction_arity_bits[i], cap.data())));
    ch.observe_elements(cap.data(), cap.size());
    uint64_t beta[2];
    ch.get_n_challenges(beta, 2);
    MP2_TRY(own(mp2gpu_fri_fold(h.fri, beta)));
    layer_caps.push_back(std::move(cap));
  }
  size_t final_len = 0;
  MP2_TRY(own(mp2gpu_fri_finish(h.fri, nullptr, &final_len)));
  std::vector<uint64_t> final_poly(final_len * 2 + 2);
  MP2_TRY(own(mp2gpu_fri_finish(h.fri, final_poly.data(), &final_len)));
  ch.observe_elements(final_poly.data(), final_len * 2);
  // fri_proof_of_work: smallest witness
  uint64_t pow_witness = 0;
  {
    uint64_t st[12];
    memcpy(st, ch.state, sizeof(st));
    for (u32 i = 0; i < ch.in_len; i++) st[i] = ch.in[i];
    const u32 min_lz = cf->proof_of_work_bits;  // + (64 - F::order().bits()) = 0
    MP2_TRY(own(mp2gpu_fri_proof_of_work(st, ch.in_len, min_lz, kind, &pow_witness)));
    ch.observe_element(pow_witness);
    const uint64_t response = ch.get_challenge();
    if (clz64(response) < min_lz) return "prove: proof of work response does not have the required leading zeros";
  }
  // query rounds: one gather per tree for all rounds
  const u32 nq = cf->num_query_rounds;
  std::vector<uint64_t> x(nq);
  for (u32 q = 0; q < nq; q++) x[q] = ch.get_challenge() & (((uint64_t)1 << lde_bits) - 1);  // canonical % 2^lde_bits
  if (!ch.err.empty()) return ch.err;
  const u32 h0 = lde_bits - cap_height;
  std::vector<uint64_t> rows[4], sibs[4];
  for (int o = 0; o < 4; o++) {
    rows[o].resize((size_t)nq * cols[o] + 1);
    sibs[o].resize((size_t)nq * h0 * 4 + 1);
    MP2_TRY(own(mp2gpu_batch_open(oracles[o], x.data(), nq, rows[o].data(),
                                  h0 ? sibs[o].data() : nullptr)));
  }
  struct Layer {
    size_t leaf_len, hsib;
    std::vector<uint64_t> leaves, sib;
  };
  std::vector<Layer> layers(cf->num_reductions);
  {
    std::vector<uint64_t> idx = x;
    for (u32 i = 0; i < cf->num_reductions; i++) {
      for (uint64_t &v : idx) v >>= cf->reduction_arity_bits[i];
      size_t nl = 0, ll = 0, nd = 0, ncp = 0;
      MP2_TRY(own(mp2gpu_fri_layer_shape(h.fri, i, &nl, &ll, &nd, &ncp)));
      Layer &L = layers[i];
      L.leaf_len = ll;
      L.hsib = (size_t)(log2_exact(nl) - log2_exact(ncp));
      L.leaves.resize((size_t)nq * ll + 1);
      L.sib.resize((size_t)nq * L.hsib * 4 + 1);
      MP2_TRY(own(mp2gpu_fri_open_layer(h.fri, i, idx.data(), nq, L.leaves.data(),
                                        L.hsib ? L.sib.data() : nullptr)));
    }
  }
  // 6. bincode(ProofWithPublicInputs): Proof { wires_cap, plonk_zs_partial_products_cap, quotient_polys_cap,
  //    openings: OpeningSet, opening_proof: FriProof }, public_inputs
  Writer w;
  w.b.reserve(1 << 18);
  w.vec(cap_w.data(), ncap, 4);
  w.vec(cap_z.data(), ncap, 4);
  w.vec(cap_q.data(), ncap, 4);
  {
    const uint64_t *e = at_zeta.data();
    const size_t w0 = cols[0], z0 = w0 + cols[1], q0 = z0 + cols[2];
    w.vec(e, nc, 2);                                        // constants
    w.vec(e + 2 * nc, w0 - nc, 2);                          // plonk_sigmas
    w.vec(e + 2 * w0, cols[1], 2);                          // wires
    w.vec(e + 2 * z0, nch, 2);                              // plonk_zs
    w.vec(zs_next.data(), nch, 2);                          // plonk_zs_next
    w.vec(e + 2 * (z0 + nch), (size_t)nch * npp, 2);        // partial_products
    w.vec(e + 2 * q0, cols[3], 2);                          // quotient_polys
    w.u64le(0);                                             // lookup_zs
    w.u64le(0);                                             // lookup_zs_next
  }
  w.u64le(layer_caps.size());                               // FriProof.commit_phase_merkle_caps
  for (auto &cap : layer_caps) w.vec(cap.data(), ncap, 4);
  w.u64le(nq);                                              // query_round_proofs
  for (u32 q = 0; q < nq; q++) {
    w.u64le(4);                                             // FriInitialTreeProof { evals_proofs }
    for (int o = 0; o < 4; o++) {
      w.vec(rows[o].data() + (size_t)q * cols[o], cols[o], 1);
      w.vec(sibs[o].data() + (size_t)q * h0 * 4, h0, 4);
    }
    w.u64le(layers.size());                                 // steps
    for (auto &L : layers) {
      w.vec(L.leaves.data() + (size_t)q * L.leaf_len, L.leaf_len / 2, 2);
      w.vec(L.sib.data() + (size_t)q * L.hsib * 4, L.hsib, 4);
    }
  }
  w.vec(final_poly.data(), final_len, 2);                   // final_poly: PolynomialCoeffs<F::Extension>
  w.u64le(pow_witness % kP);
  w.vec(public_inputs, npi, 1);
  *bytes = std::move(w.b);
  return "";
}

}  // namespace
}  // namespace mp2

extern "C" const char *mp2gpu_prove(const mp2gpu_circuit *circuit, const mp2gpu_prove_config *config,
                                    const mp2gpu_batch *constants_sigmas, const uint64_t *circuit_digest,
                                    const uint64_t *const *wires_values, const uint64_t *public_inputs,
                                    size_t num_public_inputs, const uint64_t *public_inputs_hash, uint8_t **proof_out,
                                    size_t *proof_len_out) {
  mp2::Status s;
  std::vector<uint8_t> bytes;
  if (!proof_out || !proof_len_out) {
    s = "prove: null output";
  } else {
    try {
      s = mp2::prove(circuit, config, constants_sigmas, circuit_digest, wires_values, public_inputs, num_public_inputs,
                     public_inputs_hash, &bytes);
    } catch (const std::exception &e) {
      s = std::string("exception: ") + e.what();
    }
  }
  if (s.empty()) {
    uint8_t *p = (uint8_t *)malloc(bytes.size() ? bytes.size() : 1);
    if (!p) {
      s = "prove: out of host memory";
    } else {
      memcpy(p, bytes.data(), bytes.size());
      *proof_out = p;
      *proof_len_out = bytes.size();
      return nullptr;
    }
  }
  char *m = (char *)malloc(s.size() + 1);
  if (m) memcpy(m, s.c_str(), s.size() + 1);
  return m;
}

extern "C" void mp2gpu_free_bytes(uint8_t *p) { free(p); }
