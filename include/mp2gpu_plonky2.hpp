// mp2gpu_plonky2.hpp -- C++ host-side mirror of the plonky2 surface the reference reaches the hot path
// through, written above the C ABI of mp2gpu.h (the reference's host language, Rust, has no toolchain in
// this image; the Rust binding itself is in INTEGRATION.md).  Same names, argument meaning and failure
// behaviour as plonky2 0.2.2:
//
//   PolynomialBatch::from_values / from_coeffs / get_lde_values   (plonky2 fri/oracle.rs; reached via
//       circuit_data.prove at recursion-framework/src/circuit_builder.rs:308 and builder.build at :177)
//   MerkleTree::new_ / prove / get, MerkleCap, MerkleProof          (plonky2 hash/merkle_tree.rs; called
//       directly at recursion-framework/src/universal_verifier_gadget/circuit_set.rs:189, :216)
//   FriInstanceInfo / FriBatchInfo / FriPolynomialInfo, the head of PolynomialBatch::prove_openings and the
//       fri_committed_trees loop (plonky2 fri/{structure,oracle,prover}.rs; same call sites as from_values)
//
// Where Rust panics (MerkleTree::new with cap_height > log2(len), non power-of-two lengths) this throws
// mp2gpu::Panic.  Header only; link with -lmp2gpu.
#pragma once
#include <array>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "mp2gpu.h"

namespace mp2gpu {

using F = uint64_t;                 // GoldilocksField (repr(transparent) u64), canonical on output
using HashOut = std::array<F, 4>;   // HashOut<F>

struct Panic : std::runtime_error {
  using std::runtime_error::runtime_error;
};

inline void check(const char *err) {  // handle_c_result of gnark-utils/src/utils.rs:9-20
  if (err) {
    std::string msg(err);
    mp2gpu_free_string(err);
    throw Panic(msg);
  }
}

enum class Hasher : uint32_t { Poseidon = MP2GPU_HASH_POSEIDON, Poseidon2 = MP2GPU_HASH_POSEIDON2 };

inline void init(int device = 0) { check(mp2gpu_init(device)); }

struct MerkleCap {
  std::vector<HashOut> hashes;  // MerkleCap(pub Vec<H::Hash>)
  size_t height() const {
    size_t h = 0;
    while ((size_t(1) << h) < hashes.size()) h++;
    return h;
  }
  size_t len() const { return hashes.size(); }
};

struct MerkleProof {
  std::vector<HashOut> siblings;  // bottom-up
  size_t len() const { return siblings.size(); }
};

struct PolynomialValues {
  std::vector<F> values;
};
struct PolynomialCoeffs {
  std::vector<F> coeffs;
};

template <Hasher H>
struct MerkleTree {
  std::vector<std::vector<F>> leaves;
  std::vector<HashOut> digests;
  MerkleCap cap;

  // MerkleTree::new(leaves, cap_height)
  static MerkleTree new_(std::vector<std::vector<F>> leaves, size_t cap_height) {
    MerkleTree t;
    const size_t n = leaves.size();
    std::vector<const uint64_t *> ptrs(n);
    std::vector<size_t> lens(n);
    for (size_t i = 0; i < n; i++) {
      ptrs[i] = leaves[i].data();
      lens[i] = leaves[i].size();
    }
    // sizes are validated by the library (it reports plonky2's assertion text); allocate defensively
    const size_t ncap = cap_height < 63 ? size_t(1) << cap_height : 0;
    t.digests.resize(n > ncap ? 2 * (n - ncap) : 0);
    t.cap.hashes.resize(ncap ? ncap : 1);
    check(mp2gpu_merkle_new_ragged(ptrs.data(), lens.data(), n, (uint32_t)cap_height, (uint32_t)H,
                                   t.digests.empty() ? nullptr : t.digests[0].data(), t.cap.hashes[0].data()));
    t.leaves = std::move(leaves);
    return t;
  }
  const std::vector<F> &get(size_t i) const { return leaves[i]; }
  // MerkleTree::prove(leaf_index)
  MerkleProof prove(size_t leaf_index) const {
    MerkleProof p;
    size_t lg = 0;
    while ((size_t(1) << lg) < leaves.size()) lg++;
    p.siblings.resize(lg - cap.height() ? lg - cap.height() : 1);
    size_t cnt = 0;
    check(mp2gpu_merkle_prove(digests.empty() ? nullptr : digests[0].data(), leaves.size(), (uint32_t)cap.height(),
                              leaf_index, p.siblings[0].data(), &cnt));
    p.siblings.resize(cnt);
    return p;
  }
};

inline size_t reverse_bits(size_t x, size_t bits) {
  size_t r = 0;
  for (size_t i = 0; i < bits; i++) r |= ((x >> i) & 1) << (bits - 1 - i);
  return r;
}

template <Hasher H>
struct PolynomialBatch {
  std::vector<PolynomialCoeffs> polynomials;
  MerkleTree<H> merkle_tree;
  size_t degree_log = 0, rate_bits = 0;
  bool blinding = false;
  // the copy of this batch that stays in HBM for prove_openings and the query rounds (released with the last
  // C++ copy of the batch); empty when the batch was built with keep_on_device = false
  std::shared_ptr<mp2gpu_batch> device;

  // PolynomialBatch::from_values(values, rate_bits, blinding, cap_height, timing, fft_root_table):
  // timing / fft_root_table have no GPU counterpart (twiddles are device resident) and are omitted.
  static PolynomialBatch from_values(const std::vector<PolynomialValues> &values, size_t rate_bits, bool blinding,
                                     size_t cap_height, bool keep_on_device = true) {
    std::vector<const uint64_t *> cols(values.size());
    for (size_t c = 0; c < values.size(); c++) cols[c] = values[c].values.data();
    return commit(cols, values.empty() ? 0 : values[0].values.size(), rate_bits, blinding, cap_height, false,
                  keep_on_device);
  }
  static PolynomialBatch from_coeffs(const std::vector<PolynomialCoeffs> &polys, size_t rate_bits, bool blinding,
                                     size_t cap_height, bool keep_on_device = true) {
    std::vector<const uint64_t *> cols(polys.size());
    for (size_t c = 0; c < polys.size(); c++) cols[c] = polys[c].coeffs.data();
    return commit(cols, polys.empty() ? 0 : polys[0].coeffs.size(), rate_bits, blinding, cap_height, true,
                  keep_on_device);
  }
  // OpeningSet::new's `polynomials.par_iter().map(|p| p.to_extension().eval(z))` for several points at once, on
  // the device-resident coefficients: result[point][polynomial] = [a0, a1]
  std::vector<std::vector<std::array<F, 2>>> eval(const std::vector<std::array<F, 2>> &points) const {
    if (!device) throw Panic("batch was not kept on the device");
    std::vector<uint64_t> flat(points.size() * polynomials.size() * 2);
    if (!points.empty()) check(mp2gpu_batch_eval(device.get(), points[0].data(), points.size(), flat.data()));
    std::vector<std::vector<std::array<F, 2>>> out(points.size(), std::vector<std::array<F, 2>>(polynomials.size()));
    for (size_t p = 0; p < points.size(); p++)
      for (size_t c = 0; c < polynomials.size(); c++)
        out[p][c] = {flat[2 * (p * polynomials.size() + c)], flat[2 * (p * polynomials.size() + c) + 1]};
    return out;
  }
  // get_lde_values(index, step): leaves[reverse_bits(index * step, degree_log + rate_bits)]
  const std::vector<F> &get_lde_values(size_t index, size_t step) const {
    return merkle_tree.leaves[reverse_bits(index * step, degree_log + rate_bits)];
  }

 private:
  static PolynomialBatch commit(const std::vector<const uint64_t *> &cols, size_t n, size_t rate_bits, bool blinding,
                                size_t cap_height, bool from_coeffs, bool keep_on_device) {
    if (blinding) throw Panic("blinding (salted) batches are not supported on the GPU path");
    size_t n_log = 0;
    while ((size_t(1) << n_log) < n) n_log++;
    if (cols.empty() || n == 0 || (size_t(1) << n_log) != n) throw Panic("polynomial length must be a power of two");
    const size_t N = n << rate_bits, ncols = cols.size(), ncap = size_t(1) << cap_height;
    PolynomialBatch b;
    b.degree_log = n_log;
    b.rate_bits = rate_bits;
    b.polynomials.resize(ncols);
    std::vector<uint64_t *> coeff_ptrs(ncols);
    for (size_t c = 0; c < ncols; c++) {
      b.polynomials[c].coeffs.resize(n);
      coeff_ptrs[c] = b.polynomials[c].coeffs.data();
    }
    std::vector<F> flat(N * ncols);
    b.merkle_tree.digests.resize(N > ncap ? 2 * (N - ncap) : 0);
    b.merkle_tree.cap.hashes.resize(ncap);
    auto fn = from_coeffs ? mp2gpu_commit_from_coeffs : mp2gpu_commit_from_values;
    mp2gpu_batch *handle = nullptr;
    check(fn(cols.data(), ncols, (uint32_t)n_log, (uint32_t)rate_bits, (uint32_t)cap_height, (uint32_t)H,
             coeff_ptrs.data(), flat.data(),
             b.merkle_tree.digests.empty() ? nullptr : b.merkle_tree.digests[0].data(),
             b.merkle_tree.cap.hashes[0].data(), keep_on_device ? &handle : nullptr));
    if (handle) b.device.reset(handle, mp2gpu_batch_free);
    b.merkle_tree.leaves.resize(N);
    for (size_t i = 0; i < N; i++) b.merkle_tree.leaves[i].assign(flat.begin() + i * ncols, flat.begin() + (i + 1) * ncols);
    return b;
  }
};

// ---- FRI: instance description (plonky2 fri/structure.rs) and the prover's commit phase ----------------
using Ext = std::array<F, 2>;  // QuadraticExtension<GoldilocksField>: a0 + a1 X, X^2 = 7
struct FriPolynomialInfo {
  size_t oracle_index, polynomial_index;
};
struct FriBatchInfo {
  Ext point;
  std::vector<FriPolynomialInfo> polynomials;
};
struct FriInstanceInfo {
  std::vector<FriBatchInfo> batches;  // `oracles: Vec<FriOracleInfo>` only carries blinding flags (always false here)
};

// The polynomial FRI runs on and the fri_committed_trees loop over it, device resident.  The challenger stays
// with the caller: observe the cap commit_layer returns, draw beta, fold.
template <Hasher H>
struct FriCommitPhase {
  std::shared_ptr<mp2gpu_fri> state;
  size_t cap_height = 0;
  std::vector<Ext> final_poly;  // filled by prove_openings_begin when asked for

  // lde_polynomial_coeffs of fri_proof, without its zero padding
  static FriCommitPhase begin(const std::vector<Ext> &coeffs, size_t rate_bits, size_t cap_height) {
    size_t n_log = 0;
    while ((size_t(1) << n_log) < coeffs.size()) n_log++;
    if (coeffs.empty() || (size_t(1) << n_log) != coeffs.size()) throw Panic("polynomial length must be a power of two");
    mp2gpu_fri *f = nullptr;
    check(mp2gpu_fri_begin(coeffs[0].data(), (uint32_t)n_log, (uint32_t)rate_bits, (uint32_t)cap_height, (uint32_t)H, &f));
    FriCommitPhase ph;
    ph.state.reset(f, mp2gpu_fri_free);
    ph.cap_height = cap_height;
    return ph;
  }
  // MerkleTree::new(chunked values, cap_height).cap of the next reduction layer
  MerkleCap commit_layer(size_t arity_bits) {
    MerkleCap cap;
    cap.hashes.resize(size_t(1) << cap_height);
    check(mp2gpu_fri_commit_layer(state.get(), (uint32_t)arity_bits, cap.hashes[0].data()));
    return cap;
  }
  void fold(const Ext &beta) { check(mp2gpu_fri_fold(state.get(), beta.data())); }
  std::vector<Ext> finish() {
    size_t len = 0;
    check(mp2gpu_fri_finish(state.get(), nullptr, &len));
    std::vector<Ext> out(len);
    check(mp2gpu_fri_finish(state.get(), out[0].data(), &len));
    return out;
  }
};

// Challenger<F, H> (plonky2 iop/challenger.rs): the overwrite-mode duplex sponge of the Fiat-Shamir transcript.
// `Permute` is any callable void(uint64_t state[12]); DevicePermute sends the state through mp2gpu_permute_batch
// (round 1's only option: a ~60 us round trip per duplexing); HostPermute below keeps the transcript on the host.
template <Hasher H>
struct DevicePermute {
  void operator()(uint64_t *state) const { check(mp2gpu_permute_batch(state, 1, (uint32_t)H)); }
};
// The transcript's usual permutation: on the host (mp2gpu_transcript_permute), as plonky2's own challenger does.
template <Hasher H>
struct HostPermute {
  void operator()(uint64_t *state) const { check(mp2gpu_transcript_permute(state, (uint32_t)H)); }
};
template <typename Permute>
struct Challenger {
  static constexpr size_t kWidth = 12, kRate = 8;
  static constexpr uint64_t kOrder = 0xFFFFFFFF00000001ULL;
  std::array<F, kWidth> sponge_state{};
  std::vector<F> input_buffer, output_buffer;
  Permute permute;

  explicit Challenger(Permute p = Permute()) : permute(p) {}
  void observe_element(F x) {
    output_buffer.clear();  // any buffered outputs are now invalid
    input_buffer.push_back(x >= kOrder ? x - kOrder : x);
    if (input_buffer.size() == kRate) duplexing();
  }
  void observe_elements(const F *xs, size_t n) {
    for (size_t i = 0; i < n; i++) observe_element(xs[i]);
  }
  void observe_extension_element(const Ext &e) { observe_elements(e.data(), 2); }
  void observe_hash(const HashOut &h) { observe_elements(h.data(), 4); }
  void observe_cap(const MerkleCap &cap) {
    for (auto &h : cap.hashes) observe_hash(h);
  }
  F get_challenge() {
    if (!input_buffer.empty() || output_buffer.empty()) duplexing();
    F c = output_buffer.back();  // challenges are popped from the END of the squeezed rate
    output_buffer.pop_back();
    return c;
  }
  Ext get_extension_challenge() {
    F a = get_challenge();
    return {a, get_challenge()};
  }
  void duplexing() {
    for (size_t i = 0; i < input_buffer.size(); i++) sponge_state[i] = input_buffer[i];
    input_buffer.clear();
    permute(sponge_state.data());
    output_buffer.assign(sponge_state.begin(), sponge_state.begin() + kRate);
  }
  // fri_proof_of_work's view of the transcript: the state with the pending inputs written, and where the witness goes
  std::array<F, kWidth> pow_intermediate_state(size_t *witness_pos) const {
    std::array<F, kWidth> st = sponge_state;
    for (size_t i = 0; i < input_buffer.size(); i++) st[i] = input_buffer[i];
    *witness_pos = input_buffer.size();
    return st;
  }
};

// FriConfig of standard_recursion_config (mp2-common/src/lib.rs:45-47) and FriReductionStrategy::ConstantArityBits
struct FriConfig {
  size_t rate_bits = 3, cap_height = 4, proof_of_work_bits = 16, num_query_rounds = 28;
  size_t arity_bits = 4, final_poly_bits = 5;
  std::vector<size_t> reduction_arity_bits(size_t degree_bits) const {
    std::vector<size_t> out;
    while (degree_bits > final_poly_bits && degree_bits + rate_bits - cap_height > arity_bits) {
      out.push_back(arity_bits);
      degree_bits -= arity_bits;
    }
    return out;
  }
};

// PolynomialBatch::prove_openings(instance, oracles, challenger, fri_params, timing) up to its call of fri_proof:
// alpha = challenger.get_extension_challenge() is drawn by the caller; the alpha-batched quotient is built from the
// oracles' device-resident coefficients and stays in HBM as the commit phase's polynomial.
template <Hasher H>
FriCommitPhase<H> prove_openings_begin(const FriInstanceInfo &instance, const std::vector<const PolynomialBatch<H> *> &oracles,
                                       const Ext &alpha, size_t cap_height, bool want_final_poly = false) {
  std::vector<const mp2gpu_batch *> handles;
  for (auto *o : oracles) {
    if (!o || !o->device) throw Panic("prove_openings needs device-resident oracles (keep_on_device)");
    handles.push_back(o->device.get());
  }
  std::vector<uint64_t> points;
  std::vector<uint32_t> sizes, oi, pi;
  for (auto &b : instance.batches) {
    points.push_back(b.point[0]);
    points.push_back(b.point[1]);
    sizes.push_back((uint32_t)b.polynomials.size());
    for (auto &p : b.polynomials) {
      oi.push_back((uint32_t)p.oracle_index);
      pi.push_back((uint32_t)p.polynomial_index);
    }
  }
  FriCommitPhase<H> ph;
  ph.cap_height = cap_height;
  if (want_final_poly && !oracles.empty() && oracles[0]) ph.final_poly.resize(size_t(1) << oracles[0]->degree_log);
  mp2gpu_fri *f = nullptr;
  check(mp2gpu_fri_begin_openings(handles.data(), handles.size(), points.data(), sizes.data(), sizes.size(), oi.data(),
                                  pi.data(), alpha.data(), (uint32_t)cap_height, (uint32_t)H,
                                  ph.final_poly.empty() ? nullptr : ph.final_poly[0].data(), &f));
  ph.state.reset(f, mp2gpu_fri_free);
  return ph;
}

// The CommonCircuitData fields the device-side prover needs (mp2gpu_circuit with owned storage)
struct GateInfo {
  uint32_t kind = MP2GPU_GATE_NOOP, num_ops = 0, param = 0;   // see the MP2GPU_GATE_* notes in mp2gpu.h
};
struct CircuitDesc {
  size_t degree_bits = 0, num_wires = 135, num_routed_wires = 80, num_constants = 0;
  size_t quotient_degree_bits = 3, num_challenges = 2;
  std::vector<GateInfo> gates;                       // CommonCircuitData::gates order
  std::vector<size_t> selector_indices;              // SelectorsInfo::selector_indices
  std::vector<std::pair<size_t, size_t>> groups;     // SelectorsInfo::groups (gate index ranges)

  // fills `storage` and returns a descriptor pointing into it (valid while `storage` lives)
  mp2gpu_circuit c_desc(std::vector<mp2gpu_gate> &storage) const {
    if (selector_indices.size() != gates.size()) throw Panic("CircuitDesc: one selector index per gate");
    storage.resize(gates.size());
    for (size_t g = 0; g < gates.size(); g++) {
      if (selector_indices[g] >= groups.size()) throw Panic("CircuitDesc: selector index out of range");
      const auto &grp = groups[selector_indices[g]];
      storage[g] = mp2gpu_gate{gates[g].kind, gates[g].num_ops, (uint32_t)selector_indices[g], (uint32_t)grp.first,
                               (uint32_t)grp.second, gates[g].param};
    }
    return mp2gpu_circuit{(uint32_t)degree_bits, (uint32_t)quotient_degree_bits, (uint32_t)num_challenges, (uint32_t)num_wires,
                          (uint32_t)num_routed_wires, (uint32_t)num_constants, (uint32_t)groups.size(),
                          (uint32_t)gates.size(), storage.data()};
  }
};

// plonky2 `prove_with_partition_witness` after witness generation, as one call on the device (mp2gpu_prove):
// -> bincode(ProofWithPublicInputs) bytes (mp2-common/src/proof.rs:86-90 `serialize_proof`).  `constants_sigmas` is the
// device-resident batch committed at circuit-build time; `wires` are the num_wires witness columns.
template <Hasher H>
std::vector<uint8_t> prove(const CircuitDesc &circuit, const PolynomialBatch<H> &constants_sigmas,
                           const std::array<F, 4> &circuit_digest, const std::vector<PolynomialValues> &wires,
                           const std::vector<F> &public_inputs, const std::array<F, 4> &public_inputs_hash,
                           const FriConfig &config = FriConfig()) {
  if (!constants_sigmas.device) throw Panic("prove needs a device-resident constants_sigmas batch (keep_on_device)");
  if (wires.size() != circuit.num_wires) throw Panic("prove: one column per wire");
  for (auto &w : wires)
    if (w.values.size() != (size_t(1) << circuit.degree_bits)) throw Panic("prove: wire columns must have 2^degree_bits values");
  std::vector<mp2gpu_gate> storage;
  const mp2gpu_circuit cd = circuit.c_desc(storage);
  std::vector<uint32_t> arity;
  for (size_t a : config.reduction_arity_bits(circuit.degree_bits)) arity.push_back((uint32_t)a);
  const mp2gpu_prove_config cf{(uint32_t)config.rate_bits, (uint32_t)config.cap_height, (uint32_t)H,
                               (uint32_t)config.proof_of_work_bits, (uint32_t)config.num_query_rounds,
                               (uint32_t)arity.size(), arity.data()};
  std::vector<const uint64_t *> cols(wires.size());
  for (size_t c = 0; c < wires.size(); c++) cols[c] = wires[c].values.data();
  uint8_t *bytes = nullptr;
  size_t len = 0;
  check(mp2gpu_prove(&cd, &cf, constants_sigmas.device.get(), circuit_digest.data(), cols.data(),
                     public_inputs.empty() ? nullptr : public_inputs.data(), public_inputs.size(), public_inputs_hash.data(),
                     &bytes, &len));
  std::vector<uint8_t> out(bytes, bytes + len);
  mp2gpu_free_bytes(bytes);
  return out;
}

}  // namespace mp2gpu
