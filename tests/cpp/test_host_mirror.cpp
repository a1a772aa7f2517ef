// C++ host mirror (include/mp2gpu_plonky2.hpp) against the CPU oracle.  Built and run by
// tests/test_gpu_cpp_mirror.py on the GPU box; the oracle is linked here as the checker only.
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/mp2gpu_plonky2.hpp"
#include "../../oracle/mp2_oracle.h"

using namespace mp2gpu;

static uint64_t splitmix(uint64_t &s) {
  uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
#define REQUIRE(c)                                              \
  do {                                                          \
    if (!(c)) {                                                 \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
      return 1;                                                 \
    }                                                           \
  } while (0)

template <Hasher H>
static int run(uint32_t kind) {
  uint64_t seed = 0x6d7032 + kind;
  const size_t ncols = 9, n_log = 8, n = 1 << n_log, rate_bits = 3, cap_height = 4, N = n << rate_bits;
  std::vector<PolynomialValues> values(ncols);
  for (auto &v : values) {
    v.values.resize(n);
    for (auto &x : v.values) x = splitmix(seed);  // any u64, non-canonical included
  }
  auto pb = PolynomialBatch<H>::from_values(values, rate_bits, false, cap_height);
  // oracle
  std::vector<const uint64_t *> cols(ncols);
  for (size_t c = 0; c < ncols; c++) cols[c] = values[c].values.data();
  std::vector<uint64_t> coeffs(ncols * n), leaves(N * ncols), digests(2 * (N - 16) * 4), cap(16 * 4);
  REQUIRE(orc_commit(cols.data(), ncols, n_log, rate_bits, cap_height, kind, 0, coeffs.data(), leaves.data(),
                     digests.data(), cap.data(), 2) == 0);
  for (size_t c = 0; c < ncols; c++) REQUIRE(!std::memcmp(pb.polynomials[c].coeffs.data(), &coeffs[c * n], 8 * n));
  for (size_t i = 0; i < N; i++) REQUIRE(!std::memcmp(pb.merkle_tree.leaves[i].data(), &leaves[i * ncols], 8 * ncols));
  REQUIRE(!std::memcmp(pb.merkle_tree.digests[0].data(), digests.data(), digests.size() * 8));
  REQUIRE(!std::memcmp(pb.merkle_tree.cap.hashes[0].data(), cap.data(), cap.size() * 8));
  REQUIRE(pb.degree_log == n_log && pb.rate_bits == rate_bits && !pb.blinding);
  // get_lde_values(i, 8) is row reverse_bits(8 i)
  REQUIRE(pb.get_lde_values(5, 8) == pb.merkle_tree.leaves[reverse_bits(40, n_log + rate_bits)]);
  // MerkleTree::new on the circuit-set shape: 4-element digests padded with [0], cap 0
  std::vector<std::vector<F>> set_leaves;
  for (int i = 0; i < 11; i++) set_leaves.push_back({splitmix(seed) >> 1, splitmix(seed) >> 1, splitmix(seed) >> 1, splitmix(seed) >> 1});
  while (set_leaves.size() < 16) set_leaves.push_back({0});
  auto mt = MerkleTree<H>::new_(set_leaves, 0);
  REQUIRE(mt.cap.len() == 1 && mt.digests.size() == 30);
  auto proof = mt.prove(7);
  REQUIRE(proof.len() == 4);
  uint64_t root[4];
  orc_merkle_verify(set_leaves[7].data(), 4, 7, proof.siblings[0].data(), 4, kind, root);
  REQUIRE(!std::memcmp(root, mt.cap.hashes[0].data(), 32));
  // prove_openings: all polynomials at zeta, the first two again at g*zeta; then one reduction layer
  {
    FriInstanceInfo inst;
    Ext zeta = {splitmix(seed) >> 1, splitmix(seed) >> 1}, gzeta = {splitmix(seed) >> 1, splitmix(seed) >> 1};
    Ext alpha = {splitmix(seed) >> 1, splitmix(seed) >> 1}, beta = {splitmix(seed) >> 1, splitmix(seed) >> 1};
    FriBatchInfo b0{zeta, {}}, b1{gzeta, {{0, 0}, {0, 1}}};
    for (size_t c = 0; c < ncols; c++) b0.polynomials.push_back({0, c});
    inst.batches = {b0, b1};
    // OpeningSet: polynomial 3 at zeta by Horner in GF(p^2) with the oracle's extension multiplication
    auto opened = pb.eval({zeta, gzeta});
    {
      uint64_t acc[2] = {0, 0};
      for (size_t m = n; m-- > 0;) {
        uint64_t prod[2];
        orc_ext_mul(acc, zeta.data(), prod);
        acc[0] = orc_gl_add(prod[0], coeffs[3 * n + m]);
        acc[1] = prod[1];
      }
      REQUIRE(opened.size() == 2 && opened[0].size() == ncols);
      REQUIRE(opened[0][3][0] == orc_gl_canon(acc[0]) && opened[0][3][1] == orc_gl_canon(acc[1]));
    }
    auto ph = prove_openings_begin<H>(inst, {&pb}, alpha, cap_height, true);
    std::vector<const uint64_t *> polys;
    for (auto &b : inst.batches)
      for (auto &pinfo : b.polynomials) polys.push_back(&coeffs[pinfo.polynomial_index * n]);
    const uint32_t sizes[2] = {(uint32_t)ncols, 2};
    const uint64_t points[4] = {zeta[0], zeta[1], gzeta[0], gzeta[1]};
    std::vector<uint64_t> fin(2 * n);
    REQUIRE(orc_fri_combine(polys.data(), sizes, 2, points, alpha.data(), n, fin.data()) == 0);
    REQUIRE(ph.final_poly.size() == n && !std::memcmp(ph.final_poly[0].data(), fin.data(), 16 * n));
    // fri_committed_trees, one layer of arity 16: cap, fold, remaining coefficients
    auto layer_cap = ph.commit_layer(4);
    std::vector<uint64_t> padded(2 * N, 0), vals(2 * N), lv(2 * N), dg(2 * (N / 16 - 16) * 4), cp(16 * 4);
    std::memcpy(padded.data(), fin.data(), 16 * n);
    orc_coset_fft_ext(padded.data(), n_log + rate_bits, 7, vals.data());
    orc_fri_layer_leaves(vals.data(), n_log + rate_bits, 4, lv.data());
    REQUIRE(orc_merkle_new(lv.data(), N / 16, 32, cap_height, kind, dg.data(), cp.data(), 2) == 0);
    REQUIRE(!std::memcmp(layer_cap.hashes[0].data(), cp.data(), cp.size() * 8));
    ph.fold(beta);
    auto rest = ph.finish();
    std::vector<uint64_t> folded(2 * (N / 16));
    orc_fri_fold(padded.data(), N, 4, beta.data(), folded.data());
    REQUIRE(rest.size() == n / 16 && !std::memcmp(rest[0].data(), folded.data(), 16 * (n / 16)));
  }
  // panics like plonky2
  bool threw = false;
  try {
    MerkleTree<H>::new_(std::vector<std::vector<F>>(8, std::vector<F>(5, 1)), 4);
  } catch (const Panic &e) {
    threw = std::string(e.what()).find("cap_height") != std::string::npos;
  }
  REQUIRE(threw);
  threw = false;
  try {
    MerkleTree<H>::new_(std::vector<std::vector<F>>(6, std::vector<F>(5, 1)), 0);
  } catch (const Panic &) {
    threw = true;
  }
  REQUIRE(threw);
  return 0;
}

// prove(): one native call from the C++ mirror.  Synthetic (not satisfying) wires: the call does not check the witness;
// what is pinned here is the plumbing -- descriptor, config, byte buffer ownership, determinism -- the proof CONTENT is
// pinned from Python (tests/test_gpu_prove_verify.py: byte-identical to the mirror's sequence, accepted by a verifier).
template <Hasher H>
static int run_prove() {
  uint64_t seed = 4242;
  CircuitDesc c;
  c.degree_bits = 6;
  c.num_wires = 12;
  c.num_routed_wires = 8;
  c.num_constants = 1 + 2;
  c.gates = {GateInfo{MP2GPU_GATE_ARITHMETIC, 2, 0}, GateInfo{MP2GPU_GATE_CONSTANT, 2, 0}, GateInfo{MP2GPU_GATE_NOOP, 0, 0},
             GateInfo{MP2GPU_GATE_PUBLIC_INPUT, 0, 0}};
  c.selector_indices = {0, 0, 0, 0};
  c.groups = {{0, 4}};
  const size_t n = size_t(1) << c.degree_bits;
  FriConfig cfg;
  std::vector<PolynomialValues> cs(c.num_constants + c.num_routed_wires), wires(c.num_wires);
  for (auto &col : cs) {
    col.values.resize(n);
    for (auto &v : col.values) v = splitmix(seed) % 0xFFFFFFFF00000001ULL;
  }
  for (auto &v : cs[0].values) v %= 4;  // the selector column
  for (auto &col : wires) {
    col.values.resize(n);
    for (auto &v : col.values) v = splitmix(seed) % 0xFFFFFFFF00000001ULL;
  }
  auto b_cs = PolynomialBatch<H>::from_values(cs, cfg.rate_bits, false, cfg.cap_height);
  const std::array<F, 4> digest{1, 2, 3, 4}, pi_hash{5, 6, 7, 8};
  const std::vector<F> pis{9, 10};
  const std::vector<uint8_t> a = prove<H>(c, b_cs, digest, wires, pis, pi_hash, cfg);
  const std::vector<uint8_t> b = prove<H>(c, b_cs, digest, wires, pis, pi_hash, cfg);
  REQUIRE(a == b);                       // smallest PoW witness: reproducible bytes
  REQUIRE(a.size() > 3 * (8 + 16 * 32));
  uint64_t ncap = 0;
  for (int i = 0; i < 8; i++) ncap |= (uint64_t)a[i] << (8 * i);
  REQUIRE(ncap == (uint64_t(1) << cfg.cap_height));   // Vec<HashOut> length of wires_cap
  uint64_t last = 0, npis = 0;
  for (int i = 0; i < 8; i++) last |= (uint64_t)a[a.size() - 8 + i] << (8 * i);
  for (int i = 0; i < 8; i++) npis |= (uint64_t)a[a.size() - 24 + i] << (8 * i);
  REQUIRE(last == 10 && npis == 2);      // ... and the public inputs at the end
  // errors come back as Panic
  CircuitDesc bad = c;
  bad.gates[0].kind = 99;
  bool threw = false;
  try {
    prove<H>(bad, b_cs, digest, wires, pis, pi_hash, cfg);
  } catch (const Panic &) {
    threw = true;
  }
  REQUIRE(threw);
  return 0;
}

int main() {
  try {
    init(0);
  } catch (const Panic &e) {
    std::printf("init failed: %s\n", e.what());
    return 2;
  }
  if (run<Hasher::Poseidon>(0) || run<Hasher::Poseidon2>(1)) return 1;
  if (run_prove<Hasher::Poseidon>() || run_prove<Hasher::Poseidon2>()) return 1;
  std::printf("cpp host mirror OK\n");
  return 0;
}
