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

int main() {
  try {
    init(0);
  } catch (const Panic &e) {
    std::printf("init failed: %s\n", e.what());
    return 2;
  }
  if (run<Hasher::Poseidon>(0) || run<Hasher::Poseidon2>(1)) return 1;
  std::printf("cpp host mirror OK\n");
  return 0;
}
