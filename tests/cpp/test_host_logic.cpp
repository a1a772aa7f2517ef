// CPU-only check of the C++ host mirror's transcript logic (include/mp2gpu_plonky2.hpp): Challenger over the CPU
// oracle's permutation against a straight-line restatement of the duplex rules, and the FRI reduction schedule.
// Links libmp2gpu.so only because the header declares its symbols; no device call is made.
#include <cstdio>
#include <vector>

#include "../../include/mp2gpu_plonky2.hpp"
#include "../../oracle/mp2_oracle.h"

using namespace mp2gpu;

#define REQUIRE(c)                                              \
  do {                                                          \
    if (!(c)) {                                                 \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
      return 1;                                                 \
    }                                                           \
  } while (0)

struct OraclePermute {
  uint32_t kind;
  void operator()(uint64_t *state) const { orc_permute(kind, state); }
};

static uint64_t splitmix(uint64_t &s) {
  uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

static int run(uint32_t kind) {
  uint64_t seed = 77 + kind;
  Challenger<OraclePermute> ch(OraclePermute{kind});
  // restatement: absorb in chunks of 8 by overwriting, permute, squeeze state[0..8) from the end
  uint64_t st[12] = {0};
  std::vector<uint64_t> pending, out;
  auto duplex = [&]() {
    for (size_t i = 0; i < pending.size(); i++) st[i] = pending[i];
    pending.clear();
    orc_permute(kind, st);
    out.assign(st, st + 8);
  };
  for (int step = 0; step < 60; step++) {
    if (splitmix(seed) % 3) {
      size_t cnt = 1 + splitmix(seed) % 11;
      for (size_t i = 0; i < cnt; i++) {
        uint64_t x = splitmix(seed);
        ch.observe_element(x);
        out.clear();
        pending.push_back(orc_gl_canon(x));
        if (pending.size() == 8) duplex();
      }
    } else {
      if (!pending.empty() || out.empty()) duplex();
      uint64_t want = out.back();
      out.pop_back();
      REQUIRE(ch.get_challenge() == want);
    }
    REQUIRE(ch.input_buffer.size() < 8);
  }
  size_t pos = 99;
  auto inter = ch.pow_intermediate_state(&pos);
  REQUIRE(pos == ch.input_buffer.size());
  for (size_t i = 0; i < pos; i++) REQUIRE(inter[i] == ch.input_buffer[i]);
  for (size_t i = pos; i < 12; i++) REQUIRE(inter[i] == ch.sponge_state[i]);
  return 0;
}

// HostPermute (mp2gpu_transcript_permute, the transcript's host-side permutation) == the oracle's, both hashers
static int host_permute_matches_oracle() {
  uint64_t seed = 4242;
  for (int it = 0; it < 50; it++) {
    uint64_t a[12], b[12], c[12];
    for (int i = 0; i < 12; i++) a[i] = b[i] = c[i] = it < 2 ? (it ? ~0ULL : 0ULL) : splitmix(seed);
    uint64_t a2[12];
    for (int i = 0; i < 12; i++) a2[i] = a[i];
    HostPermute<Hasher::Poseidon>()(a);
    orc_permute(0, b);
    HostPermute<Hasher::Poseidon2>()(a2);
    orc_permute(1, c);
    for (int i = 0; i < 12; i++) {
      REQUIRE(a[i] == orc_gl_canon(b[i]));
      REQUIRE(a2[i] == orc_gl_canon(c[i]));
    }
  }
  return 0;
}

int main() {
  if (run(0) || run(1)) return 1;
  if (host_permute_matches_oracle()) return 1;
  FriConfig cfg;
  REQUIRE((cfg.reduction_arity_bits(14) == std::vector<size_t>{4, 4, 4}));
  REQUIRE((cfg.reduction_arity_bits(13) == std::vector<size_t>{4, 4}));
  REQUIRE((cfg.reduction_arity_bits(12) == std::vector<size_t>{4, 4}));
  REQUIRE(cfg.reduction_arity_bits(5).empty());
  // CircuitDesc -> mp2gpu_circuit: selector groups resolved per gate, inconsistent descriptors refused
  CircuitDesc cd;
  cd.degree_bits = 5;
  cd.num_constants = 4;
  cd.gates = {GateInfo{MP2GPU_GATE_ARITHMETIC, 20, 0}, GateInfo{MP2GPU_GATE_NOOP, 0, 0},
              GateInfo{MP2GPU_GATE_COSET_INTERPOLATION, 4, 6}};
  cd.selector_indices = {0, 0, 1};
  cd.groups = {{0, 2}, {2, 3}};
  std::vector<mp2gpu_gate> storage;
  const mp2gpu_circuit c = cd.c_desc(storage);
  REQUIRE(c.num_gates == 3 && c.num_selectors == 2 && c.gates == storage.data() && c.num_wires == 135 && c.num_routed_wires == 80);
  REQUIRE(storage[1].group_begin == 0 && storage[1].group_end == 2 && storage[2].selector_index == 1);
  REQUIRE(storage[2].kind == MP2GPU_GATE_COSET_INTERPOLATION && storage[2].num_ops == 4 && storage[2].param == 6);
  cd.selector_indices = {0, 0, 2};
  bool threw = false;
  try {
    cd.c_desc(storage);
  } catch (const Panic &) {
    threw = true;
  }
  REQUIRE(threw);
  std::printf("cpp host logic OK\n");
  return 0;
}
