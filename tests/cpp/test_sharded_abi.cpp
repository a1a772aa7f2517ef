// Multi-GPU C ABI (mp2gpu_comm_init + mp2gpu_commit_from_values_sharded, include/mp2gpu.h) against the CPU oracle
// and against the single-GPU entry point.  Built and run by tests/test_gpu_sharded.py on a box with >= 2 GPUs
// (argv[1] = number of devices to use; 1 exercises the degenerate communicator).  The oracle is linked as the
// checker only.  Shapes cover the single-pass transform (n = 2^12) and the four-step one (n = 2^15), both
// hashers, from_values and from_coeffs.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/mp2gpu.h"
#include "../../oracle/mp2_oracle.h"

static uint64_t splitmix(uint64_t &s) {
  uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
#define REQUIRE(c)                                               \
  do {                                                           \
    if (!(c)) {                                                  \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
      return 1;                                                  \
    }                                                            \
  } while (0)
#define OK(call)                                          \
  do {                                                    \
    const char *_e = (call);                              \
    if (_e) {                                             \
      std::printf("FAILED %s: %s\n", #call, _e);          \
      mp2gpu_free_string(_e);                             \
      return 1;                                           \
    }                                                     \
  } while (0)

static int run(mp2gpu_comm *comm, size_t ncols, uint32_t n_log, uint32_t kind, int from_coeffs) {
  const uint32_t rate_bits = 3, cap_height = 4;
  const size_t n = (size_t)1 << n_log, N = n << rate_bits, ndig = 2 * (N - 16);
  uint64_t seed = 0x6d7033 + 31 * n_log + kind;
  std::vector<uint64_t> in(ncols * n);
  for (auto &x : in) x = splitmix(seed);  // any u64, non-canonical included
  std::vector<const uint64_t *> cols(ncols);
  for (size_t c = 0; c < ncols; c++) cols[c] = in.data() + c * n;
  // oracle
  std::vector<uint64_t> r_coeffs(ncols * n), r_leaves(N * ncols), r_dig(ndig * 4), r_cap(64);
  REQUIRE(orc_commit(cols.data(), ncols, n_log, rate_bits, cap_height, kind, from_coeffs, r_coeffs.data(), r_leaves.data(),
                     r_dig.data(), r_cap.data(), 0) == 0);
  // sharded
  std::vector<uint64_t> coeffs(ncols * n, ~0ull), leaves(N * ncols, ~0ull), dig(ndig * 4, ~0ull), cap(64, ~0ull);
  std::vector<uint64_t *> cout_(ncols);
  for (size_t c = 0; c < ncols; c++) cout_[c] = coeffs.data() + c * n;
  OK(mp2gpu_commit_from_values_sharded(comm, cols.data(), ncols, n_log, rate_bits, cap_height, kind, from_coeffs,
                                       cout_.data(), leaves.data(), dig.data(), cap.data()));
  REQUIRE(coeffs == r_coeffs);
  REQUIRE(leaves == r_leaves);
  REQUIRE(dig == r_dig);
  REQUIRE(cap == r_cap);
  // a second call reuses the communicator's buffers; outputs optional
  std::vector<uint64_t> cap2(64, ~0ull);
  OK(mp2gpu_commit_from_values_sharded(comm, cols.data(), ncols, n_log, rate_bits, cap_height, kind, from_coeffs, nullptr,
                                       nullptr, nullptr, cap2.data()));
  REQUIRE(cap2 == r_cap);
  return 0;
}

int main(int argc, char **argv) {
  const int ndev = argc > 1 ? std::atoi(argv[1]) : 2;
  int have = 0;
  OK(mp2gpu_device_count(&have));
  if (have < ndev) {
    std::printf("SKIP: %d device(s), need %d\n", have, ndev);
    return 0;
  }
  std::vector<int> devs(ndev);
  for (int i = 0; i < ndev; i++) devs[i] = i;
  mp2gpu_comm *comm = nullptr;
  OK(mp2gpu_comm_init(ndev, devs.data(), &comm));
  for (uint32_t kind = 0; kind < 2; kind++) {
    if (run(comm, 16, 12, kind, 0)) return 1;
    if (run(comm, 8, 15, kind, 0)) return 1;
    if (run(comm, 8, 10, kind, 1)) return 1;
  }
  // argument errors come back as strings, never as crashes
  {
    std::vector<uint64_t> in(3 * 16, 1), cap(64);
    std::vector<const uint64_t *> cols = {in.data(), in.data() + 16, in.data() + 32};
    const char *e = mp2gpu_commit_from_values_sharded(comm, cols.data(), 3, 4, 3, 4, 0, 0, nullptr, nullptr, nullptr, cap.data());
    if (ndev > 1) {
      REQUIRE(e != nullptr && std::strstr(e, "split evenly"));
    }
    if (e) mp2gpu_free_string(e);
    e = mp2gpu_commit_from_values_sharded(nullptr, cols.data(), 3, 4, 3, 4, 0, 0, nullptr, nullptr, nullptr, cap.data());
    REQUIRE(e != nullptr);
    mp2gpu_free_string(e);
  }
  mp2gpu_comm_free(comm);
  mp2gpu_comm *bad = nullptr;
  const char *e = mp2gpu_comm_init(3, nullptr, &bad);
  REQUIRE(e != nullptr && bad == nullptr);
  mp2gpu_free_string(e);
  std::printf("sharded C ABI OK on %d device(s)\n", ndev);
  return 0;
}
