// Host-side Poseidon / Poseidon2 permutation for the Fiat-Shamir TRANSCRIPT only (plonky2 iop/challenger.rs
// `Challenger::duplexing`): a proof needs a few dozen strictly sequential permutations of a 12-element state whose
// inputs are caps and challenges that live on the host.  In the reference the challenger is host code (it is part of
// plonky2's prover, reached from recursion-framework/src/circuit_builder.rs:308); sending each of these permutations
// through the GPU cost a ~60 us round trip apiece, 6.7 ms per FRI proof.  This is NOT a data-path
// fallback: leaves, Merkle nodes, proof-of-work and every batched hash run on the device only (merkle.cu), and
// nothing in those paths calls this file.  Plain by-definition rounds (A.6 / A.7 of SURVEY.md), constants from the
// same generated table the kernels use.
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "../../include/mp2gpu.h"
#include "poseidon_constants.h"

namespace {
typedef unsigned long long u64;
typedef unsigned __int128 u128;
const u64 P = 0xFFFFFFFF00000001ULL;

inline u64 red(u128 x) { return (u64)(x % P); }
inline u64 mul(u64 a, u64 b) { return red((u128)a * b); }
inline u64 pow7(u64 x) {
  const u64 x2 = mul(x, x), x3 = mul(x2, x), x4 = mul(x2, x2);
  return mul(x3, x4);
}

const u64 kPosRc[MP2_POSEIDON_RC_LEN] = {MP2_POSEIDON_RC_LIST};
const u64 kP2Rc[MP2_POSEIDON2_RC_LEN] = {MP2_POSEIDON2_RC_LIST};
const u64 kP2Diag[MP2_POSEIDON2_DIAG_LEN] = {MP2_POSEIDON2_DIAG_LIST};

void poseidon(u64 (&s)[12]) {
  static const u64 CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
  for (int r = 0; r < 30; r++) {
    for (int i = 0; i < 12; i++) s[i] = red((u128)s[i] + kPosRc[12 * r + i]);
    if (r < 4 || r >= 26) {
      for (int i = 0; i < 12; i++) s[i] = pow7(s[i]);
    } else {
      s[0] = pow7(s[0]);
    }
    u64 t[12];
    for (int row = 0; row < 12; row++) {
      u128 acc = 0;  // 12 * 2^64 * 41 + 8 * 2^64 < 2^74
      for (int i = 0; i < 12; i++) acc += (u128)s[(i + row) % 12] * CIRC[i];
      if (row == 0) acc += (u128)s[0] * 8;
      t[row] = red(acc);
    }
    memcpy(s, t, sizeof(t));
  }
}

void p2_external(u64 (&s)[12]) {
  static const u64 M4[4][4] = {{5, 7, 1, 3}, {4, 6, 1, 1}, {1, 3, 5, 7}, {1, 1, 4, 6}};
  u64 y[12];
  for (int c = 0; c < 12; c += 4)
    for (int r = 0; r < 4; r++) {
      u128 acc = 0;
      for (int k = 0; k < 4; k++) acc += (u128)s[c + k] * M4[r][k];
      y[c + r] = red(acc);
    }
  for (int i = 0; i < 12; i++) s[i] = red((u128)y[i] + y[i % 4] + y[4 + i % 4] + y[8 + i % 4]);
}

void poseidon2(u64 (&s)[12]) {
  const u64 *rc = kP2Rc;
  p2_external(s);
  for (int r = 0; r < 4; r++, rc += 12) {
    for (int i = 0; i < 12; i++) s[i] = pow7(red((u128)s[i] + rc[i]));
    p2_external(s);
  }
  for (int r = 0; r < 22; r++, rc += 1) {
    s[0] = pow7(red((u128)s[0] + rc[0]));
    u128 tot = 0;
    for (int i = 0; i < 12; i++) tot += s[i];
    const u64 sum = red(tot);
    for (int i = 0; i < 12; i++) s[i] = red((u128)s[i] * kP2Diag[i] + sum);
  }
  for (int r = 0; r < 4; r++, rc += 12) {
    for (int i = 0; i < 12; i++) s[i] = pow7(red((u128)s[i] + rc[i]));
    p2_external(s);
  }
}
}  // namespace

extern "C" const char *mp2gpu_transcript_permute(uint64_t *state12, uint32_t hash_kind) {
  const char *msg = nullptr;
  if (!state12) msg = "null state";
  else if (hash_kind > 1) msg = "unknown hash_kind";
  if (msg) {
    char *m = (char *)malloc(strlen(msg) + 1);
    if (m) strcpy(m, msg);
    return m;
  }
  u64 s[12];
  for (int i = 0; i < 12; i++) s[i] = state12[i] % P;
  if (hash_kind == 1) poseidon2(s);
  else poseidon(s);
  for (int i = 0; i < 12; i++) state12[i] = s[i];
  return nullptr;
}

// Challenger::observe_elements in one call: every element clears the output buffer and joins the input buffer; a full
// rate (8) triggers a duplexing (overwrite-absorb + permutation).  state12 / buffer8 / buffer_len are the
// challenger's sponge_state and input_buffer, updated in place; *duplexed_last_out = 1 iff the LAST element
// completed a duplexing (then the output buffer is state[0..8), otherwise it is empty).
extern "C" const char *mp2gpu_transcript_observe(uint64_t *state12, uint64_t *buffer8, uint32_t *buffer_len,
                                                 const uint64_t *elems, size_t n, uint32_t hash_kind,
                                                 uint32_t *duplexed_last_out) {
  const char *msg = nullptr;
  if (!state12 || !buffer8 || !buffer_len || (!elems && n) || !duplexed_last_out) msg = "null argument";
  else if (hash_kind > 1) msg = "unknown hash_kind";
  else if (*buffer_len >= 8) msg = "input buffer already holds a full rate";
  if (msg) {
    char *m = (char *)malloc(strlen(msg) + 1);
    if (m) strcpy(m, msg);
    return m;
  }
  u64 s[12];
  for (int i = 0; i < 12; i++) s[i] = state12[i] % P;
  uint32_t len = *buffer_len, last = 0;
  for (size_t k = 0; k < n; k++) {
    buffer8[len++] = elems[k] % P;
    last = 0;
    if (len == 8) {
      for (int i = 0; i < 8; i++) s[i] = buffer8[i];
      if (hash_kind == 1) poseidon2(s);
      else poseidon(s);
      len = 0;
      last = 1;
    }
  }
  for (int i = 0; i < 12; i++) state12[i] = s[i];
  *buffer_len = len;
  *duplexed_last_out = last;
  return nullptr;
}
