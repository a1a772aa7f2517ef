/*
 * mp2_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).  See mp2_oracle.h
 * for provenance, pinning status and who may call this.  Every function cites the
 * SURVEY.md Appendix-A clause (restating plonky2 0.2.2, which is not vendored in the
 * reference tree) and the reference call site that constrains it.
 */
#include "mp2_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;

#define GL_P 0xFFFFFFFF00000001ULL /* mp2-common/src/group_hashing/utils.rs:51 */
#define GL_EPS 0xFFFFFFFFULL

/* ------------------------------------------------------------------------- */
/* A.1 field: p = 2^64 - 2^32 + 1                                             */
/* ------------------------------------------------------------------------- */
uint64_t orc_gl_canon(uint64_t a) { return a >= GL_P ? a - GL_P : a; }

static inline uint64_t gl_add(uint64_t a, uint64_t b) { /* canonical in -> canonical out */
  uint64_t s = a + b;
  if (s < a || s >= GL_P) s -= GL_P;
  return s;
}
static inline uint64_t gl_sub(uint64_t a, uint64_t b) {
  uint64_t d = a - b;
  if (a < b) d += GL_P;
  return d;
}
/* reduce128 exactly as A.1: result in [0,2^64), then canonicalised */
static inline uint64_t gl_reduce128(u128 x) {
  uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
  uint64_t hh = hi >> 32, hl = hi & GL_EPS;
  uint64_t t0 = lo - hh;
  t0 -= (0 - (uint64_t)(lo < hh)) & GL_EPS; /* branchless: the borrow is a coin flip, a branch mispredicts */
  uint64_t t1 = hl * GL_EPS;
  uint64_t r = t0 + t1;
  r += (0 - (uint64_t)(r < t0)) & GL_EPS;
  return r >= GL_P ? r - GL_P : r;
}
static inline uint64_t gl_mul(uint64_t a, uint64_t b) { return gl_reduce128((u128)a * b); }

uint64_t orc_gl_add(uint64_t a, uint64_t b) { return gl_add(orc_gl_canon(a), orc_gl_canon(b)); }
uint64_t orc_gl_sub(uint64_t a, uint64_t b) { return gl_sub(orc_gl_canon(a), orc_gl_canon(b)); }
uint64_t orc_gl_mul(uint64_t a, uint64_t b) { return gl_mul(a, b); }
uint64_t orc_gl_pow(uint64_t a, uint64_t e) {
  uint64_t r = 1, b = orc_gl_canon(a);
  while (e) {
    if (e & 1) r = gl_mul(r, b);
    b = gl_mul(b, b);
    e >>= 1;
  }
  return r;
}
uint64_t orc_gl_inv(uint64_t a) { return orc_gl_pow(a, GL_P - 2); }

/* POWER_OF_TWO_GENERATOR = 7^((p-1)/2^32); primitive_root_of_unity(k) = that^(2^(32-k)) */
uint64_t orc_gl_root_of_unity(uint32_t log_n) {
  uint64_t w = orc_gl_pow(7, (GL_P - 1) >> 32);
  for (uint32_t i = log_n; i < 32; i++) w = gl_mul(w, w);
  return w;
}

/* ------------------------------------------------------------------------- */
/* A.6 Poseidon (width 12, x^7, 4+22+4) -- constants regenerated, naive rounds */
/* ------------------------------------------------------------------------- */
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static inline uint32_t rotr32(uint32_t x, int r) { return r ? (x >> r) | (x << (32 - r)) : x; }

/* rand_chacha ChaCha8Rng: 8 rounds, 64-bit block counter in words 12-13, stream 0 */
static void chacha8_block(const uint32_t key[8], uint64_t counter, uint32_t out[16]) {
  uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
  for (int i = 0; i < 8; i++) in[4 + i] = key[i];
  in[12] = (uint32_t)counter;
  in[13] = (uint32_t)(counter >> 32);
  in[14] = 0;
  in[15] = 0;
  uint32_t x[16];
  memcpy(x, in, sizeof x);
#define QR(a, b, c, d)                                                                             \
  x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 12);      \
  x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 7);
  for (int r = 0; r < 4; r++) {
    QR(0, 4, 8, 12) QR(1, 5, 9, 13) QR(2, 6, 10, 14) QR(3, 7, 11, 15)
    QR(0, 5, 10, 15) QR(1, 6, 11, 12) QR(2, 7, 8, 13) QR(3, 4, 9, 14)
  }
#undef QR
  for (int i = 0; i < 16; i++) out[i] = x[i] + in[i];
}

static uint64_t POS_RC[360];
static int pos_rc_ready = 0;

/* RC[k] = ChaCha8Rng::seed_from_u64(0).gen_range(0..p), k = 0..359 */
void orc_poseidon_round_constants(uint64_t out[360]) {
  /* seed_from_u64(0): eight PCG32 outputs, little endian, form the 32-byte key */
  uint32_t key[8];
  uint64_t st = 0;
  for (int i = 0; i < 8; i++) {
    st = st * 6364136223846793005ULL + 11634580027462260723ULL;
    uint32_t xs = (uint32_t)(((st >> 18) ^ st) >> 27);
    key[i] = rotr32(xs, (int)(st >> 59));
  }
  uint32_t blk[16];
  uint64_t ctr = 0;
  int pos = 16;
  for (int k = 0; k < 360;) {
    if (pos == 16) {
      chacha8_block(key, ctr++, blk);
      pos = 0;
    }
    uint64_t v = (uint64_t)blk[pos] | ((uint64_t)blk[pos + 1] << 32); /* next_u64: lo then hi */
    pos += 2;
    u128 m = (u128)v * GL_P; /* UniformInt::sample_single, zone = p - 1 */
    if ((uint64_t)m <= GL_P - 1) out[k++] = (uint64_t)(m >> 64);
  }
}

static void pos_init(void) {
  if (pos_rc_ready) return;
#pragma omp critical(orc_pos_init)
  {
    if (!pos_rc_ready) {
      orc_poseidon_round_constants(POS_RC);
      __atomic_store_n(&pos_rc_ready, 1, __ATOMIC_RELEASE);
    }
  }
}

static const uint64_t POS_CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
static const uint64_t POS_DIAG[12] = {8, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

/* loose arithmetic (values in [0,2^64), canonicalised once per permutation) keeps the CPU baseline
 * honest: this is the same lazy-reduction discipline plonky2's scalar code uses */
static inline uint64_t gl_reduce128_loose(u128 x) {
  uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
  uint64_t hh = hi >> 32, hl = hi & GL_EPS;
  uint64_t t0 = lo - hh;
  t0 -= (0 - (uint64_t)(lo < hh)) & GL_EPS; /* branchless: the borrow is a coin flip, a branch mispredicts */
  uint64_t t1 = (hl << 32) - hl;
  uint64_t r = t0 + t1;
  r += (0 - (uint64_t)(r < t0)) & GL_EPS;
  return r;
}
static inline uint64_t gl_mul_loose(uint64_t a, uint64_t b) { return gl_reduce128_loose((u128)a * b); }
static inline uint64_t sbox7_loose(uint64_t x) {
  uint64_t x2 = gl_mul_loose(x, x), x4 = gl_mul_loose(x2, x2), x3 = gl_mul_loose(x, x2);
  return gl_mul_loose(x3, x4);
}
static inline uint64_t sbox7(uint64_t x) { return orc_gl_canon(sbox7_loose(x)); }
/* the full-round S-box layer, one multiplication step at a time over all 12 lanes: twelve independent
 * products per step instead of twelve serial x -> x^7 chains (the out-of-order core overlaps them) */
static inline void sbox7_layer_loose(uint64_t s[12]) {
  uint64_t x2[12], x3[12], x4[12];
  for (int i = 0; i < 12; i++) x2[i] = gl_mul_loose(s[i], s[i]);
  for (int i = 0; i < 12; i++) x4[i] = gl_mul_loose(x2[i], x2[i]);
  for (int i = 0; i < 12; i++) x3[i] = gl_mul_loose(s[i], x2[i]);
  for (int i = 0; i < 12; i++) s[i] = gl_mul_loose(x3[i], x4[i]);
}

/* out[r] = sum_i state[(i+r)%12] * CIRC[i] + state[r] * DIAG[r] + rc[r]   (mds_row_shf).
 * The matrix entries are tiny (<= 41, row sum 264), so -- as plonky2's scalar code does -- the state is cut
 * into 32-bit halves and each half goes through the matrix in plain 64-bit arithmetic (2^32 * 264 < 2^41, no
 * overflow); the two planes are recombined as lo + hi * 2^32 (< 2^74) and reduced once.  gcc vectorises the
 * two inner products.  (The first version accumulated twelve 128-bit products per row: ~2.5x slower.) */
static inline void pos_mds_rc(uint64_t s[12], const uint64_t *rc) {
  uint32_t lo[24], hi[24];
  uint64_t al[12], ah[12];
  for (int i = 0; i < 12; i++) {
    lo[i] = lo[i + 12] = (uint32_t)s[i];
    hi[i] = hi[i + 12] = (uint32_t)(s[i] >> 32);
    al[i] = ah[i] = 0;
  }
  al[0] = (uint64_t)lo[0] * (uint32_t)POS_DIAG[0];
  ah[0] = (uint64_t)hi[0] * (uint32_t)POS_DIAG[0];
  for (int i = 0; i < 12; i++) { /* outer-product order: the inner loop is 12 independent 32x32->64 multiply-adds */
    const uint32_t c = (uint32_t)POS_CIRC[i];
    for (int r = 0; r < 12; r++) {
      al[r] += (uint64_t)lo[i + r] * c;
      ah[r] += (uint64_t)hi[i + r] * c;
    }
  }
  for (int r = 0; r < 12; r++) s[r] = gl_reduce128_loose((u128)al[r] + ((u128)ah[r] << 32) + rc[r]);
}

static const uint64_t POS_ZERO_RC[12] = {0};

void orc_poseidon_permute(uint64_t s[12]) {
  pos_init();
  /* round r: +RC[r], S-box, MDS.  The constants of round r+1 are added right after the MDS of
   * round r (same values, fewer passes over the state). */
  for (int i = 0; i < 12; i++) s[i] = gl_add(orc_gl_canon(s[i]), POS_RC[i]);
  for (int round = 0; round < 30; round++) {
    if (round < 4 || round >= 26) {
      sbox7_layer_loose(s);
    } else {
      s[0] = sbox7_loose(s[0]);
    }
    pos_mds_rc(s, round < 29 ? POS_RC + 12 * (round + 1) : POS_ZERO_RC);
  }
  for (int i = 0; i < 12; i++) s[i] = orc_gl_canon(s[i]);
}

/* ------------------------------------------------------------------------- */
/* A.7 Poseidon2 (Horizen-Labs Goldilocks t=12 instance; provenance unconfirmed) */
/* ------------------------------------------------------------------------- */
static uint64_t P2_RC[118];
static int p2_rc_ready = 0;
static const uint64_t P2_DIAG[12] = {
    0xc3b6c08e23ba9300ULL, 0xd84b5de94a324fb6ULL, 0x0d0c371c5b35b84fULL, 0x7964f570e7188037ULL,
    0x5daf18bbd996604bULL, 0x6743bc47b9595257ULL, 0x5528b9362c59bb70ULL, 0xac45e25b7127b68bULL,
    0xa2077d7dfbb606b5ULL, 0xf3faac6faee378aeULL, 0x0c6388b51545e883ULL, 0xd27dbb6944917b60ULL};

void orc_poseidon2_diag(uint64_t out[12]) { memcpy(out, P2_DIAG, sizeof P2_DIAG); }

/* Poseidon Grain LFSR (80 bits) parameterised for field=1, sbox=0, n=64, t=12, R_F=8, R_P=22 */
typedef struct {
  uint8_t b[80];
  int head;
} grain_t;
static int grain_step(grain_t *g) {
  uint8_t *b = g->b;
  int h = g->head;
#define GB(i) b[(h + (i)) % 80]
  int nb = GB(62) ^ GB(51) ^ GB(38) ^ GB(23) ^ GB(13) ^ GB(0);
#undef GB
  b[h] = (uint8_t)nb; /* overwrite oldest, advance head => new bit becomes position 79 */
  g->head = (h + 1) % 80;
  return nb;
}
static int grain_bit(grain_t *g) {
  for (;;) {
    int a = grain_step(g);
    int c = grain_step(g);
    if (a) return c;
  }
}
void orc_poseidon2_round_constants(uint64_t out[118]) {
  grain_t g;
  g.head = 0;
  int pos = 0;
  const uint32_t fields[6][2] = {{1, 2}, {0, 4}, {64, 12}, {12, 12}, {8, 10}, {22, 10}};
  for (int f = 0; f < 6; f++)
    for (int i = (int)fields[f][1] - 1; i >= 0; i--) g.b[pos++] = (fields[f][0] >> i) & 1;
  while (pos < 80) g.b[pos++] = 1;
  for (int i = 0; i < 160; i++) grain_step(&g);
  for (int k = 0; k < 118;) {
    uint64_t v = 0;
    for (int i = 0; i < 64; i++) v = (v << 1) | (uint64_t)grain_bit(&g);
    if (v < GL_P) out[k++] = v;
  }
}
static void p2_init(void) {
  if (p2_rc_ready) return;
#pragma omp critical(orc_p2_init)
  {
    if (!p2_rc_ready) {
      orc_poseidon2_round_constants(P2_RC);
      __atomic_store_n(&p2_rc_ready, 1, __ATOMIC_RELEASE);
    }
  }
}

/* M_E: M4 = [[5,7,1,3],[4,6,1,1],[1,3,5,7],[1,1,4,6]] on each 4-lane chunk, then add column sums.  Row sums are
 * <= 64, so the network runs on the 32-bit halves of the (loose) state in plain 64-bit adds and every lane is
 * recombined and reduced once. */
static inline void p2_m4_half(uint64_t x[12]) {
  for (int c = 0; c < 12; c += 4) {
    uint64_t x0 = x[c], x1 = x[c + 1], x2 = x[c + 2], x3 = x[c + 3];
    uint64_t t0 = x0 + x1, t1 = x2 + x3;
    uint64_t t2 = 2 * x1 + t1, t3 = 2 * x3 + t0;
    uint64_t t4 = 4 * t1 + t3, t5 = 4 * t0 + t2;
    x[c] = t3 + t5;
    x[c + 1] = t5;
    x[c + 2] = t2 + t4;
    x[c + 3] = t4;
  }
  uint64_t col[4];
  for (int l = 0; l < 4; l++) col[l] = x[l] + x[4 + l] + x[8 + l];
  for (int i = 0; i < 12; i++) x[i] += col[i % 4];
}
static void p2_external(uint64_t s[12]) {
  uint64_t lo[12], hi[12];
  for (int i = 0; i < 12; i++) {
    lo[i] = s[i] & GL_EPS;
    hi[i] = s[i] >> 32;
  }
  p2_m4_half(lo);
  p2_m4_half(hi);
  for (int i = 0; i < 12; i++) s[i] = gl_reduce128_loose((u128)lo[i] + ((u128)hi[i] << 32));
}
/* M_I: out[i] = state[i] * mu_i + sum(state); the sum rides on the 128-bit product */
static void p2_internal(uint64_t s[12]) {
  u128 acc = 0;
  for (int i = 0; i < 12; i++) acc += s[i];
  const uint64_t sum = gl_reduce128_loose(acc);
  for (int i = 0; i < 12; i++) s[i] = gl_reduce128_loose((u128)s[i] * P2_DIAG[i] + sum);
}
/* loose arithmetic throughout (values in [0, 2^64)), canonicalised once at the end -- as in orc_poseidon_permute */
void orc_poseidon2_permute(uint64_t s[12]) {
  p2_init();
  const uint64_t *rc = P2_RC;
  p2_external(s);
  for (int r = 0; r < 4; r++) {
    for (int i = 0; i < 12; i++) s[i] = gl_reduce128_loose((u128)s[i] + *rc++);
    sbox7_layer_loose(s);
    p2_external(s);
  }
  for (int r = 0; r < 22; r++) {
    s[0] = sbox7_loose(gl_reduce128_loose((u128)s[0] + *rc++));
    p2_internal(s);
  }
  for (int r = 0; r < 4; r++) {
    for (int i = 0; i < 12; i++) s[i] = gl_reduce128_loose((u128)s[i] + *rc++);
    sbox7_layer_loose(s);
    p2_external(s);
  }
  for (int i = 0; i < 12; i++) s[i] = orc_gl_canon(s[i]);
}

void orc_permute(uint32_t kind, uint64_t s[12]) {
  if (kind == ORC_HASH_POSEIDON2) orc_poseidon2_permute(s);
  else orc_poseidon_permute(s);
}

/* ------------------------------------------------------------------------- */
/* A.5 sponge wrapper: rate 8, capacity 4, OVERWRITE absorb, squeeze state[0..4] */
/* (mirrors mp2-common/src/poseidon.rs:151-171 and mp2-common/src/hash.rs:24-45) */
/* ------------------------------------------------------------------------- */
void orc_hash_no_pad(uint32_t kind, const uint64_t *in, size_t len, uint64_t out[4]) {
  uint64_t st[12] = {0};
  for (size_t off = 0; off < len; off += 8) {
    size_t m = len - off < 8 ? len - off : 8;
    for (size_t i = 0; i < m; i++) st[i] = orc_gl_canon(in[off + i]);
    orc_permute(kind, st);
  }
  for (int i = 0; i < 4; i++) out[i] = st[i];
}
/* hash_pad: append 1, zeros until (len+1) % 8 == 0, then 1 (circuit_set.rs:149-151 uses hash_pad(&[])) */
void orc_hash_pad(uint32_t kind, const uint64_t *in, size_t len, uint64_t out[4]) {
  size_t padded = len + 1;
  while ((padded + 1) % 8 != 0) padded++;
  padded++;
  uint64_t *buf = (uint64_t *)calloc(padded, sizeof(uint64_t));
  if (len) memcpy(buf, in, len * sizeof(uint64_t));
  buf[len] = 1;
  buf[padded - 1] = 1;
  orc_hash_no_pad(kind, buf, padded, out);
  free(buf);
}
void orc_hash_or_noop(uint32_t kind, const uint64_t *in, size_t len, uint64_t out[4]) {
  if (len <= 4) {
    for (size_t i = 0; i < 4; i++) out[i] = i < len ? orc_gl_canon(in[i]) : 0;
  } else {
    orc_hash_no_pad(kind, in, len, out);
  }
}
void orc_two_to_one(uint32_t kind, const uint64_t a[4], const uint64_t b[4], uint64_t out[4]) {
  uint64_t st[12] = {0};
  for (int i = 0; i < 4; i++) {
    st[i] = orc_gl_canon(a[i]);
    st[4 + i] = orc_gl_canon(b[i]);
  }
  orc_permute(kind, st);
  for (int i = 0; i < 4; i++) out[i] = st[i];
}

/* ------------------------------------------------------------------------- */
/* A.2 transforms                                                             */
/* ------------------------------------------------------------------------- */
static inline size_t bitrev(size_t x, uint32_t bits) {
  if (bits == 0) return 0;
  uint64_t v = (uint64_t)x; /* byte swap + in-byte swaps: O(1) instead of a loop over the bits */
  v = __builtin_bswap64(v);
  v = ((v & 0xF0F0F0F0F0F0F0F0ULL) >> 4) | ((v & 0x0F0F0F0F0F0F0F0FULL) << 4);
  v = ((v & 0xCCCCCCCCCCCCCCCCULL) >> 2) | ((v & 0x3333333333333333ULL) << 2);
  v = ((v & 0xAAAAAAAAAAAAAAAAULL) >> 1) | ((v & 0x5555555555555555ULL) << 1);
  return (size_t)(v >> (64 - bits));
}

/* FftRootTable (plonky2_field fft_root_table): per layer s the powers w_{2^s}^k, k < 2^(s-1), computed once per
 * size and shared by every column and thread -- what plonky2 passes down as `fft_root_table`.  Layout: the
 * table of layer s starts at offset 2^(s-1) - 1 (layers 1..32 would need 2^32 entries; sizes are built on demand
 * up to the largest transform seen). */
static uint64_t *g_roots = NULL;
static uint32_t g_roots_log = 0;
static const uint64_t *root_table(uint32_t log_n) {
  uint32_t have = __atomic_load_n(&g_roots_log, __ATOMIC_ACQUIRE);
  if (have >= log_n && g_roots) return g_roots;
#pragma omp critical(orc_root_table)
  {
    if (g_roots_log < log_n || !g_roots) {
      uint64_t *t = (uint64_t *)malloc(sizeof(uint64_t) * ((size_t)1 << log_n));
      for (uint32_t s = 1; s <= log_n; s++) {
        size_t half = (size_t)1 << (s - 1);
        uint64_t w = orc_gl_root_of_unity(s), *row = t + half - 1;
        row[0] = 1;
        for (size_t k = 1; k < half; k++) row[k] = gl_mul(row[k - 1], w);
      }
      /* the old table (if any) is leaked on purpose: another thread may still be reading it */
      __atomic_store_n(&g_roots, t, __ATOMIC_RELEASE);
      __atomic_store_n(&g_roots_log, log_n, __ATOMIC_RELEASE);
    }
  }
  return g_roots;
}

/* radix-2 decimation in time on the bit-reversed input (fft_classic).  `zero_log`: the caller promises that
 * only the first 2^(log_n - zero_log) inputs are non-zero (fft_with_options' zero_factor): after the bit reversal
 * every aligned group of 2^zero_log entries holds one value followed by zeros, so the first zero_log layers are a
 * broadcast -- plonky2 skips them the same way. */
static void fft_zero_padded(uint64_t *v, uint32_t log_n, uint32_t zero_log) {
  size_t n = (size_t)1 << log_n;
  for (size_t i = 0; i < n; i++) {
    size_t j = bitrev(i, log_n);
    if (i < j) {
      uint64_t t = v[i];
      v[i] = v[j];
      v[j] = t;
    }
  }
  for (size_t i = 0; i < n; i++) v[i] = orc_gl_canon(v[i]);
  const uint64_t *roots = root_table(log_n);
  if (zero_log) {
    size_t g = (size_t)1 << zero_log;
    for (size_t base = 0; base < n; base += g)
      for (size_t k = 1; k < g; k++) v[base + k] = v[base];
  }
  for (uint32_t s = zero_log + 1; s <= log_n; s++) {
    size_t half = (size_t)1 << (s - 1);
    const uint64_t *tw = roots + half - 1;
    for (size_t base = 0; base < n; base += 2 * half)
      for (size_t k = 0; k < half; k++) {
        uint64_t a = v[base + k], b = gl_mul(v[base + k + half], tw[k]);
        v[base + k] = gl_add(a, b);
        v[base + k + half] = gl_sub(a, b);
      }
  }
}

void orc_fft(uint64_t *v, uint32_t log_n) { fft_zero_padded(v, log_n, 0); }

/* ifft = fft, then coeffs[i] = fft[(n-i)%n] * n^-1 */
void orc_ifft(uint64_t *v, uint32_t log_n) {
  size_t n = (size_t)1 << log_n;
  orc_fft(v, log_n);
  uint64_t ninv = orc_gl_inv((uint64_t)n);
  for (size_t i = 1; i < n - i; i++) {
    uint64_t t = v[i];
    v[i] = v[n - i];
    v[n - i] = t;
  }
  for (size_t i = 0; i < n; i++) v[i] = gl_mul(v[i], ninv);
}

/* lde(r) then coset_fft(shift): out[i] = P(shift * w_N^i), natural order */
void orc_coset_lde(const uint64_t *coeffs, uint32_t log_n, uint32_t rate_bits, uint64_t shift,
                   uint64_t *out) {
  size_t n = (size_t)1 << log_n, N = n << rate_bits;
  uint64_t pw = 1;
  for (size_t j = 0; j < n; j++) {
    out[j] = gl_mul(coeffs[j], pw);
    pw = gl_mul(pw, shift);
  }
  memset(out + n, 0, (N - n) * sizeof(uint64_t));
  fft_zero_padded(out, log_n + rate_bits, rate_bits);
}

void orc_eval_naive(const uint64_t *coeffs, size_t n, uint64_t shift, uint64_t w, size_t n_out,
                    uint64_t *out) {
  uint64_t x = orc_gl_canon(shift);
  for (size_t i = 0; i < n_out; i++) {
    uint64_t acc = 0;
    for (size_t j = n; j-- > 0;) acc = gl_add(gl_mul(acc, x), orc_gl_canon(coeffs[j]));
    out[i] = acc;
    x = gl_mul(x, w);
  }
}

/* ------------------------------------------------------------------------- */
/* A.4 MerkleTree::new / prove -- plonky2's interleaved digest layout          */
/* (called directly at circuit_set.rs:189; proof consumed at circuit_set.rs:216) */
/* ------------------------------------------------------------------------- */
static void fill_subtree(uint64_t *buf, size_t buf_len /* digests */, const uint64_t *leaves,
                         size_t nleaves, size_t leaf_len, uint32_t kind, uint64_t out[4],
                         int depth) {
  if (buf_len == 0) {
    orc_hash_or_noop(kind, leaves, leaf_len, out);
    return;
  }
  size_t half = buf_len / 2;
  uint64_t *lbuf = buf, *rbuf = buf + 4 * half;
  uint64_t *lmem = lbuf + 4 * (half - 1); /* last digest of the left half */
  uint64_t *rmem = rbuf;                  /* first digest of the right half */
  uint64_t ld[4], rd[4];
  if (depth > 0 && nleaves >= 64) {
#pragma omp task shared(ld) firstprivate(lbuf, half, leaves, nleaves, leaf_len, kind, depth)
    fill_subtree(lbuf, half - 1, leaves, nleaves / 2, leaf_len, kind, ld, depth - 1);
#pragma omp task shared(rd) firstprivate(rbuf, half, leaves, nleaves, leaf_len, kind, depth)
    fill_subtree(rbuf + 4, half - 1, leaves + (nleaves / 2) * leaf_len, nleaves / 2, leaf_len,
                 kind, rd, depth - 1);
#pragma omp taskwait
  } else {
    fill_subtree(lbuf, half - 1, leaves, nleaves / 2, leaf_len, kind, ld, 0);
    fill_subtree(rbuf + 4, half - 1, leaves + (nleaves / 2) * leaf_len, nleaves / 2, leaf_len,
                 kind, rd, 0);
  }
  memcpy(lmem, ld, 32);
  memcpy(rmem, rd, 32);
  orc_two_to_one(kind, ld, rd, out);
}

static int log2_exact(size_t n) {
  if (n == 0 || (n & (n - 1))) return -1;
  int l = 0;
  while (((size_t)1 << l) < n) l++;
  return l;
}

int orc_merkle_new(const uint64_t *leaves, size_t nleaves, size_t leaf_len, uint32_t cap_height,
                   uint32_t kind, uint64_t *digests_out, uint64_t *cap_out, int nthreads) {
  int lg = log2_exact(nleaves);
  if (lg < 0 || (int)cap_height > lg) return -1;
  pos_init();
  p2_init();
  size_t ncap = (size_t)1 << cap_height;
  size_t sub_leaves = nleaves >> cap_height;
  size_t sub_digests = 2 * (sub_leaves - 1); /* digests.len() / ncap */
  if (nthreads < 1) nthreads = 1;
  int depth = 0;
  while (((size_t)1 << depth) * ncap < (size_t)nthreads * 4 && depth < 16) depth++;
#pragma omp parallel num_threads(nthreads)
#pragma omp single
  {
    for (size_t s = 0; s < ncap; s++) {
#pragma omp task firstprivate(s)
      fill_subtree(digests_out + 4 * s * sub_digests, sub_digests,
                   leaves + s * sub_leaves * leaf_len, sub_leaves, leaf_len, kind, cap_out + 4 * s,
                   depth);
    }
  }
  return 0;
}

/* closed-form sibling indices (A.4): layer i pair k at digests 2q, 2q+1, q = (k<<(i+1)) + (1<<i) - 1 */
int orc_merkle_prove(const uint64_t *digests, size_t nleaves, uint32_t cap_height,
                     size_t leaf_index, uint64_t *siblings_out) {
  int lg = log2_exact(nleaves);
  if (lg < 0 || (int)cap_height > lg || leaf_index >= nleaves) return -1;
  uint32_t h = (uint32_t)lg - cap_height;
  size_t sub_digests = 2 * (((size_t)1 << h) - 1);
  const uint64_t *sub = digests + 4 * (leaf_index >> h) * sub_digests;
  size_t pair_index = leaf_index & (((size_t)1 << h) - 1);
  for (uint32_t i = 0; i < h; i++) {
    size_t parity = pair_index & 1;
    pair_index >>= 1;
    size_t q = (pair_index << (i + 1)) + ((size_t)1 << i) - 1;
    memcpy(siblings_out + 4 * i, sub + 4 * (2 * q + (1 - parity)), 32);
  }
  return (int)h;
}

int orc_merkle_verify(const uint64_t *leaf, size_t leaf_len, size_t leaf_index,
                      const uint64_t *siblings, size_t nsib, uint32_t kind, uint64_t root[4]) {
  uint64_t cur[4], nxt[4];
  orc_hash_or_noop(kind, leaf, leaf_len, cur);
  size_t idx = leaf_index;
  for (size_t i = 0; i < nsib; i++) {
    if (idx & 1) orc_two_to_one(kind, siblings + 4 * i, cur, nxt);
    else orc_two_to_one(kind, cur, siblings + 4 * i, nxt);
    memcpy(cur, nxt, 32);
    idx >>= 1;
  }
  memcpy(root, cur, 32);
  return (int)idx; /* cap index */
}

/* ------------------------------------------------------------------------- */
/* a1/a2 PolynomialBatch::from_values / from_coeffs                            */
/* ------------------------------------------------------------------------- */
int orc_commit(const uint64_t *const *cols, size_t ncols, uint32_t log_n, uint32_t rate_bits,
               uint32_t cap_height, uint32_t kind, int from_coeffs, uint64_t *coeffs_out,
               uint64_t *leaves_out, uint64_t *digests_out, uint64_t *cap_out, int nthreads) {
  size_t n = (size_t)1 << log_n, N = n << rate_bits;
  uint32_t log_N = log_n + rate_bits;
  if (cap_height > log_N || ncols == 0) return -1;
  if (nthreads < 1) nthreads = 1;
  pos_init();
  p2_init();
  uint64_t *lde = (uint64_t *)malloc(sizeof(uint64_t) * N * ncols); /* column-major */
  uint64_t *leaves = leaves_out ? leaves_out : (uint64_t *)malloc(sizeof(uint64_t) * N * ncols);
  if (!lde || !leaves) return -1;
  /* "IFFT" + "FFT + blinding" scopes: one task per column, like rayon over polynomials */
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (size_t c = 0; c < ncols; c++) {
    uint64_t *tmp = (uint64_t *)malloc(sizeof(uint64_t) * n);
    memcpy(tmp, cols[c], sizeof(uint64_t) * n);
    if (!from_coeffs) orc_ifft(tmp, log_n);
    else
      for (size_t i = 0; i < n; i++) tmp[i] = orc_gl_canon(tmp[i]);
    if (coeffs_out) memcpy(coeffs_out + c * n, tmp, sizeof(uint64_t) * n);
    orc_coset_lde(tmp, log_n, rate_bits, 7 /* coset_shift() = MULTIPLICATIVE_GROUP_GENERATOR */,
                  lde + c * N);
    free(tmp);
  }
  /* "transpose LDEs" + reverse_index_bits_in_place: leaves[i] = row bitrev(i)  (A.3) */
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (size_t i = 0; i < N; i++) {
    size_t src = bitrev(i, log_N);
    for (size_t c = 0; c < ncols; c++) leaves[i * ncols + c] = lde[c * N + src];
  }
  free(lde);
  int rc = orc_merkle_new(leaves, N, ncols, cap_height, kind, digests_out, cap_out, nthreads);
  if (!leaves_out) free(leaves);
  return rc;
}

/* ------------------------------------------------------------------------- */
/* a10 / 8(f).2: FRI commit phase pieces (plonky2 fri/prover.rs, restated)      */
/* ------------------------------------------------------------------------- */
void orc_ext_mul(const uint64_t a[2], const uint64_t b[2], uint64_t out[2]) { /* X^2 = W = 7 */
  uint64_t a0 = orc_gl_canon(a[0]), a1 = orc_gl_canon(a[1]), b0 = orc_gl_canon(b[0]), b1 = orc_gl_canon(b[1]);
  uint64_t c0 = gl_add(gl_mul(a0, b0), gl_mul(7, gl_mul(a1, b1)));
  uint64_t c1 = gl_add(gl_mul(a0, b1), gl_mul(a1, b0));
  out[0] = c0;
  out[1] = c1;
}

void orc_fri_fold(const uint64_t *coeffs, size_t m, uint32_t arity_bits, const uint64_t beta[2], uint64_t *out) {
  size_t arity = (size_t)1 << arity_bits;
  for (size_t j = 0; j < m >> arity_bits; j++) {
    uint64_t acc[2] = {0, 0}; /* reduce_with_powers: Horner from the last term */
    for (size_t t = arity; t-- > 0;) {
      uint64_t prod[2];
      orc_ext_mul(acc, beta, prod);
      acc[0] = gl_add(prod[0], orc_gl_canon(coeffs[2 * ((j << arity_bits) + t)]));
      acc[1] = gl_add(prod[1], orc_gl_canon(coeffs[2 * ((j << arity_bits) + t) + 1]));
    }
    out[2 * j] = acc[0];
    out[2 * j + 1] = acc[1];
  }
}

/* PolynomialBatch::prove_openings, up to its call of fri_proof (plonky2 0.2.2 fri/oracle.rs), with the
 * ReducingFactor bookkeeping of util/reducing.rs:
 *   for each batch (point z, polynomials f_0..f_{c-1}):
 *     composition = reduce_polys_base = sum_j alpha^j f_j            (count += c)
 *     quotient    = composition.divide_by_linear(z); quotient.push(0)
 *     final_poly  = final_poly * alpha^count  (shift_poly; count = 0)  +  quotient            */
int orc_fri_combine(const uint64_t *const *polys, const uint32_t *batch_sizes, size_t nbatches,
                    const uint64_t *points, const uint64_t alpha[2], size_t n, uint64_t *out) {
  if (!n) return -1;
  uint64_t *comp = (uint64_t *)malloc(sizeof(uint64_t) * 2 * n);
  memset(out, 0, sizeof(uint64_t) * 2 * n);
  size_t at = 0;
  for (size_t i = 0; i < nbatches; i++) {
    const uint64_t *z = points + 2 * i;
    uint64_t pw[2] = {1, 0}, shift[2];
    memset(comp, 0, sizeof(uint64_t) * 2 * n);
    for (uint32_t j = 0; j < batch_sizes[i]; j++, at++) {
      for (size_t m = 0; m < n; m++) { /* poly.mul_extension(base_power) */
        uint64_t v = orc_gl_canon(polys[at][m]);
        comp[2 * m] = gl_add(comp[2 * m], gl_mul(v, pw[0]));
        comp[2 * m + 1] = gl_add(comp[2 * m + 1], gl_mul(v, pw[1]));
      }
      uint64_t nx[2];
      orc_ext_mul(pw, alpha, nx);
      pw[0] = nx[0], pw[1] = nx[1];
    }
    shift[0] = pw[0], shift[1] = pw[1]; /* alpha^count */
    /* divide_by_linear: b_k = b_(k+1) * z + a_k from the top; the quotient's coefficient k is b_(k+1) */
    uint64_t acc[2] = {0, 0};
    for (size_t m = n; m-- > 0;) {
      uint64_t scaled[2], prod[2];
      orc_ext_mul(out + 2 * m, shift, scaled);
      out[2 * m] = gl_add(scaled[0], acc[0]); /* final[m] * alpha^count + quotient[m]; quotient[n-1] = 0 */
      out[2 * m + 1] = gl_add(scaled[1], acc[1]);
      orc_ext_mul(acc, z, prod);
      acc[0] = gl_add(prod[0], comp[2 * m]);
      acc[1] = gl_add(prod[1], comp[2 * m + 1]);
    }
  }
  for (size_t m = 0; m < 2 * n; m++) out[m] = orc_gl_canon(out[m]);
  free(comp);
  return 0;
}

/* the extension's roots of unity of order <= 2^32 are the base field's (EXT_POWER_OF_TWO_GENERATOR^2 =
 * POWER_OF_TWO_GENERATOR), so the transform acts on the two components separately */
void orc_coset_fft_ext(const uint64_t *coeffs, uint32_t log_m, uint64_t shift, uint64_t *values) {
  size_t m = (size_t)1 << log_m;
  uint64_t *tmp = (uint64_t *)malloc(sizeof(uint64_t) * m);
  for (int comp = 0; comp < 2; comp++) {
    uint64_t pw = 1;
    for (size_t j = 0; j < m; j++) {
      tmp[j] = gl_mul(coeffs[2 * j + comp], pw);
      pw = gl_mul(pw, shift);
    }
    orc_fft(tmp, log_m);
    for (size_t j = 0; j < m; j++) values[2 * j + comp] = tmp[j];
  }
  free(tmp);
}

void orc_fri_layer_leaves(const uint64_t *values, uint32_t log_m, uint32_t arity_bits, uint64_t *leaves) {
  size_t m = (size_t)1 << log_m;
  for (size_t i = 0; i < m; i++) { /* leaf i>>ab, slot i & (arity-1): value bitrev(i), flattened [a0,a1] */
    size_t src = bitrev(i, log_m);
    leaves[2 * i] = orc_gl_canon(values[2 * src]);
    leaves[2 * i + 1] = orc_gl_canon(values[2 * src + 1]);
  }
  (void)arity_bits; /* consecutive groups of 2^arity_bits pairs are the leaves */
}

int orc_fri_pow(uint32_t kind, const uint64_t state[12], uint32_t pos, uint32_t min_lz, uint64_t start,
                uint64_t limit, uint64_t *witness) {
  for (uint64_t c = start; c < limit; c++) {
    uint64_t st[12];
    memcpy(st, state, sizeof st);
    st[pos] = c;
    orc_permute(kind, st);
    uint64_t v = orc_gl_canon(st[7]); /* duplex_state.squeeze().iter().last() with RATE = 8 */
    uint32_t lz = v ? (uint32_t)__builtin_clzll(v) : 64;
    if (lz >= min_lz) {
      *witness = c;
      return 0;
    }
  }
  return -1;
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
