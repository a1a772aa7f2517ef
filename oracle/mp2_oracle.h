/*
 * mp2_oracle.h -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * Plain-C restatement of the plonky2 0.2.2 polynomial-batch commitment path that
 * every proof of Lagrange-Labs/mapreduce-plonky2 goes through.  The algorithm
 * itself lives in third-party crates that are NOT vendored under /root/reference:
 *   plonky2 / plonky2_field / plonky2_util 0.2.x  and  poseidon2_plonky2 0.1.0,
 *   git+https://github.com/Lagrange-Labs/plonky2?branch=upstream
 *   #22c42f64367e8f087e565bdb664525910a62fc76   (Cargo.toml:63,114-117; Cargo.lock:4716-4719)
 * so this file restates the published algorithm and is
 * anchored on the reference's call sites:
 *   recursion-framework/src/universal_verifier_gadget/circuit_set.rs:173-237 (MerkleTree::new / prove)
 *   recursion-framework/src/universal_verifier_gadget/circuit_set.rs:136-158 (circuit digest formula)
 *   mp2-common/src/poseidon.rs:136-172, mp2-common/src/hash.rs:16-46       (sponge semantics)
 *   mp2-common/src/lib.rs:36-47                                            (F, D, C, standard config)
 *   mp2-common/src/group_hashing/utils.rs:11,51                            (field order, 2/3 mod p)
 *
 * PARITY STATUS: "parity unpinned" at the commitment boundary -- the reference's
 * own tests hold no golden vectors for NTT/LDE/Merkle.  What IS
 * pinned: the 360 Poseidon round constants (regenerated from ChaCha8Rng seed 0 and
 * checked against the published table anchors), plonky2's three Poseidon
 * permutation test vectors, the Horizen-Labs Poseidon2 t=12 KAT, and the
 * Goldilocks generators (tests/test_oracle_pins.py).  Structurally (the kind of pin the reference's own
 * tests use -- prove then verify): a FRI proof assembled from this library's commitments, prove_openings
 * combination, commit phase and Merkle paths is accepted by a by-definition restatement of plonky2's FRI
 * verifier, and tampered proofs are rejected (tests/test_fri_prove_verify.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may call into this library.
 */
#ifndef MP2_ORACLE_H
#define MP2_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_HASH_POSEIDON 0u
#define ORC_HASH_POSEIDON2 1u

/* ---- field (A.1) ---- */
uint64_t orc_gl_add(uint64_t a, uint64_t b);
uint64_t orc_gl_sub(uint64_t a, uint64_t b);
uint64_t orc_gl_mul(uint64_t a, uint64_t b);
uint64_t orc_gl_pow(uint64_t a, uint64_t e);
uint64_t orc_gl_inv(uint64_t a);
uint64_t orc_gl_canon(uint64_t a);
uint64_t orc_gl_root_of_unity(uint32_t log_n); /* primitive_root_of_unity(log_n) */

/* ---- permutations (A.6, A.7) ---- */
void orc_poseidon_round_constants(uint64_t out[360]); /* regenerated, ChaCha8Rng(0) */
void orc_poseidon_permute(uint64_t state[12]);        /* naive schedule */
void orc_poseidon2_round_constants(uint64_t out[118]); /* Grain LFSR */
void orc_poseidon2_diag(uint64_t out[12]);
void orc_poseidon2_permute(uint64_t state[12]);
void orc_permute(uint32_t hash_kind, uint64_t state[12]);

/* ---- sponge wrapper (A.5) ---- */
void orc_hash_no_pad(uint32_t hash_kind, const uint64_t *in, size_t len, uint64_t out[4]);
void orc_hash_pad(uint32_t hash_kind, const uint64_t *in, size_t len, uint64_t out[4]);
void orc_hash_or_noop(uint32_t hash_kind, const uint64_t *in, size_t len, uint64_t out[4]);
void orc_two_to_one(uint32_t hash_kind, const uint64_t a[4], const uint64_t b[4], uint64_t out[4]);

/* ---- transforms (A.2), in place on one column ---- */
void orc_fft(uint64_t *v, uint32_t log_n);  /* v[i] <- sum_j v[j] w^(ij), natural order */
void orc_ifft(uint64_t *v, uint32_t log_n); /* inverse, natural order */
/* coeffs (n = 2^log_n) -> out (N = n << rate_bits): out[i] = P(shift * w_N^i) */
void orc_coset_lde(const uint64_t *coeffs, uint32_t log_n, uint32_t rate_bits, uint64_t shift,
                   uint64_t *out);
/* O(n^2) evaluation by definition, for cross-checks on tiny sizes */
void orc_eval_naive(const uint64_t *coeffs, size_t n, uint64_t shift, uint64_t w, size_t n_out,
                    uint64_t *out);

/* ---- Merkle tree (A.4) ---- */
/* leaves: row-major nleaves x leaf_len.  digests_out: 2*(nleaves - 2^cap_height) x 4.
 * cap_out: 2^cap_height x 4.  Returns 0 on success, -1 on bad arguments
 * (nleaves not a power of two, cap_height > log2(nleaves)). */
int orc_merkle_new(const uint64_t *leaves, size_t nleaves, size_t leaf_len, uint32_t cap_height,
                   uint32_t hash_kind, uint64_t *digests_out, uint64_t *cap_out, int nthreads);
/* siblings_out: (log2(nleaves) - cap_height) x 4; returns number of siblings or -1 */
int orc_merkle_prove(const uint64_t *digests, size_t nleaves, uint32_t cap_height,
                     size_t leaf_index, uint64_t *siblings_out);
/* recompute the cap entry index and value from a leaf and its proof */
int orc_merkle_verify(const uint64_t *leaf, size_t leaf_len, size_t leaf_index,
                      const uint64_t *siblings, size_t nsiblings, uint32_t hash_kind,
                      uint64_t root_out[4]);

/* ---- PolynomialBatch::from_values / from_coeffs (a1, a2) ----
 * cols: ncols pointers to n = 2^log_n u64 each (values or coeffs).
 * coeffs_out: ncols x n (column-major, natural order) or NULL.
 * leaves_out: (n << rate_bits) x ncols row-major, bit-reversed row order, or NULL.
 * digests_out / cap_out as orc_merkle_new.  Returns 0 / -1. */
int orc_commit(const uint64_t *const *cols, size_t ncols, uint32_t log_n, uint32_t rate_bits,
               uint32_t cap_height, uint32_t hash_kind, int from_coeffs, uint64_t *coeffs_out,
               uint64_t *leaves_out, uint64_t *digests_out, uint64_t *cap_out, int nthreads);

/* ---- FRI commit phase (a10; plonky2 fri/prover.rs fri_committed_trees) -- "next" row 8(f).2 --------
 * Extension field GF(p^2) = F[X]/(X^2 - 7) (QuadraticExtension<GoldilocksField>, D = 2 at
 * mp2-common/src/lib.rs:36); elements are interleaved pairs [a0, a1]. */
void orc_ext_mul(const uint64_t a[2], const uint64_t b[2], uint64_t out[2]);
/* coeffs' = chunks_exact(2^arity_bits).map(|c| reduce_with_powers(c, beta)):  out[j] = sum_t in[(j<<ab)+t]*beta^t */
void orc_fri_fold(const uint64_t *coeffs, size_t m, uint32_t arity_bits, const uint64_t beta[2], uint64_t *out);
/* PolynomialCoeffs<F::Extension>::coset_fft(shift): values[i] = P(shift * w_m^i), natural order */
void orc_coset_fft_ext(const uint64_t *coeffs, uint32_t log_m, uint64_t shift, uint64_t *values);
/* reverse_index_bits_in_place(values); chunks(2^arity_bits).map(flatten): (m >> ab) leaves of (2 << ab) elements */
void orc_fri_layer_leaves(const uint64_t *values, uint32_t log_m, uint32_t arity_bits, uint64_t *leaves);

/* PolynomialBatch::prove_openings before fri_proof: the alpha-batched quotient
 *   final_poly = sum_i alpha^(k_i) (F_i(X) - F_i(z_i)) / (X - z_i),   F_i = sum_j alpha^j f_ij
 * polys: the batches' polynomials concatenated (pointers to n base-field coefficients); points: nbatches x 2;
 * out: n x 2 (canonical).  Returns 0 / -1. */
int orc_fri_combine(const uint64_t *const *polys, const uint32_t *batch_sizes, size_t nbatches,
                    const uint64_t *points, const uint64_t alpha[2], size_t n, uint64_t *out);

/* fri_proof_of_work: smallest candidate c >= start such that, with state[pos] = c, permute(state)[7] (the
 * last squeezed rate element) has >= min_leading_zeros leading zero bits in canonical form.  plonky2 searches
 * with rayon find_any (any witness is valid); "smallest" is the deterministic rule both sides use here
 * (SURVEY.md 0.7).  Returns 0 and *witness, or -1 if none below `limit`. */
int orc_fri_pow(uint32_t hash_kind, const uint64_t state[12], uint32_t pos, uint32_t min_leading_zeros,
                uint64_t start, uint64_t limit, uint64_t *witness);

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
