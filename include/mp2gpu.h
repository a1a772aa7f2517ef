/*
 * mp2gpu.h -- C ABI of the B200 (sm_100a) polynomial-batch commitment library.
 *
 * This is the drop-in boundary for the ONE data-parallel hot path of
 * Lagrange-Labs/mapreduce-plonky2: plonky2's PolynomialBatch::from_values / from_coeffs and
 * MerkleTree::new.  The reference reaches that path only through the plonky2
 * crate it patches in at Cargo.toml:114-117; a fork of that crate forwards its fri/oracle.rs and
 * hash/merkle_tree.rs bodies to these symbols (binding shown in INTEGRATION.md).
 *
 * Conventions (modelled on the reference's only extern "C" interface,
 * gnark-utils/src/lib.rs:17-52 and gnark-utils/src/utils.rs:9-20):
 *   - every entry point returns NULL on success, else a heap-allocated, NUL-terminated error string
 *     that the caller releases with mp2gpu_free_string();  nothing ever unwinds across the boundary;
 *   - the caller owns all host buffers; the library never keeps a host pointer past the call;
 *   - inputs may hold non-canonical field elements (>= p = 2^64 - 2^32 + 1); every output is
 *     canonical (< p), so ==, serde and to_bytes agree with the CPU path;
 *   - entry points are re-entrant: each calling thread gets its own CUDA stream, the device is the
 *     one last chosen by that thread with mp2gpu_init() (default 0);
 *   - there is NO CPU fallback: without a usable CUDA device every data-path call returns an error string (the two
 *     mp2gpu_transcript_* entry points and mp2gpu_merkle_prove are host-only by nature and say so).
 *
 * Layouts:
 *   coeffs   ncols x n            column-major, natural order
 *   leaves   N x ncols            row-major, N = n << rate_bits, row i = LDE row bitrev(i),
 *                                 i.e. leaves[i][c] = P_c(7 * w_N^bitrev(i))
 *   digests  2*(N - 2^cap) x 4    plonky2's interleaved per-subtree layout
 *   cap      2^cap_height x 4
 */
#ifndef MP2GPU_H
#define MP2GPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MP2GPU_HASH_POSEIDON 0u  /* PoseidonGoldilocksConfig  (feature original_poseidon, and WrapC) */
#define MP2GPU_HASH_POSEIDON2 1u /* Poseidon2GoldilocksConfig (default C, mp2-common/src/lib.rs:37-40) */

typedef struct mp2gpu_batch mp2gpu_batch; /* device-resident PolynomialBatch */

/* ---- lifetime ------------------------------------------------------------------------------- */
/* Binds the calling thread to `device`, uploads the constant tables on first use. */
const char *mp2gpu_init(int device);
const char *mp2gpu_device_count(int *count_out);
/* Releases an error string (the FreeString of gnark-utils/src/lib.rs:51). */
void mp2gpu_free_string(const char *s);
/* "major.minor.patch (sm_100a)" -- static storage, do not free. */
const char *mp2gpu_version(void);
/* Page-locked host buffers (optional; any host pointer is accepted, pinned ones copy faster). */
const char *mp2gpu_host_alloc(void **ptr_out, size_t bytes);
const char *mp2gpu_host_free(void *ptr);

/* ---- PolynomialBatch::from_values / from_coeffs (plonky2 fri/oracle.rs; reached from
 *      recursion-framework/src/circuit_builder.rs:177,308 and
 *      recursion-framework/src/universal_verifier_gadget/wrap_circuit.rs:98,143) -------------------
 * cols[c]       n = 2^n_log field elements of column c (values resp. coefficients), host memory.
 * coeffs_out[c] n elements (may be NULL as a whole: coefficients are then not returned).
 * leaves_out    N*ncols elements or NULL (keep leaves on the device only).
 * digests_out   2*(N-2^cap_height)*4 elements or NULL.
 * cap_out       2^cap_height*4 elements (required).
 * handle_out    if non-NULL receives a device-resident handle (release with mp2gpu_batch_free).
 * blinding (salt columns) is not supported: the reference never enables it
 * (zero_knowledge = false, mp2-common/src/lib.rs:45-47). */
const char *mp2gpu_commit_from_values(const uint64_t *const *cols, size_t ncols, uint32_t n_log,
                                      uint32_t rate_bits, uint32_t cap_height, uint32_t hash_kind,
                                      uint64_t *const *coeffs_out, uint64_t *leaves_out,
                                      uint64_t *digests_out, uint64_t *cap_out,
                                      mp2gpu_batch **handle_out);
const char *mp2gpu_commit_from_coeffs(const uint64_t *const *cols, size_t ncols, uint32_t n_log,
                                      uint32_t rate_bits, uint32_t cap_height, uint32_t hash_kind,
                                      uint64_t *const *coeffs_out, uint64_t *leaves_out,
                                      uint64_t *digests_out, uint64_t *cap_out,
                                      mp2gpu_batch **handle_out);

/* ---- MerkleTree::new (plonky2 hash/merkle_tree.rs; called directly at
 *      recursion-framework/src/universal_verifier_gadget/circuit_set.rs:189) -----------------------
 * leaves: nleaves x leaf_len row-major, host memory.  Errors (where plonky2 panics): nleaves not a
 * power of two, cap_height > log2(nleaves).  leaf_len <= 4 takes hash_or_noop's no-op branch. */
const char *mp2gpu_merkle_new(const uint64_t *leaves, size_t nleaves, size_t leaf_len,
                              uint32_t cap_height, uint32_t hash_kind, uint64_t *digests_out,
                              uint64_t *cap_out);
/* Vec<Vec<F>> with per-leaf lengths (the circuit-set tree pads with vec![F::ZERO],
 * circuit_set.rs:184-185).  leaves[i] points to leaf_lens[i] elements. */
const char *mp2gpu_merkle_new_ragged(const uint64_t *const *leaves, const size_t *leaf_lens,
                                     size_t nleaves, uint32_t cap_height, uint32_t hash_kind,
                                     uint64_t *digests_out, uint64_t *cap_out);
/* MerkleTree::prove (circuit_set.rs:216): siblings bottom-up from a host digests array.
 * siblings_out: (log2(nleaves) - cap_height) x 4.  Pure index arithmetic (no GPU work). */
const char *mp2gpu_merkle_prove(const uint64_t *digests, size_t nleaves, uint32_t cap_height,
                                size_t leaf_index, uint64_t *siblings_out, size_t *nsiblings_out);

/* ---- Hasher::{hash_no_pad, hash_or_noop, two_to_one} in batches (plonky2 hash/hashing.rs; native
 *      uses at mp2-common/src/poseidon.rs:49-51, mp2-common/src/utils.rs:294-315) -------------------
 * inputs: count x input_len row-major; out: count x 4. */
const char *mp2gpu_hash_no_pad_batch(const uint64_t *inputs, size_t count, size_t input_len,
                                     uint32_t hash_kind, uint64_t *out);
/* a, b, out: count x 4. */
const char *mp2gpu_two_to_one_batch(const uint64_t *a, const uint64_t *b, size_t count,
                                    uint32_t hash_kind, uint64_t *out);
/* states: count x 12, permuted in place (PlonkyPermutation::permute). */
const char *mp2gpu_permute_batch(uint64_t *states, size_t count, uint32_t hash_kind);

/* ---- device-resident PolynomialBatch handle ------------------------------------------------ */
/* get_lde_values(index, step): out[r] = leaves[row_idx[r]] (ncols elements each); row_idx are LEAF
 * indices (callers apply bitrev(index*step) as plonky2 does). */
const char *mp2gpu_batch_fetch_rows(const mp2gpu_batch *b, const uint64_t *row_idx, size_t nrows,
                                    uint64_t *out);
/* MerkleTree::prove on the device-resident digests. */
const char *mp2gpu_batch_prove(const mp2gpu_batch *b, size_t leaf_index, uint64_t *siblings_out,
                               size_t *nsiblings_out);
/* Query-phase openings (fri_prover_query_round reads MerkleTree::get + MerkleTree::prove from every oracle
 * for each of the 28 query indices): rows_out[q] = leaf row leaf_idx[q] (ncols elements), siblings_out[q] =
 * its (log2 N - cap_height) x 4 sibling digests, bottom-up; gathered on the device, two copies back. */
const char *mp2gpu_batch_open(const mp2gpu_batch *b, const uint64_t *leaf_idx, size_t count,
                              uint64_t *rows_out, uint64_t *siblings_out);
/* OpeningSet::new's `c.polynomials.par_iter().map(|p| p.to_extension().eval(z))` (plonky2 plonk/proof.rs; step 7 of
 * prove(), between the quotient commitment and prove_openings) on the resident coefficients.
 * points: npoints x 2 (extension elements); out: npoints x ncols x 2, out[(p * ncols + c)] = polynomial c at point p. */
const char *mp2gpu_batch_eval(const mp2gpu_batch *b, const uint64_t *points, size_t npoints, uint64_t *out);
/* Any of the outputs may be NULL. Sizes as for mp2gpu_commit_from_values. */
const char *mp2gpu_batch_fetch(const mp2gpu_batch *b, uint64_t *const *coeffs_out,
                               uint64_t *leaves_out, uint64_t *digests_out, uint64_t *cap_out);
const char *mp2gpu_batch_shape(const mp2gpu_batch *b, size_t *ncols, uint32_t *n_log,
                               uint32_t *rate_bits, uint32_t *cap_height, uint32_t *hash_kind);
void mp2gpu_batch_free(mp2gpu_batch *b);

/* ---- FRI commit phase (plonky2 fri/prover.rs fri_committed_trees; runs inside every prove(), e.g.
 *      recursion-framework/src/circuit_builder.rs:308) -- device-resident loop state --------------------
 * Extension elements (D = 2, mp2-common/src/lib.rs:36) are interleaved pairs [a0, a1].  Usage, per proof:
 *   mp2gpu_fri_begin(final_poly coefficients, degree_log, rate_bits, cap_height, hash_kind, &f);
 *   for arity_bits in fri_params.reduction_arity_bits:
 *       mp2gpu_fri_commit_layer(f, arity_bits, cap);   // MerkleTree::new(chunked values, cap_height)
 *       challenger.observe_cap(cap); beta = challenger.get_extension_challenge();   // host, tiny
 *       mp2gpu_fri_fold(f, beta);                       // reduce_with_powers + coset_fft(shift^arity)
 *   mp2gpu_fri_finish(f, final_coeffs, &len);           // already truncated by rate_bits
 * The layer trees stay on the device for the query phase (mp2gpu_fri_fetch_layer).  The values of the first
 * layer are computed here from the coefficients (plonky2 passes both; they are the same function of them). */
typedef struct mp2gpu_fri mp2gpu_fri;
/* coeffs_ext: 2^n_log extension coefficients (the non-padded final polynomial; its lde(rate_bits) is implied). */
const char *mp2gpu_fri_begin(const uint64_t *coeffs_ext, uint32_t n_log, uint32_t rate_bits,
                             uint32_t cap_height, uint32_t hash_kind, mp2gpu_fri **out);
/* PolynomialBatch::prove_openings up to its call of fri_proof (plonky2 fri/oracle.rs; ReducingFactor of
 * util/reducing.rs): the alpha-batched quotient
 *     final_poly = sum_i alpha^(k_i) (F_i(X) - F_i(z_i)) / (X - z_i),    F_i = sum_j alpha^j f_ij
 * computed from the coefficients the committed batches keep on the device, and left there as the state of a new
 * commit phase (*out, exactly what mp2gpu_fri_begin builds from host coefficients; rate_bits is the oracles').
 * oracles[noracles]: handles from mp2gpu_commit_from_values / _coeffs (plonky2's FRI_ORACLES order).  Batch i
 * opens batch_sizes[i] polynomials at the extension point points[2i..2i+2]; the polynomials of all batches are
 * listed one after the other as FriPolynomialInfo { oracle_index, polynomial_index }.
 * final_poly_out: optional, 2^n_log x 2 (interleaved, canonical, last coefficient 0 as divide_by_linear's
 * padding leaves it). */
const char *mp2gpu_fri_begin_openings(const mp2gpu_batch *const *oracles, size_t noracles,
                                      const uint64_t *points, const uint32_t *batch_sizes, size_t nbatches,
                                      const uint32_t *oracle_index, const uint32_t *polynomial_index,
                                      const uint64_t alpha[2], uint32_t cap_height, uint32_t hash_kind,
                                      uint64_t *final_poly_out, mp2gpu_fri **out);
/* cap_out: 2^cap_height x 4. */
const char *mp2gpu_fri_commit_layer(mp2gpu_fri *f, uint32_t arity_bits, uint64_t *cap_out);
const char *mp2gpu_fri_fold(mp2gpu_fri *f, const uint64_t beta[2]);
/* The same for a FRI layer tree (leaves_out[q] = 2 << arity_bits elements). */
const char *mp2gpu_fri_open_layer(const mp2gpu_fri *f, uint32_t layer, const uint64_t *leaf_idx,
                                  size_t count, uint64_t *leaves_out, uint64_t *siblings_out);
const char *mp2gpu_fri_layer_shape(const mp2gpu_fri *f, uint32_t layer, size_t *nleaves, size_t *leaf_len,
                                   size_t *ndigests, size_t *ncap);
/* Any output may be NULL; sizes from mp2gpu_fri_layer_shape. */
const char *mp2gpu_fri_fetch_layer(const mp2gpu_fri *f, uint32_t layer, uint64_t *leaves_out,
                                   uint64_t *digests_out, uint64_t *cap_out);
/* final_coeffs_out: *len_out extension coefficients (interleaved); may be NULL to query the length. */
const char *mp2gpu_fri_finish(mp2gpu_fri *f, uint64_t *final_coeffs_out, size_t *len_out);
void mp2gpu_fri_free(mp2gpu_fri *f);

/* fri_proof_of_work (plonky2 fri/prover.rs): duplex_state = the challenger's sponge state with its pending
 * inputs already written (12 elements), witness_pos = challenger.input_buffer.len().  Finds the SMALLEST
 * candidate c such that, with state[witness_pos] = c, permute(state)[7] has >= min_leading_zeros leading zero
 * bits in canonical form (min_leading_zeros = proof_of_work_bits + 64 - 64 = 16 under
 * standard_recursion_config).  plonky2 accepts any witness (rayon find_any); the smallest one is the
 * deterministic rule that makes proofs byte-comparable. */
const char *mp2gpu_fri_proof_of_work(const uint64_t *duplex_state, uint32_t witness_pos,
                                     uint32_t min_leading_zeros, uint32_t hash_kind,
                                     uint64_t *witness_out);

/* ---- device-pointer stages (inputs already resident in HBM; asynchronous on `stream`, a
 *      cudaStream_t passed as void*: NULL is CUDA's legacy default stream, MP2GPU_STREAM_THREAD the
 *      calling thread's private library stream).  These are what the multi-GPU driver and bench.py's
 *      device-resident leg call. ------------------------------------------------------------------ */
#define MP2GPU_STREAM_THREAD ((void *)(~(uintptr_t)0))
/* values (ncols x n, column c at values + c*in_stride) -> coefficients (same shape, out_stride). */
const char *mp2gpu_dev_intt(const uint64_t *values, size_t in_stride, uint64_t *coeffs,
                            size_t out_stride, size_t ncols, uint32_t n_log, void *stream);
/* coefficients -> coset LDE on 7*<w_N>, written LEAF-ordered and column-major:
 * element (column c, leaf L) at lde[(L >> shard_log') ...] -- precisely
 *   lde[(L / Ls) * shard_stride + c * lde_stride + (L % Ls)],  Ls = N >> shard_log,
 * so that with shard_log = log2(G) the block destined for rank g of a row-sharded exchange is the
 * contiguous range [g*shard_stride, g*shard_stride + ncols*lde_stride).  shard_log = 0 gives plain
 * column-major (shard_stride ignored). */
const char *mp2gpu_dev_coset_lde(const uint64_t *coeffs, size_t in_stride, uint64_t *lde,
                                 size_t lde_stride, size_t ncols, uint32_t n_log,
                                 uint32_t rate_bits, uint32_t shard_log, size_t shard_stride,
                                 void *stream);
/* The same transform with the row-shard exchange fused into its store: shard g (the leaves owned by
 * rank g) is written straight to shard_bases[g] + c*lde_stride -- pointers into the other ranks' HBM
 * mapped over NVLink/NVSwitch (peer or symmetric memory), so no all-to-all follows.  shard_bases is a
 * HOST array of 2^shard_log (<= 16) device pointers.  Callers order the kernel against the peers with
 * their own barrier (sharded.py: symmetric-memory barrier before and after).  first_shard: the destination this
 * launch stores to first, the others following in rotated order -- every rank passes its own index, so that at
 * any moment the ranks write to different peers (all starting with shard 0 serialises the exchange on one
 * NVLink ingress).  scratch: ncols * (n << rate_bits) elements of LOCAL device memory for the four-step
 * intermediate when n > 2^14 (NULL: taken from the stream-ordered pool for the call -- avoid that in a loop, a
 * multi-GB pool allocation per step is what stalled the first step after an idle period in round 1). */
const char *mp2gpu_dev_coset_lde_peer(const uint64_t *coeffs, size_t in_stride,
                                      uint64_t *const *shard_bases, size_t lde_stride, size_t ncols,
                                      uint32_t n_log, uint32_t rate_bits, uint32_t shard_log,
                                      uint32_t first_shard, uint64_t *scratch, void *stream);
/* The two halves of mp2gpu_dev_merkle_colmajor, for callers that overlap the hashing of one leaf range with
 * the device->host copy of the previous one: leaf digests (and optional row-major rows) of leaves
 * [leaf_begin, leaf_end) only; then the inner levels + cap once every leaf has been hashed. */
const char *mp2gpu_dev_merkle_colmajor_leaves(const uint64_t *lde, size_t lde_stride, size_t ncols, size_t nleaves,
                                              uint32_t cap_height, uint32_t hash_kind, size_t leaf_begin,
                                              size_t leaf_end, uint64_t *leaves_out, uint64_t *digests_out,
                                              uint64_t *cap_out, void *stream);
const char *mp2gpu_dev_merkle_levels(size_t nleaves, uint32_t cap_height, uint32_t hash_kind, uint64_t *digests,
                                     uint64_t *cap, void *stream);
/* Leaf-ordered column-major LDE (column c at lde + c*lde_stride, nleaves elements) -> optional
 * row-major leaves (nleaves x ncols), digests and cap of a tree with `nleaves` leaves.  A rank of a
 * G-way row-sharded batch passes nleaves = N/G and cap_height - log2(G): its digests/cap are the
 * contiguous slices of the global arrays. */
const char *mp2gpu_dev_merkle_colmajor(const uint64_t *lde, size_t lde_stride, size_t ncols,
                                       size_t nleaves, uint32_t cap_height, uint32_t hash_kind,
                                       uint64_t *leaves_out, uint64_t *digests_out,
                                       uint64_t *cap_out, void *stream);
/* Row-major leaves already on the device -> digests and cap. */
const char *mp2gpu_dev_merkle_rowmajor(const uint64_t *leaves, size_t nleaves, size_t leaf_len,
                                       uint32_t cap_height, uint32_t hash_kind,
                                       uint64_t *digests_out, uint64_t *cap_out, void *stream);
/* Whole commitment on device buffers: cols_dev is ncols x n column-major (stride n).
 * coeffs_dev (ncols x n), lde_dev (ncols x N scratch, leaf-ordered column-major), leaves_dev
 * (N x ncols, may be NULL), digests_dev, cap_dev are device buffers owned by the caller. */
const char *mp2gpu_dev_commit(const uint64_t *cols_dev, size_t ncols, uint32_t n_log,
                              uint32_t rate_bits, uint32_t cap_height, uint32_t hash_kind,
                              int from_coeffs, uint64_t *coeffs_dev, uint64_t *lde_dev,
                              uint64_t *leaves_dev, uint64_t *digests_dev, uint64_t *cap_dev,
                              void *stream);
/* Blocks until `stream` (see above for NULL / MP2GPU_STREAM_THREAD) has drained. */
const char *mp2gpu_sync(void *stream);
/* Canonicalises `count` field elements (x >= p -> x - p); in == out allowed. */
const char *mp2gpu_dev_canonicalize(const uint64_t *in, uint64_t *out, size_t count, void *stream);
/* Number of kernel launches issued by this library since load (bench.py's gpu_launches). */
uint64_t mp2gpu_launch_count(void);

/* ---- one wide batch over several GPUs of ONE process --------------------
 * The multi-GPU form of PolynomialBatch::from_values / from_coeffs for a prover process that owns G devices
 * (the reference runs one prover process per machine: mp2-v1/src/api.rs:154-165 hands it opaque tasks).
 * mp2gpu_comm_init binds G = 2^k devices (G <= 2^cap_height at commit time), enables peer access between
 * every pair and creates one worker context (streams, receive buffer) per device.
 * mp2gpu_commit_from_values_sharded then runs, with one host thread per device:
 *     columns [g*c/G, (g+1)*c/G) -> upload -> iNTT -> coset LDE whose stores go straight into the owners' HBM
 *     over NVLink (leaf L belongs to device L / (N/G); no separate all-to-all) -> barrier -> leaf hashing of
 *     rows [g*N/G, (g+1)*N/G) in blocks, each block copied to the host while the next is hashed ->
 *     the device's 2^cap/G subtrees -> its slice of digests / cap.
 * Arguments and outputs are exactly those of mp2gpu_commit_from_values (host buffers, global layouts:
 * coeffs_out[c], leaves_out N x ncols row-major, digests_out in plonky2's layout, cap_out); any of
 * coeffs_out / leaves_out / digests_out may be NULL.  ncols must be a multiple of G. */
typedef struct mp2gpu_comm mp2gpu_comm;
const char *mp2gpu_comm_init(int ndev, const int *devs, mp2gpu_comm **comm_out);
void mp2gpu_comm_free(mp2gpu_comm *comm);
const char *mp2gpu_commit_from_values_sharded(mp2gpu_comm *comm, const uint64_t *const *cols, size_t ncols,
                                              uint32_t n_log, uint32_t rate_bits, uint32_t cap_height,
                                              uint32_t hash_kind, int from_coeffs, uint64_t *const *coeffs_out,
                                              uint64_t *leaves_out, uint64_t *digests_out, uint64_t *cap_out);

/* ---- Fiat-Shamir transcript (host) ----------------------------------------------------------------
 * One Poseidon / Poseidon2 permutation of a 12-element state ON THE HOST, for `Challenger::duplexing` (plonky2
 * iop/challenger.rs) only: the transcript is a few dozen strictly sequential permutations per proof over caps and
 * challenges that already live on the host, and it is host code in the reference (part of plonky2's prover, reached
 * from recursion-framework/src/circuit_builder.rs:308).  Not a fallback for anything on the data path: leaves,
 * nodes, proof-of-work and all batched hashing run on the device only.  A Rust integration keeps plonky2's own
 * challenger and never calls this; the Python / C++ host mirrors use it (tests check it against
 * mp2gpu_permute_batch).  In place; inputs may be non-canonical, outputs are canonical. */
const char *mp2gpu_transcript_permute(uint64_t *state12, uint32_t hash_kind);
/* `Challenger::observe_elements` for n elements in one call (same host-only scope): state12 = sponge_state,
 * buffer8 / *buffer_len = input_buffer (fewer than 8 pending elements), all updated in place; every element clears
 * the output buffer, a full rate of 8 triggers a duplexing.  *duplexed_last_out = 1 iff the last element completed
 * one (the output buffer is then state12[0..8), otherwise it is empty). */
const char *mp2gpu_transcript_observe(uint64_t *state12, uint64_t *buffer8, uint32_t *buffer_len,
                                      const uint64_t *elems, size_t n, uint32_t hash_kind,
                                      uint32_t *duplexed_last_out);

/* ---- quotient polynomials on the device ------------------------------------
 * plonky2 0.2.2 `compute_quotient_polys` (plonk/prover.rs, with plonk/vanishing_poly.rs
 * eval_vanishing_poly_base_batch) followed by the prover's `PolynomialBatch::from_coeffs(all_quotient_poly_chunks,
 * rate_bits, blinding = false, cap_height, ...)`: step 5 of every `circuit_data.prove(pw)` the reference issues
 * (recursion-framework/src/circuit_builder.rs:308, .../universal_verifier_gadget/wrap_circuit.rs:143).  The three
 * input batches are the device-resident handles returned by mp2gpu_commit_from_values (constants+sigmas, wires,
 * Zs + partial products); their LDE rows are read in HBM -- with this call no batch needs `leaves_out`.
 * Gate set: the staged subset of mp2-common/src/serialization/circuit_data_serialization.rs:234-266 below; any
 * other kind is rejected with an error (never silently skipped).  No lookups, no blinding (zero_knowledge = false,
 * mp2-common/src/lib.rs:45-47). */
#define MP2GPU_GATE_NOOP 0u          /* NoopGate: no constraints */
#define MP2GPU_GATE_ARITHMETIC 1u    /* ArithmeticGate{num_ops}: w[4i+3] - (c0 w[4i] w[4i+1] + c1 w[4i+2]) */
#define MP2GPU_GATE_CONSTANT 2u      /* ConstantGate{num_consts = num_ops}: c_i - w_i */
#define MP2GPU_GATE_PUBLIC_INPUT 3u  /* PublicInputGate: w_i - public_inputs_hash[i], i < 4 */
#define MP2GPU_GATE_POSEIDON 4u      /* PoseidonGate (gates/poseidon.rs), 135 wires: input 0..11, output 12..23, swap 24,
                                        delta 25..28, S-box inputs of full rounds 1..3 at 29 + 12 (round - 1) + i, of the
                                        22 partial rounds at 65 + r, of the last full rounds at 87 + 12 round + i; 123
                                        constraints: swap bit, 4 deltas, 36 + 22 + 48 S-box inputs, 12 outputs */
#define MP2GPU_GATE_ARITHMETIC_EXT 5u /* ArithmeticExtensionGate{num_ops}, D = 2: per op the 2-wire groups m0 | m1 | addend |
                                        output at 8i; output - (c0 m0 m1 + c1 addend), 2 constraints per op */
#define MP2GPU_GATE_MUL_EXT 6u       /* MulExtensionGate{num_ops}: m0 | m1 | output at 6i; output - c0 m0 m1 */
#define MP2GPU_GATE_BASE_SUM 7u      /* BaseSumGate<B = param>{num_limbs = num_ops}: wire 0 = sum, limbs from wire 1;
                                        sum_i limb_i B^i - sum, then prod_{k < B} (limb_i - k) per limb */
#define MP2GPU_GATE_REDUCING 8u      /* ReducingGate{num_coeffs = num_ops}, D = 2: output 0..2 | alpha 2..4 | old_acc 4..6 | base
                                        coefficients from 6 | accumulators after them (the last one is the output wires);
                                        acc_{i-1} alpha + coeff_i - acc_i, 2 constraints each */
#define MP2GPU_GATE_REDUCING_EXT 9u  /* ReducingExtensionGate{num_coeffs = num_ops}: the same with extension coefficients */
#define MP2GPU_GATE_RANDOM_ACCESS 10u /* RandomAccessGate{bits = param & 0xFF, num_copies = num_ops, num_extra_constants =
                                        param >> 8}: per copy access_index | claimed_element | 2^bits list items, then the
                                        extra constants, then (unrouted) the index bits; per copy b(b-1) per bit, index
                                        reconstruction, folded list - claimed_element; then constant_i - wire_i */
#define MP2GPU_GATE_EXPONENTIATION 11u /* ExponentiationGate{num_power_bits = num_ops}: base 0 | bits (LE) from 1 | output 1+n |
                                         intermediate values from 2+n; (i ? iv_{i-1}^2 : 1)(b base + 1 - b) - iv_i with
                                         b = bit_{n-1-i}; then output - iv_{n-1} */
#define MP2GPU_GATE_POSEIDON_MDS 12u   /* PoseidonMdsGate: 12 extension inputs at 2i, outputs at 24 + 2i; output - MDS(input) */
#define MP2GPU_GATE_COSET_INTERPOLATION 13u /* CosetInterpolationGate{subgroup_bits = num_ops, degree = param}, D = 2: shift 0 |
                                         2^bits extension values from 1 | evaluation_point | evaluation_value (routed up to
                                         here) | intermediate evals | intermediate prods | shifted_evaluation_point;
                                         barycentric fold eval' = eval (x - w^k) + value_k prod weight_k, prod' = prod (x - w^k)
                                         cut after `degree` points, then every degree - 1; constraints (2 each):
                                         point - shift * shifted_point, per cut (eval wire - eval, prod wire - prod), value - eval */
typedef struct mp2gpu_gate {
  uint32_t kind;            /* MP2GPU_GATE_* */
  uint32_t num_ops;         /* see the kinds above */
  uint32_t selector_index;  /* SelectorsInfo::selector_indices[gate] */
  uint32_t group_begin;     /* SelectorsInfo::groups[selector_index] = group_begin..group_end (gate indices) */
  uint32_t group_end;
  uint32_t param;           /* BaseSumGate: the base B; RandomAccessGate: bits | num_extra_constants << 8;
                               CosetInterpolationGate: degree; 0 otherwise */
} mp2gpu_gate;
typedef struct mp2gpu_circuit {   /* the CommonCircuitData fields the vanishing polynomial depends on */
  uint32_t degree_bits;
  uint32_t quotient_degree_bits;  /* log2(quotient_degree_factor), <= the batches' rate_bits */
  uint32_t num_challenges;
  uint32_t num_wires, num_routed_wires;
  uint32_t num_constants;         /* selectors + gate constants (columns before the sigmas) */
  uint32_t num_selectors;
  uint32_t num_gates;
  const mp2gpu_gate *gates;       /* in CommonCircuitData::gates order (the order selector values refer to) */
} mp2gpu_circuit;
/* betas / gammas / alphas: num_challenges elements each; public_inputs_hash: 4 elements (may be NULL without a
 * PublicInputGate).  Outputs: chunks_out[c * 2^qb + k] = chunk k of challenge c (n coefficients; the array or any
 * entry may be NULL), then the outputs of mp2gpu_commit_from_coeffs for that batch (leaves_out / digests_out may be
 * NULL, cap_out is required, quotient_batch_out optionally receives the device-resident batch). */
const char *mp2gpu_quotient_polys(const mp2gpu_circuit *circuit, const mp2gpu_batch *constants_sigmas,
                                  const mp2gpu_batch *wires, const mp2gpu_batch *zs_partial_products,
                                  const uint64_t *betas, const uint64_t *gammas, const uint64_t *alphas,
                                  const uint64_t *public_inputs_hash, uint32_t rate_bits, uint32_t cap_height,
                                  uint32_t hash_kind, uint64_t *const *chunks_out, uint64_t *leaves_out,
                                  uint64_t *digests_out, uint64_t *cap_out, mp2gpu_batch **quotient_batch_out);

/* ---- the permutation argument's running products on the device ------------------------------------
 * plonky2 0.2.2 `all_wires_permutation_partial_products` / `wires_permutation_partial_products_and_zs`
 * (plonk/prover.rs) followed by the prover's second commitment `PolynomialBatch::from_values(zs_partial_products,
 * rate_bits, blinding = false, cap_height, ...)`: step 3 of every `circuit_data.prove(pw)` the reference issues
 * (recursion-framework/src/circuit_builder.rs:308).  Inputs are the device-resident constants+sigmas and wires
 * batches; the wire and sigma VALUES on the subgroup are recomputed from the coefficients they keep in HBM.
 * Columns of the result: [Z_0 .. Z_{nch-1}, partial products of challenge 0 (num_partial_products of them), of
 * challenge 1, ...] -- the layout the quotient step and the verifier's `zs_range` / `partial_products_range` expect.
 * values_out (optional): that many host columns of n canonical values; the other outputs are those of
 * mp2gpu_commit_from_values.  A zero denominator w + beta*sigma + gamma is an error (plonky2 panics there). */
const char *mp2gpu_partial_products_and_zs(const mp2gpu_circuit *circuit, const mp2gpu_batch *constants_sigmas,
                                           const mp2gpu_batch *wires, const uint64_t *betas, const uint64_t *gammas,
                                           uint32_t rate_bits, uint32_t cap_height, uint32_t hash_kind,
                                           uint64_t *const *values_out, uint64_t *leaves_out, uint64_t *digests_out,
                                           uint64_t *cap_out, mp2gpu_batch **zs_partial_products_batch_out);

/* ---- the whole prove() from the witness on, as one call ---------------------------------------------
 * plonky2 0.2.2 `prove_with_partition_witness` (plonk/prover.rs) after witness generation -- what every
 * `circuit_data.prove(pw)` of the reference runs (recursion-framework/src/circuit_builder.rs:308,
 * .../universal_verifier_gadget/wrap_circuit.rs:143): wires commitment, challenger, Z / partial products and their
 * commitment, quotient polynomials and their commitment, OpeningSet at zeta and g*zeta, prove_openings (FRI commit
 * phase, smallest proof-of-work witness, query rounds).  Data-path steps run on the device, the Fiat-Shamir
 * transcript on the host; the rows, coefficients and digests of the four batches never leave HBM.
 * constants_sigmas: the device-resident batch committed at circuit-build time (same rate_bits / cap_height).
 * wires_values: num_wires host columns of n = 2^degree_bits values.  public_inputs_hash: 4 elements (the caller
 * hashes the public inputs with the config's InnerHasher, as plonky2 does before this point).
 * *proof_out receives bincode(ProofWithPublicInputs) -- mp2-common/src/proof.rs:86-90 `serialize_proof` -- in a
 * buffer of *proof_len_out bytes that the caller releases with mp2gpu_free_bytes.  Gate coverage is
 * mp2gpu_quotient_polys'; no lookups, no blinding. */
typedef struct mp2gpu_prove_config {
  uint32_t rate_bits, cap_height, hash_kind;  /* FriConfig::rate_bits / cap_height; MP2GPU_HASH_* */
  uint32_t proof_of_work_bits, num_query_rounds;
  uint32_t num_reductions;                    /* FriParams::reduction_arity_bits for this degree */
  const uint32_t *reduction_arity_bits;
} mp2gpu_prove_config;
const char *mp2gpu_prove(const mp2gpu_circuit *circuit, const mp2gpu_prove_config *config,
                         const mp2gpu_batch *constants_sigmas, const uint64_t *circuit_digest,
                         const uint64_t *const *wires_values, const uint64_t *public_inputs, size_t num_public_inputs,
                         const uint64_t *public_inputs_hash, uint8_t **proof_out, size_t *proof_len_out);
void mp2gpu_free_bytes(uint8_t *p);

/* Returns the calling thread's cached device blocks, the device's cached twiddle tables and the unused part of its
 * stream-ordered pool to the driver (the library keeps freed scratch for reuse: a prover repeats the same shapes;
 * tables are rebuilt on demand).  Call it when another allocator in the process needs the memory and no other
 * thread has library work in flight on this device. */
const char *mp2gpu_trim(void);

/* ---- measurement hooks (bench.py) ----------------------------------------------------------- */
/* While enabled, every kernel launch is bracketed by CUDA events on its own stream. */
const char *mp2gpu_profile_enable(int on);
/* Synchronises the device and writes "kernel_name launches total_ms\n" lines for everything
 * recorded since the last report into buf (NUL terminated). */
const char *mp2gpu_profile_report(char *buf, size_t buf_len);
/* Live probe of the 32-bit integer multiply-add issue rate (the Poseidon roofline denominator):
 * thread-instructions per clock per SM, SM clock held during the probe (MHz), and T IMAD/s. */
const char *mp2gpu_debug_int_pipe_peak(double *imad_per_clk_per_sm, double *sm_clock_mhz,
                                       double *t_imad_per_s);

/* Device self-test of the Goldilocks primitives (add, add-canonical, sub, mul, sqr, mul-add, reduce128, x^7,
 * the three shift twiddles, every 2^(12 j) shift -- in that order) against 128-bit arithmetic by definition,
 * over all pairs of 48 corner values around 0 / 2^32 / 2^63 / p / 2^64 and 976 pseudo-random ones.
 * mismatches_out[i] = number of wrong results of test i (ntests >= 12).  (The 8- and 16-point butterflies are
 * checked from the host through mp2gpu_debug_dft.)  Replaces plonky2_field's goldilocks_field unit tests for the
 * GPU arithmetic. */
const char *mp2gpu_debug_field_selftest(uint64_t *mismatches_out, size_t ntests);
/* Runs `count` independent 2^log_points-point shift-twiddle butterflies (log_points = 3 or 4; csrc/dft.cuh) in
 * place on io (count * 2^log_points elements, any u64 in, canonical out): output position j holds the DFT value of
 * frequency bitrev(j) for plonky2's primitive_root_of_unity(log_points).  Test hook. */
const char *mp2gpu_debug_dft(uint64_t *io, uint32_t log_points, size_t count);
/* Register-only throughput of the two arithmetic inner loops, thread-level operations per (nominal) clock per
 * SM: out[0] = x^7 S-boxes, out[1] = radix-8 butterfly elements (each butterfly + 7 twiddle multiplications). */
const char *mp2gpu_debug_field_probe(double *ops_per_clk_per_sm_out);

#ifdef __cplusplus
}
#endif
#endif /* MP2GPU_H */
