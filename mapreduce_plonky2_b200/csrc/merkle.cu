// Poseidon / Poseidon2 leaf hashing and Merkle tree + cap construction in plonky2's digest layout.
//
// Replaces plonky2::hash::merkle_tree::MerkleTree::new (fill_digests_buf / fill_subtree) and
// Hasher::{hash_or_noop, hash_no_pad, two_to_one} -- SURVEY.md 8(a) a5/a6/a7, Appendix A.4/A.5.
// Reference call sites: recursion-framework/src/universal_verifier_gadget/circuit_set.rs:189
// (MerkleTree::new) and :216 (prove); sponge semantics restated in-circuit at
// mp2-common/src/poseidon.rs:136-172 and mp2-common/src/hash.rs:16-46.
//
// Digest layout (A.4): the digests array is split into 2^cap_height equal chunks, one per cap
// subtree of 2^h leaves.  Inside a chunk, node m of layer i (layer 0 = leaf digests, i < h) lives
// at digest index 2*(((m>>1) << (i+1)) + (1<<i) - 1) + (m&1).  The subtree roots form the cap.
#include <cstdlib>
#include <vector>

#include "internal.h"
#include "poseidon.cuh"

#ifndef MP2_HASH_BLOCK
#define MP2_HASH_BLOCK 128
#endif

namespace mp2 {

// index (in digests) of node m of layer i inside one subtree chunk
GL_DEV size_t node_slot(u32 layer, size_t m) {
  return 2 * (((m >> 1) << (layer + 1)) + ((size_t)1 << layer) - 1) + (m & 1);
}

GL_DEV void store_digest(u64 *dst, const u64 (&s)[12]) {
  ulonglong2 a = make_ulonglong2(gl_canon(s[0]), gl_canon(s[1]));
  ulonglong2 b = make_ulonglong2(gl_canon(s[2]), gl_canon(s[3]));
  reinterpret_cast<ulonglong2 *>(dst)[0] = a;
  reinterpret_cast<ulonglong2 *>(dst)[1] = b;
}

// where the digest of leaf L goes: layer-0 slot of its subtree, or the cap if the tree is all cap
GL_DEV u64 *leaf_digest_ptr(size_t L, u32 h, u64 *digests, u64 *cap) {
  if (h == 0) return cap + 4 * L;
  size_t sub = L >> h, l = L & (((size_t)1 << h) - 1);
  size_t per = 2 * (((size_t)1 << h) - 1);
  return digests + 4 * (sub * per + node_slot(0, l));
}

// ---- K4: leaf sponge.  One thread per leaf; rate 8, overwrite absorb (A.5). ---------------------
// COLMAJOR: element (column c, leaf L) at in[c*stride + L]  -> loads coalesce across the warp.
// !COLMAJOR: element at in[L*stride + c] (row-major user leaves, FRI layers).
// 5 CTAs of 128 threads per SM (<= 96 registers): the FP64 partial rounds fit without spills (ptxas -v)
#ifndef MP2_HASH_MIN_CTAS
#define MP2_HASH_MIN_CTAS 5
#endif
template <u32 KIND, bool COLMAJOR, int BLOCK>
__global__ void __launch_bounds__(BLOCK, BLOCK == 128 ? MP2_HASH_MIN_CTAS : 1)
k_leaf_hash(const u64 *__restrict__ in, size_t stride, u32 ncols, size_t leaf_begin, size_t leaf_end, u32 h,
            u64 *__restrict__ leaves_out, u64 *__restrict__ digests, u64 *__restrict__ cap) {
  size_t L = leaf_begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = L < leaf_end;  // inactive threads still walk the rounds (block-wide barriers)
  if (!active) L = leaf_end - 1;
  u64 st[12];
#pragma unroll
  for (int i = 0; i < 12; i++) st[i] = 0;
  const u64 *src = COLMAJOR ? in + L : in + L * stride;
  const size_t step = COLMAJOR ? stride : 1;
  u64 *row = (leaves_out && active) ? leaves_out + L * (size_t)ncols : nullptr;
  if (ncols <= 4) {  // hash_or_noop: no permutation, zero padded
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (j < ncols) {
        u64 v = gl_canon(src[j * step]);
        st[j] = v;
        if (row) row[j] = v;
      }
  } else {
    for (u32 c0 = 0; c0 < ncols; c0 += 8) {
      u32 m = ncols - c0 < 8 ? ncols - c0 : 8;
      u64 v[8];
#pragma unroll
      for (int j = 0; j < 8; j++)
        if (j < m) v[j] = src[(size_t)(c0 + j) * step];
#pragma unroll
      for (int j = 0; j < 8; j++)
        if (j < m) {
          st[j] = v[j];
          if (row) row[c0 + j] = gl_canon(v[j]);
        }
      permute<KIND, true>(st);
    }
  }
  if (active) store_digest(leaf_digest_ptr(L, h, digests, cap), st);
}

// ragged leaves: leaf L = flat[off[L] .. off[L+1])
template <u32 KIND>
__global__ void __launch_bounds__(MP2_HASH_BLOCK)
k_leaf_hash_ragged(const u64 *__restrict__ flat, const u64 *__restrict__ off, size_t nleaves, u32 h,
                   u64 *__restrict__ digests, u64 *__restrict__ cap) {
  size_t L = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (L >= nleaves) return;
  u64 st[12];
#pragma unroll
  for (int i = 0; i < 12; i++) st[i] = 0;
  const u64 *src = flat + off[L];
  size_t len = off[L + 1] - off[L];
  if (len <= 4) {
#pragma unroll
    for (int j = 0; j < 4; j++)
      if ((size_t)j < len) st[j] = gl_canon(src[j]);
  } else {
    for (size_t c0 = 0; c0 < len; c0 += 8) {
      size_t m = len - c0 < 8 ? len - c0 : 8;
#pragma unroll
      for (int j = 0; j < 8; j++)
        if ((size_t)j < m) st[j] = src[c0 + j];
      permute<KIND>(st);
    }
  }
  store_digest(leaf_digest_ptr(L, h, digests, cap), st);
}

// ---- K5: one Merkle layer.  Thread t -> node m of layer `layer` (1..h) of subtree sub. ---------
template <u32 KIND>
__global__ void __launch_bounds__(MP2_HASH_BLOCK)
k_merkle_layer(u64 *__restrict__ digests, u64 *__restrict__ cap, u32 h, u32 layer, size_t nnodes) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = t < nnodes;
  if (!active) t = nnodes - 1;
  u32 width_log = h - layer;  // nodes of this layer per subtree = 2^width_log
  size_t sub = t >> width_log, m = t & (((size_t)1 << width_log) - 1);
  size_t per = 2 * (((size_t)1 << h) - 1);
  u64 *base = digests + 4 * sub * per;
  // children: nodes 2m, 2m+1 of layer-1 share one pair slot -> 8 contiguous u64
  const ulonglong2 *ch = reinterpret_cast<const ulonglong2 *>(base + 4 * node_slot(layer - 1, 2 * m));
  ulonglong2 c0 = ch[0], c1 = ch[1], c2 = ch[2], c3 = ch[3];
  u64 st[12] = {c0.x, c0.y, c1.x, c1.y, c2.x, c2.y, c3.x, c3.y, 0, 0, 0, 0};
  permute<KIND, true>(st);
  u64 *dst = layer == h ? cap + 4 * sub : base + 4 * node_slot(layer, m);
  if (active) store_digest(dst, st);
}

// ---- K5': one Merkle layer, 16 lanes per node (12 active), for the SPARSE upper levels ----------------
// One permutation is ~24 k instructions on one thread -- ~30 us when a warp has the SM to itself -- and the
// levels with fewer nodes than the GPU has warps are pure latency.  Here a node's 12 state elements live in
// 12 lanes: every lane does its own S-box, and the linear layers are row-wise dot products whose operands
// are fetched from the sibling lanes with warp shuffles (3 limb shuffles per term).  ~5.7 k instructions per
// lane and permutation: 4-5x lower latency for 4x more issue slots, so it is used only below
// kCoopLevelNodes nodes per level.  Results are bit-identical to the one-thread form (same field arithmetic).
static __device__ const u64 g_pos_rc[MP2_POSEIDON_RC_LEN] = {MP2_POSEIDON_RC_LIST};
static __device__ const u64 g_p2_rc[MP2_POSEIDON2_RC_LEN] = {MP2_POSEIDON2_RC_LIST};
static __device__ const u64 g_p2_diag[MP2_POSEIDON2_DIAG_LEN] = {MP2_POSEIDON2_DIAG_LIST};

GL_DEV u32 shfl32(u32 v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// Poseidon: y_r = sum_i CIRC[i]*x[(r+i)%12] + 8*x[0]*[r==0] + rc
GL_DEV u64 coop_pos_mds(u64 x, int r, int gbase, u64 rc_next) {
  constexpr u32 CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
  u32 l0, l1, l2;
  pos_split3(x, l0, l1, l2);
  u32 a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    int src = r + i;
    src = src >= 24 ? src - 24 : src >= 12 ? src - 12 : src;
    src += gbase;
    a0 += shfl32(l0, src) * CIRC[i];
    a1 += shfl32(l1, src) * CIRC[i];
    a2 += shfl32(l2, src) * CIRC[i];
  }
  if (r == 0) {
    a0 += l0 << 3;
    a1 += l1 << 3;
    a2 += l2 << 3;
  }
  return gl_add_c(pos_merge3(a0, a1, a2), rc_next);
}
GL_DEV u64 coop_poseidon(u64 x, int r, int gbase) {
  const bool act = r < 12;
  x = act ? gl_add_c(x, g_pos_rc[r]) : 0;
#pragma unroll 1
  for (int round = 0; round < 30; round++) {
    // the next round's constant is requested before the S-box so that its latency hides behind it
    const u64 rcn = (act && round < 29) ? __ldg(g_pos_rc + 12 * (round + 1) + r) : 0;
    const bool full = round < 4 || round >= 26;
    if (full || r == 0) x = gl_pow7(x);
    x = coop_pos_mds(x, r, gbase, rcn);  // __shfl_sync re-converges the warp
  }
  return x;
}

// Poseidon2 external layer: M_E[r][j] = M4[r%4][j%4] * (2 if r/4 == j/4 else 1)
GL_DEV u64 coop_p2_ext(u64 x, int gbase, const u32 (&m)[4], int rblock, u64 rc_next) {
  u32 l0, l1, l2;
  pos_split3(x, l0, l1, l2);
  u32 a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
  for (int j = 0; j < 12; j++) {
    const u32 coef = m[j & 3] << (rblock == (j >> 2) ? 1 : 0);
    a0 += shfl32(l0, gbase + j) * coef;
    a1 += shfl32(l1, gbase + j) * coef;
    a2 += shfl32(l2, gbase + j) * coef;
  }
  return gl_add_c(pos_merge3(a0, a1, a2), rc_next);
}
// Poseidon2 internal layer: x*mu + sum(state)
GL_DEV u64 coop_p2_int(u64 x, int gbase, u64 mu) {
  u32 s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
  for (int j = 0; j < 12; j++) {
    const u32 lo = shfl32(lo32(x), gbase + j), hi = shfl32(hi32(x), gbase + j);
    asm("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.u32 %2, %2, 0;" : "+r"(s0), "+r"(s1), "+r"(s2) : "r"(lo), "r"(hi));
  }
  return gl_mul_add(x, mu, gl_reduce128w(s0, s1, s2, 0u));
}
GL_DEV u64 coop_poseidon2(u64 x, int r, int gbase) {
  const bool act = r < 12;
  const int rr = act ? r : 0;
  constexpr u32 M4[4][4] = {{5, 7, 1, 3}, {4, 6, 1, 1}, {1, 3, 5, 7}, {1, 1, 4, 6}};
  u32 m[4];
#pragma unroll
  for (int k = 0; k < 4; k++) m[k] = (rr & 3) == 0 ? M4[0][k] : (rr & 3) == 1 ? M4[1][k] : (rr & 3) == 2 ? M4[2][k] : M4[3][k];
  const int rblock = rr >> 2;
  const u64 mu = g_p2_diag[rr];
  if (!act) x = 0;
  x = coop_p2_ext(x, gbase, m, rblock, act ? g_p2_rc[rr] : 0);
#pragma unroll 1
  for (int k = 0; k < 4; k++) {
    const u64 rcn = (act && k < 3) ? __ldg(g_p2_rc + 12 * (k + 1) + rr) : 0;
    x = gl_pow7(x);
    x = coop_p2_ext(x, gbase, m, rblock, rcn);
  }
  u64 rci = __ldg(g_p2_rc + 48);
#pragma unroll 1
  for (int t = 0; t < 22; t++) {
    const u64 rc_now = rci;
    rci = __ldg(g_p2_rc + 48 + (t < 21 ? t + 1 : t));
    if (r == 0) x = gl_pow7(gl_add_c(x, rc_now));
    x = coop_p2_int(x, gbase, mu);
  }
  x = act ? gl_add_c(x, g_p2_rc[70 + rr]) : 0;
#pragma unroll 1
  for (int k = 0; k < 4; k++) {
    const u64 rcn = (act && k < 3) ? __ldg(g_p2_rc + 70 + 12 * (k + 1) + rr) : 0;
    x = gl_pow7(x);
    x = coop_p2_ext(x, gbase, m, rblock, rcn);
  }
  return x;
}

template <u32 KIND>
__global__ void __launch_bounds__(128)
k_merkle_layer_coop(u64 *__restrict__ digests, u64 *__restrict__ cap, u32 h, u32 layer, size_t nnodes) {
  size_t t = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;  // node
  const int r = threadIdx.x & 15, gbase = threadIdx.x & 16;
  const bool active = t < nnodes;
  if (!active) t = nnodes - 1;
  u32 width_log = h - layer;
  size_t sub = t >> width_log, m = t & (((size_t)1 << width_log) - 1);
  size_t per = 2 * (((size_t)1 << h) - 1);
  u64 *base = digests + 4 * sub * per;
  const u64 *ch = base + 4 * node_slot(layer - 1, 2 * m);  // the two children: 8 contiguous u64
  u64 x = r < 8 ? ch[r] : 0;
  x = KIND == MP2_HASH_POSEIDON2 ? coop_poseidon2(x, r, gbase) : coop_poseidon(x, r, gbase);
  u64 *dst = layer == h ? cap + 4 * sub : base + 4 * node_slot(layer, m);
  if (active && r < 4) dst[r] = gl_canon(x);
}

template <u32 KIND>
__global__ void __launch_bounds__(MP2_HASH_BLOCK)
k_two_to_one(const u64 *__restrict__ a, const u64 *__restrict__ b, size_t count, u64 *__restrict__ out) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = t < count;
  if (!active) t = count - 1;
  u64 st[12];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    st[i] = a[4 * t + i];
    st[4 + i] = b[4 * t + i];
    st[8 + i] = 0;
  }
  permute<KIND, true>(st);
  if (active) store_digest(out + 4 * t, st);
}

template <u32 KIND>
__global__ void __launch_bounds__(MP2_HASH_BLOCK) k_permute(u64 *__restrict__ states, size_t count) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = t < count;
  if (!active) t = count - 1;
  u64 st[12];
#pragma unroll
  for (int i = 0; i < 12; i++) st[i] = states[12 * t + i];
  permute<KIND, true>(st);
  if (active) {
#pragma unroll
    for (int i = 0; i < 12; i++) states[12 * t + i] = gl_canon(st[i]);
  }
}

// hash_no_pad over row-major inputs (never the no-op branch); len == 0 -> zeros
template <u32 KIND>
__global__ void __launch_bounds__(MP2_HASH_BLOCK)
k_hash_no_pad(const u64 *__restrict__ in, size_t count, size_t len, u64 *__restrict__ out) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = t < count;
  if (!active) t = count - 1;
  u64 st[12];
#pragma unroll
  for (int i = 0; i < 12; i++) st[i] = 0;
  const u64 *src = in + t * len;
  for (size_t c0 = 0; c0 < len; c0 += 8) {
    size_t m = len - c0 < 8 ? len - c0 : 8;
#pragma unroll
    for (int j = 0; j < 8; j++)
      if ((size_t)j < m) st[j] = src[c0 + j];
    permute<KIND, true>(st);
  }
  if (active) store_digest(out + 4 * t, st);
}

// fri_proof_of_work: one candidate per thread, smallest hit wins
struct PowState {
  u64 s[12];
};
template <u32 KIND>
__global__ void __launch_bounds__(MP2_HASH_BLOCK)
k_pow_search(PowState init, u32 pos, u32 min_lz, u64 start, u64 count, u64 *__restrict__ best) {
  const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const u64 cand = start + (t < count ? t : count - 1);
  u64 st[12];
#pragma unroll
  for (int i = 0; i < 12; i++) st[i] = (u32)i == pos ? cand : init.s[i];
  permute<KIND, true>(st);
  const u64 v = gl_canon(st[7]);  // duplex_state.squeeze().iter().last(), RATE = 8
  if (t < count && (u32)__clzll((long long)v) >= min_lz) atomicMin(best, cand);
}

// get_lde_values: out[r][c] = leaf row row_idx[r]
__global__ void k_gather_rows(const u64 *__restrict__ rowmajor, const u64 *__restrict__ colmajor,
                              size_t stride, u32 ncols, const u64 *__restrict__ row_idx, size_t nrows,
                              u64 *__restrict__ out) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nrows * ncols) return;
  size_t r = t / ncols, c = t % ncols;
  size_t L = row_idx[r];
  out[t] = rowmajor ? rowmajor[L * ncols + c] : colmajor[c * stride + L];
}

// out[t] = digests[slot[t]] (4 x u64 each)
__global__ void k_gather_digests(const u64 *__restrict__ digests, const u64 *__restrict__ slot, size_t n,
                                 u64 *__restrict__ out) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 4 * n) return;
  out[t] = digests[4 * slot[t >> 2] + (t & 3)];
}

static inline unsigned grid_for(size_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }


// ---- leaf-hash launch shape ----------------------------------------------------------------------
// 128-thread CTAs (4 warps kept in step by the per-round barrier), as many resident as the register
// file allows (5-6).  A leaf is an indivisible unit, so 2^17 leaves on 148 SMs quantise into 1.4 "waves";
// capping the resident CTAs to even that out was measured to change nothing (the SM's throughput scales
// with its resident warps), so the only knob left is MP2_HASH_CTAS for experiments: it caps the resident
// CTAs per SM by requesting dynamic shared memory that is never touched.
struct HashLaunch {
  size_t smem;  // dynamic shared memory that enforces the cap (0 = none)
};

static int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

template <typename K>
static Status plan_hash_launch(K kernel, HashLaunch *out) {
  out->smem = 0;
  const int forced = env_int("MP2_HASH_CTAS", 0);
  if (forced > 0) {
    size_t per = (size_t)(228 * 1024) / forced - 1024;  // each CTA also reserves 1 KB
    per &= ~(size_t)127;
    if (per > 227 * 1024) per = 227 * 1024;
    MP2_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)per));
    out->smem = per;
  }
  return "";
}

template <u32 KIND, bool COLMAJOR, int BLOCK>
static Status launch_leaf_hash_b(const u64 *in, size_t stride, u32 ncols, size_t leaf_begin, size_t leaf_end, u32 h,
                                 u64 *leaves_out, u64 *digests, u64 *cap, cudaStream_t st) {
  HashLaunch hl;
  const size_t nleaves = leaf_end - leaf_begin;
  MP2_TRY(plan_hash_launch(k_leaf_hash<KIND, COLMAJOR, BLOCK>, &hl));
  { ProfScope _p("k_leaf_hash", st); k_leaf_hash<KIND, COLMAJOR, BLOCK><<<grid_for(nleaves, BLOCK), BLOCK, hl.smem, st>>>(in, stride, ncols, leaf_begin, leaf_end, h,
                                                                                      leaves_out, digests, cap); }
  MP2_LAUNCH_CHECK();
  return "";
}

template <u32 KIND, bool COLMAJOR>
static Status launch_leaf_hash(const u64 *in, size_t stride, u32 ncols, size_t leaf_begin, size_t leaf_end, u32 h,
                               u64 *leaves_out, u64 *digests, u64 *cap, cudaStream_t st) {
  switch (env_int("MP2_HASH_BLOCK", MP2_HASH_BLOCK)) {
    case 64: return launch_leaf_hash_b<KIND, COLMAJOR, 64>(in, stride, ncols, leaf_begin, leaf_end, h, leaves_out, digests, cap, st);
    case 256: return launch_leaf_hash_b<KIND, COLMAJOR, 256>(in, stride, ncols, leaf_begin, leaf_end, h, leaves_out, digests, cap, st);
    case 512: return launch_leaf_hash_b<KIND, COLMAJOR, 512>(in, stride, ncols, leaf_begin, leaf_end, h, leaves_out, digests, cap, st);
    case 640: return launch_leaf_hash_b<KIND, COLMAJOR, 640>(in, stride, ncols, leaf_begin, leaf_end, h, leaves_out, digests, cap, st);
    default: return launch_leaf_hash_b<KIND, COLMAJOR, 128>(in, stride, ncols, leaf_begin, leaf_end, h, leaves_out, digests, cap, st);
  }
}

#ifndef MP2_COOP_LEVEL_NODES
#define MP2_COOP_LEVEL_NODES 2048
#endif

// Levels of `nsub` consecutive cap subtrees whose digest chunks start at `digests` (and cap entries at `cap`).
template <u32 KIND>
static Status build_levels(u64 *digests, u64 *cap, u32 h, size_t nsub, cudaStream_t st) {
  const size_t coop_below = (size_t)env_int("MP2_COOP_LEVEL_NODES", MP2_COOP_LEVEL_NODES);
  for (u32 layer = 1; layer <= h; layer++) {
    size_t nnodes = ((size_t)1 << (h - layer)) * nsub;
    if (nnodes <= coop_below) {
      ProfScope _p("k_merkle_layer_coop", st);
      k_merkle_layer_coop<KIND><<<grid_for(nnodes * 16, 128), 128, 0, st>>>(digests, cap, h, layer, nnodes);
    } else {
      ProfScope _p("k_merkle_layer", st);
      k_merkle_layer<KIND><<<grid_for(nnodes, MP2_HASH_BLOCK), MP2_HASH_BLOCK, 0, st>>>(digests, cap, h, layer, nnodes);
    }
    MP2_LAUNCH_CHECK();
  }
  return "";
}

static Status check_tree_args(size_t nleaves, u32 cap_height, u32 hash_kind, u32 *h_out) {
  int lg = log2_exact(nleaves);
  if (lg < 0) return "MerkleTree::new: number of leaves (" + std::to_string(nleaves) + ") is not a power of two";
  if ((int)cap_height > lg)
    return "MerkleTree::new: cap_height=" + std::to_string(cap_height) + " should be at most log2(leaves.len())=" + std::to_string(lg);
  if (hash_kind > 1) return "unknown hash_kind " + std::to_string(hash_kind);
  *h_out = (u32)lg - cap_height;
  return "";
}

Status merkle_colmajor_leaves(const u64 *lde, size_t lde_stride, size_t ncols, size_t nleaves, u32 cap_height,
                              u32 hash_kind, size_t leaf_begin, size_t leaf_end, u64 *leaves_out, u64 *digests,
                              u64 *cap, cudaStream_t st) {
  u32 h;
  MP2_TRY(check_tree_args(nleaves, cap_height, hash_kind, &h));
  if (ncols == 0 || ncols > 0xFFFFFFFFu) return "bad number of columns";
  if (leaf_begin >= leaf_end || leaf_end > nleaves) return "bad leaf range";
  if (hash_kind == MP2_HASH_POSEIDON2)
    return launch_leaf_hash<MP2_HASH_POSEIDON2, true>(lde, lde_stride, (u32)ncols, leaf_begin, leaf_end, h, leaves_out, digests, cap, st);
  return launch_leaf_hash<MP2_HASH_POSEIDON, true>(lde, lde_stride, (u32)ncols, leaf_begin, leaf_end, h, leaves_out, digests, cap, st);
}

Status merkle_levels(size_t nleaves, u32 cap_height, u32 hash_kind, u64 *digests, u64 *cap, cudaStream_t st) {
  u32 h;
  MP2_TRY(check_tree_args(nleaves, cap_height, hash_kind, &h));
  return hash_kind == MP2_HASH_POSEIDON2 ? build_levels<MP2_HASH_POSEIDON2>(digests, cap, h, (size_t)1 << cap_height, st)
                                         : build_levels<MP2_HASH_POSEIDON>(digests, cap, h, (size_t)1 << cap_height, st);
}

// the levels of cap subtrees [sub0, sub0 + nsub) only (their digest chunks are contiguous): lets a caller copy one
// group's digests out while the next group is still being built
Status merkle_levels_subtrees(size_t nleaves, u32 cap_height, u32 hash_kind, u64 *digests, u64 *cap, size_t sub0,
                              size_t nsub, cudaStream_t st) {
  u32 h;
  MP2_TRY(check_tree_args(nleaves, cap_height, hash_kind, &h));
  if (sub0 + nsub > ((size_t)1 << cap_height)) return "subtree range out of bounds";
  const size_t per = 2 * (((size_t)1 << h) - 1);
  u64 *d = digests + 4 * sub0 * per, *c = cap + 4 * sub0;
  return hash_kind == MP2_HASH_POSEIDON2 ? build_levels<MP2_HASH_POSEIDON2>(d, c, h, nsub, st)
                                         : build_levels<MP2_HASH_POSEIDON>(d, c, h, nsub, st);
}

Status merkle_colmajor(const u64 *lde, size_t lde_stride, size_t ncols, size_t nleaves, u32 cap_height,
                       u32 hash_kind, u64 *leaves_out, u64 *digests, u64 *cap, cudaStream_t st) {
  MP2_TRY(merkle_colmajor_leaves(lde, lde_stride, ncols, nleaves, cap_height, hash_kind, 0, nleaves, leaves_out, digests,
                                 cap, st));
  return merkle_levels(nleaves, cap_height, hash_kind, digests, cap, st);
}

Status merkle_rowmajor(const u64 *leaves, size_t nleaves, size_t leaf_len, u32 cap_height, u32 hash_kind,
                       u64 *digests, u64 *cap, cudaStream_t st) {
  u32 h;
  MP2_TRY(check_tree_args(nleaves, cap_height, hash_kind, &h));
  if (leaf_len > 0xFFFFFFFFu) return "leaf too long";
  if (hash_kind == MP2_HASH_POSEIDON2)
    MP2_TRY((launch_leaf_hash<MP2_HASH_POSEIDON2, false>(leaves, leaf_len, (u32)leaf_len, 0, nleaves, h, nullptr, digests, cap, st)));
  else
    MP2_TRY((launch_leaf_hash<MP2_HASH_POSEIDON, false>(leaves, leaf_len, (u32)leaf_len, 0, nleaves, h, nullptr, digests, cap, st)));
  return hash_kind == MP2_HASH_POSEIDON2 ? build_levels<MP2_HASH_POSEIDON2>(digests, cap, h, (size_t)1 << cap_height, st)
                                         : build_levels<MP2_HASH_POSEIDON>(digests, cap, h, (size_t)1 << cap_height, st);
}

Status merkle_ragged(const u64 *flat, const u64 *offsets, size_t nleaves, u32 cap_height, u32 hash_kind,
                     u64 *digests, u64 *cap, cudaStream_t st) {
  u32 h;
  MP2_TRY(check_tree_args(nleaves, cap_height, hash_kind, &h));
  unsigned g = grid_for(nleaves, MP2_HASH_BLOCK);
  if (hash_kind == MP2_HASH_POSEIDON2)
    { ProfScope _p("k_leaf_hash_ragged", st); k_leaf_hash_ragged<MP2_HASH_POSEIDON2><<<g, MP2_HASH_BLOCK, 0, st>>>(flat, offsets, nleaves, h, digests, cap); }
  else
    { ProfScope _p("k_leaf_hash_ragged", st); k_leaf_hash_ragged<MP2_HASH_POSEIDON><<<g, MP2_HASH_BLOCK, 0, st>>>(flat, offsets, nleaves, h, digests, cap); }
  MP2_LAUNCH_CHECK();
  return hash_kind == MP2_HASH_POSEIDON2 ? build_levels<MP2_HASH_POSEIDON2>(digests, cap, h, (size_t)1 << cap_height, st)
                                         : build_levels<MP2_HASH_POSEIDON>(digests, cap, h, (size_t)1 << cap_height, st);
}

Status hash_no_pad_batch(const u64 *inputs, size_t count, size_t input_len, u32 hash_kind, u64 *out,
                         cudaStream_t st) {
  if (hash_kind > 1) return "unknown hash_kind " + std::to_string(hash_kind);
  if (count == 0) return "";
  unsigned g = grid_for(count, MP2_HASH_BLOCK);
  if (hash_kind == MP2_HASH_POSEIDON2) { ProfScope _p("k_hash_no_pad", st); k_hash_no_pad<MP2_HASH_POSEIDON2><<<g, MP2_HASH_BLOCK, 0, st>>>(inputs, count, input_len, out); }
  else { ProfScope _p("k_hash_no_pad", st); k_hash_no_pad<MP2_HASH_POSEIDON><<<g, MP2_HASH_BLOCK, 0, st>>>(inputs, count, input_len, out); }
  MP2_LAUNCH_CHECK();
  return "";
}

Status two_to_one_batch(const u64 *a, const u64 *b, size_t count, u32 hash_kind, u64 *out, cudaStream_t st) {
  if (hash_kind > 1) return "unknown hash_kind " + std::to_string(hash_kind);
  if (count == 0) return "";
  unsigned g = grid_for(count, MP2_HASH_BLOCK);
  if (hash_kind == MP2_HASH_POSEIDON2) { ProfScope _p("k_two_to_one", st); k_two_to_one<MP2_HASH_POSEIDON2><<<g, MP2_HASH_BLOCK, 0, st>>>(a, b, count, out); }
  else { ProfScope _p("k_two_to_one", st); k_two_to_one<MP2_HASH_POSEIDON><<<g, MP2_HASH_BLOCK, 0, st>>>(a, b, count, out); }
  MP2_LAUNCH_CHECK();
  return "";
}

Status permute_batch(u64 *states, size_t count, u32 hash_kind, cudaStream_t st) {
  if (hash_kind > 1) return "unknown hash_kind " + std::to_string(hash_kind);
  if (count == 0) return "";
  unsigned g = grid_for(count, MP2_HASH_BLOCK);
  if (hash_kind == MP2_HASH_POSEIDON2) { ProfScope _p("k_permute", st); k_permute<MP2_HASH_POSEIDON2><<<g, MP2_HASH_BLOCK, 0, st>>>(states, count); }
  else { ProfScope _p("k_permute", st); k_permute<MP2_HASH_POSEIDON><<<g, MP2_HASH_BLOCK, 0, st>>>(states, count); }
  MP2_LAUNCH_CHECK();
  return "";
}

Status pow_search(const u64 *state12_host, u32 pos, u32 min_lz, u32 hash_kind, u64 start, u64 count, u64 *found,
                  cudaStream_t st) {
  if (hash_kind > 1) return "unknown hash_kind " + std::to_string(hash_kind);
  if (pos >= 8) return "witness position must be inside the sponge rate";
  PowState init;
  for (int i = 0; i < 12; i++) init.s[i] = state12_host[i];
  DevBuf best;
  MP2_TRY(best.alloc(1, st));
  u64 *d_best = best.p;
  MP2_CUDA(cudaMemsetAsync(d_best, 0xFF, sizeof(u64), st));
  unsigned g = grid_for(count, MP2_HASH_BLOCK);
  if (hash_kind == MP2_HASH_POSEIDON2) { ProfScope _p("k_pow_search", st); k_pow_search<MP2_HASH_POSEIDON2><<<g, MP2_HASH_BLOCK, 0, st>>>(init, pos, min_lz, start, count, d_best); }
  else { ProfScope _p("k_pow_search", st); k_pow_search<MP2_HASH_POSEIDON><<<g, MP2_HASH_BLOCK, 0, st>>>(init, pos, min_lz, start, count, d_best); }
  MP2_LAUNCH_CHECK();
  MP2_CUDA(cudaMemcpyAsync(found, d_best, sizeof(u64), cudaMemcpyDeviceToHost, st));
  MP2_CUDA(cudaStreamSynchronize(st));
  return "";
}

Status gather_rows(const u64 *rowmajor, const u64 *colmajor, size_t stride, size_t ncols, const u64 *row_idx,
                   size_t nrows, u64 *out, cudaStream_t st) {
  if (nrows == 0) return "";
  { ProfScope _p("k_gather_rows", st); k_gather_rows<<<grid_for(nrows * ncols, 256), 256, 0, st>>>(rowmajor, colmajor, stride, (u32)ncols, row_idx, nrows, out); }
  MP2_LAUNCH_CHECK();
  return "";
}

}  // namespace mp2

namespace mp2 {
Status merkle_open(const u64 *rowmajor, const u64 *colmajor, size_t stride, size_t leaf_len, const u64 *digests,
                   size_t nleaves, u32 cap_height, const u64 *idx_host, size_t count, u64 *rows_out,
                   u64 *sib_out, cudaStream_t st) {
  if (count == 0) return "";
  int lg = log2_exact(nleaves);
  if (lg < 0 || (int)cap_height > lg) return "MerkleTree::prove: bad tree shape";
  const u32 h = (u32)lg - cap_height;
  const size_t per = 2 * (((size_t)1 << h) - 1);
  std::vector<u64> slots(count * (h ? h : 1));
  for (size_t q = 0; q < count; q++) {
    const size_t leaf = idx_host[q];
    if (leaf >= nleaves) return "MerkleTree::prove: leaf_index out of range";
    size_t pair_index = leaf & (((size_t)1 << h) - 1);
    const size_t base = (leaf >> h) * per;
    for (u32 i = 0; i < h; i++) {  // closed-form sibling slots
      size_t parity = pair_index & 1;
      pair_index >>= 1;
      slots[q * h + i] = base + 2 * ((pair_index << (i + 1)) + ((size_t)1 << i) - 1) + (1 - parity);
    }
  }
  DevBuf d_idx, d_rows, d_slots, d_sib;
  MP2_TRY(d_idx.alloc(count, st));
  MP2_CUDA(cudaMemcpyAsync(d_idx.p, idx_host, sizeof(u64) * count, cudaMemcpyHostToDevice, st));
  if (rows_out && leaf_len) {
    MP2_TRY(d_rows.alloc(count * leaf_len, st));
    MP2_TRY(gather_rows(rowmajor, colmajor, stride, leaf_len, d_idx.p, count, d_rows.p, st));
    MP2_CUDA(cudaMemcpyAsync(rows_out, d_rows.p, sizeof(u64) * count * leaf_len, cudaMemcpyDeviceToHost, st));
  }
  if (sib_out && h) {
    MP2_TRY(d_slots.alloc(count * h, st));
    MP2_TRY(d_sib.alloc(4 * count * h, st));
    MP2_CUDA(cudaMemcpyAsync(d_slots.p, slots.data(), sizeof(u64) * count * h, cudaMemcpyHostToDevice, st));
    { ProfScope _p("k_gather_digests", st); k_gather_digests<<<grid_for(4 * count * h, 256), 256, 0, st>>>(digests, d_slots.p, count * h, d_sib.p); }
    MP2_LAUNCH_CHECK();
    MP2_CUDA(cudaMemcpyAsync(sib_out, d_sib.p, sizeof(u64) * 4 * count * h, cudaMemcpyDeviceToHost, st));
  }
  MP2_CUDA(cudaStreamSynchronize(st));
  return "";
}
}  // namespace mp2
