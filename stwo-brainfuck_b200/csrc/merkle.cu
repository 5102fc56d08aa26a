// Blake2s Merkle layer hashing and proof-of-work grinding for sm_100a.
//
// Replaces SimdBackend's MerkleOps<Blake2sMerkleHasher>::commit_on_layer and GrindOps<Blake2sChannel>::grind
// (stwo-prover 0.1.1 @ 31e8dbc core/backend/simd/{blake2s.rs,grind.rs}, core/vcs/blake2_merkle.rs),
// reached from tree_builder.commit at crates/brainfuck_prover/src/brainfuck_air/mod.rs:500,583,723 and from every FRI
// layer commit / the grind inside prover::prove (:732).
//
// Node function: state = 0^8; if children: state = F(state, left||right); then F(state, 16 column words) per chunk
// (zero padded), all counters/flags zero.  One thread per row: a warp covers 32 consecutive rows, so every column
// load is one coalesced 128-byte line and the 32-byte digest store fills whole sectors.  The whole compression is
// unrolled with compile-time sigma so message and state words live in registers (no local memory).
#include "kernels.cuh"
#include "blake2s.cuh"
#include <cstdlib>

namespace sb {

// rep_log > 0: every column repeats each value 2^rep_log times and so do the children, hence so do the nodes of this layer;
// a thread hashes the first node of its group and stores the digest 2^rep_log times (rows = number of groups).
// NC4: the layer injects at most four columns (every FRI layer, the composition tree, one IsFirst column per preprocessed
// layer, most interaction layers): one column block whose message words 4..15 are compile-time zeros.
template <bool HAS_PREV, bool NC4>
__global__ void __launch_bounds__(256) commit_layer_kernel(uint32_t rows, const uint32_t* __restrict__ prev,
                                                           const uint32_t* const* __restrict__ cols, uint32_t ncols,
                                                           uint32_t* __restrict__ out, uint32_t one, uint32_t rep_log) {
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= rows) return;
  const uint32_t i = g << rep_log;
  uint32_t h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint32_t m[16];
  if (HAS_PREV) {
    const uint4* pc = reinterpret_cast<const uint4*>(prev) + (size_t)i * 4;
    uint4 a = __ldg(pc), b = __ldg(pc + 1), c = __ldg(pc + 2), d = __ldg(pc + 3);
    m[0] = a.x; m[1] = a.y; m[2] = a.z; m[3] = a.w; m[4] = b.x; m[5] = b.y; m[6] = b.z; m[7] = b.w;
    m[8] = c.x; m[9] = c.y; m[10] = c.z; m[11] = c.w; m[12] = d.x; m[13] = d.y; m[14] = d.z; m[15] = d.w;
    b2s_compress(h, m, one);
  }
  if (NC4) {
    if (ncols) {
#pragma unroll
      for (uint32_t j = 0; j < 4; j++) m[j] = (j < ncols) ? __ldg(cols[j] + i) : 0u;
      b2s_compress<4>(h, m, one);
    }
  } else {
    for (uint32_t c0 = 0; c0 < ncols; c0 += 16) {
#pragma unroll
      for (uint32_t j = 0; j < 16; j++) m[j] = (c0 + j < ncols) ? __ldg(cols[c0 + j] + i) : 0u;
      b2s_compress(h, m, one);
    }
  }
  uint4* o = reinterpret_cast<uint4*>(out) + (size_t)i * 2;
  const uint4 lo = make_uint4(h[0], h[1], h[2], h[3]), hi = make_uint4(h[4], h[5], h[6], h[7]);
  for (uint32_t k = 0; k < (1u << rep_log); k++) { o[2 * k] = lo; o[2 * k + 1] = hi; }
}

int launch_commit_layer(uint32_t log_size, const uint32_t* prev, const uint32_t* const* cols, uint32_t ncols,
                        uint32_t* out, cudaStream_t st, uint32_t rep_log) {
  if (rep_log > log_size) rep_log = log_size;
  uint32_t rows = 1u << (log_size - rep_log);
  uint32_t threads = rows < 256 ? (rows < 32 ? 32 : rows) : 256;
  uint32_t blocks = (rows + threads - 1) / threads;
  static const bool generic_only = getenv("SC_MERKLE_GENERIC") != nullptr;   // A/B switch for tools/merkle_bench.py
  const bool nc4 = ncols <= 4 && !generic_only;
  if (prev) {
    if (nc4) commit_layer_kernel<true, true><<<blocks, threads, 0, st>>>(rows, prev, cols, ncols, out, 1u, rep_log);
    else commit_layer_kernel<true, false><<<blocks, threads, 0, st>>>(rows, prev, cols, ncols, out, 1u, rep_log);
  } else {
    if (nc4) commit_layer_kernel<false, true><<<blocks, threads, 0, st>>>(rows, prev, cols, ncols, out, 1u, rep_log);
    else commit_layer_kernel<false, false><<<blocks, threads, 0, st>>>(rows, prev, cols, ncols, out, 1u, rep_log);
  }
  g_launch_count++;
  return (int)cudaGetLastError();
}

// The top of a tree in one launch.  Layers of <= 2^TOP_LOG nodes take 2-3 us each and as long again to launch; a tree has
// TOP_LOG + 1 of them and a proof about thirty trees (four commitments + one per FRI layer).  One CTA walks the layers
// from `top_log` down to the root, a thread per node, __syncthreads between layers (the children were written by this CTA,
// so they are re-read with ordinary loads, not through the read-only path).
struct TopArgs {
  const uint32_t* prev;            // layer top_log + 1 (NULL when top_log is the deepest layer)
  const uint32_t* const* cols;     // column pointers of layers top_log, top_log-1, ..., 0, concatenated
  uint32_t col_off[MERKLE_TOP_LOG + 2];  // cols of layer top_log - k are cols[col_off[k] .. col_off[k+1])
  uint32_t* out[MERKLE_TOP_LOG + 1];     // out[k] = layer top_log - k
  uint32_t top_log;
  uint32_t one;
};
__global__ void __launch_bounds__(1 << MERKLE_TOP_LOG) commit_top_kernel(TopArgs a) {
  const uint32_t i = threadIdx.x;
  const uint32_t* prev = a.prev;
  for (uint32_t k = 0; k <= a.top_log; k++) {
    const uint32_t lg = a.top_log - k;
    if (i < (1u << lg)) {
      uint32_t h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      uint32_t m[16];
      if (prev) {
        const uint4* pc = reinterpret_cast<const uint4*>(prev) + (size_t)i * 4;
        uint4 x = pc[0], y = pc[1], z = pc[2], w = pc[3];
        m[0] = x.x; m[1] = x.y; m[2] = x.z; m[3] = x.w; m[4] = y.x; m[5] = y.y; m[6] = y.z; m[7] = y.w;
        m[8] = z.x; m[9] = z.y; m[10] = z.z; m[11] = z.w; m[12] = w.x; m[13] = w.y; m[14] = w.z; m[15] = w.w;
        b2s_compress(h, m, a.one);
      }
      const uint32_t c_lo = a.col_off[k], c_hi = a.col_off[k + 1];
      for (uint32_t c0 = c_lo; c0 < c_hi; c0 += 16) {
#pragma unroll
        for (uint32_t j = 0; j < 16; j++) m[j] = (c0 + j < c_hi) ? __ldg(a.cols[c0 + j] + i) : 0u;
        b2s_compress(h, m, a.one);
      }
      uint4* o = reinterpret_cast<uint4*>(a.out[k]) + (size_t)i * 2;
      o[0] = make_uint4(h[0], h[1], h[2], h[3]);
      o[1] = make_uint4(h[4], h[5], h[6], h[7]);
    }
    __syncthreads();
    prev = a.out[k];
  }
}
// The middle of a tree in one launch: CTA b owns nodes [b << S, (b + 1) << S) of layer L and walks its sub-tree up to the one
// node of layer L - S, handing digests from layer to layer through shared memory (every layer is also written to global
// memory: decommitment reads them).  Layers of 2^10 .. 2^19 nodes cost a launch each otherwise — 5-6 us apiece against
// 0.7 us of hashing once they stop filling the machine — and a proof has some thirty trees.
struct SubtreeArgs {
  const uint32_t* prev;                     // layer L + 1 (NULL when L is the deepest layer)
  const uint32_t* const* cols;              // column pointers of layers L, L-1, ..., L-S, concatenated
  uint32_t col_off[MERKLE_SUB_MAX + 2];     // columns of layer L - k are cols[col_off[k] .. col_off[k+1])
  uint32_t* out[MERKLE_SUB_MAX + 1];        // out[k] = layer L - k
  uint32_t L, S, one;
};
__device__ __forceinline__ void subtree_columns(uint32_t h[8], uint32_t m[16], const uint32_t* const* cols, uint32_t c_lo, uint32_t c_hi,
                                                uint32_t node, uint32_t one) {
  if (c_hi - c_lo <= 4) {
    if (c_hi > c_lo) {
#pragma unroll
      for (uint32_t j = 0; j < 4; j++) m[j] = (c_lo + j < c_hi) ? __ldg(cols[c_lo + j] + node) : 0u;
      b2s_compress<4>(h, m, one);
    }
    return;
  }
  for (uint32_t c0 = c_lo; c0 < c_hi; c0 += 16) {
#pragma unroll
    for (uint32_t j = 0; j < 16; j++) m[j] = (c0 + j < c_hi) ? __ldg(cols[c0 + j] + node) : 0u;
    b2s_compress(h, m, one);
  }
}
__global__ void __launch_bounds__(256) commit_subtree_kernel(SubtreeArgs a) {
  // digests of the layer below, word-major (word w of node j at [w * stride + j]): the eight stores of a node and the eight
  // 64-bit loads of a parent's two children are conflict-free (node-major uint4 rows cost 2- and 4-way bank conflicts: ncu
  // counted 1.04 M conflicts in 1.6 M shared wavefronts)
  extern __shared__ uint32_t sub_sm[];
  const uint32_t stride = 1u << a.S;
  uint32_t* cur = sub_sm;
  uint32_t* nxt = sub_sm + 8u * stride;
  for (uint32_t k = 0; k <= a.S; k++) {
    const uint32_t cnt = 1u << (a.S - k), base = blockIdx.x << (a.S - k);
    for (uint32_t j = threadIdx.x; j < cnt; j += blockDim.x) {
      const uint32_t node = base + j;
      uint32_t h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      uint32_t m[16];
      if (k == 0) {
        if (a.prev) {
          const uint4* pc = reinterpret_cast<const uint4*>(a.prev) + (size_t)node * 4;
          const uint4 x = __ldg(pc), y = __ldg(pc + 1), z = __ldg(pc + 2), w = __ldg(pc + 3);
          m[0] = x.x; m[1] = x.y; m[2] = x.z; m[3] = x.w; m[4] = y.x; m[5] = y.y; m[6] = y.z; m[7] = y.w;
          m[8] = z.x; m[9] = z.y; m[10] = z.z; m[11] = z.w; m[12] = w.x; m[13] = w.y; m[14] = w.z; m[15] = w.w;
          b2s_compress(h, m, a.one);
        }
      } else {
#pragma unroll
        for (uint32_t w = 0; w < 8; w++) {
          const uint2 c = *reinterpret_cast<const uint2*>(cur + w * stride + 2 * j);   // word w of children 2j and 2j + 1
          m[w] = c.x; m[8 + w] = c.y;
        }
        b2s_compress(h, m, a.one);
      }
      subtree_columns(h, m, a.cols, a.col_off[k], a.col_off[k + 1], node, a.one);
      uint4* o = reinterpret_cast<uint4*>(a.out[k]) + (size_t)node * 2;
      o[0] = make_uint4(h[0], h[1], h[2], h[3]); o[1] = make_uint4(h[4], h[5], h[6], h[7]);
#pragma unroll
      for (uint32_t w = 0; w < 8; w++) nxt[w * stride + j] = h[w];
    }
    __syncthreads();
    uint32_t* t = cur; cur = nxt; nxt = t;
  }
}
// layers L, L-1, ..., L-S (S <= MERKLE_SUB_MAX); cols / col_off / out as in SubtreeArgs (col_off and out are host arrays)
int launch_commit_subtree(uint32_t L, uint32_t S, const uint32_t* prev, const uint32_t* const* cols, const uint32_t* col_off,
                          uint32_t* const* out, cudaStream_t st) {
  if (S > MERKLE_SUB_MAX || S > L) return -1;
  SubtreeArgs a;
  a.prev = prev; a.cols = cols; a.L = L; a.S = S; a.one = 1u;
  for (uint32_t k = 0; k <= S + 1; k++) a.col_off[k] = col_off[k];
  for (uint32_t k = 0; k <= S; k++) a.out[k] = out[k];
  const uint32_t per_cta = 1u << S, threads = per_cta < 256 ? (per_cta < 32 ? 32 : per_cta) : 256;
  const size_t smem = (size_t)(16u << S) * sizeof(uint32_t);   // two word-major buffers of 2^S nodes
  commit_subtree_kernel<<<1u << (L - S), threads, smem, st>>>(a); g_launch_count++;
  return (int)cudaGetLastError();
}

// cols: device array of the column pointers (layer top_log first); col_off / out as in TopArgs (host arrays).
int launch_commit_top(uint32_t top_log, const uint32_t* prev, const uint32_t* const* cols, const uint32_t* col_off,
                      uint32_t* const* out, cudaStream_t st) {
  if (top_log > MERKLE_TOP_LOG) return -1;
  TopArgs a;
  a.prev = prev; a.cols = cols; a.top_log = top_log; a.one = 1u;
  for (uint32_t k = 0; k <= top_log + 1; k++) a.col_off[k] = col_off[k];
  for (uint32_t k = 0; k <= top_log; k++) a.out[k] = out[k];
  uint32_t threads = 1u << top_log;
  if (threads < 32) threads = 32;
  commit_top_kernel<<<1, threads, 0, st>>>(a); g_launch_count++;
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------- grind
// Smallest nonce with trailing_zeros(F(digest, [nonce_lo, nonce_hi, 0...])) >= pow_bits (first 128 bits, LE).
__global__ void grind_kernel(uint32_t d0, uint32_t d1, uint32_t d2, uint32_t d3, uint32_t d4, uint32_t d5, uint32_t d6,
                             uint32_t d7, uint32_t pow_bits, unsigned long long base, unsigned long long* result, uint32_t one) {
  unsigned long long nonce = base + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t h[8] = {d0, d1, d2, d3, d4, d5, d6, d7};
  uint32_t m[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  m[0] = (uint32_t)nonce; m[1] = (uint32_t)(nonce >> 32);
  b2s_compress(h, m, one);
  uint32_t tz;
  if (h[0]) tz = __ffs(h[0]) - 1;
  else if (h[1]) tz = 32 + __ffs(h[1]) - 1;
  else if (h[2]) tz = 64 + __ffs(h[2]) - 1;
  else if (h[3]) tz = 96 + __ffs(h[3]) - 1;
  else tz = 128;
  if (tz >= pow_bits) atomicMin(result, nonce);
}

// Searches [0, 2^40) in batches; *d_result must be initialised to ~0ull by the caller; returns after the first
// batch that produced a hit (batches are scanned in increasing order, atomicMin keeps the smallest nonce).
int launch_grind(const uint32_t digest[8], uint32_t pow_bits, unsigned long long* d_result, cudaStream_t st) {
  // batches grow 2^14 -> 2^22: with the reference's pow_bits (5) the first 16 384 nonces contain a solution with probability
  // 1 - (31/32)^16384, and hashing four million of them first cost 0.17 ms per proof
  unsigned long long batch = 1ull << 14;
  for (unsigned long long base = 0; base < (1ull << 40); base += batch, batch = batch < (1ull << 22) ? batch << 4 : batch) {
    grind_kernel<<<(unsigned)(batch / 256), 256, 0, st>>>(digest[0], digest[1], digest[2], digest[3], digest[4], digest[5],
                                                          digest[6], digest[7], pow_bits, base, d_result, 1u); g_launch_count++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    unsigned long long r;
    e = cudaMemcpyAsync(&r, d_result, 8, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return (int)e;
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return (int)e;
    if (r != ~0ull) return 0;
  }
  return -2;
}

}  // namespace sb
