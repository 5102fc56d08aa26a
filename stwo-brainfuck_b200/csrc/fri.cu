// FriProver::commit with the Fiat–Shamir channel on the device (stwo-prover 0.1.1 @ 31e8dbc core/fri.rs `FriProver::commit`,
// core/channel/blake2s.rs, core/vcs/blake2_merkle.rs `mix_root`; reached from prover::prove at
// crates/brainfuck_prover/src/brainfuck_air/mod.rs:732).
//
// Upstream's commit phase is a chain  fold -> Merkle tree -> root -> channel.mix_root -> channel.draw_felt -> next fold:
// every layer waits for a 32-byte read-back and a host-side hash before its folding coefficient exists (25 layers for a
// 2^26-row quotient).  Here the channel state lives in device memory: a one-thread kernel mixes each root and draws the next
// coefficient, the fold kernels read the coefficient from device memory, and the host replays the transcript ONCE from the
// roots it reads back at the end (same digests, same coefficients: the proof is unchanged).  The layers of <= 2^FRI_TAIL_LOG
// values — folds, leaf hashes, trees, channel — are one persistent CTA working out of shared memory (fri_tail_kernel).
#include "blake2s.cuh"
#include "kernels.cuh"

namespace sb {

// ---------------------------------------------------------------- Blake2s-256 (the real hash: IV, parameter block, counter, final flag)
// out = Blake2s-256 of the 64-byte message m (one final block: t = 64, f0 = ~0) on the unrolled compression of blake2s.cuh
__device__ __noinline__ void chan_hash64(const uint32_t m[16], uint32_t out[8]) {
  uint32_t h[8] = {0x6A09E667u ^ 0x01010020u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
  b2s_compress(h, m, 1u, 64u, 0xFFFFFFFFu);
  for (int i = 0; i < 8; i++) out[i] = h[i];
}
// Blake2sMerkleChannel::mix_root followed by Blake2sChannel::draw_felt: digest <- H(digest || root); then H(digest || counter)
// with counter 0, 1, ... until all eight words are < 2P; the first four, reduced, are the felt.
__device__ void chan_mix_root_draw(uint32_t* digest, const uint32_t* root, uint32_t* alpha_out) {
  uint32_t m[16], h[8], w[8];
  for (int i = 0; i < 8; i++) { m[i] = digest[i]; m[8 + i] = root[i]; }
  chan_hash64(m, h);
  for (int i = 0; i < 8; i++) digest[i] = h[i];
  for (uint32_t n_sent = 0;; n_sent++) {
    for (int i = 0; i < 8; i++) { m[i] = h[i]; m[8 + i] = 0; }
    m[8] = n_sent;
    chan_hash64(m, w);
    bool ok = true;
    for (int i = 0; i < 8; i++) ok &= w[i] < 2u * P;
    if (ok) break;
  }
  for (int i = 0; i < 4; i++) alpha_out[i] = w[i] >= P ? w[i] - P : w[i];
}
__global__ void fri_channel_kernel(uint32_t* __restrict__ digest, const uint32_t* __restrict__ root, uint32_t* __restrict__ alpha_out,
                                   uint32_t* __restrict__ root_copy) {
  if (threadIdx.x || blockIdx.x) return;
  uint32_t d[8], r[8], a[4];
  for (int i = 0; i < 8; i++) { d[i] = digest[i]; r[i] = root[i]; root_copy[i] = r[i]; }
  chan_mix_root_draw(d, r, a);
  for (int i = 0; i < 8; i++) digest[i] = d[i];
  for (int i = 0; i < 4; i++) alpha_out[i] = a[i];
}
int launch_fri_channel(uint32_t* d_digest, const uint32_t* d_root, uint32_t* d_alpha_out, uint32_t* d_root_copy, cudaStream_t st) {
  fri_channel_kernel<<<1, 32, 0, st>>>(d_digest, d_root, d_alpha_out, d_root_copy); g_launch_count++;
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------- folds with the coefficient in device memory
struct FPtr4 { uint32_t* p[4]; };
struct FCPtr4 { const uint32_t* p[4]; };
__device__ __forceinline__ QM31 load_q(const uint32_t* p) { return q_make(p[0], p[1], p[2], p[3]); }
// same arithmetic as ops.cu fold_line_kernel / fold_circle_kernel
__global__ void __launch_bounds__(256) fold_line_dev_kernel(FCPtr4 s, FPtr4 d, uint32_t log, const uint32_t* __restrict__ alpha_p,
                                                            const uint32_t* __restrict__ itw_end) {
  const size_t half = (size_t)1 << (log - 1);
  const uint32_t* itw = itw_end - ((size_t)1 << log);
  const QM31 alpha = load_q(alpha_p);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += (size_t)gridDim.x * blockDim.x) {
    uint2 c0 = reinterpret_cast<const uint2*>(s.p[0])[i], c1 = reinterpret_cast<const uint2*>(s.p[1])[i];
    uint2 c2 = reinterpret_cast<const uint2*>(s.p[2])[i], c3 = reinterpret_cast<const uint2*>(s.p[3])[i];
    QM31 a = q_make(c0.x, c1.x, c2.x, c3.x), b = q_make(c0.y, c1.y, c2.y, c3.y);
    QM31 f0 = q_add(a, b), f1 = q_mulm(q_sub(a, b), __ldg(itw + i));
    QM31 r = q_add(f0, q_mul(alpha, f1));
    d.p[0][i] = r.a.a; d.p[1][i] = r.a.b; d.p[2][i] = r.b.a; d.p[3][i] = r.b.b;
  }
}
__device__ __forceinline__ uint32_t circle_itw(const uint32_t* l1, size_t i) {  // 1/y of pair i: [x, y] -> [y, -y, -x, x]
  const size_t pair = (i >> 2) * 2;
  const uint32_t x = l1[pair], y = l1[pair + 1], sel = (uint32_t)i & 3u;
  uint32_t t = sel < 2 ? y : x;
  if (sel == 1 || sel == 2) t = P - t;
  return t;
}
// first != 0: dst is not read (it would be all zeros: the first column folded into a fresh line evaluation)
__global__ void __launch_bounds__(256) fold_circle_dev_kernel(FCPtr4 s, FPtr4 d, uint32_t log, const uint32_t* __restrict__ alpha_p,
                                                              const uint32_t* __restrict__ itw_end, uint32_t first) {
  const size_t half = (size_t)1 << (log - 1);
  const uint32_t* l1 = itw_end - ((size_t)1 << (log - 1));
  const QM31 alpha = load_q(alpha_p), alpha_sq = q_mul(alpha, alpha);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t t = circle_itw(l1, i);
    uint2 c0 = reinterpret_cast<const uint2*>(s.p[0])[i], c1 = reinterpret_cast<const uint2*>(s.p[1])[i];
    uint2 c2 = reinterpret_cast<const uint2*>(s.p[2])[i], c3 = reinterpret_cast<const uint2*>(s.p[3])[i];
    QM31 a = q_make(c0.x, c1.x, c2.x, c3.x), b = q_make(c0.y, c1.y, c2.y, c3.y);
    QM31 f0 = q_add(a, b), f1 = q_mulm(q_sub(a, b), t);
    QM31 r = q_add(f0, q_mul(alpha, f1));
    if (!first) r = q_add(q_mul(q_make(d.p[0][i], d.p[1][i], d.p[2][i], d.p[3][i]), alpha_sq), r);
    d.p[0][i] = r.a.a; d.p[1][i] = r.a.b; d.p[2][i] = r.b.a; d.p[3][i] = r.b.b;
  }
}
static inline unsigned fold_grid(size_t n) { size_t b = (n + 255) / 256; return (unsigned)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b)); }
int launch_fold_line_dev(const uint32_t* const src[4], uint32_t log, const uint32_t* d_alpha, uint32_t* const dst[4], const uint32_t* itw_end,
                         cudaStream_t st) {
  if (log < 1) return -1;
  FCPtr4 s{{src[0], src[1], src[2], src[3]}};
  FPtr4 d{{dst[0], dst[1], dst[2], dst[3]}};
  fold_line_dev_kernel<<<fold_grid((size_t)1 << (log - 1)), 256, 0, st>>>(s, d, log, d_alpha, itw_end); g_launch_count++;
  return (int)cudaGetLastError();
}
int launch_fold_circle_dev(const uint32_t* const src[4], uint32_t log, const uint32_t* d_alpha, uint32_t* const dst[4], const uint32_t* itw_end,
                           bool first, cudaStream_t st) {
  if (log < 3) return -1;
  FCPtr4 s{{src[0], src[1], src[2], src[3]}};
  FPtr4 d{{dst[0], dst[1], dst[2], dst[3]}};
  fold_circle_dev_kernel<<<fold_grid((size_t)1 << (log - 1)), 256, 0, st>>>(s, d, log, d_alpha, itw_end, first ? 1u : 0u); g_launch_count++;
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------- the tail: every layer of <= 2^FRI_TAIL_LOG values in one CTA
// Shared memory: the line evaluation (4 coordinates x cap words), two digest buffers for the tree walk.  Per layer lg:
//   circle-fold the quotient column of log lg + 1 into the evaluation (if there is one); store the evaluation; hash the
//   leaves (4 values per node); walk the tree to the root; thread 0 mixes the root and draws the coefficient; fold.
__global__ void __launch_bounds__(1024) fri_tail_kernel(FriTailArgs a) {
  extern __shared__ uint32_t sm[];
  const uint32_t cap = 1u << a.start_log, tid = threadIdx.x, nt = blockDim.x;
  uint32_t* ev = sm;                 // ev[k * cap + i]
  uint32_t* dg0 = ev + 4 * cap;      // 8 words per node
  uint32_t* dg1 = dg0 + 8 * cap;
  __shared__ uint32_t s_alpha[4], s_digest[8], s_calpha[4];
  for (uint32_t i = tid; i < cap; i += nt)
    for (int k = 0; k < 4; k++) ev[k * cap + i] = a.layer_in[k] ? a.layer_in[k][i] : 0u;
  if (tid < 8) s_digest[tid] = a.digest[tid];
  if (tid < 4) s_calpha[tid] = a.circle_alpha[tid];
  __syncthreads();
  const QM31 ca = load_q(s_calpha), ca2 = q_mul(ca, ca);
  uint32_t off = 0, li = 0;
  for (uint32_t lg = a.start_log; lg > a.last_log; lg--, li++) {
    const uint32_t n = 1u << lg;
    // ---- circle column of log lg + 1 folded in with the FIRST layer's coefficient (fri.rs: fold_circle_into_line)
    const uint32_t* q0 = a.quot[li * 4];
    if (q0) {
      const uint32_t* l1 = a.itw_end - ((size_t)1 << lg);
      for (uint32_t i = tid; i < n; i += nt) {
        const uint32_t t = circle_itw(l1, i);
        uint32_t x[4], y[4];
        for (int k = 0; k < 4; k++) { const uint2 c = reinterpret_cast<const uint2*>(a.quot[li * 4 + k])[i]; x[k] = c.x; y[k] = c.y; }
        const QM31 p = q_make(x[0], x[1], x[2], x[3]), q = q_make(y[0], y[1], y[2], y[3]);
        const QM31 f0 = q_add(p, q), f1 = q_mulm(q_sub(p, q), t);
        const QM31 acc = q_make(ev[i], ev[cap + i], ev[2 * cap + i], ev[3 * cap + i]);
        const QM31 r = q_add(q_mul(acc, ca2), q_add(f0, q_mul(ca, f1)));
        ev[i] = r.a.a; ev[cap + i] = r.a.b; ev[2 * cap + i] = r.b.a; ev[3 * cap + i] = r.b.b;
      }
      __syncthreads();
    }
    // ---- the committed evaluation and its leaf hashes
    for (uint32_t i = tid; i < n; i += nt) {
      uint32_t h[8] = {0, 0, 0, 0, 0, 0, 0, 0}, m[16];
#pragma unroll
      for (int k = 0; k < 4; k++) { m[k] = ev[k * cap + i]; a.eval_out[li * 4 + k][i] = m[k]; }
      b2s_compress<4>(h, m, a.one);
      uint32_t* o = a.tree_out[off] + (size_t)i * 8;
#pragma unroll
      for (int k = 0; k < 8; k++) { dg0[i * 8 + k] = h[k]; o[k] = h[k]; }
    }
    __syncthreads();
    uint32_t *cur = dg0, *nxt = dg1;
    for (uint32_t k = lg; k-- > 0;) {
      for (uint32_t i = tid; i < (1u << k); i += nt) {
        uint32_t h[8] = {0, 0, 0, 0, 0, 0, 0, 0}, m[16];
#pragma unroll
        for (int j = 0; j < 16; j++) m[j] = cur[i * 16 + j];
        b2s_compress(h, m, a.one);
        uint32_t* o = a.tree_out[off + (lg - k)] + (size_t)i * 8;
#pragma unroll
        for (int j = 0; j < 8; j++) { nxt[i * 8 + j] = h[j]; o[j] = h[j]; }
      }
      __syncthreads();
      uint32_t* x = cur; cur = nxt; nxt = x;
    }
    if (tid == 0) {
      for (int j = 0; j < 8; j++) a.roots_out[li * 8 + j] = cur[j];
      chan_mix_root_draw(s_digest, cur, s_alpha);
    }
    __syncthreads();
    off += lg + 1;
    // ---- fold_line with the coefficient just drawn: n / 2 <= 512 outputs, at most one per thread
    const QM31 al = load_q(s_alpha);
    QM31 r = q_make(0, 0, 0, 0);
    const bool mine = tid < n / 2;
    if (mine) {
      const uint32_t* itw = a.itw_end - ((size_t)1 << lg);
      const QM31 p = q_make(ev[2 * tid], ev[cap + 2 * tid], ev[2 * cap + 2 * tid], ev[3 * cap + 2 * tid]);
      const QM31 q = q_make(ev[2 * tid + 1], ev[cap + 2 * tid + 1], ev[2 * cap + 2 * tid + 1], ev[3 * cap + 2 * tid + 1]);
      r = q_add(q_add(p, q), q_mul(al, q_mulm(q_sub(p, q), itw[tid])));
    }
    __syncthreads();
    if (mine) { ev[tid] = r.a.a; ev[cap + tid] = r.a.b; ev[2 * cap + tid] = r.b.a; ev[3 * cap + tid] = r.b.b; }
    __syncthreads();
  }
  for (uint32_t i = tid; i < (1u << a.last_log); i += nt)
    for (int k = 0; k < 4; k++) a.last_out[k][i] = ev[k * cap + i];
  if (tid < 8) a.digest[tid] = s_digest[tid];
}
int launch_fri_tail(const FriTailArgs& a, cudaStream_t st) {
  if (a.start_log > FRI_TAIL_LOG || a.start_log <= a.last_log) return -1;
  const size_t smem = (size_t)(4 + 8 + 8) * 4 << a.start_log;   // 80 KB at 2^10
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(fri_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)(4 + 8 + 8) * 4 << FRI_TAIL_LOG));
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  fri_tail_kernel<<<1, 1024, smem, st>>>(a); g_launch_count++;
  return (int)cudaGetLastError();
}

}  // namespace sb
