// DEEP quotient accumulation for sm_100a.
//
// Replaces QuotientOps::accumulate_quotients (stwo-prover 0.1.1 @ 31e8dbc core/backend/{cpu,simd}/quotients.rs with the
// helpers of core/pcs/quotients.rs), reached from compute_fri_quotients inside prover::prove at
// crates/brainfuck_prover/src/brainfuck_air/mod.rs:732.
//
// Per row with domain point (px,py), for each sample batch (point S, entries j):
//   numerator   = sum_j alpha^(j+1) * (c*col_j(row) - (a_j*py + b_j)),   (a,b,c) = (conj(v)-v, v*c - a*S.y, conj(S.y)-S.y)
//   denominator = (Re(S.x)-px)*Im(S.y) - (Re(S.y)-py)*Im(S.x)    in CM31
//   acc = acc * alpha^len(batch) + numerator / denominator
// SimdBackend evaluates on a sub-domain and re-extends; the result is the same function on the same domain, so the
// kernel evaluates directly.  HBM-bound: every column word is read once (128-bit loads, 4 consecutive rows per thread),
// 16 bytes per row are written.  Four consecutive bit-reversed rows are (x,y),(x,-y),(-x,-y),(-x,y): one point per thread.
// The numerator is split as sum_j C_j*col_j - (py*A + B) with A = sum a_j, B = sum b_j folded on the host, and the
// sum of products is carried in 64-bit lanes with one partial reduction (a single IMAD.WIDE) every three terms.
#include <mutex>
#include "kernels.cuh"

namespace sb {

__constant__ Pt c_qgen_pow[31];

__device__ __forceinline__ Pt q_point_at_index(uint32_t idx) {
  Pt r = {1u, 0u};
#pragma unroll 1
  for (int k = 0; k < 31; k++)
    if ((idx >> k) & 1u) r = p_add(r, c_qgen_pow[k]);
  return r;
}

// partial fold of a 64-bit lane: the high word re-enters with weight 2^32 == 2 (mod P) — one IMAD.WIDE; the result is below
// 3 * 2^32, so three more products of two 31-bit factors fit before the next fold (3 * 2^62 + 3 * 2^32 < 2^64)
__device__ __forceinline__ uint64_t fold64(uint64_t v) { return (uint64_t)(uint32_t)(v >> 32) * 2u + (uint32_t)v; }
__device__ __forceinline__ uint32_t red64(uint64_t v) { return m_red_wide(v); }  // full reduction of any 64-bit value

__global__ void __launch_bounds__(128) quotients_kernel(uint32_t log, const uint32_t* const* __restrict__ cols,
                                                        const QuotBatch* __restrict__ batches, uint32_t nb,
                                                        const QuotEntry* __restrict__ entries, uint32_t* o0, uint32_t* o1,
                                                        uint32_t* o2, uint32_t* o3, uint32_t k_off, uint32_t nq) {
  // rows [4*k_off, 4*(k_off+nq)) of the domain; columns and outputs are indexed by the LOCAL quad index k
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < nq; k += gridDim.x * blockDim.x) {
    // base point: half_odds(log-1).at(bitrev(kg, log-2)): index 2^(30-log) + j0 * 2^(32-log)
    const uint32_t kg = k + k_off;
    uint32_t j0 = (log > 2) ? (__brev(kg) >> (34 - log)) : 0;
    uint32_t idx = ((1u << (30 - log)) + (uint32_t)(((uint64_t)j0 << (32 - log)) & 0x7fffffffu)) & 0x7fffffffu;
    Pt bp = q_point_at_index(idx);
    const uint32_t px[4] = {bp.x, bp.x, m_neg(bp.x), m_neg(bp.x)};
    const uint32_t py[4] = {bp.y, m_neg(bp.y), m_neg(bp.y), bp.y};
    QM31 acc[4] = {q_zero(), q_zero(), q_zero(), q_zero()};
    for (uint32_t b = 0; b < nb; b++) {
      const QuotBatch qb = batches[b];
      // denominators for the 4 rows, batch-inverted (CM31)
      CM31 den[4], pre[4];
      CM31 run = {1u, 0u};
#pragma unroll
      for (int r = 0; r < 4; r++) {
        CM31 dx = c_sub(qb.prx, CM31{px[r], 0u}), dy = c_sub(qb.pry, CM31{py[r], 0u});
        den[r] = c_sub(c_mul(dx, qb.piy), c_mul(dy, qb.pix));
        pre[r] = run;
        run = c_mul(run, den[r]);
      }
      CM31 inv = c_inv(run);
      CM31 dinv[4];
#pragma unroll
      for (int r = 3; r >= 0; r--) { dinv[r] = c_mul(inv, pre[r]); inv = c_mul(inv, den[r]); }
      // sum_j C_j * col_j(row): 64-bit lanes
      uint64_t s[4][4];
#pragma unroll
      for (int r = 0; r < 4; r++) { s[r][0] = s[r][1] = s[r][2] = s[r][3] = 0; }
      uint32_t pend = 0;
      for (uint32_t e = qb.first; e < qb.first + qb.count; e++) {
        const QuotEntry en = entries[e];
        uint4 v = __ldg(reinterpret_cast<const uint4*>(cols[en.col]) + k);
        const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int r = 0; r < 4; r++) {
          s[r][0] += (uint64_t)en.c[0] * vv[r]; s[r][1] += (uint64_t)en.c[1] * vv[r];
          s[r][2] += (uint64_t)en.c[2] * vv[r]; s[r][3] += (uint64_t)en.c[3] * vv[r];
        }
        if (++pend == 3u) {
          pend = 0;
#pragma unroll
          for (int r = 0; r < 4; r++) { s[r][0] = fold64(s[r][0]); s[r][1] = fold64(s[r][1]); s[r][2] = fold64(s[r][2]); s[r][3] = fold64(s[r][3]); }
        }
      }
#pragma unroll
      for (int r = 0; r < 4; r++) {
        QM31 num = q_make(red64(s[r][0]), red64(s[r][1]), red64(s[r][2]), red64(s[r][3]));
        QM31 lin = q_add(q_mulm(qb.suma, py[r]), qb.sumb);
        num = q_sub(num, lin);
        acc[r] = q_add(q_mul(acc[r], qb.coeff), q_mulc(num, dinv[r]));
      }
    }
    reinterpret_cast<uint4*>(o0)[k] = make_uint4(acc[0].a.a, acc[1].a.a, acc[2].a.a, acc[3].a.a);
    reinterpret_cast<uint4*>(o1)[k] = make_uint4(acc[0].a.b, acc[1].a.b, acc[2].a.b, acc[3].a.b);
    reinterpret_cast<uint4*>(o2)[k] = make_uint4(acc[0].b.a, acc[1].b.a, acc[2].b.a, acc[3].b.a);
    reinterpret_cast<uint4*>(o3)[k] = make_uint4(acc[0].b.b, acc[1].b.b, acc[2].b.b, acc[3].b.b);
  }
}

// ---------------------------------------------------------------- v2: table-driven points, one inversion per thread
// The per-thread fixed cost of the kernel above (a 31-step point ladder and one CM31 inversion per sample batch) is several
// times the useful work when a size has few columns (the 2^26-row launch carries the four composition columns only).
//  * points: quad kg = (kb << 7) | t has domain index base + brev(kg)*step, and brev splits, so its point is Q[t] + R[kb]
//    with Q (128 entries) and R (one per 128 quads) written by a small pre-kernel: one group addition per thread;
//  * denominators: den_r = c0 -/+ piy*x +/- pix*y with c0 = prx*piy - pry*pix folded on the host; the four rows of a quad
//    and up to QV2_NB batches share ONE base-field inversion (Montgomery's trick on the norms re^2 + im^2);
//  * the first batch skips the acc*coeff product, and py*suma is formed once per batch.
constexpr int QV2_NB = 2;
constexpr uint32_t QV2_TLOG = 7;

__global__ void quot_points_kernel(uint32_t log, uint32_t kb0, uint32_t nkb, Pt* __restrict__ Q, Pt* __restrict__ Rt) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (1u << QV2_TLOG)) {
    // base + brev7(t) << 23   (the top 7 bits of brev(kg) times the step 2^(32-log))
    uint32_t idx = ((1u << (30 - log)) + ((__brev(i) >> (32 - QV2_TLOG)) << (31 - QV2_TLOG - 1))) & 0x7fffffffu;
    Q[i] = q_point_at_index(idx);
  } else if (i - (1u << QV2_TLOG) < nkb) {
    const uint32_t kb = kb0 + (i - (1u << QV2_TLOG));
    const uint32_t hb = log - 2 - QV2_TLOG;                       // bits of kb
    uint32_t j = hb ? (__brev(kb) >> (32 - hb)) : 0;
    uint32_t idx = (uint32_t)(((uint64_t)j << (32 - log)) & 0x7fffffffu);
    Rt[i - (1u << QV2_TLOG)] = q_point_at_index(idx);
  }
}

// (128, 8): 64 registers with ~300 B of spills measured 5 % faster than the unconstrained 127-register build
__global__ void __launch_bounds__(128, 8) quotients_kernel2(const uint32_t* const* __restrict__ cols, const QuotBatch* __restrict__ batches,
                                                         uint32_t nb, const QuotEntry* __restrict__ entries, uint32_t* o0, uint32_t* o1,
                                                         uint32_t* o2, uint32_t* o3, uint32_t k_off, uint32_t nq,
                                                         const Pt* __restrict__ Q, const Pt* __restrict__ Rt, uint32_t kb0) {
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < nq; k += gridDim.x * blockDim.x) {
    const uint32_t kg = k + k_off;
    const Pt bp = p_add(Q[kg & ((1u << QV2_TLOG) - 1u)], Rt[(kg >> QV2_TLOG) - kb0]);
    const uint32_t x = bp.x, y = bp.y;                    // rows: (x,y) (x,-y) (-x,-y) (-x,y)
    QM31 acc[4];
    for (uint32_t b0 = 0; b0 < nb; b0 += QV2_NB) {
      const uint32_t nbc = nb - b0 < (uint32_t)QV2_NB ? nb - b0 : (uint32_t)QV2_NB;
      CM31 den[QV2_NB][4];
      uint32_t pre[QV2_NB][4];
      uint32_t run = 1u;
#pragma unroll
      for (int bb = 0; bb < QV2_NB; bb++) {
        if ((uint32_t)bb < nbc) {
          const QuotBatch& qb = batches[b0 + bb];
          const CM31 u = c_mulm(qb.piy, x), v = c_mulm(qb.pix, y);
          const CM31 lo = c_sub(qb.c0, u), hi = c_add(qb.c0, u);
          den[bb][0] = c_add(lo, v); den[bb][1] = c_sub(lo, v); den[bb][2] = c_sub(hi, v); den[bb][3] = c_add(hi, v);
#pragma unroll
          for (int r = 0; r < 4; r++) {
            pre[bb][r] = run;
            run = m_mul(run, m_add(m_sqr(den[bb][r].a), m_sqr(den[bb][r].b)));
          }
        }
      }
      uint32_t inv = m_inv(run);
      CM31 dinv[QV2_NB][4];
#pragma unroll
      for (int bb = QV2_NB - 1; bb >= 0; bb--) {
        if ((uint32_t)bb < nbc) {
#pragma unroll
          for (int r = 3; r >= 0; r--) {
            const uint32_t ninv = m_mul(inv, pre[bb][r]);                     // 1 / norm
            inv = m_mul(inv, m_add(m_sqr(den[bb][r].a), m_sqr(den[bb][r].b)));
            dinv[bb][r] = CM31{m_mul(den[bb][r].a, ninv), m_mul(m_neg(den[bb][r].b), ninv)};   // conj / norm
          }
        }
      }
#pragma unroll
      for (int bb = 0; bb < QV2_NB; bb++) {
        if ((uint32_t)bb < nbc) {
          const QuotBatch& qb = batches[b0 + bb];
          uint64_t s[4][4];
#pragma unroll
          for (int r = 0; r < 4; r++) { s[r][0] = s[r][1] = s[r][2] = s[r][3] = 0; }
          const uint32_t e0 = qb.first, e1 = qb.first + qb.count;
          uint32_t pend = 0;
          for (uint32_t e = e0; e < e1; e++) {
            const QuotEntry en = entries[e];
            uint4 v = __ldg(reinterpret_cast<const uint4*>(cols[en.col]) + k);
            const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int r = 0; r < 4; r++) {
              s[r][0] += (uint64_t)en.c[0] * vv[r]; s[r][1] += (uint64_t)en.c[1] * vv[r];
              s[r][2] += (uint64_t)en.c[2] * vv[r]; s[r][3] += (uint64_t)en.c[3] * vv[r];
            }
            if (++pend == 3u) {
              pend = 0;
#pragma unroll
              for (int r = 0; r < 4; r++) { s[r][0] = fold64(s[r][0]); s[r][1] = fold64(s[r][1]); s[r][2] = fold64(s[r][2]); s[r][3] = fold64(s[r][3]); }
            }
          }
          const QM31 ya = q_mulm(qb.suma, y);
          const QM31 linp = q_add(qb.sumb, ya), linm = q_sub(qb.sumb, ya);   // py = +y rows 0,3 ; -y rows 1,2
#pragma unroll
          for (int r = 0; r < 4; r++) {
            QM31 num = q_make(red64(s[r][0]), red64(s[r][1]), red64(s[r][2]), red64(s[r][3]));
            num = q_sub(num, (r == 0 || r == 3) ? linp : linm);
            const QM31 term = q_mulc(num, dinv[bb][r]);
            acc[r] = (b0 + bb == 0) ? term : q_add(q_mul(acc[r], qb.coeff), term);
          }
        }
      }
    }
    reinterpret_cast<uint4*>(o0)[k] = make_uint4(acc[0].a.a, acc[1].a.a, acc[2].a.a, acc[3].a.a);
    reinterpret_cast<uint4*>(o1)[k] = make_uint4(acc[0].a.b, acc[1].a.b, acc[2].a.b, acc[3].a.b);
    reinterpret_cast<uint4*>(o2)[k] = make_uint4(acc[0].b.a, acc[1].b.a, acc[2].b.a, acc[3].b.a);
    reinterpret_cast<uint4*>(o3)[k] = make_uint4(acc[0].b.b, acc[1].b.b, acc[2].b.b, acc[3].b.b);
  }
}

size_t quotients_scratch_words(uint32_t log, uint64_t row_off, uint64_t nrows);
// ---------------------------------------------------------------- IsFirst on the extended domain, in closed form
// The inverse transform of e_0 (1 at row 0) only ever meets twiddle index 0, so coefficient i of the IsFirst polynomial of log
// size L is 2^-L * prod_{l in bits(i)} t_l (is_first_coeffs_kernel, ops.cu) — a rank-one tensor — and the polynomial is
//   f(p) = 2^-L * (1 + t_0 y)(1 + t_1 x) * prod_{l >= 2} (1 + t_l pi^(l-1)(x)),   pi(x) = 2x^2 - 1.
// Its low-degree extension is therefore a row-local function: no transform, no HBM pass beside the store, and on several GPUs no
// column->row exchange for the preprocessed tree — every rank writes its own row range.  The four rows of a quad
// (x,y) (x,-y) (-x,-y) (-x,y) share the factors l >= 2: about 3 L multiplications per quad.  Points as in quotients_kernel2.
__global__ void __launch_bounds__(128) is_first_lde_kernel(uint32_t* __restrict__ out, uint32_t L, uint32_t dom_log, const uint32_t* __restrict__ itw_end,
                                                           uint32_t ninv, uint32_t k_off, uint32_t nq, const Pt* __restrict__ Q,
                                                           const Pt* __restrict__ Rt, uint32_t kb0) {
  __shared__ uint32_t t[32];
  if (threadIdx.x < L) {
    const uint32_t l = threadIdx.x;
    const uint32_t* l1 = itw_end - ((size_t)1 << (L - 1));
    t[l] = l == 0 ? l1[1] : (l == 1 ? l1[0] : *(itw_end - ((size_t)1 << (L - l))));
  }
  __syncthreads();
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < nq; k += gridDim.x * blockDim.x) {
    const uint32_t kg = k + k_off;
    Pt bp;
    if (Q) bp = p_add(Q[kg & ((1u << QV2_TLOG) - 1u)], Rt[(kg >> QV2_TLOG) - kb0]);
    else {
      const uint32_t j0 = (dom_log > 2) ? (__brev(kg) >> (34 - dom_log)) : 0;
      bp = q_point_at_index(((1u << (30 - dom_log)) + (uint32_t)(((uint64_t)j0 << (32 - dom_log)) & 0x7fffffffu)) & 0x7fffffffu);
    }
    uint32_t c = ninv, z = bp.x;
    for (uint32_t l = 2; l < L; l++) {
      z = m_sub(m_mul(m_add(z, z), z), 1u);          // pi^(l-1)(x)
      c = m_mul(c, m_add(1u, m_mul(t[l], z)));
    }
    const uint32_t ux = m_mul(t[1], bp.x), uy = m_mul(t[0], bp.y);
    const uint32_t ap = m_mul(c, m_add(1u, ux)), am = m_mul(c, m_sub(1u, ux));   // x, -x
    const uint32_t bpv = m_add(1u, uy), bm = m_sub(1u, uy);                      // y, -y
    reinterpret_cast<uint4*>(out)[k] = make_uint4(m_mul(ap, bpv), m_mul(ap, bm), m_mul(am, bm), m_mul(am, bpv));
  }
}
// rows [row_off, row_off + nrows) of IsFirst(log L) on CanonicCoset(dom_log).circle_domain(), bit-reversed; d_scratch as in
// launch_accumulate_quotients (quotients_scratch_words(dom_log, row_off, nrows) words, may be NULL for small domains)
int launch_is_first_lde(uint32_t* out, uint32_t L, uint32_t dom_log, uint64_t row_off, uint64_t nrows, const uint32_t* itw_plain_end,
                        cudaStream_t st, uint32_t* d_scratch) {
  if (L < 3 || L > 31 || dom_log < L || dom_log > 30 || (row_off & 3) || (nrows & 3) || row_off + nrows > ((uint64_t)1 << dom_log)) return -1;
  {
    static std::mutex mu;
    static bool init[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    if (dev >= 0 && dev < 64 && !init[dev]) {
      Pt g[31];
      g[0] = {GEN_X, GEN_Y};
      for (int k = 1; k < 31; k++) g[k] = p_dbl(g[k - 1]);
      cudaError_t e = cudaMemcpyToSymbol(c_qgen_pow, g, sizeof(g));
      if (e != cudaSuccess) return (int)e;
      init[dev] = true;
    }
  }
  const uint32_t nq = (uint32_t)(nrows >> 2), k0 = (uint32_t)(row_off >> 2);
  if (!nq) return 0;
  uint32_t blocks = (nq + 127) / 128;
  if (blocks > 148u * 16u) blocks = 148u * 16u;
  const uint32_t ninv = m_inv(m_pow(2, L));
  if (d_scratch && quotients_scratch_words(dom_log, row_off, nrows)) {
    const uint32_t kb0 = k0 >> QV2_TLOG, nkb = ((k0 + nq + 127) >> QV2_TLOG) - kb0;
    Pt* Q = reinterpret_cast<Pt*>(d_scratch);
    Pt* Rt = Q + 128;
    quot_points_kernel<<<(128 + nkb + 127) / 128, 128, 0, st>>>(dom_log, kb0, nkb, Q, Rt); g_launch_count++;
    is_first_lde_kernel<<<blocks, 128, 0, st>>>(out, L, dom_log, itw_plain_end, ninv, k0, nq, Q, Rt, kb0); g_launch_count++;
  } else {
    is_first_lde_kernel<<<blocks, 128, 0, st>>>(out, L, dom_log, itw_plain_end, ninv, k0, nq, nullptr, nullptr, 0); g_launch_count++;
  }
  return (int)cudaGetLastError();
}

// words of device scratch launch_accumulate_quotients needs for its point tables (0: the small-domain kernel is used)
size_t quotients_scratch_words(uint32_t log, uint64_t row_off, uint64_t nrows) {
  if (log < 2 + QV2_TLOG + 1 || nrows == 0) return 0;
  uint64_t k0 = row_off >> 2, k1 = (row_off + nrows) >> 2;
  uint64_t nkb = ((k1 + 127) >> QV2_TLOG) - (k0 >> QV2_TLOG);
  return (size_t)(128 + nkb) * (sizeof(Pt) / 4);
}

int launch_accumulate_quotients(uint32_t log, uint64_t row_off, uint64_t nrows, const uint32_t* const* d_cols,
                                const QuotBatch* d_batches, uint32_t nb, const QuotEntry* d_entries, uint32_t* const out[4],
                                cudaStream_t st, uint32_t* d_scratch) {
  // __constant__ memory is per device: initialise it once for every device this process uses (one flag per ordinal;
  // the mutex covers contexts of different devices created from different host threads)
  {
    static std::mutex mu;
    static bool init[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    if (dev >= 0 && dev < 64 && !init[dev]) {
      Pt g[31];
      g[0] = {GEN_X, GEN_Y};
      for (int k = 1; k < 31; k++) g[k] = p_dbl(g[k - 1]);
      cudaError_t e = cudaMemcpyToSymbol(c_qgen_pow, g, sizeof(g));
      if (e != cudaSuccess) return (int)e;
      init[dev] = true;
    }
  }
  if (log < 2 || log > 30 || (row_off & 3) || (nrows & 3) || row_off + nrows > ((uint64_t)1 << log)) return -1;
  uint32_t nq = (uint32_t)(nrows >> 2);
  uint32_t blocks = (nq + 127) / 128;
  if (blocks > 148u * 16u) blocks = 148u * 16u;
  if (d_scratch && quotients_scratch_words(log, row_off, nrows)) {
    const uint32_t k0 = (uint32_t)(row_off >> 2), kb0 = k0 >> QV2_TLOG;
    const uint32_t nkb = ((k0 + nq + 127) >> QV2_TLOG) - kb0;
    Pt* Q = reinterpret_cast<Pt*>(d_scratch);
    Pt* Rt = Q + 128;
    quot_points_kernel<<<(128 + nkb + 127) / 128, 128, 0, st>>>(log, kb0, nkb, Q, Rt); g_launch_count++;
    quotients_kernel2<<<blocks, 128, 0, st>>>(d_cols, d_batches, nb, d_entries, out[0], out[1], out[2], out[3], k0, nq, Q, Rt, kb0); g_launch_count++;
    return (int)cudaGetLastError();
  }
  quotients_kernel<<<blocks, 128, 0, st>>>(log, d_cols, d_batches, nb, d_entries, out[0], out[1], out[2], out[3],
                                           (uint32_t)(row_off >> 2), nq); g_launch_count++;
  return (int)cudaGetLastError();
}

}  // namespace sb
