// Row-local Backend ops for sm_100a: ColumnOps::bit_reverse_column, FieldOps::batch_inverse, FriOps::{fold_line,
// fold_circle_into_line}, AccumulationOps::accumulate, gen_is_first, the LogUp coset-order prefix sum and
// PolyOps::eval_at_point.  Definitions: stwo-prover 0.1.1 @ 31e8dbc core/backend/{cpu,simd}/{fri,accumulation,
// bit_reverse,prefix_sum,circle}.rs; all are reached from prover::prove at
// crates/brainfuck_prover/src/brainfuck_air/mod.rs:732 or from interaction_trace_evaluation (e.g.
// crates/brainfuck_prover/src/components/processor/table.rs:456-533).
// Every kernel here is HBM-bound: one coalesced read and one coalesced write per element, grid = multiple of 148 SMs.
#include <algorithm>
#include "kernels.cuh"

namespace sb {

static inline unsigned grid_for(size_t n, unsigned threads, unsigned per_thread = 1) {
  size_t b = (n + (size_t)threads * per_thread - 1) / ((size_t)threads * per_thread);
  size_t cap = 148u * 16u;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ---------------------------------------------------------------- bit reverse (in place), 32x32 tiles through smem
// Index = hi(5) | mid | lo(5): swap with rev(lo) | rev(mid) | rev(hi).  A CTA owns the pair of mid values (mid, rev(mid))
// and moves both 32x32 tiles through shared memory so that global reads and writes are 128-byte rows.
__global__ void bit_reverse_tiled_kernel(uint32_t* __restrict__ v, uint32_t log) {
  __shared__ uint32_t ta[32][33], tb[32][33];
  const uint32_t mlog = log - 10;
  const uint32_t nmid = 1u << mlog;
  const uint32_t tx = threadIdx.x & 31u, ty = threadIdx.x >> 5;  // 32 x 8
  for (uint32_t mid = blockIdx.x; mid < nmid; mid += gridDim.x) {
    uint32_t rmid = mlog ? (__brev(mid) >> (32 - mlog)) : 0;
    if (rmid < mid) continue;
    // tile(mid)[hi][lo] = v[hi<<(log-5) | mid<<5 | lo]
    for (uint32_t r = ty; r < 32; r += 8) {
      ta[r][tx] = v[((size_t)r << (log - 5)) | (mid << 5) | tx];
      if (rmid != mid) tb[r][tx] = v[((size_t)r << (log - 5)) | (rmid << 5) | tx];
    }
    __syncthreads();
    // element (hi,lo) of tile(mid) goes to index rev5(lo)<<(log-5) | rmid<<5 | rev5(hi): row rev5(lo) of tile(rmid), col rev5(hi)
    for (uint32_t r = ty; r < 32; r += 8) {
      uint32_t rr = __brev(r) >> 27, rc = __brev(tx) >> 27;
      // write row r of tile(rmid): column tx takes source (hi = rev5(tx), lo = rev5(r)) of tile(mid)
      v[((size_t)r << (log - 5)) | (rmid << 5) | tx] = ta[rc][rr];
      if (rmid != mid) v[((size_t)r << (log - 5)) | (mid << 5) | tx] = tb[rc][rr];
    }
    __syncthreads();
  }
}
__global__ void bit_reverse_small_kernel(uint32_t* __restrict__ v, uint32_t log) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (1u << log)) return;
  uint32_t j = bitrev32(i, log);
  if (i < j) { uint32_t a = v[i], b = v[j]; v[i] = b; v[j] = a; }
}
int launch_bit_reverse(uint32_t* v, uint32_t log, cudaStream_t st) {
  if (log >= 10) {
    uint32_t nmid = 1u << (log - 10);
    bit_reverse_tiled_kernel<<<nmid < 148u * 8 ? nmid : 148u * 8, 256, 0, st>>>(v, log); g_launch_count++;
  } else {
    uint32_t n = 1u << log;
    bit_reverse_small_kernel<<<(n + 255) / 256, 256, 0, st>>>(v, log); g_launch_count++;
  }
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------- batch inverse (Fermat per lane; ALU is free next to HBM)
__global__ void inv_m31_kernel(const uint32_t* __restrict__ s, uint32_t* __restrict__ d, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = m_inv(s[i]);
}
struct Ptr4 { uint32_t* p[4]; };
struct CPtr4 { const uint32_t* p[4]; };
__global__ void inv_qm31_kernel(CPtr4 s, Ptr4 d, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    QM31 r = q_inv(q_make(s.p[0][i], s.p[1][i], s.p[2][i], s.p[3][i]));
    d.p[0][i] = r.a.a; d.p[1][i] = r.a.b; d.p[2][i] = r.b.a; d.p[3][i] = r.b.b;
  }
}
int launch_batch_inverse_m31(const uint32_t* src, uint32_t* dst, size_t n, cudaStream_t st) {
  inv_m31_kernel<<<grid_for(n, 256), 256, 0, st>>>(src, dst, n); g_launch_count++;
  return (int)cudaGetLastError();
}
int launch_batch_inverse_qm31(const uint32_t* const src[4], uint32_t* const dst[4], size_t n, cudaStream_t st) {
  CPtr4 s{{src[0], src[1], src[2], src[3]}};
  Ptr4 d{{dst[0], dst[1], dst[2], dst[3]}};
  inv_qm31_kernel<<<grid_for(n, 256), 256, 0, st>>>(s, d, n); g_launch_count++;
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------- FRI folds
// fold_line: out[i] = (e[2i]+e[2i+1]) + alpha * (e[2i]-e[2i+1]) / x_i ; 1/x_i = itw level of coset log `log`, index i.
__global__ void fold_line_kernel(CPtr4 s, Ptr4 d, uint32_t log, QM31 alpha, const uint32_t* __restrict__ itw_end) {
  const size_t half = (size_t)1 << (log - 1);
  const uint32_t* itw = itw_end - ((size_t)1 << log);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += (size_t)gridDim.x * blockDim.x) {
    uint2 c0 = reinterpret_cast<const uint2*>(s.p[0])[i], c1 = reinterpret_cast<const uint2*>(s.p[1])[i];
    uint2 c2 = reinterpret_cast<const uint2*>(s.p[2])[i], c3 = reinterpret_cast<const uint2*>(s.p[3])[i];
    QM31 a = q_make(c0.x, c1.x, c2.x, c3.x), b = q_make(c0.y, c1.y, c2.y, c3.y);
    QM31 f0 = q_add(a, b), f1 = q_mulm(q_sub(a, b), __ldg(itw + i));
    QM31 r = q_add(f0, q_mul(alpha, f1));
    d.p[0][i] = r.a.a; d.p[1][i] = r.a.b; d.p[2][i] = r.b.a; d.p[3][i] = r.b.b;
  }
}
int launch_fold_line(const uint32_t* const src[4], uint32_t log, QM31 alpha, uint32_t* const dst[4], const uint32_t* itw_end,
                     cudaStream_t st) {
  if (log < 1) return -1;
  CPtr4 s{{src[0], src[1], src[2], src[3]}};
  Ptr4 d{{dst[0], dst[1], dst[2], dst[3]}};
  fold_line_kernel<<<grid_for((size_t)1 << (log - 1), 256), 256, 0, st>>>(s, d, log, alpha, itw_end); g_launch_count++;
  return (int)cudaGetLastError();
}

// fold_circle_into_line: dst[i] = dst[i]*alpha^2 + (f0 + alpha*f1), (f0,f1) = ibutterfly(src[2i], src[2i+1], 1/p_i.y);
// 1/p_i.y is the inverse circle-layer twiddle of the size-`log` FFT: from itw line layer 1, [x,y] -> [y,-y,-x,x].
__global__ void fold_circle_kernel(CPtr4 s, Ptr4 d, uint32_t log, QM31 alpha, QM31 alpha_sq, const uint32_t* __restrict__ itw_end) {
  const size_t half = (size_t)1 << (log - 1);
  const uint32_t* l1 = itw_end - ((size_t)1 << (log - 1));
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += (size_t)gridDim.x * blockDim.x) {
    size_t pair = (i >> 2) * 2;
    uint32_t x = __ldg(l1 + pair), y = __ldg(l1 + pair + 1);
    uint32_t sel = (uint32_t)i & 3u;
    uint32_t t = sel < 2 ? y : x;
    if (sel == 1 || sel == 2) t = P - t;
    uint2 c0 = reinterpret_cast<const uint2*>(s.p[0])[i], c1 = reinterpret_cast<const uint2*>(s.p[1])[i];
    uint2 c2 = reinterpret_cast<const uint2*>(s.p[2])[i], c3 = reinterpret_cast<const uint2*>(s.p[3])[i];
    QM31 a = q_make(c0.x, c1.x, c2.x, c3.x), b = q_make(c0.y, c1.y, c2.y, c3.y);
    QM31 f0 = q_add(a, b), f1 = q_mulm(q_sub(a, b), t);
    QM31 acc = q_make(d.p[0][i], d.p[1][i], d.p[2][i], d.p[3][i]);
    QM31 r = q_add(q_mul(acc, alpha_sq), q_add(f0, q_mul(alpha, f1)));
    d.p[0][i] = r.a.a; d.p[1][i] = r.a.b; d.p[2][i] = r.b.a; d.p[3][i] = r.b.b;
  }
}
int launch_fold_circle_into_line(const uint32_t* const src[4], uint32_t log, QM31 alpha, uint32_t* const dst[4],
                                 const uint32_t* itw_end, cudaStream_t st) {
  if (log < 3) return -1;
  CPtr4 s{{src[0], src[1], src[2], src[3]}};
  Ptr4 d{{dst[0], dst[1], dst[2], dst[3]}};
  fold_circle_kernel<<<grid_for((size_t)1 << (log - 1), 256), 256, 0, st>>>(s, d, log, alpha, q_mul(alpha, alpha), itw_end); g_launch_count++;
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------- accumulate / fill / is_first
__global__ void accumulate_kernel(Ptr4 d, CPtr4 s, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      uint4 a = reinterpret_cast<uint4*>(d.p[k])[i], b = reinterpret_cast<const uint4*>(s.p[k])[i];
      a.x = m_add(a.x, b.x); a.y = m_add(a.y, b.y); a.z = m_add(a.z, b.z); a.w = m_add(a.w, b.w);
      reinterpret_cast<uint4*>(d.p[k])[i] = a;
    }
  }
}
int launch_accumulate(uint32_t* const dst[4], const uint32_t* const src[4], size_t n, cudaStream_t st) {
  if (n % 4) return -1;
  Ptr4 d{{dst[0], dst[1], dst[2], dst[3]}};
  CPtr4 s{{src[0], src[1], src[2], src[3]}};
  accumulate_kernel<<<grid_for(n / 4, 256), 256, 0, st>>>(d, s, n / 4); g_launch_count++;
  return (int)cudaGetLastError();
}
__global__ void fill_kernel(uint32_t* __restrict__ v, size_t n, uint32_t value, uint32_t first) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    v[i] = i == 0 ? first : value;
}
int launch_fill(uint32_t* v, size_t n, uint32_t value, cudaStream_t st) {
  fill_kernel<<<grid_for(n, 256), 256, 0, st>>>(v, n, value, value); g_launch_count++;
  return (int)cudaGetLastError();
}
// Coefficients of the IsFirst column (1 at row 0, else 0) in closed form.  The inverse transform of e_0 only ever meets
// twiddle index 0 (entry i of a layer is non-zero only below 2^(l+1), i.e. in group 0), so coefficient i is the product of
// t_l over the set bits l of i, times 1/2^log, with t_0 the first circle twiddle (y of the first layer-1 pair), t_1 the
// first layer-1 twiddle and t_l the first twiddle of line layer l: a write-only kernel instead of a fill and 1-2 FFT passes.
__global__ void __launch_bounds__(256) is_first_coeffs_kernel(uint32_t* __restrict__ out, uint32_t log, const uint32_t* __restrict__ itw_end,
                                                              uint32_t ninv) {
  __shared__ uint32_t t[32];
  __shared__ uint32_t hp;
  if (threadIdx.x < log) {
    const uint32_t l = threadIdx.x;
    const uint32_t* l1 = itw_end - ((size_t)1 << (log - 1));
    t[l] = l == 0 ? l1[1] : (l == 1 ? l1[0] : *(itw_end - ((size_t)1 << (log - l))));
  }
  __syncthreads();
  // the 256 coefficients of a block share their high bits: one product for those (thread 0), one table entry per thread for
  // the low eight bits, one multiplication per coefficient
  uint32_t low = 1u;
  for (uint32_t l = 0; l < 8 && l < log; l++) if ((threadIdx.x >> l) & 1u) low = m_mul(low, t[l]);
  const size_t n = (size_t)1 << log;
  for (size_t base = (size_t)blockIdx.x * 256; base < n; base += (size_t)gridDim.x * 256) {  // capped grid
    if (threadIdx.x == 0) {
      uint32_t v = ninv;
      for (uint32_t l = 8; l < log; l++) if ((base >> l) & 1u) v = m_mul(v, t[l]);
      hp = v;
    }
    __syncthreads();
    if (base + threadIdx.x < n) out[base + threadIdx.x] = m_mul(hp, low);
    __syncthreads();
  }
}
int launch_is_first_coeffs(uint32_t* out, uint32_t log, const uint32_t* itw_plain_end, cudaStream_t st) {
  if (log < 3 || log > 31) return -1;
  uint32_t ninv = m_inv(m_pow(2, log));
  is_first_coeffs_kernel<<<grid_for((size_t)1 << log, 256), 256, 0, st>>>(out, log, itw_plain_end, ninv); g_launch_count++;
  return (int)cudaGetLastError();
}
int launch_gen_is_first(uint32_t* v, uint32_t log, cudaStream_t st) {
  fill_kernel<<<grid_for((size_t)1 << log, 256), 256, 0, st>>>(v, (size_t)1 << log, 0u, 1u); g_launch_count++;
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------- LogUp prefix sum in trace-coset order
// Storage index s <-> coset index i:  s = 2*brev(m) -> i = 2m ;  s = 2*brev(m)+1 -> i = 2^log - 1 - 2m   (m < 2^(log-1)).
__device__ __forceinline__ uint32_t coset_to_storage(uint32_t i, uint32_t log) {
  uint32_t m = (i & 1u) ? (((1u << log) - 1u - i) >> 1) : (i >> 1);
  uint32_t r = (log > 1) ? (__brev(m) >> (33 - log)) : 0;
  return 2u * r + (i & 1u);
}
constexpr uint32_t SCAN_CHUNK_LOG = 11;  // 256 threads x 8 elements

__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long x, unsigned long long* total) {
  __shared__ unsigned long long wsum[8];
  const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  unsigned long long inc = x;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned long long y = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (uint32_t)o) inc += y;
  }
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  unsigned long long base = 0, tot = 0;
#pragma unroll
  for (uint32_t k = 0; k < 8; k++) { if (k < w) base += wsum[k]; tot += wsum[k]; }
  __syncthreads();
  *total = tot;
  return base + inc - x;
}

__global__ void __launch_bounds__(256) scan_gather_kernel(const uint32_t* __restrict__ v, uint32_t log, uint32_t* __restrict__ nat,
                                                          unsigned long long* __restrict__ chunk_sums) {
  const uint32_t n = 1u << log;
  const uint32_t base = blockIdx.x << SCAN_CHUNK_LOG;
  unsigned long long s = 0;
#pragma unroll
  for (uint32_t k = 0; k < 8; k++) {
    uint32_t i = base + k * 256 + threadIdx.x;
    if (i < n) { uint32_t x = v[coset_to_storage(i, log)]; nat[i] = x; s += x; }
  }
  unsigned long long tot;
  block_exclusive_scan(s, &tot);
  if (threadIdx.x == 0) chunk_sums[blockIdx.x] = tot % P;
}
__global__ void __launch_bounds__(256) scan_chunks_kernel(unsigned long long* __restrict__ chunk_sums, uint32_t nchunks) {
  // single CTA: exclusive scan (mod P) of the chunk sums, sequential over 256-wide slabs
  unsigned long long carry = 0;
  for (uint32_t b = 0; b < nchunks; b += 256) {
    uint32_t i = b + threadIdx.x;
    unsigned long long x = i < nchunks ? chunk_sums[i] : 0, tot;
    unsigned long long ex = block_exclusive_scan(x, &tot);
    if (i < nchunks) chunk_sums[i] = (carry + ex) % P;
    carry = (carry + tot) % P;
    __syncthreads();
  }
}
__global__ void __launch_bounds__(256) scan_scatter_kernel(uint32_t* __restrict__ v, uint32_t log, const uint32_t* __restrict__ nat,
                                                           const unsigned long long* __restrict__ chunk_sums) {
  const uint32_t n = 1u << log;
  const uint32_t base = (blockIdx.x << SCAN_CHUNK_LOG) + threadIdx.x * 8;  // 8 consecutive coset rows per thread
  unsigned long long x[8], s = 0;
#pragma unroll
  for (uint32_t k = 0; k < 8; k++) { x[k] = (base + k < n) ? nat[base + k] : 0; s += x[k]; }
  unsigned long long tot;
  unsigned long long run = block_exclusive_scan(s, &tot) + chunk_sums[blockIdx.x];
#pragma unroll
  for (uint32_t k = 0; k < 8; k++) {
    run += x[k];
    if (base + k < n) v[coset_to_storage(base + k, log)] = (uint32_t)(run % P);
  }
}
// scratch: 2^log words for the natural-order copy + (2^log >> 11) + 1 u64 chunk sums (8-byte aligned, placed first).
int launch_prefix_sum_bitrev(uint32_t* v, uint32_t log, uint32_t* scratch, cudaStream_t st) {
  uint32_t n = 1u << log;
  uint32_t nchunks = (n + (1u << SCAN_CHUNK_LOG) - 1) >> SCAN_CHUNK_LOG;
  unsigned long long* sums = reinterpret_cast<unsigned long long*>(scratch);
  uint32_t* nat = scratch + 2 * (size_t)((nchunks + 1) & ~1u) + 2;
  scan_gather_kernel<<<nchunks, 256, 0, st>>>(v, log, nat, sums); g_launch_count++;
  scan_chunks_kernel<<<1, 256, 0, st>>>(sums, nchunks); g_launch_count++;
  scan_scatter_kernel<<<nchunks, 256, 0, st>>>(v, log, nat, sums); g_launch_count++;
  return (int)cudaGetLastError();
}

// The four coordinate columns of one LogUp batch in one launch set (grid.y = coordinate); scratch holds four regions of
// `words` words laid out as above.
struct Scan4 { uint32_t* v[4]; uint32_t* nat[4]; unsigned long long* sums[4]; };
__global__ void __launch_bounds__(256) scan_gather4_kernel(Scan4 a, uint32_t log) {
  const uint32_t n = 1u << log;
  const uint32_t base = blockIdx.x << SCAN_CHUNK_LOG;
  const uint32_t* __restrict__ v = a.v[blockIdx.y];
  uint32_t* __restrict__ nat = a.nat[blockIdx.y];
  unsigned long long s = 0;
#pragma unroll
  for (uint32_t k = 0; k < 8; k++) {
    uint32_t i = base + k * 256 + threadIdx.x;
    if (i < n) { uint32_t x = v[coset_to_storage(i, log)]; nat[i] = x; s += x; }
  }
  unsigned long long tot;
  block_exclusive_scan(s, &tot);
  if (threadIdx.x == 0) a.sums[blockIdx.y][blockIdx.x] = tot % P;
}
__global__ void __launch_bounds__(256) scan_chunks4_kernel(Scan4 a, uint32_t nchunks) {
  unsigned long long* __restrict__ chunk_sums = a.sums[blockIdx.x];
  unsigned long long carry = 0;
  for (uint32_t b = 0; b < nchunks; b += 256) {
    uint32_t i = b + threadIdx.x;
    unsigned long long x = i < nchunks ? chunk_sums[i] : 0, tot;
    unsigned long long ex = block_exclusive_scan(x, &tot);
    if (i < nchunks) chunk_sums[i] = (carry + ex) % P;
    carry = (carry + tot) % P;
    __syncthreads();
  }
}
__global__ void __launch_bounds__(256) scan_scatter4_kernel(Scan4 a, uint32_t log) {
  const uint32_t n = 1u << log;
  const uint32_t base = (blockIdx.x << SCAN_CHUNK_LOG) + threadIdx.x * 8;
  uint32_t* __restrict__ v = a.v[blockIdx.y];
  const uint32_t* __restrict__ nat = a.nat[blockIdx.y];
  unsigned long long x[8], s = 0;
#pragma unroll
  for (uint32_t k = 0; k < 8; k++) { x[k] = (base + k < n) ? nat[base + k] : 0; s += x[k]; }
  unsigned long long tot;
  unsigned long long run = block_exclusive_scan(s, &tot) + a.sums[blockIdx.y][blockIdx.x];
#pragma unroll
  for (uint32_t k = 0; k < 8; k++) {
    run += x[k];
    if (base + k < n) v[coset_to_storage(base + k, log)] = (uint32_t)(run % P);
  }
}
int launch_prefix_sum_bitrev4(uint32_t* const v[4], uint32_t log, uint32_t* scratch, size_t words, cudaStream_t st) {
  uint32_t n = 1u << log;
  uint32_t nchunks = (n + (1u << SCAN_CHUNK_LOG) - 1) >> SCAN_CHUNK_LOG;
  Scan4 a;
  for (int k = 0; k < 4; k++) {
    uint32_t* s = scratch + (size_t)k * words;
    a.v[k] = v[k];
    a.sums[k] = reinterpret_cast<unsigned long long*>(s);
    a.nat[k] = s + 2 * (size_t)((nchunks + 1) & ~1u) + 2;
  }
  scan_gather4_kernel<<<dim3(nchunks, 4), 256, 0, st>>>(a, log); g_launch_count++;
  scan_chunks4_kernel<<<4, 256, 0, st>>>(a, nchunks); g_launch_count++;
  scan_scatter4_kernel<<<dim3(nchunks, 4), 256, 0, st>>>(a, log); g_launch_count++;
  return (int)cudaGetLastError();
}

// ---- tiled variant (log >= 12): every global access is a full 256-byte run.
// Coset position i = 2m is stored at 2*brev(m), i = 2m+1 at 2*(~brev(m))+1 (m over log-1 bits).  Split m = (a | mid | b)
// with 5-bit a (top) and b (bottom): brev(m) = (brev5(b) | brev(mid) | brev5(a)), so for fixed (b, mid) the 32 values of a
// together with both parities are 64 consecutive storage words, and for fixed (a, mid) the 32 values of b are 64
// consecutive coset positions — once the odd positions are taken from the mirrored tile (~mid), which is why a CTA owns
// the tile pair {mid, ~mid}.  Pass A un-permutes into `nat` and records the sum of every 64-position run; two small
// kernels turn run sums into exclusive offsets (64 runs per super-run, then one CTA over the super-runs); pass C scans
// each run in a warp, adds its offset and permutes back through the same tiles.
struct PsArgs {
  uint32_t* v[4];
  uint32_t* nat[4];
  uint32_t* run[4];   // 2^(log-6) run sums -> exclusive offsets within their super-run
  uint32_t* sup[4];   // 2^(log-12) super-run sums -> exclusive offsets
};
__device__ __forceinline__ uint32_t brev5(uint32_t x) { return __brev(x) >> 27; }
__device__ __forceinline__ uint32_t ps_mod(unsigned long long v) {  // v < 2^40
  uint32_t s = (uint32_t)(v >> 31) + ((uint32_t)v & P);
  return s >= P ? s - P : s;
}
template <bool BACK>
__global__ void __launch_bounds__(256) ps_tile_kernel(PsArgs a, uint32_t log) {
  __shared__ uint32_t T[2][32][65];
  const uint32_t Lp = log - 1, mb = Lp - 10;                 // bits of mid
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t mid[2] = {blockIdx.x, (~blockIdx.x) & ((1u << mb) - 1u)};
  uint32_t* __restrict__ v = a.v[blockIdx.y];
  uint32_t* __restrict__ nat = a.nat[blockIdx.y];
  uint32_t rowbase[2];
#pragma unroll
  for (int tt = 0; tt < 2; tt++) rowbase[tt] = (__brev(mid[tt]) >> (32 - mb)) << 6;
  if (!BACK) {
#pragma unroll
    for (int tt = 0; tt < 2; tt++)
      for (uint32_t b = warp; b < 32; b += 8) {
        const uint32_t* row = v + ((brev5(b) << (Lp - 4)) | rowbase[tt]);
        T[tt][b][lane] = row[lane];
        T[tt][b][lane + 32] = row[lane + 32];
      }
    __syncthreads();
  }
  for (uint32_t r = warp; r < 64; r += 8) {
    const uint32_t tt = r >> 5, aa = r & 31u, b = lane;
    const uint32_t m0 = (aa << (Lp - 5)) | (mid[tt] << 5);
    const uint32_t ce = 2u * brev5(aa), co = 2u * brev5((~aa) & 31u) + 1u, bo = (~b) & 31u;
    uint2* np = reinterpret_cast<uint2*>(nat + 2u * (size_t)m0) + b;
    if (!BACK) {
      const uint32_t e = T[tt][b][ce], o = T[tt ^ 1u][bo][co];
      *np = make_uint2(e, o);
      unsigned long long sum = (unsigned long long)e + o;
#pragma unroll
      for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
      if (lane == 0) a.run[blockIdx.y][m0 >> 5] = ps_mod(sum);
    } else {
      const uint2 x = *np;
      const unsigned long long loc = (unsigned long long)x.x + x.y;
      unsigned long long inc = loc;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        unsigned long long y = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (uint32_t)d) inc += y;
      }
      const uint32_t run = m0 >> 5;
      const unsigned long long off = (unsigned long long)a.run[blockIdx.y][run] + a.sup[blockIdx.y][run >> 6];
      T[tt][b][ce] = ps_mod(off + inc - x.y);
      T[tt ^ 1u][bo][co] = ps_mod(off + inc);
    }
  }
  if (BACK) {
    __syncthreads();
#pragma unroll
    for (int tt = 0; tt < 2; tt++)
      for (uint32_t b = warp; b < 32; b += 8) {
        uint32_t* row = v + ((brev5(b) << (Lp - 4)) | rowbase[tt]);
        row[lane] = T[tt][b][lane];
        row[lane + 32] = T[tt][b][lane + 32];
      }
  }
}
// 64 run sums per CTA (one warp, two runs per lane) -> exclusive offsets in place, total to sup
__global__ void __launch_bounds__(32) ps_runs_kernel(PsArgs a) {
  uint32_t* run = a.run[blockIdx.y] + (size_t)blockIdx.x * 64;
  const uint32_t lane = threadIdx.x;
  const uint2 x = reinterpret_cast<const uint2*>(run)[lane];
  unsigned long long inc = (unsigned long long)x.x + x.y;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    unsigned long long y = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= (uint32_t)d) inc += y;
  }
  const unsigned long long ex = inc - x.x - x.y;
  reinterpret_cast<uint2*>(run)[lane] = make_uint2(ps_mod(ex), ps_mod(ex + x.x));
  if (lane == 31) a.sup[blockIdx.y][blockIdx.x] = ps_mod(inc);
}
// one CTA per column: exclusive scan of the super-run sums
__global__ void __launch_bounds__(1024) ps_super_kernel(PsArgs a, uint32_t nsup) {
  __shared__ unsigned long long wsum[32];
  uint32_t* sup = a.sup[blockIdx.x];
  const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  unsigned long long carry = 0;
  for (uint32_t base = 0; base < nsup; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const unsigned long long x = i < nsup ? sup[i] : 0;
    unsigned long long inc = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      unsigned long long y = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= (uint32_t)d) inc += y;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    unsigned long long wb = 0, tot = 0;
    for (uint32_t k = 0; k < 32; k++) { if (k < w) wb += wsum[k]; tot += wsum[k]; }
    if (i < nsup) sup[i] = ps_mod(carry + wb + inc - x);
    carry = ps_mod(carry + tot);
    __syncthreads();
  }
}
size_t prefix_sum_tiled_words(uint32_t log) { return (((size_t)1 << log) + ((size_t)1 << (log - 6)) + ((size_t)1 << (log - 12)) + 3) & ~(size_t)3; }
// ncols <= 4 columns of one size (log >= 12); scratch: ncols * prefix_sum_tiled_words(log) words
int launch_prefix_sum_bitrev_tiled(uint32_t* const* v, uint32_t ncols, uint32_t log, uint32_t* scratch, cudaStream_t st) {
  if (log < 12 || ncols == 0 || ncols > 4) return -1;
  const size_t words = prefix_sum_tiled_words(log);
  PsArgs a;
  for (uint32_t k = 0; k < 4; k++) {
    uint32_t* s = scratch + (size_t)(k < ncols ? k : 0) * words;
    a.v[k] = v[k < ncols ? k : 0];
    a.nat[k] = s; a.run[k] = s + ((size_t)1 << log); a.sup[k] = a.run[k] + ((size_t)1 << (log - 6));
  }
  const uint32_t pairs = 1u << (log - 12);
  ps_tile_kernel<false><<<dim3(pairs, ncols), 256, 0, st>>>(a, log); g_launch_count++;
  ps_runs_kernel<<<dim3(1u << (log - 12), ncols), 32, 0, st>>>(a); g_launch_count++;
  ps_super_kernel<<<ncols, 1024, 0, st>>>(a, 1u << (log - 12)); g_launch_count++;
  ps_tile_kernel<true><<<dim3(pairs, ncols), 256, 0, st>>>(a, log); g_launch_count++;
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------- eval_at_point
// value = sum_i c_i * prod_k f_k^{bit_k(i)},  f = [p.y, p.x, pi(p.x), pi^2(p.x), ...]   (CpuBackend fold()).
// Stage 1: a CTA folds 2^13 base-field coefficients -> one QM31 partial.  A thread owns 32 consecutive coefficients and
// accumulates c_j * basis_j (basis_j = the monomial of f_0..f_4 for index j, 32 QM31 values in shared memory) in 64-bit
// lanes: one IMAD.WIDE per coordinate, one partial reduction every third term — ~8 instructions per coefficient instead of
// the ~60 of a QM31 Horner fold, so the kernel is bound by the 4 bytes it reads per coefficient.  The 256 thread partials
// are then folded with f_5..f_12 in shared memory.  Stage 2: one CTA per task folds the chunk partials with the rest.
typedef EvalTaskHost EvalTask;  // {coeffs, log, first_block (prefix of stage-1 blocks), f[28]}
constexpr uint32_t EV_CHUNK_LOG = 13;

// high word back in with weight 2^32 == 2 (mod P): one IMAD.WIDE, result < 3 * 2^32 (three more products fit below 2^64)
__device__ __forceinline__ uint64_t ev_fold(uint64_t v) { return (uint64_t)(uint32_t)(v >> 32) * 2u + (uint32_t)v; }
__device__ __forceinline__ uint32_t ev_red(uint64_t v) { return m_red_wide(v); }

__global__ void __launch_bounds__(256) eval_stage1_kernel(const EvalTask* __restrict__ tasks, uint32_t ntasks, QM31* __restrict__ partials) {
  // locate task by binary search over first_block
  uint32_t lo = 0, hi = ntasks - 1;
  while (lo < hi) { uint32_t mid = (lo + hi + 1) >> 1; if (tasks[mid].first_block <= blockIdx.x) lo = mid; else hi = mid - 1; }
  const EvalTask& t = tasks[lo];
  const uint32_t chunk = blockIdx.x - t.first_block;
  const uint32_t clog = t.log < EV_CHUNK_LOG ? t.log : EV_CHUNK_LOG;
  const uint32_t n = 1u << clog;
  const uint32_t* c = t.coeffs + ((size_t)chunk << EV_CHUNK_LOG);
  __shared__ QM31 sm[256];
  __shared__ uint4 basis[32];
  if (threadIdx.x < 32) {
    QM31 b = q_fromm(1);
#pragma unroll
    for (uint32_t k = 0; k < 5; k++)
      if (((threadIdx.x >> k) & 1u) && k < clog) b = q_mul(b, t.f[k]);
    basis[threadIdx.x] = make_uint4(b.a.a, b.a.b, b.b.a, b.b.b);
  }
  __syncthreads();
  const uint32_t i0 = threadIdx.x * 32;
  QM31 acc = q_zero();
  if (i0 < n) {
    uint64_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    const uint32_t cnt = n - i0 < 32 ? n - i0 : 32;   // a power of two >= 1; < 4 only for polynomials of 1 or 2 words
    if (cnt >= 4) {
#pragma unroll
      for (uint32_t j = 0; j < 32; j += 4) {
        if (j < cnt) {
          uint4 x = __ldg(reinterpret_cast<const uint4*>(c + i0 + j));
          const uint32_t xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int e = 0; e < 4; e++) {
            uint4 b = basis[j + e];
            s0 += (uint64_t)xv[e] * b.x; s1 += (uint64_t)xv[e] * b.y; s2 += (uint64_t)xv[e] * b.z; s3 += (uint64_t)xv[e] * b.w;
            if ((j + e) % 3 == 2) { s0 = ev_fold(s0); s1 = ev_fold(s1); s2 = ev_fold(s2); s3 = ev_fold(s3); }
          }
        }
      }
    } else {
      for (uint32_t j = 0; j < cnt; j++) {
        uint32_t x = __ldg(c + i0 + j);
        uint4 b = basis[j];
        s0 += (uint64_t)x * b.x; s1 += (uint64_t)x * b.y; s2 += (uint64_t)x * b.z; s3 += (uint64_t)x * b.w;
      }
    }
    acc = q_make(ev_red(s0), ev_red(s1), ev_red(s2), ev_red(s3));
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
  // levels 5..clog-1 over the 2^(clog-5) thread partials
  for (uint32_t lvl = 5; lvl < clog; lvl++) {
    uint32_t cnt = 1u << (clog - lvl - 1);
    QM31 r = q_zero();
    if (threadIdx.x < cnt) r = q_add(sm[2 * threadIdx.x], q_mul(sm[2 * threadIdx.x + 1], t.f[lvl]));
    __syncthreads();
    if (threadIdx.x < cnt) sm[threadIdx.x] = r;
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[blockIdx.x] = sm[0];
}
__global__ void __launch_bounds__(256) eval_stage2_kernel(const EvalTask* __restrict__ tasks, const QM31* __restrict__ partials,
                                                          QM31* __restrict__ out, QM31* __restrict__ work) {
  const EvalTask& t = tasks[blockIdx.x];
  if (t.log <= EV_CHUNK_LOG) { if (threadIdx.x == 0) out[blockIdx.x] = partials[t.first_block]; return; }
  const uint32_t levels = t.log - EV_CHUNK_LOG;
  QM31* w = work + t.first_block;  // in-place tree over this task's partials (copied first)
  uint32_t cnt = 1u << levels;
  for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) w[i] = partials[t.first_block + i];
  __syncthreads();
  for (uint32_t lvl = 0; lvl < levels; lvl++) {
    uint32_t half = cnt >> 1;
    // read pairs, sync, write — strided so no in-place hazard: out index i < 2i
    for (uint32_t base = 0; base < half; base += blockDim.x) {
      uint32_t i = base + threadIdx.x;
      QM31 r = q_zero();
      if (i < half) r = q_add(w[2 * i], q_mul(w[2 * i + 1], t.f[EV_CHUNK_LOG + lvl]));
      __syncthreads();
      if (i < half) w[i] = r;
      __syncthreads();
    }
    cnt = half;
  }
  if (threadIdx.x == 0) out[blockIdx.x] = w[0];
}

int launch_eval_at_point_tasks(const void* d_tasks, uint32_t ntasks, uint32_t total_blocks, QM31* d_partials, QM31* d_work,
                               QM31* d_out, cudaStream_t st) {
  if (!ntasks) return 0;
  eval_stage1_kernel<<<total_blocks, 256, 0, st>>>(reinterpret_cast<const EvalTask*>(d_tasks), ntasks, d_partials); g_launch_count++;
  eval_stage2_kernel<<<ntasks, 256, 0, st>>>(reinterpret_cast<const EvalTask*>(d_tasks), d_partials, d_out, d_work); g_launch_count++;
  return (int)cudaGetLastError();
}

// decommitment gather: one thread per word
__global__ void gather_kernel(const uint32_t* const* __restrict__ src, uint32_t n, uint32_t words, uint32_t* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n * words) out[i] = src[i / words][i % words];
}
int launch_gather(const uint32_t* const* d_src, uint32_t n, uint32_t words, uint32_t* d_out, cudaStream_t st) {
  gather_kernel<<<(n * words + 255) / 256, 256, 0, st>>>(d_src, n, words, d_out); g_launch_count++;
  return (int)cudaGetLastError();
}

// 16x lane broadcast of a trace column (reference: PackedBaseField::broadcast in every table.rs trace_evaluation).
__global__ void broadcast16_kernel(const uint32_t* __restrict__ s, uint4* __restrict__ d, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * 4; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t x = __ldg(s + (i >> 2));
    d[i] = make_uint4(x, x, x, x);
  }
}
__global__ void broadcast_cols_kernel(const uint32_t* const* __restrict__ src, uint32_t* const* __restrict__ dst, size_t n, uint32_t sh) {
  const uint32_t* __restrict__ s = src[blockIdx.y];
  uint4* __restrict__ d = reinterpret_cast<uint4*>(dst[blockIdx.y]);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n << sh); i += (size_t)gridDim.x * blockDim.x) {
    uint32_t x = __ldg(s + (i >> sh));
    d[i] = make_uint4(x, x, x, x);
  }
}
int launch_broadcast_cols(const uint32_t* const* src, uint32_t* const* dst, uint32_t ncols, size_t src_len, uint32_t rep_log, cudaStream_t st) {
  if (rep_log < 2 || rep_log > 8) return -1;
  size_t vecs = src_len << (rep_log - 2);
  unsigned bx = (unsigned)std::min<size_t>((vecs + 255) / 256, 4096);
  broadcast_cols_kernel<<<dim3(bx, ncols), 256, 0, st>>>(src, dst, src_len, rep_log - 2); g_launch_count++;
  return (int)cudaGetLastError();
}
int launch_broadcast16(const uint32_t* src, uint32_t* dst, size_t src_len, cudaStream_t st) {
  broadcast16_kernel<<<grid_for(src_len * 4, 256), 256, 0, st>>>(src, reinterpret_cast<uint4*>(dst), src_len); g_launch_count++;
  return (int)cudaGetLastError();
}

}  // namespace sb
