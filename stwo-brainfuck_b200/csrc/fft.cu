// Circle FFT for sm_100a: twiddle tree, interpolate (iFFT) and evaluate (FFT / low-degree extension).
//
// Replaces SimdBackend's PolyOps::{precompute_twiddles, interpolate(_columns), evaluate(_polynomials), extend}
// (stwo-prover 0.1.1 @ 31e8dbc core/backend/simd/{circle.rs,fft/*}; definitions SURVEY.md A.3/A.4), reached from
// crates/brainfuck_prover/src/brainfuck_air/mod.rs:480-484 (twiddles), :497,550-562,690-702 (extend_evals →
// interpolate_columns) and :500,583,723 (commit → evaluate_polynomials).
//
// Design (B200-first, not SimdBackend's cache-blocked recursion):
//  * a transform of log size n is cut into "passes"; a pass moves a tile of 2^K elements HBM→smem once, runs up to
//    K butterfly layers on it out of registers (radix-16 groups: 4 layers per smem round trip) and writes it back;
//  * the low pass owns the contiguous bits [0,K); a strided pass owns global bits [L0,L0+k) x 2^c contiguous
//    elements, so every global access is a 2^c*4-byte run (>= 64 B) issued as 128-bit vector loads/stores;
//  * columns of one size are batched through blockIdx.y, so concurrent CTAs share twiddle lines in L1/L2;
//  * the blow-up layers of an LDE (zero high coefficients) are not computed: the first pass reads index & (2^src-1).
// All values canonical in [0,P) at kernel boundaries.
#include "kernels.cuh"

namespace sb {

// ---------------------------------------------------------------- twiddle tree
// Level j of the tree rooted at half_odds(R): tw[off_j + bitrev(i, R-j-1)] = coset_j.at(i).x, i < 2^(R-j-1);
// coset_j = half_odds(R-j): initial index 2^(29-(R-j)), step 2^(31-(R-j)).  Last word is the padding 1.
__constant__ Pt c_gen_pow[31];  // G^(2^k)

__device__ __forceinline__ Pt point_at_index(uint32_t idx) {
  Pt r = {1u, 0u};
#pragma unroll 1
  for (int k = 0; k < 31; k++) {
    if ((idx >> k) & 1u) r = p_add(r, c_gen_pow[k]);
  }
  return r;
}

__global__ void twiddle_tree_kernel(uint32_t* __restrict__ tw, uint32_t* __restrict__ itw, uint32_t R) {
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)1 << R;
  if (tid >= total) return;
  uint32_t v;
  if (tid == total - 1) {
    v = 1u;
  } else {
    // find level: offsets are total - 2^(R-j) for level j  (sum_{j'<j} 2^(R-j'-1)).
    size_t rem = total - tid;                       // in (2^(R-j-1), 2^(R-j)]
    uint32_t lg = 63 - __clzll((unsigned long long)(rem - 1));  // floor(log2(rem-1)), rem>=2
    uint32_t j = R - 1 - lg;                        // level
    uint32_t hl = R - j - 1;                        // log of level size
    uint32_t pos = (uint32_t)(tid - (total - ((size_t)2 << hl)));
    uint32_t i = bitrev32(pos, hl);
    uint32_t cl = R - j;                            // coset log
    uint32_t idx = (1u << (29 - cl)) + (uint32_t)(((uint64_t)i << (31 - cl)) & 0x7fffffffu);
    v = point_at_index(idx & 0x7fffffffu).x;
  }
  tw[tid] = v;
  itw[tid] = m_inv(v);
}

int launch_twiddle_tree(uint32_t* tw, uint32_t* itw, uint32_t R, cudaStream_t st) {
  static bool init = false;
  if (!init) {
    Pt g[31];
    g[0] = {GEN_X, GEN_Y};
    for (int k = 1; k < 31; k++) g[k] = p_dbl(g[k - 1]);
    cudaError_t e = cudaMemcpyToSymbol(c_gen_pow, g, sizeof(g));
    if (e != cudaSuccess) return (int)e;
    init = true;
  }
  size_t total = (size_t)1 << R;
  unsigned blocks = (unsigned)((total + 255) / 256);
  twiddle_tree_kernel<<<blocks, 256, 0, st>>>(tw, itw, R); g_launch_count++;
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------- butterflies
__device__ __forceinline__ void bfly_fwd(uint32_t& a, uint32_t& b, uint32_t t) {
  uint32_t m = m_reduce64((uint64_t)b * t);
  uint32_t a0 = a;
  a = m_add(a0, m);
  b = m_sub(a0, m);
}
__device__ __forceinline__ void bfly_inv(uint32_t& a, uint32_t& b, uint32_t t) {
  uint32_t a0 = a;
  a = m_add(a0, b);
  b = m_reduce64((uint64_t)m_sub(a0, b) * t);
}

__device__ __forceinline__ uint32_t smpad(uint32_t i) { return i + (i >> 5); }

struct FftPass {
  const uint32_t* const* src;
  uint32_t* const* dst;
  const uint32_t* twend;  // one past the end of the (i)twiddle buffer
  uint32_t n;             // transform log size
  uint32_t src_log;       // loads read index & (2^src_log - 1); forward layers >= src_log are identity-duplications
  uint32_t K, c, L0;      // tile: 2^c contiguous x 2^(K-c) rows at stride 2^L0
  uint32_t lb_lo;         // first local bit whose layer this pass computes
  uint32_t scale;         // multiply on store (inverse normalisation), 1 = none
};

// Twiddle of layer l (>=1) at index h / circle layer 0 derived from line layer 1: [x,y] -> [y,-y,-x,x].
__device__ __forceinline__ uint32_t line_tw(const FftPass& p, uint32_t l, uint32_t h) {
  return __ldg(p.twend - ((size_t)1 << (p.n - l)) + h);
}
__device__ __forceinline__ uint32_t circle_tw(const FftPass& p, uint32_t h) {
  const uint32_t* l1 = p.twend - ((size_t)1 << (p.n - 1));
  uint32_t pair = (h >> 2) * 2;
  uint32_t x = __ldg(l1 + pair), y = __ldg(l1 + pair + 1);
  uint32_t s = h & 3u;
  uint32_t v = (s < 2) ? y : x;
  return (s == 1 || s == 2) ? (P - v) : v;  // twiddles are never 0
}

template <bool INV, int R>
__device__ __forceinline__ void radix_round(const FftPass& p, uint32_t* sm, uint32_t b, uint32_t gbase) {
  const uint32_t K = p.K, c = p.c, L0 = p.L0;
  const uint32_t gb0 = (b < c) ? b : L0 + (b - c);  // global bit of local bit b (round never straddles c unless L0==c)
  constexpr uint32_t M = 1u << R;
  for (uint32_t q = threadIdx.x; q < (1u << (K - R)); q += blockDim.x) {
    uint32_t low = q & ((1u << b) - 1u), high = q >> b;
    uint32_t li0 = low | (high << (b + R));
    uint32_t g0 = gbase | (li0 & ((1u << c) - 1u)) | ((li0 >> c) << L0);
    uint32_t v[M];
#pragma unroll
    for (uint32_t m = 0; m < M; m++) v[m] = sm[smpad(li0 | (m << b))];
#pragma unroll
    for (int ss = 0; ss < R; ss++) {
      const int s = INV ? ss : (R - 1 - ss);
      const uint32_t l = gb0 + s;
      if (!INV && l >= p.src_log) continue;  // zero-padded coefficients: (v0, 0) -> (v0, v0), done by the load
      const uint32_t hbase = g0 >> (l + 1);
#pragma unroll
      for (uint32_t j = 0; j < (M >> (s + 1)); j++) {
        uint32_t t = (l == 0) ? circle_tw(p, hbase + j) : line_tw(p, l, hbase + j);
#pragma unroll
        for (uint32_t w = 0; w < (1u << s); w++) {
          uint32_t m0 = (j << (s + 1)) | w, m1 = m0 | (1u << s);
          if (INV) bfly_inv(v[m0], v[m1], t); else bfly_fwd(v[m0], v[m1], t);
        }
      }
    }
#pragma unroll
    for (uint32_t m = 0; m < M; m++) sm[smpad(li0 | (m << b))] = v[m];
  }
}

template <bool INV>
__global__ void __launch_bounds__(256) fft_pass_kernel(FftPass p) {
  extern __shared__ uint32_t sm[];
  const uint32_t K = p.K, c = p.c, L0 = p.L0, k = K - c;
  const uint32_t tile = blockIdx.x;
  const uint32_t nlow = L0 - c;
  const uint32_t gbase = ((tile & ((1u << nlow) - 1u)) << c) | ((tile >> nlow) << (L0 + k));
  const uint32_t* __restrict__ src = p.src[blockIdx.y];
  uint32_t* __restrict__ dst = p.dst[blockIdx.y];
  const uint32_t smask = (p.src_log >= 32) ? 0xffffffffu : ((1u << p.src_log) - 1u);
  const uint32_t cm = (1u << c) - 1u;

  for (uint32_t li = threadIdx.x * 4; li < (1u << K); li += blockDim.x * 4) {
    uint32_t g = (gbase | (li & cm) | ((li >> c) << L0)) & smask;
    uint4 x = __ldg(reinterpret_cast<const uint4*>(src + g));
    uint32_t o = smpad(li);
    sm[o] = x.x; sm[o + 1] = x.y; sm[o + 2] = x.z; sm[o + 3] = x.w;
  }
  __syncthreads();

  // rounds over local bits [lb_lo, K), 4 layers at a time; inverse ascends, forward descends.
  const uint32_t nl = K - p.lb_lo;
  const uint32_t nr = (nl + 3) / 4;
  for (uint32_t r = 0; r < nr; r++) {
    uint32_t b, w;
    if (INV) { b = p.lb_lo + 4 * r; w = min(4u, K - b); }
    else { uint32_t top = K - 4 * r; w = min(4u, top - p.lb_lo); b = top - w; }
    switch (w) {
      case 4: radix_round<INV, 4>(p, sm, b, gbase); break;
      case 3: radix_round<INV, 3>(p, sm, b, gbase); break;
      case 2: radix_round<INV, 2>(p, sm, b, gbase); break;
      default: radix_round<INV, 1>(p, sm, b, gbase); break;
    }
    __syncthreads();
  }

  const uint32_t scale = p.scale;
  for (uint32_t li = threadIdx.x * 4; li < (1u << K); li += blockDim.x * 4) {
    uint32_t g = gbase | (li & cm) | ((li >> c) << L0);
    uint32_t o = smpad(li);
    uint4 x = make_uint4(sm[o], sm[o + 1], sm[o + 2], sm[o + 3]);
    if (scale != 1u) { x.x = m_mul(x.x, scale); x.y = m_mul(x.y, scale); x.z = m_mul(x.z, scale); x.w = m_mul(x.w, scale); }
    else { x.x = x.x == P ? 0 : x.x; x.y = x.y == P ? 0 : x.y; x.z = x.z == P ? 0 : x.z; x.w = x.w == P ? 0 : x.w; }
    *reinterpret_cast<uint4*>(dst + g) = x;
  }
}

// ---------------------------------------------------------------- host-side pass planner
static const uint32_t KMAX = 13;  // 2^13 words (+pad) = 33 KB smem per CTA
static const uint32_t KSTRIDE_MAX = 9, CMIN = 4;

struct PassDesc { uint32_t K, c, L0, lb_lo; };

static int plan_passes(uint32_t n, PassDesc* out) {  // ascending layer order
  int np = 0;
  uint32_t K0 = n < KMAX ? n : KMAX;
  out[np++] = {K0, K0, K0, 0};
  uint32_t rem = n - K0;
  if (rem) {
    uint32_t ns = (rem + KSTRIDE_MAX - 1) / KSTRIDE_MAX;
    uint32_t L = K0;
    for (uint32_t i = 0; i < ns; i++) {
      uint32_t k = rem / ns + (i < rem % ns ? 1 : 0);
      uint32_t c = KMAX - k; if (c > 5) c = 5; if (c < CMIN) c = CMIN;
      out[np++] = {k + c, c, L, c};
      L += k;
    }
  }
  return np;
}

static uint32_t threads_for(uint32_t K) {
  uint32_t t = K >= 12 ? 256u : (K >= 4 ? (1u << (K - 4)) : 1u);
  if (t < 32) t = 32;
  if (t > 256) t = 256;
  return t;
}

template <bool INV>
static int run_pass(const PassDesc& d, const uint32_t* const* src, uint32_t* const* dst, uint32_t ncols, uint32_t n,
                    uint32_t src_log, const uint32_t* twend, uint32_t scale, cudaStream_t st) {
  FftPass p{src, dst, twend, n, src_log, d.K, d.c, d.L0, d.lb_lo, scale};
  dim3 grid(1u << (n - d.K), ncols);
  size_t smem = ((size_t)(1u << d.K) + ((1u << d.K) >> 5) + 4) * 4;
  fft_pass_kernel<INV><<<grid, threads_for(d.K), smem, st>>>(p); g_launch_count++;
  return (int)cudaGetLastError();
}

// In-place interpolate of ncols columns of log size n (device pointer array `cols`).
int launch_interpolate(uint32_t* const* cols, uint32_t ncols, uint32_t n, const uint32_t* itw_end, cudaStream_t st) {
  if (n < 3 || ncols == 0) return n < 3 ? -1 : 0;
  PassDesc pd[8];
  int np = plan_passes(n, pd);
  uint32_t ninv = m_inv(m_pow(2, n));
  for (int i = 0; i < np; i++) {
    int e = run_pass<true>(pd[i], cols, cols, ncols, n, 32, itw_end, i == np - 1 ? ninv : 1u, st);
    if (e) return e;
  }
  return 0;
}

// coeffs (log src_log) -> evaluations on the canonic domain of log n = src_log + log_blowup, out of place.
int launch_evaluate(const uint32_t* const* coeffs, uint32_t* const* out, uint32_t ncols, uint32_t src_log, uint32_t n,
                    const uint32_t* tw_end, cudaStream_t st) {
  if (n < 3 || ncols == 0) return n < 3 ? -1 : 0;
  PassDesc pd[8];
  int np = plan_passes(n, pd);
  for (int i = np - 1; i >= 0; i--) {
    bool first = (i == np - 1);
    int e = run_pass<false>(pd[i], first ? coeffs : (const uint32_t* const*)out, out, ncols, n, first ? src_log : 32, tw_end, 1u, st);
    if (e) return e;
  }
  return 0;
}

}  // namespace sb
