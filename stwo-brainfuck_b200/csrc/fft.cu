// Circle FFT for sm_100a: twiddle tree, interpolate (iFFT) and evaluate (FFT / low-degree extension).
//
// Replaces SimdBackend's PolyOps::{precompute_twiddles, interpolate(_columns), evaluate(_polynomials), extend}
// (stwo-prover 0.1.1 @ 31e8dbc core/backend/simd/{circle.rs,fft/*}; definitions SURVEY.md A.3/A.4), reached from
// crates/brainfuck_prover/src/brainfuck_air/mod.rs:480-484 (twiddles), :497,550-562,690-702 (extend_evals →
// interpolate_columns) and :500,583,723 (commit → evaluate_polynomials).
//
// Design (B200-first, not SimdBackend's cache-blocked recursion):
//  * a transform of log size n is cut into "passes"; a pass moves a tile of 2^K <= 2^13 elements HBM->smem once (128-bit
//    loads), runs its butterfly layers out of registers in radix-32/16 rounds (5 or 4 layers per smem round trip) and
//    writes the tile back;
//  * the low pass owns the contiguous bits [0,K); a strided pass owns <= 9 higher bits x 16 contiguous words (64-byte
//    runs), or exactly 10 bits x 8 words when that saves a whole pass (log 23);
//  * everything that shapes addressing (K, round start, radix) is a template parameter: shared-memory accesses are
//    [base + immediate], twiddles of a round are fetched with 128/64-bit loads (one load serves the circle layer and line
//    layer 1), and ncu showed the round-1 version spending 13.9 instructions per element-layer against ~6 for the math;
//  * a butterfly is 7 integer instructions: twiddles are stored doubled so that the 64-bit product 2bt splits into
//    (bt >> 31, (bt & P) << 1) without a mask (IMAD.WIDE + LEA.HI), every conditional subtraction of P is ONE VIADDMNMX
//    (min(s, s + imm)), the two additions are IMADs with a runtime +-1 held in a uniform register (FMA pipe): 4 ALU-pipe +
//    3 FMA-pipe instructions (ncu: ALU 51-62 %, FMA 21-25 %, issue 58-68 %; the 10-instruction form of round 1 was bound by
//    the FMA pipe's IMAD rate);
//  * shared memory is padded per pass so that every round is bank-conflict-free (struct Pad) and the load / store phases of
//    the strided passes use 128-bit accesses;
//  * columns that repeat every value 2^r times (the whole main trace, r = 4) are transformed on their distinct values: the
//    LINE variants at the end of this file;
//  * columns of one size are batched through blockIdx.y, so concurrent CTAs share twiddle lines in L1/L2;
//  * the blow-up layers of an LDE (zero high coefficients) are not computed: the first pass reads index & (2^src-1).
// All values canonical in [0,P) at kernel boundaries.
#include <cstdlib>
#include <mutex>
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
    size_t rem = total - tid;                                   // in (2^(R-j-1), 2^(R-j)]
    uint32_t lg = 63 - __clzll((unsigned long long)(rem - 1));  // floor(log2(rem-1)), rem >= 2
    uint32_t j = R - 1 - lg;                                    // level
    uint32_t hl = R - j - 1;                                    // log of level size
    uint32_t pos = (uint32_t)(tid - (total - ((size_t)2 << hl)));
    uint32_t i = bitrev32(pos, hl);
    uint32_t cl = R - j;                                        // coset log
    uint32_t idx = (1u << (29 - cl)) + (uint32_t)(((uint64_t)i << (31 - cl)) & 0x7fffffffu);
    v = point_at_index(idx & 0x7fffffffu).x;
  }
  const uint32_t iv = m_inv(v);
  tw[tid] = v;
  itw[tid] = iv;
  // second half of both buffers: the same tree doubled (2t < 2^32), which is what the butterflies multiply by (mulred)
  tw[total + tid] = 2u * v;
  itw[total + tid] = 2u * iv;
}

int launch_twiddle_tree(uint32_t* tw, uint32_t* itw, uint32_t R, cudaStream_t st) {
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
      cudaError_t e = cudaMemcpyToSymbol(c_gen_pow, g, sizeof(g));
      if (e != cudaSuccess) return (int)e;
      init[dev] = true;
    }
  }
  size_t total = (size_t)1 << R;
  unsigned blocks = (unsigned)((total + 255) / 256);
  twiddle_tree_kernel<<<blocks, 256, 0, st>>>(tw, itw, R); g_launch_count++;
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------- butterflies (pipe-balanced)
// x*1+y on the FMA pipe; `one` is a kernel argument so ptxas keeps the IMAD instead of folding it into an ALU-pipe IADD.
__device__ __forceinline__ uint32_t fadd(uint32_t x, uint32_t y, uint32_t one) {
  uint32_t r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(one), "r"(y));
  return r;
}
// y - x as x*(-1)+y, again on the FMA pipe (mone = 0 - one)
__device__ __forceinline__ uint32_t fsub(uint32_t y, uint32_t x, uint32_t mone) {
  uint32_t r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(mone), "r"(y));
  return r;
}
// Butterfly = 7 instructions.  sm_100a has VIADDMNMX (min(s, s + imm) in ONE ALU-pipe instruction), which is what a
// conditional subtraction of P is; with it a butterfly is IMAD.WIDE + LEA.HI + VIADDMNMX (b*t), IMAD + VIADDMNMX (a + bt),
// IMAD + VIADDMNMX (a - bt): 4 ALU-pipe and 3 FMA-pipe instructions.  (Round 1/2 spelled the reduction as IMAD + VIMNMX to
// keep the ALU pipe free: 10 instructions, 6 of them on the FMA pipe, whose IMAD rate — one warp instruction every two
// cycles per sub-partition — then bounded the kernel at 12 cycles per butterfly; now the ALU pipe bounds it at 8.)
// The two remaining additions stay IMADs with the +-1 multiplicand in a UNIFORM register (tools/exp/gmix.cu: the
// three-vector-register form costs 9 % of the pipe); FK arrives as kernel arguments so that `one` never needs a vector register.
struct FK { uint32_t one, mone; };
static const FK FK_HOST{1u, 0xffffffffu};
// canonical reduction of s in [0, 2P): min(s, s - P) as unsigned
__device__ __forceinline__ uint32_t cred(uint32_t s) { return min(s, s + 0x80000001u); }
// b*t mod P for b in [0,P], t in [0,P) given as t2 = 2t: the 64-bit product 2bt has (bt >> 31) in its high word and
// (bt & P) << 1 in its low word, so bt = hi + (lo >> 1) (mod P): LEA.HI, no shifts or masks.
__device__ __forceinline__ uint32_t mulred(uint32_t b, uint32_t t2) {
  const uint64_t pr = (uint64_t)b * t2;
  const uint32_t lo = (uint32_t)pr, hi = (uint32_t)(pr >> 32);
  uint32_t s;
  asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(s) : "r"(lo), "r"(0x80000000u), "r"(hi));  // ptxas turns this into LEA.HI (ALU pipe)
  return cred(s);
}
__device__ __forceinline__ void bfly_fwd(uint32_t& a, uint32_t& b, uint32_t t2, const FK& k) {
  const uint32_t m = mulred(b, t2);
  const uint32_t a0 = a;
  a = cred(fadd(a0, m, k.one));
  const uint32_t d = fsub(a0, m, k.mone);                 // wraps when a0 < m
  b = min(d, d + P);
}
__device__ __forceinline__ void bfly_inv(uint32_t& a, uint32_t& b, uint32_t t2, const FK& k) {
  const uint32_t a0 = a;
  a = cred(fadd(a0, b, k.one));
  const uint32_t d = fsub(a0, b, k.mone);
  b = mulred(min(d, d + P), t2);
}

// ---------------------------------------------------------------- compile-time round partition
// A pass processes `nbits` consecutive local bits in ceil(nbits/5) register rounds of 3..5 bits.
__host__ __device__ constexpr int n_rounds(int nbits) { return (nbits + 4) / 5; }
__host__ __device__ constexpr int round_size(int nbits, int i) { return nbits / n_rounds(nbits) + (i < nbits % n_rounds(nbits) ? 1 : 0); }
__host__ __device__ constexpr int round_start(int nbits, int i) {
  int s = 0;
  for (int j = 0; j < i; j++) s += round_size(nbits, j);
  return s;
}
constexpr int STRIDED_C = 4;  // log2 of the contiguous words per row of a strided tile (64-byte runs); 3 for the 10-layer pass

// Shared-memory index of local element i (padding against bank conflicts).
//  * low pass: i + i / 32 — the first round's threads own 32 consecutive words each (stride 33: conflict-free), later rounds
//    read 32 consecutive words per warp;
//  * strided pass: a warp of the first round (local bits [SC, SC + R0)) touches 2^(5-SC) runs of 2^SC words that lie
//    2^(SC+R0) words apart; 2^SC words of padding per 2^(SC+R0)-word block puts consecutive runs 2^SC banks apart, so the
//    runs of a warp tile the 32 banks (ncu on the i + i/32 padding of round 1: 9.0 M conflicts in 34.8 M shared wavefronts on
//    fft_kernel<1,10,0,0,4>).  The padding is a multiple of four words, so the load / store phases use 128-bit accesses.
// PAD is additive over the disjoint bit fields the rounds use (a thread's base + a compile-time offset per element).
template <bool LOW, int K, int SC>
struct Pad {
  static constexpr int BLK = LOW ? 5 : SC + round_size(K - SC, 0);   // log2 of the block that is followed by padding
  static constexpr int PW = LOW ? 0 : SC;                            // log2 of the padding words per block
  static __host__ __device__ constexpr uint32_t at(uint32_t i) { return i + ((i >> BLK) << PW); }
  static constexpr uint32_t words = (1u << K) + (((1u << K) >> BLK) << PW);
};

struct FftArgs {
  const uint32_t* const* src;
  uint32_t* const* dst;
  const uint32_t* twend;  // one past the end of the DOUBLED (i)twiddle tree (second half of the buffer)
  uint32_t n;             // transform log size
  uint32_t src_log;       // loads read index & (2^src_log - 1); forward layers >= src_log are copies (zero-padded coeffs)
  uint32_t L0;            // strided pass: global bit of local bit STRIDED_C
  uint32_t scale;         // multiply on store (inverse normalisation), 1 = none; the kernels multiply by 2*scale (mulred)
  FK k;                   // runtime 1, -1, P, 2^32 - P (see fadd, FK)
};

// One register round over local bits [B, B+R) of a 2^K tile.  gb = global bit of local bit B; T = tile index bits above the tile.
template <bool INV, int K, int B, int R, bool CIRCLE, bool LOW, int SC>
__device__ __forceinline__ void fft_round(uint32_t* __restrict__ sm, const FftArgs& a, uint32_t T, uint32_t gb) {
  typedef Pad<LOW, K, SC> PD;
  constexpr int M = 1 << R;
  constexpr int NG = 1 << (K - R);
  const FK one = a.k;
  for (int q = threadIdx.x; q < NG; q += blockDim.x) {
    const uint32_t low = q & ((1u << B) - 1u), high = (uint32_t)q >> B;
    const uint32_t li0 = low | (high << (B + R));
    uint32_t* p = sm + PD::at(li0);
    uint32_t v[M];
#pragma unroll
    for (int m = 0; m < M; m++) v[m] = p[PD::at((uint32_t)m << B)];
    const uint32_t H = (T << (K - B - R)) | high;
    // CIRCLE (low pass, B == 0): line layer 1's 2^(R-2) twiddles also define the 2^(R-1) circle twiddles: [x,y] -> [y,-y,-x,x];
    // fetched right before the first layer that needs them (the forward round walks s = R-1 .. 0: live for two layers only)
    uint32_t w1[CIRCLE ? (M / 4) : 1];
#pragma unroll
    for (int ss = 0; ss < R; ss++) {
      const int s = INV ? ss : (R - 1 - ss);
      const uint32_t l = gb + s;
      if (CIRCLE && s == (INV ? 0 : 1)) {
        const uint32_t* l1 = a.twend - ((size_t)1 << (a.n - 1)) + ((size_t)H << (R - 2));
        if (R >= 4) {
#pragma unroll
          for (int i = 0; i < M / 16; i++) {
            uint4 x = __ldg(reinterpret_cast<const uint4*>(l1) + i);
            w1[4 * i] = x.x; w1[4 * i + 1] = x.y; w1[4 * i + 2] = x.z; w1[4 * i + 3] = x.w;
          }
        } else {
          uint2 x = __ldg(reinterpret_cast<const uint2*>(l1));
          w1[0] = x.x; w1[1] = x.y;
        }
      }
      if (!INV && l >= a.src_log) continue;  // (v0, 0) -> (v0, v0): already done by the masked load
      const int NT = M >> (s + 1);           // distinct twiddles of this layer in the group
      uint32_t tw[M / 2];
      if (CIRCLE && s == 0) {
#pragma unroll
        for (int j = 0; j < M / 2; j++) {
          uint32_t x = w1[2 * (j >> 2)], y = w1[2 * (j >> 2) + 1];
          uint32_t t = ((j & 3) < 2) ? y : x;
          tw[j] = ((j & 3) == 1 || (j & 3) == 2) ? (2u * P - t) : t;   // doubled twiddles: -t is 2P - 2t
        }
      } else if (CIRCLE && s == 1) {
#pragma unroll
        for (int j = 0; j < M / 4; j++) tw[j] = w1[j];
      } else {
        const uint32_t* base = a.twend - ((size_t)1 << (a.n - l)) + ((size_t)H << (R - 1 - s));
        if (NT >= 4) {
#pragma unroll
          for (int i = 0; i < NT / 4; i++) {
            uint4 x = __ldg(reinterpret_cast<const uint4*>(base) + i);
            tw[4 * i] = x.x; tw[4 * i + 1] = x.y; tw[4 * i + 2] = x.z; tw[4 * i + 3] = x.w;
          }
        } else if (NT == 2) {
          uint2 x = __ldg(reinterpret_cast<const uint2*>(base));
          tw[0] = x.x; tw[1] = x.y;
        } else {
          tw[0] = __ldg(base);
        }
      }
#pragma unroll
      for (int j = 0; j < NT; j++) {
#pragma unroll
        for (int w = 0; w < (1 << s); w++) {
          const int m0 = (j << (s + 1)) | w, m1 = m0 | (1 << s);
          if (INV) bfly_inv(v[m0], v[m1], tw[j], one); else bfly_fwd(v[m0], v[m1], tw[j], one);
        }
      }
    }
#pragma unroll
    for (int m = 0; m < M; m++) p[PD::at((uint32_t)m << B)] = v[m];
  }
}

// Runs round I of the pass (ascending bit order); inverse walks I = 0..NR-1, forward NR-1..0.
template <bool INV, int K, bool LOW, bool LINE, int SC, int I>
__device__ __forceinline__ void run_round(uint32_t* sm, const FftArgs& a, uint32_t T) {
  constexpr int C0 = LOW ? 0 : SC;
  constexpr int NB = K - C0;
  constexpr int R = round_size(NB, I);
  constexpr int B = C0 + round_start(NB, I);
  const uint32_t gb = LOW ? (uint32_t)B : a.L0 + (uint32_t)(B - SC);
  fft_round<INV, K, B, R, (LOW && B == 0 && !LINE), LOW, SC>(sm, a, T, gb);
  __syncthreads();
}
template <bool INV, int K, bool LOW, bool LINE, int SC, int I, int NR>
struct Rounds {
  static __device__ __forceinline__ void run(uint32_t* sm, const FftArgs& a, uint32_t T) {
    run_round<INV, K, LOW, LINE, SC, (INV ? I : NR - 1 - I)>(sm, a, T);
    Rounds<INV, K, LOW, LINE, SC, I + 1, NR>::run(sm, a, T);
  }
};
template <bool INV, int K, bool LOW, bool LINE, int SC, int NR>
struct Rounds<INV, K, LOW, LINE, SC, NR, NR> {
  static __device__ __forceinline__ void run(uint32_t*, const FftArgs&, uint32_t) {}
};

// LINE: every layer is a line layer (layer l uses the 2^(n-1-l) twiddles at twend - 2^(n-l), l = 0 included).  That is the
// transform of a column whose evaluations repeat each value 2^r times, restricted to its 2^n distinct values: the first r
// layers of the circle transform of log n+r only scale (inverse) or replicate (forward), see launch_interpolate_repeated.
// (256, 4): 64 registers, four CTAs per SM; the forward low pass otherwise takes 80 and runs three (LDE 6-7 % slower, measured)
// K = 15: the one strided pass that finishes a transform of log 24..26 after the 13-layer low pass (two HBM round trips instead
// of three): 2^(15-SC) rows x 2^SC contiguous words = 128 KB of shared memory, one CTA of 1024 threads per SM (the same 32 warps
// per SM as four 256-thread CTAs).  Rows are only 16..64 bytes long; the rest of each DRAM sector is consumed by the CTA that
// owns the neighbouring tile, which runs at the same time, so the sector is served from L2.
template <bool INV, int K, bool LOW, bool LINE = false, int SC = STRIDED_C>
__global__ void __launch_bounds__(K >= 14 ? 1024 : 256, K >= 14 ? 1 : 4) fft_kernel(FftArgs a) {
  extern __shared__ __align__(16) uint32_t sm[];
  constexpr int C = LOW ? K : SC;
  const uint32_t L0 = LOW ? (uint32_t)K : a.L0;
  const uint32_t tile = blockIdx.x;
  const uint32_t nlow = L0 - C;
  const uint32_t T = tile >> nlow;
  const uint32_t gbase = ((tile & ((1u << nlow) - 1u)) << C) | (T << (L0 + K - C));
  const uint32_t* __restrict__ src = a.src[blockIdx.y];
  uint32_t* __restrict__ dst = a.dst[blockIdx.y];
  const uint32_t smask = (a.src_log >= 32) ? 0xffffffffu : ((1u << a.src_log) - 1u);
  constexpr uint32_t cm = (1u << C) - 1u;

  typedef Pad<LOW, K, SC> PD;
  for (uint32_t li = threadIdx.x * 4; li < (1u << K); li += blockDim.x * 4) {
    uint32_t g = (gbase | (li & cm) | ((li >> C) << L0)) & smask;
    uint4 x = __ldg(reinterpret_cast<const uint4*>(src + g));
    uint32_t o = PD::at(li);
    if (PD::PW >= 2) *reinterpret_cast<uint4*>(sm + o) = x;   // padding in multiples of four words: the quad stays aligned
    else { sm[o] = x.x; sm[o + 1] = x.y; sm[o + 2] = x.z; sm[o + 3] = x.w; }
  }
  __syncthreads();

  Rounds<INV, K, LOW, LINE, SC, 0, n_rounds(LOW ? K : K - SC)>::run(sm, a, T);

  const uint32_t scale = a.scale, scale2 = a.scale << 1;
  for (uint32_t li = threadIdx.x * 4; li < (1u << K); li += blockDim.x * 4) {
    uint32_t g = gbase | (li & cm) | ((li >> C) << L0);
    uint32_t o = PD::at(li);
    uint4 x = PD::PW >= 2 ? *reinterpret_cast<const uint4*>(sm + o) : make_uint4(sm[o], sm[o + 1], sm[o + 2], sm[o + 3]);
    if (scale != 1u) { x.x = mulred(x.x, scale2); x.y = mulred(x.y, scale2); x.z = mulred(x.z, scale2); x.w = mulred(x.w, scale2); }
    *reinterpret_cast<uint4*>(dst + g) = x;
  }
}

// ---------------------------------------------------------------- host-side pass planner
static const uint32_t KMAX = 13;          // 2^13 words (+pad) = 33 KB smem per CTA
static const uint32_t KSTRIDE_MAX = 9;    // layers per strided pass (tile 2^(9+4))

struct PassDesc { uint32_t K, L0; bool low; uint32_t sc; };

static int plan_passes(uint32_t n, PassDesc* out) {  // ascending layer order
  int np = 0;
  uint32_t K0 = n < KMAX ? n : KMAX;
  out[np++] = {K0, K0, true, 0};
  uint32_t rem = n - K0;
  // Measured on B200 (tools/fft_bench.py, profiles/r2_fft_two_pass_ab.json): the single 2^15-word strided pass LOSES to two
  // 2^10..2^11-word passes — log 25: interpolate 1.16 vs 1.06 ms, LDE 2.94 vs 2.26 ms — because one 1024-thread CTA per SM
  // serialises its load, butterfly and store phases where four small CTAs overlap them, and the transform is bound by integer
  // issue (about 5 instructions per element and layer), not by the third HBM round trip.  Kept behind SC_FFT_TWO_PASS=1.
  static const bool three_pass = getenv("SC_FFT_TWO_PASS") == nullptr;
  if (rem == KSTRIDE_MAX + 1) {
    // ten layers left (log 23): one strided pass of 2^10 rows x 8 words instead of two passes of 16-word rows
    out[np++] = {rem + 3, K0, false, 3};
  } else if (rem >= 11 && rem <= 13 && !three_pass) {
    // log 24..26: ONE strided pass over a 2^15-word tile (2^rem rows x 2^(15-rem) words) instead of two
    out[np++] = {15, K0, false, 15 - rem};
  } else if (rem) {
    uint32_t ns = (rem + KSTRIDE_MAX - 1) / KSTRIDE_MAX;
    uint32_t L = K0;
    for (uint32_t i = 0; i < ns; i++) {
      uint32_t k = rem / ns + (i < rem % ns ? 1 : 0);
      out[np++] = {k + STRIDED_C, L, false, (uint32_t)STRIDED_C};
      L += k;
    }
  }
  return np;
}

static uint32_t threads_for(uint32_t K) {
  if (K >= 14) return 1024u;
  uint32_t t = K >= 12 ? 256u : (K >= 4 ? (1u << (K - 4)) : 1u);
  if (t < 32) t = 32;
  return t;
}

template <bool INV, int K, bool LOW, bool LINE = false, int SC = STRIDED_C>
static int launch_one(const FftArgs& a, dim3 grid, cudaStream_t st) {
  size_t smem = ((size_t)Pad<LOW, K, SC>::words + 4) * 4;
  if (K >= 14) {   // above the 48 KB default: opt in once per instantiation
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(fft_kernel<INV, K, LOW, LINE, SC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
      configured = true;
    }
  }
  fft_kernel<INV, K, LOW, LINE, SC><<<grid, threads_for(K), smem, st>>>(a); g_launch_count++;
  return (int)cudaGetLastError();
}

template <bool INV>
static int run_pass(const PassDesc& d, const uint32_t* const* src, uint32_t* const* dst, uint32_t ncols, uint32_t n,
                    uint32_t src_log, const uint32_t* twend, uint32_t scale, cudaStream_t st, bool line = false) {
  FftArgs a{src, dst, twend, n, src_log, d.L0, scale, FK_HOST};
  dim3 grid(1u << (n - d.K), ncols);
  if (d.low && line) {
    switch (d.K) {
      case 3: return launch_one<INV, 3, true, true>(a, grid, st);
      case 4: return launch_one<INV, 4, true, true>(a, grid, st);
      case 5: return launch_one<INV, 5, true, true>(a, grid, st);
      case 6: return launch_one<INV, 6, true, true>(a, grid, st);
      case 7: return launch_one<INV, 7, true, true>(a, grid, st);
      case 8: return launch_one<INV, 8, true, true>(a, grid, st);
      case 9: return launch_one<INV, 9, true, true>(a, grid, st);
      case 10: return launch_one<INV, 10, true, true>(a, grid, st);
      case 11: return launch_one<INV, 11, true, true>(a, grid, st);
      case 12: return launch_one<INV, 12, true, true>(a, grid, st);
      case 13: return launch_one<INV, 13, true, true>(a, grid, st);
    }
    return -1;
  }
  if (!d.low && d.K == 15) {
    switch (d.sc) {
      case 2: return launch_one<INV, 15, false, false, 2>(a, grid, st);
      case 3: return launch_one<INV, 15, false, false, 3>(a, grid, st);
      case 4: return launch_one<INV, 15, false, false, 4>(a, grid, st);
    }
    return -1;
  }
  if (!d.low && d.sc == 3) return d.K == 13 ? launch_one<INV, 13, false, false, 3>(a, grid, st) : -1;
  if (d.low) {
    switch (d.K) {
      case 3: return launch_one<INV, 3, true>(a, grid, st);
      case 4: return launch_one<INV, 4, true>(a, grid, st);
      case 5: return launch_one<INV, 5, true>(a, grid, st);
      case 6: return launch_one<INV, 6, true>(a, grid, st);
      case 7: return launch_one<INV, 7, true>(a, grid, st);
      case 8: return launch_one<INV, 8, true>(a, grid, st);
      case 9: return launch_one<INV, 9, true>(a, grid, st);
      case 10: return launch_one<INV, 10, true>(a, grid, st);
      case 11: return launch_one<INV, 11, true>(a, grid, st);
      case 12: return launch_one<INV, 12, true>(a, grid, st);
      case 13: return launch_one<INV, 13, true>(a, grid, st);
    }
  } else {
    switch (d.K) {
      case 5: return launch_one<INV, 5, false>(a, grid, st);
      case 6: return launch_one<INV, 6, false>(a, grid, st);
      case 7: return launch_one<INV, 7, false>(a, grid, st);
      case 8: return launch_one<INV, 8, false>(a, grid, st);
      case 9: return launch_one<INV, 9, false>(a, grid, st);
      case 10: return launch_one<INV, 10, false>(a, grid, st);
      case 11: return launch_one<INV, 11, false>(a, grid, st);
      case 12: return launch_one<INV, 12, false>(a, grid, st);
      case 13: return launch_one<INV, 13, false>(a, grid, st);
    }
  }
  return -1;
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

// coeffs (log src_log) -> evaluations on the canonic domain of log n = src_log + log_blowup (<= 1 here), out of place.
int launch_evaluate(const uint32_t* const* coeffs, uint32_t* const* out, uint32_t ncols, uint32_t src_log, uint32_t n,
                    const uint32_t* tw_end, cudaStream_t st) {
  if (n < 3 || ncols == 0) return n < 3 ? -1 : 0;
  if (n - src_log > 1) return -1;  // larger blow-ups are zero-extended by the caller (capi.cu)
  PassDesc pd[8];
  int np = plan_passes(n, pd);
  for (int i = np - 1; i >= 0; i--) {
    bool first = (i == np - 1);
    int e = run_pass<false>(pd[i], first ? coeffs : (const uint32_t* const*)out, out, ncols, n, first ? src_log : 32, tw_end, 1u, st);
    if (e) return e;
  }
  return 0;
}

// ---------------------------------------------------------------- lane-repeated columns (line transforms)
// log size < 3: one thread per column does the few butterflies serially
template <bool INV>
__global__ void line_fft_small_kernel(FftArgs a, uint32_t ncols) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncols) return;
  const uint32_t n = a.n, size = 1u << n;
  const uint32_t smask = (a.src_log >= 32) ? 0xffffffffu : ((1u << a.src_log) - 1u);
  const FK fk = a.k;
  uint32_t v[4];
  for (uint32_t i = 0; i < size; i++) v[i] = a.src[c][i & smask];
  for (uint32_t ss = 0; ss < n; ss++) {
    const uint32_t s = INV ? ss : n - 1 - ss;
    if (!INV && s >= a.src_log) continue;
    const uint32_t* tw = a.twend - ((size_t)1 << (n - s));
    for (uint32_t i = 0; i < size; i++) {
      if ((i >> s) & 1u) continue;
      if (INV) bfly_inv(v[i], v[i + (1u << s)], tw[i >> (s + 1)], fk); else bfly_fwd(v[i], v[i + (1u << s)], tw[i >> (s + 1)], fk);
    }
  }
  for (uint32_t i = 0; i < size; i++) a.dst[c][i] = a.scale != 1u ? m_mul(v[i], a.scale) : v[i];
}

// src -> cols (may be the same arrays): the 2^n distinct values of each column -> the non-zero coefficients of the polynomial interpolating the
// column with every value repeated 2^r times (any r: the r skipped layers contribute the factor 2^r that turns the
// 1/2^(n+r) normalisation into 1/2^n).  Coefficient i of the result is coefficient i << r of the full vector.
int launch_interpolate_repeated(const uint32_t* const* src, uint32_t* const* cols, uint32_t ncols, uint32_t n, const uint32_t* itw_end,
                                cudaStream_t st) {
  if (ncols == 0) return 0;
  uint32_t ninv = m_inv(m_pow(2, n));
  if (n < 3) {
    FftArgs a{src, cols, itw_end, n, 32, 0, ninv, FK_HOST};
    line_fft_small_kernel<true><<<(ncols + 63) / 64, 64, 0, st>>>(a, ncols); g_launch_count++;
    return (int)cudaGetLastError();
  }
  PassDesc pd[8];
  int np = plan_passes(n, pd);
  for (int i = 0; i < np; i++) {
    int e = run_pass<true>(pd[i], i == 0 ? src : (const uint32_t* const*)cols, cols, ncols, n, 32, itw_end, i == np - 1 ? ninv : 1u, st, true);
    if (e) return e;
  }
  return 0;
}

// Compact coefficients (log src_log) -> the 2^n distinct evaluations (n = src_log + log_blowup, log_blowup <= 1) of the
// repeated column on the larger domain; the caller replicates each 2^r times.
int launch_evaluate_repeated(const uint32_t* const* coeffs, uint32_t* const* out, uint32_t ncols, uint32_t src_log, uint32_t n,
                             const uint32_t* tw_end, cudaStream_t st) {
  if (ncols == 0) return 0;
  if (n - src_log > 1) return -1;
  if (n < 3) {
    FftArgs a{coeffs, out, tw_end, n, src_log, 0, 1u, FK_HOST};
    line_fft_small_kernel<false><<<(ncols + 63) / 64, 64, 0, st>>>(a, ncols); g_launch_count++;
    return (int)cudaGetLastError();
  }
  PassDesc pd[8];
  int np = plan_passes(n, pd);
  for (int i = np - 1; i >= 0; i--) {
    bool first = (i == np - 1);
    int e = run_pass<false>(pd[i], first ? coeffs : (const uint32_t* const*)out, out, ncols, n, first ? src_log : 32, tw_end, 1u, st, true);
    if (e) return e;
  }
  return 0;
}

}  // namespace sb
