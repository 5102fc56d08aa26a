// Integer-pipe micro-benchmark for sm_100a: the measured denominator of the Blake2s (Merkle) roofline.
//
// BASELINE.md §3.3 / SURVEY.md §8(d): "INT32 ALU pipe (≈ #SM × 128 lanes × clock; measure)".  Each kernel below issues a
// long stream of ONE instruction class from eight independent register chains per thread (so neither latency nor
// occupancy limits it) and reports lane-operations per clock per SM from the SM's own cycle counter:
//   kind 0  LOP3   (xor3)                 ALU pipe        kind 3  IADD3 (add.u32 x2 -> one IADD3)   ALU pipe
//   kind 1  SHF    (rotate by 12)         ALU pipe        kind 4  IMAD  (mad.lo.u32)                FMA pipe
//   kind 2  PRMT   (rotate by 16)         ALU pipe        kind 5  ALU+IMAD 1:1 interleaved          both pipes
//   kind 6  the Blake2s G mix exactly as merkle.cu issues it: per G 4 LOP3 + 2 SHF + 2 PRMT (ALU) and 6 IMAD (FMA, the
//           multiplicand in a uniform register), four independent G columns per thread, i.e. the compression without its
//           loads and stores.
// The instruction classes are fixed with inline PTX; `cuobjdump -sass` of this file shows the SASS each one became
// (profiles/r2_microbench_sass.txt).  Not on the proving path.
#include "kernels.cuh"

namespace sb {

template <int KIND>
__global__ void __launch_bounds__(1024) int_pipe_kernel(uint32_t iters, uint32_t seed, uint32_t one, uint32_t* sink,
                                                        unsigned long long* cycles) {
  uint32_t x0 = seed + threadIdx.x, x1 = x0 * 3u + 1u, x2 = x0 * 5u + 2u, x3 = x0 * 7u + 3u;
  uint32_t x4 = x0 * 11u + 4u, x5 = x0 * 13u + 5u, x6 = x0 * 17u + 6u, x7 = x0 * 19u + 7u;
  const uint32_t k0 = seed ^ 0x9E3779B9u, k1 = seed * 0x85EBCA6Bu + 1u;
  uint32_t y0 = x0 ^ 0x243F6A88u, y1 = x1 ^ 0x85A308D3u, y2 = x2 ^ 0x13198A2Eu, y3 = x3 ^ 0x03707344u;   // KIND 6: four G columns
  uint32_t y4 = x4 ^ 0xA4093822u, y5 = x5 ^ 0x299F31D0u, y6 = x6 ^ 0x082EFA98u, y7 = x7 ^ 0xEC4E6C89u;
  const uint32_t m0 = k0 + threadIdx.x, m1 = k1 ^ threadIdx.x;                                          // message words: vector registers
  __syncthreads();
  const long long t0 = clock64();
#define OP8(INS)                                                                                      \
  INS(x0) INS(x1) INS(x2) INS(x3) INS(x4) INS(x5) INS(x6) INS(x7)
#define I_LOP3(x) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(k0), "r"(k1));
#define I_SHF(x) asm volatile("shf.r.wrap.b32 %0, %0, %0, 12;" : "+r"(x));
#define I_PRMT(x) asm volatile("prmt.b32 %0, %0, %0, 0x1032;" : "+r"(x));
#define I_IADD3(x) asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(x) : "r"(k0), "r"(k1));
#define I_IMAD(x) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(k0), "r"(k1));
  for (uint32_t it = 0; it < iters; it++) {
    if (KIND == 0) { OP8(I_LOP3) OP8(I_LOP3) OP8(I_LOP3) OP8(I_LOP3) }
    if (KIND == 1) { OP8(I_SHF) OP8(I_SHF) OP8(I_SHF) OP8(I_SHF) }
    if (KIND == 2) { OP8(I_PRMT) OP8(I_PRMT) OP8(I_PRMT) OP8(I_PRMT) }
    if (KIND == 3) { OP8(I_IADD3) OP8(I_IADD3) OP8(I_IADD3) OP8(I_IADD3) }
    if (KIND == 4) { OP8(I_IMAD) OP8(I_IMAD) OP8(I_IMAD) OP8(I_IMAD) }
    if (KIND == 5) {
      I_LOP3(x0) I_IMAD(x1) I_LOP3(x2) I_IMAD(x3) I_LOP3(x4) I_IMAD(x5) I_LOP3(x6) I_IMAD(x7)
      I_IMAD(x0) I_LOP3(x1) I_IMAD(x2) I_LOP3(x3) I_IMAD(x4) I_LOP3(x5) I_IMAD(x6) I_LOP3(x7)
      I_LOP3(x0) I_IMAD(x1) I_LOP3(x2) I_IMAD(x3) I_LOP3(x4) I_IMAD(x5) I_LOP3(x6) I_IMAD(x7)
      I_IMAD(x0) I_LOP3(x1) I_IMAD(x2) I_LOP3(x3) I_IMAD(x4) I_LOP3(x5) I_IMAD(x6) I_LOP3(x7)
    }
    if (KIND == 6) {
      // four G columns on (x0..x3), (x4..x7), (y0..y3), (y4..y7) as (a, b, c, d), message words in vector registers, the
      // IMAD multiplicand `one` in a UNIFORM register (a kernel argument, never paired with an immediate addend): the form
      // blake2s.cuh compiles to.  With `one` in a vector register (three vector operands per IMAD, the round-1 form) the
      // same mix reaches 88 % of the ALU pipe instead of 96 %.
#define FADD(r, p, q) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(p), "r"(one), "r"(q));
#define XROT_P(r, p, q, sel) asm volatile("{ .reg .u32 t; xor.b32 t, %1, %2; prmt.b32 %0, t, t, " #sel "; }" : "=r"(r) : "r"(p), "r"(q));
#define XROT_S(r, p, q, n) asm volatile("{ .reg .u32 t; xor.b32 t, %1, %2; shf.r.wrap.b32 %0, t, t, " #n "; }" : "=r"(r) : "r"(p), "r"(q));
#define GMIX(a, b, c, d)                                                       \
  FADD(a, b, a) FADD(a, m0, a) XROT_P(d, d, a, 0x1032) FADD(c, d, c) XROT_S(b, b, c, 12) \
  FADD(a, b, a) FADD(a, m1, a) XROT_P(d, d, a, 0x0321) FADD(c, d, c) XROT_S(b, b, c, 7)
      GMIX(x0, x1, x2, x3) GMIX(x4, x5, x6, x7) GMIX(y0, y1, y2, y3) GMIX(y4, y5, y6, y7)
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  uint32_t r = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
  if (KIND == 6) r ^= y0 ^ y1 ^ y2 ^ y3 ^ y4 ^ y5 ^ y6 ^ y7;
  if (r == 0x12345678u) sink[0] = r;  // keeps the chains alive; practically never taken
}

// ops issued per thread and iteration (lane-operations), ALU-pipe and FMA-pipe instructions separately
static void ops_per_iter(int kind, uint32_t* alu, uint32_t* fma) {
  *alu = *fma = 0;
  if (kind <= 3) *alu = 32;
  else if (kind == 4) *fma = 32;
  else if (kind == 5) { *alu = 16; *fma = 16; }
  else { *alu = 4 * 8; *fma = 4 * 6; }   // four G per iteration
}

// Runs `kind` on every SM at full occupancy (2 CTAs x 1024 threads per SM, one wave).  out[0] = ALU-pipe lane-ops per clock
// per SM, out[1] = FMA-pipe lane-ops per clock per SM (both from the per-CTA cycle counters), out[2] = kernel milliseconds
// (CUDA events), out[3] = number of SMs.  d_scratch needs 8 * ctas + 4 bytes.
int launch_int_pipe_bench(int kind, uint32_t iters, int n_sm, void* d_scratch, double out[4], cudaStream_t st) {
  const int ctas = 2 * n_sm;
  unsigned long long* d_cycles = reinterpret_cast<unsigned long long*>(d_scratch);
  uint32_t* d_sink = reinterpret_cast<uint32_t*>(d_cycles + ctas);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; rep++) {  // the first run warms the instruction cache and the clocks
    cudaEventRecord(e0, st);
    switch (kind) {
      case 0: int_pipe_kernel<0><<<ctas, 1024, 0, st>>>(iters, 12345u, 1u, d_sink, d_cycles); break;
      case 1: int_pipe_kernel<1><<<ctas, 1024, 0, st>>>(iters, 12345u, 1u, d_sink, d_cycles); break;
      case 2: int_pipe_kernel<2><<<ctas, 1024, 0, st>>>(iters, 12345u, 1u, d_sink, d_cycles); break;
      case 3: int_pipe_kernel<3><<<ctas, 1024, 0, st>>>(iters, 12345u, 1u, d_sink, d_cycles); break;
      case 4: int_pipe_kernel<4><<<ctas, 1024, 0, st>>>(iters, 12345u, 1u, d_sink, d_cycles); break;
      case 5: int_pipe_kernel<5><<<ctas, 1024, 0, st>>>(iters, 12345u, 1u, d_sink, d_cycles); break;
      case 6: int_pipe_kernel<6><<<ctas, 1024, 0, st>>>(iters, 12345u, 1u, d_sink, d_cycles); break;
      default: cudaEventDestroy(e0); cudaEventDestroy(e1); return -1;
    }
    g_launch_count++;
    cudaEventRecord(e1, st);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return (int)e;
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  unsigned long long* h = new unsigned long long[ctas];
  e = cudaMemcpy(h, d_cycles, 8 * (size_t)ctas, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { delete[] h; return (int)e; }
  double mean = 0;
  for (int i = 0; i < ctas; i++) mean += (double)h[i];
  mean /= ctas;
  delete[] h;
  uint32_t alu, fma;
  ops_per_iter(kind, &alu, &fma);
  // two co-resident CTAs of 1024 threads per SM run for `mean` cycles each
  out[0] = 2.0 * 1024.0 * iters * alu / mean;
  out[1] = 2.0 * 1024.0 * iters * fma / mean;
  out[2] = ms;
  out[3] = n_sm;
  return 0;
}

}  // namespace sb
