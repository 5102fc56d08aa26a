// Experiment (not on the proving path): how close can the Blake2s G instruction mix get to the ALU-pipe bound?
// Variants: independent G chains per thread (ILP 2/4/8), warps per SM sub-partition (8/12/16), instruction order
// (chain-major: one G after the other; step-major: step k of every chain, then step k+1).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/gmix tools/exp/gmix.cu ; run: build/gmix
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define FADD(r, p, q) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(p), "r"(one), "r"(q));
#define XOR_(r, p, q) asm volatile("xor.b32 %0, %1, %2;" : "=r"(r) : "r"(p), "r"(q));
#define PRMT_(r, sel) asm volatile("prmt.b32 %0, %0, %0, " #sel ";" : "+r"(r));
#define SHF_(r, n) asm volatile("shf.r.wrap.b32 %0, %0, %0, " #n ";" : "+r"(r));
// the 14 steps of a G on chain j = (x[4j], x[4j+1], x[4j+2], x[4j+3]) as (a,b,c,d)
#define A(j) x[4*(j)]
#define B(j) x[4*(j)+1]
#define C(j) x[4*(j)+2]
#define D(j) x[4*(j)+3]
#define S0(j) FADD(A(j), B(j), A(j))
#define S1(j) FADD(A(j), k0, A(j))
#define S2(j) XOR_(D(j), D(j), A(j))
#define S3(j) PRMT_(D(j), 0x1032)
#define S4(j) FADD(C(j), D(j), C(j))
#define S5(j) XOR_(B(j), B(j), C(j))
#define S6(j) SHF_(B(j), 12)
#define S7(j) FADD(A(j), B(j), A(j))
#define S8(j) FADD(A(j), k1, A(j))
#define S9(j) XOR_(D(j), D(j), A(j))
#define S10(j) PRMT_(D(j), 0x0321)
#define S11(j) FADD(C(j), D(j), C(j))
#define S12(j) XOR_(B(j), B(j), C(j))
#define S13(j) SHF_(B(j), 7)
#define GALL(j) S0(j) S1(j) S2(j) S3(j) S4(j) S5(j) S6(j) S7(j) S8(j) S9(j) S10(j) S11(j) S12(j) S13(j)

// NCH chains, ORDER 0 = chain-major, 1 = step-major, 2 = skewed (chain j runs step k-j at position k)
template <int NCH, int ORDER>
__global__ void __launch_bounds__(512) gmix_kernel(uint32_t iters, uint32_t seed, uint32_t one, uint32_t* sink, unsigned long long* cycles) {
  uint32_t x[32];
#pragma unroll
  for (int i = 0; i < 32; i++) x[i] = seed * (2 * i + 1) + threadIdx.x * (i + 3);
  const uint32_t k0 = seed ^ 0x9E3779B9u, k1 = seed * 0x85EBCA6Bu + 1u;
  __syncthreads();
  const long long t0 = clock64();
  for (uint32_t it = 0; it < iters; it++) {
#define CH(j) j
    if (ORDER == 0) {
      GALL(CH(0)) GALL(CH(1))
      if (NCH >= 4) { GALL(CH(2)) GALL(CH(3)) }
      if (NCH >= 8) { GALL(CH(4)) GALL(CH(5)) GALL(CH(6)) GALL(CH(7)) }
    } else if (ORDER == 1) {
#define STEP(S) S(CH(0)) S(CH(1)) if (NCH >= 4) { S(CH(2)) S(CH(3)) } if (NCH >= 8) { S(CH(4)) S(CH(5)) S(CH(6)) S(CH(7)) }
      STEP(S0) STEP(S1) STEP(S2) STEP(S3) STEP(S4) STEP(S5) STEP(S6) STEP(S7) STEP(S8) STEP(S9) STEP(S10) STEP(S11) STEP(S12) STEP(S13)
    } else {
      // skewed by pairs: chains (0,1) lead chains (2,3) by 3 steps (so that ALU and FMA steps meet), 4 chains only
#define P01(S) S(CH(0)) S(CH(1))
#define P23(S) S(CH(2)) S(CH(3))
      P01(S0) P23(S11) P01(S1) P23(S12) P01(S2) P23(S13) P01(S3) P23(S0) P01(S4) P23(S1) P01(S5) P23(S2) P01(S6) P23(S3)
      P01(S7) P23(S4) P01(S8) P23(S5) P01(S9) P23(S6) P01(S10) P23(S7) P01(S11) P23(S8) P01(S12) P23(S9) P01(S13) P23(S10)
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < 4 * NCH; i++) r ^= x[i];
  if (r == 0x12345678u) sink[0] = r;
}

template <int NCH, int ORDER>
static void run(const char* name, int n_sm, int ctas_per_sm, int threads, uint32_t* d_sink, unsigned long long* d_cyc) {
  const uint32_t iters = 4096;
  const int ctas = n_sm * ctas_per_sm;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; rep++) { cudaEventRecord(e0); gmix_kernel<NCH, ORDER><<<ctas, threads>>>(iters, 12345u, 1u, d_sink, d_cyc); cudaEventRecord(e1); }
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  static unsigned long long h[4096];
  cudaMemcpy(h, d_cyc, 8 * ctas, cudaMemcpyDeviceToHost);
  double mean = 0;   // the LONGEST CTA: warps are not served fairly, the mean over CTAs under-counts
  for (int i = 0; i < ctas; i++) if ((double)h[i] > mean) mean = (double)h[i];
  const double tops = (double)ctas * threads * iters * NCH * 8 / (ms * 1e-3) / 1e12;
  const double alu = (double)ctas_per_sm * threads * iters * NCH * 8 / mean;   // ALU lane-ops per clk per SM
  printf("{\"variant\": \"%s\", \"chains\": %d, \"order\": %d, \"warps_per_smsp\": %d, \"alu_lanes_per_clk_sm\": %.2f, \"frac_of_64\": %.4f, \"alu_Tops_events\": %.3f, \"ms\": %.3f}\n",
         name, NCH, ORDER, ctas_per_sm * threads / 128, alu, alu / 64.0, tops, ms);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int n_sm = p.multiProcessorCount;
  uint32_t* d_sink; unsigned long long* d_cyc;
  cudaMalloc(&d_sink, 64); cudaMalloc(&d_cyc, 8 * 4096);
  // warps per SMSP = ctas_per_sm * threads / 128
  run<2, 0>("ilp2 chain-major", n_sm, 4, 512, d_sink, d_cyc);
  run<2, 0>("ilp2 chain-major", n_sm, 3, 512, d_sink, d_cyc);
  run<2, 0>("ilp2 chain-major", n_sm, 2, 512, d_sink, d_cyc);
  run<2, 1>("ilp2 step-major", n_sm, 4, 512, d_sink, d_cyc);
  run<4, 0>("ilp4 chain-major", n_sm, 4, 512, d_sink, d_cyc);
  run<4, 0>("ilp4 chain-major", n_sm, 3, 512, d_sink, d_cyc);
  run<4, 0>("ilp4 chain-major", n_sm, 2, 512, d_sink, d_cyc);
  run<4, 0>("ilp4 chain-major", n_sm, 1, 512, d_sink, d_cyc);
  run<4, 1>("ilp4 step-major", n_sm, 4, 512, d_sink, d_cyc);
  run<4, 1>("ilp4 step-major", n_sm, 3, 512, d_sink, d_cyc);
  run<4, 1>("ilp4 step-major", n_sm, 2, 512, d_sink, d_cyc);
  run<4, 1>("ilp4 step-major", n_sm, 1, 512, d_sink, d_cyc);
  run<4, 2>("ilp4 skewed", n_sm, 4, 512, d_sink, d_cyc);
  run<4, 2>("ilp4 skewed", n_sm, 3, 512, d_sink, d_cyc);
  run<4, 2>("ilp4 skewed", n_sm, 2, 512, d_sink, d_cyc);
  run<4, 2>("ilp4 skewed", n_sm, 1, 512, d_sink, d_cyc);
  run<8, 1>("ilp8 step-major", n_sm, 2, 512, d_sink, d_cyc);
  run<8, 1>("ilp8 step-major", n_sm, 1, 512, d_sink, d_cyc);
  run<8, 0>("ilp8 chain-major", n_sm, 2, 512, d_sink, d_cyc);
  return 0;
}
