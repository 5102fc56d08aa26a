// Blake2s compression for sm_100a as the Merkle hasher uses it (state h, 16 message words, zero counters and flags):
// fully unrolled with compile-time sigma so state and message stay in registers.  Shared by merkle.cu and fri.cu.
#pragma once
#include <cstdint>

namespace sb {

__device__ __forceinline__ uint32_t rotr16(uint32_t x) { return __byte_perm(x, x, 0x1032); }
__device__ __forceinline__ uint32_t rotr8(uint32_t x) { return __byte_perm(x, x, 0x0321); }
__device__ __forceinline__ uint32_t rotr12(uint32_t x) { return __funnelshift_r(x, x, 12); }
__device__ __forceinline__ uint32_t rotr7(uint32_t x) { return __funnelshift_r(x, x, 7); }

// Pipe balance (ncu, round 1): with plain adds the compression is 12 ALU-pipe ops per G (IADD3, LOP3, SHF, PRMT) and the
// ALU pipe sits at 93-97 % while the FMA pipe idles at 10 %.  The additions are therefore issued as IMAD (x*1+y, `one` is a
// runtime 1 so ptxas keeps the multiply): 8 ALU + 6 FMA ops per G, which lowers the pipe bound from 12/16 to 8/16 cycles.
__device__ __forceinline__ uint32_t fadd(uint32_t x, uint32_t y, uint32_t one) {
  uint32_t r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(one), "r"(y));
  return r;
}
// Tried and rejected (tools/merkle_bench.py, B200): rotating by 16 on the FMA pipe as x*2^16 -> lo+hi (IMAD.WIDE + IMAD) to take
// one of the eight ALU ops per G off the ALU pipe: 17.8-20.0 G compressions/s against 19.8-22.3 with the PRMT below.
// NZ: message words NZ..15 are known to be zero (a node that injects at most NZ column values): their additions vanish at
// compile time — 12 of the 16 words for the four-column layers (FRI, composition, quotients) that make up most of a proof.
#define B2S_G(a, b, c, d, ix, iy)            \
  do {                                       \
    a = fadd(b, a, one);                     \
    if ((ix) < NZ) a = fadd(m[ix], a, one);  \
    d = rotr16(d ^ a);                       \
    c = fadd(d, c, one);                     \
    b = rotr12(b ^ c);                       \
    a = fadd(b, a, one);                     \
    if ((iy) < NZ) a = fadd(m[iy], a, one);  \
    d = rotr8(d ^ a);                        \
    c = fadd(d, c, one);                     \
    b = rotr7(b ^ c);                        \
  } while (0)

// The first column step of round 1 (B2S_G0): c is still an IV constant at its first update, and an IMAD cannot take both a
// uniform-register multiplicand and an immediate addend — with fadd there ptxas keeps `one` in a VECTOR register for the
// whole kernel and every IMAD reads three vector registers.  With a plain add for those four instructions `one` stays in a
// uniform register (IMAD R, R, UR, R): tools/exp/gmix.cu measured the G mix at 96.5 % of the ALU pipe in that form against
// 88 % with three vector-register operands.
#define B2S_G0(a, b, c, d, ix, iy)            \
  do {                                       \
    a = fadd(b, a, one);                     \
    if ((ix) < NZ) a = fadd(m[ix], a, one);  \
    d = rotr16(d ^ a);                       \
    c = d + c;                               \
    b = rotr12(b ^ c);                       \
    a = fadd(b, a, one);                     \
    if ((iy) < NZ) a = fadd(m[iy], a, one);  \
    d = rotr8(d ^ a);                        \
    c = fadd(d, c, one);                     \
    b = rotr7(b ^ c);                        \
  } while (0)

#define B2S_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
  B2S_G(v0, v4, v8, v12, s0, s1);                                                        \
  B2S_G(v1, v5, v9, v13, s2, s3);                                                        \
  B2S_G(v2, v6, v10, v14, s4, s5);                                                       \
  B2S_G(v3, v7, v11, v15, s6, s7);                                                       \
  B2S_G(v0, v5, v10, v15, s8, s9);                                                       \
  B2S_G(v1, v6, v11, v12, s10, s11);                                                     \
  B2S_G(v2, v7, v8, v13, s12, s13);                                                      \
  B2S_G(v3, v4, v9, v14, s14, s15);

#define B2S_ROUND0(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
  B2S_G0(v0, v4, v8, v12, s0, s1);                                                        \
  B2S_G0(v1, v5, v9, v13, s2, s3);                                                        \
  B2S_G0(v2, v6, v10, v14, s4, s5);                                                       \
  B2S_G0(v3, v7, v11, v15, s6, s7);                                                       \
  B2S_G(v0, v5, v10, v15, s8, s9);                                                       \
  B2S_G(v1, v6, v11, v12, s10, s11);                                                     \
  B2S_G(v2, v7, v8, v13, s12, s13);                                                      \
  B2S_G(v3, v4, v9, v14, s14, s15);

// h <- F(h, m, 0, 0, 0, 0)
// t0 / f0: byte counter and final-block flag of the real Blake2s hash (the channel); zero — and folded away — for Merkle nodes
template <int NZ = 16>
__device__ __forceinline__ void b2s_compress(uint32_t h[8], const uint32_t m[16], uint32_t one, uint32_t t0 = 0, uint32_t f0 = 0) {
  uint32_t v0 = h[0], v1 = h[1], v2 = h[2], v3 = h[3], v4 = h[4], v5 = h[5], v6 = h[6], v7 = h[7];
  uint32_t v8 = 0x6A09E667u, v9 = 0xBB67AE85u, v10 = 0x3C6EF372u, v11 = 0xA54FF53Au;
  uint32_t v12 = 0x510E527Fu ^ t0, v13 = 0x9B05688Cu, v14 = 0x1F83D9ABu ^ f0, v15 = 0x5BE0CD19u;
  B2S_ROUND0(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
  B2S_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3)
  B2S_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4)
  B2S_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8)
  B2S_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13)
  B2S_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9)
  B2S_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11)
  B2S_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10)
  B2S_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5)
  B2S_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0)
  h[0] ^= v0 ^ v8;  h[1] ^= v1 ^ v9;  h[2] ^= v2 ^ v10; h[3] ^= v3 ^ v11;
  h[4] ^= v4 ^ v12; h[5] ^= v5 ^ v13; h[6] ^= v6 ^ v14; h[7] ^= v7 ^ v15;
}

}  // namespace sb
