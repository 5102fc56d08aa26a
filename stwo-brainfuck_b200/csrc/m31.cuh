// M31 / CM31 / QM31 arithmetic in 32-bit lanes, host + device.
// Definitions follow stwo-prover 0.1.1 @ 31e8dbc core/fields/{m31,cm31,qm31}.rs:
//   P = 2^31-1, CM31 = M31[i]/(i^2+1), QM31 = CM31[u]/(u^2-(2+i)).
// All values are canonical (in [0,P)) at rest; kernels may keep lazy forms internally.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define SB_HD __host__ __device__ __forceinline__
#else
#define SB_HD inline
#endif

namespace sb {

constexpr uint32_t P = 0x7fffffffu;

SB_HD uint32_t m_add(uint32_t a, uint32_t b) {
  uint32_t s = a + b;
  uint32_t t = s - P;
  return t < s ? t : s;  // min(s, s-P) as unsigned
}
SB_HD uint32_t m_sub(uint32_t a, uint32_t b) {
  uint32_t d = a - b;
  uint32_t t = d + P;
  return t < d ? t : d;
}
SB_HD uint32_t m_neg(uint32_t a) { return a ? P - a : 0; }
// Partial reduce of a 62-bit product to [0, 2^32): (v >> 31) + (v & P); then one conditional subtract.
SB_HD uint32_t m_reduce64(uint64_t v) {
  uint32_t s = (uint32_t)(v >> 31) + ((uint32_t)v & P);  // <= 2P
  uint32_t t = s - P;
  return t < s ? t : s;  // note: s == 2P -> t == P -> handled below
}
// Device form (sm_100a, 4 instructions: SHL/IADD, IMAD.WIDE, LEA.HI, VIADDMNMX): the 64-bit product a * 2b has
// (ab >> 31) in its high word and (ab & P) << 1 in its low word, so ab = hi + (lo >> 1) (mod P), at most 2P - 2, and one
// conditional subtraction — min(s, s - P) as unsigned, a single VIADDMNMX — makes it canonical.  a, b in [0, P].
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ uint32_t m_hl(uint32_t hi, uint32_t lo) {   // hi + (lo >> 1) as LEA.HI
  uint32_t s;
  asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(s) : "r"(lo), "r"(0x80000000u), "r"(hi));
  return s;
}
#endif
SB_HD uint32_t m_mul(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  const uint64_t pr = (uint64_t)a * (b << 1);
  const uint32_t s = m_hl((uint32_t)(pr >> 32), (uint32_t)pr);
  return min(s, s - P);
#else
  uint32_t r = m_reduce64((uint64_t)a * b);
  return r == P ? 0 : r;
#endif
}
SB_HD uint32_t m_sqr(uint32_t a) { return m_mul(a, a); }
SB_HD uint32_t m_pow(uint32_t a, uint32_t e) {
  uint32_t r = 1;
  while (e) {
    if (e & 1) r = m_mul(r, a);
    a = m_sqr(a);
    e >>= 1;
  }
  return r;
}
SB_HD uint32_t m_sqn(uint32_t a, int n) {
  for (int i = 0; i < n; i++) a = m_sqr(a);
  return a;
}
// a^(P-2) = a^(2^31-3) by an addition chain: 30 squarings + 8 multiplications.
SB_HD uint32_t m_inv(uint32_t a) {
  uint32_t t2 = m_mul(m_sqr(a), a);          // 2^2-1
  uint32_t t4 = m_mul(m_sqn(t2, 2), t2);     // 2^4-1
  uint32_t t8 = m_mul(m_sqn(t4, 4), t4);     // 2^8-1
  uint32_t t16 = m_mul(m_sqn(t8, 8), t8);    // 2^16-1
  uint32_t t24 = m_mul(m_sqn(t16, 8), t8);   // 2^24-1
  uint32_t t28 = m_mul(m_sqn(t24, 4), t4);   // 2^28-1
  uint32_t t29 = m_mul(m_sqr(t28), a);       // 2^29-1
  return m_mul(m_sqn(t29, 2), a);            // 4*(2^29-1)+1 = 2^31-3
}

struct CM31 {
  uint32_t a, b;
};
SB_HD CM31 c_add(CM31 x, CM31 y) { return {m_add(x.a, y.a), m_add(x.b, y.b)}; }
SB_HD CM31 c_sub(CM31 x, CM31 y) { return {m_sub(x.a, y.a), m_sub(x.b, y.b)}; }
SB_HD CM31 c_neg(CM31 x) { return {m_neg(x.a), m_neg(x.b)}; }
// Any 64-bit value -> canonical [0,P).  Host: two folds by 2^31 == 1 (mod P) and one conditional subtraction.  Device: the
// high word folds in with weight 2^32 == 2 (one IMAD.WIDE), what is left is below 2^34: one fold and one VIADDMNMX.
SB_HD uint32_t m_red_wide(uint64_t v) {
#if defined(__CUDA_ARCH__)
  const uint64_t t = (uint64_t)(uint32_t)(v >> 32) * 2u + (uint32_t)v;   // < 3 * 2^32
  const uint32_t s = (uint32_t)(t >> 31) + ((uint32_t)t & P);            // <= P + 5
  return min(s, s - P);
#else
  v = (v >> 31) + (v & P);                               // < 2^33 + 2^31
  uint32_t s = (uint32_t)(v >> 31) + ((uint32_t)v & P);  // < P + 8
  return s >= P ? s - P : s;
#endif
}
#if defined(__CUDA_ARCH__)
// x*y2 + z*w2 for DOUBLED second factors (y2 = 2y, w2 = 2w; x, y, z, w in [0,P]): the sum is 2T with T < 2^63, its high
// word T >> 31 is below 2P, so: conditional subtraction, LEA.HI with the low word, conditional subtraction — 2 IMAD.WIDE + 3.
__device__ __forceinline__ uint32_t m_dot2(uint32_t x, uint32_t y2, uint32_t z, uint32_t w2) {
  const uint64_t pr = (uint64_t)x * y2 + (uint64_t)z * w2;
  uint32_t hi = (uint32_t)(pr >> 32);
  hi = min(hi, hi - P);
  const uint32_t s = m_hl(hi, (uint32_t)pr);
  return min(s, s - P);
}
#endif
SB_HD CM31 c_mul(CM31 x, CM31 y) {
#if defined(__CUDA_ARCH__)
  const uint32_t ya2 = y.a << 1, yb2 = y.b << 1, nyb2 = (P - y.b) << 1;
  return {m_dot2(x.a, ya2, x.b, nyb2), m_dot2(x.a, yb2, x.b, ya2)};
#else
  // (a+bi)(c+di) = (ac - bd) + (ad + bc)i; each coordinate is ONE 64-bit sum of two products (-bd as b(P-d)), reduced once
  uint64_t re = (uint64_t)x.a * y.a + (uint64_t)x.b * (P - y.b);
  uint64_t im = (uint64_t)x.a * y.b + (uint64_t)x.b * y.a;
  return {m_red_wide(re), m_red_wide(im)};
#endif
}
SB_HD CM31 c_mulm(CM31 x, uint32_t m) { return {m_mul(x.a, m), m_mul(x.b, m)}; }
SB_HD CM31 c_inv(CM31 x) {
  uint32_t n = m_inv(m_add(m_sqr(x.a), m_sqr(x.b)));
  return {m_mul(x.a, n), m_mul(m_neg(x.b), n)};
}

struct QM31 {
  CM31 a, b;
};
SB_HD QM31 q_make(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return {{a, b}, {c, d}}; }
SB_HD QM31 q_fromm(uint32_t a) { return {{a, 0}, {0, 0}}; }
SB_HD QM31 q_zero() { return {{0, 0}, {0, 0}}; }
SB_HD QM31 q_add(QM31 x, QM31 y) { return {c_add(x.a, y.a), c_add(x.b, y.b)}; }
SB_HD QM31 q_sub(QM31 x, QM31 y) { return {c_sub(x.a, y.a), c_sub(x.b, y.b)}; }
SB_HD QM31 q_neg(QM31 x) { return {c_neg(x.a), c_neg(x.b)}; }
SB_HD CM31 c_mulR(CM31 x) {  // * (2 + i)
  return {m_sub(m_add(x.a, x.a), x.b), m_add(m_add(x.b, x.b), x.a)};
}
SB_HD QM31 q_mul(QM31 x, QM31 y) {
  // x = A + Bu, y = C + Du, u^2 = 2 + i:  (AC + (2+i)BD) + (AD + BC)u.
#if defined(__CUDA_ARCH__)
  // AC and BD as two-product sums with doubled factors (m_dot2: 2 IMAD.WIDE + 3 each), AD + BC as four-product 64-bit sums
  const uint32_t c02 = y.a.a << 1, c12 = y.a.b << 1, nc12 = (P - y.a.b) << 1;
  const CM31 ac = {m_dot2(x.a.a, c02, x.a.b, nc12), m_dot2(x.a.a, c12, x.a.b, c02)};
  const CM31 bd = c_mul(x.b, y.b);
  const uint32_t nc1 = P - y.a.b, nd1 = P - y.b.b;
  const uint64_t ure = (uint64_t)x.a.a * y.b.a + (uint64_t)x.a.b * nd1 + (uint64_t)x.b.a * y.a.a + (uint64_t)x.b.b * nc1;
  const uint64_t uim = (uint64_t)x.a.a * y.b.b + (uint64_t)x.a.b * y.b.a + (uint64_t)x.b.a * y.a.b + (uint64_t)x.b.b * y.a.a;
  return {c_add(ac, c_mulR(bd)), {m_red_wide(ure), m_red_wide(uim)}};
#else
  // BD is reduced first; every other coordinate is one 64-bit sum (at most four products, 4P^2 < 2^64) and one reduction:
  // 16 multiplications and 6 reductions in all.
  const CM31 bd = c_mul(x.b, y.b);
  uint64_t re = (uint64_t)x.a.a * y.a.a + (uint64_t)x.a.b * (P - y.a.b) + 2ull * bd.a + (P - bd.b);
  uint64_t im = (uint64_t)x.a.a * y.a.b + (uint64_t)x.a.b * y.a.a + bd.a + 2ull * bd.b;
  uint64_t ure = (uint64_t)x.a.a * y.b.a + (uint64_t)x.a.b * (P - y.b.b) + (uint64_t)x.b.a * y.a.a + (uint64_t)x.b.b * (P - y.a.b);
  uint64_t uim = (uint64_t)x.a.a * y.b.b + (uint64_t)x.a.b * y.b.a + (uint64_t)x.b.a * y.a.b + (uint64_t)x.b.b * y.a.a;
  return {{m_red_wide(re), m_red_wide(im)}, {m_red_wide(ure), m_red_wide(uim)}};
#endif
}
SB_HD QM31 q_mulm(QM31 x, uint32_t m) { return {c_mulm(x.a, m), c_mulm(x.b, m)}; }
SB_HD QM31 q_mulc(QM31 x, CM31 c) { return {c_mul(x.a, c), c_mul(x.b, c)}; }
SB_HD QM31 q_sqr(QM31 x) { return q_mul(x, x); }
SB_HD QM31 q_inv(QM31 x) {
  CM31 d = c_inv(c_sub(c_mul(x.a, x.a), c_mulR(c_mul(x.b, x.b))));
  return {c_mul(x.a, d), c_neg(c_mul(x.b, d))};
}
SB_HD QM31 q_conj(QM31 x) { return {x.a, c_neg(x.b)}; }
SB_HD bool q_eq(QM31 x, QM31 y) { return x.a.a == y.a.a && x.a.b == y.a.b && x.b.a == y.b.a && x.b.b == y.b.b; }
SB_HD QM31 q_pow(QM31 x, uint64_t e) {
  QM31 r = q_fromm(1);
  while (e) {
    if (e & 1) r = q_mul(r, x);
    x = q_sqr(x);
    e >>= 1;
  }
  return r;
}

// Circle group over M31.
struct Pt {
  uint32_t x, y;
};
SB_HD Pt p_add(Pt p, Pt q) {
  return {m_sub(m_mul(p.x, q.x), m_mul(p.y, q.y)), m_add(m_mul(p.x, q.y), m_mul(p.y, q.x))};
}
SB_HD Pt p_dbl(Pt p) { return p_add(p, p); }
SB_HD Pt p_conj(Pt p) { return {p.x, m_neg(p.y)}; }
constexpr uint32_t GEN_X = 2u, GEN_Y = 1268011823u;  // M31_CIRCLE_GEN, order 2^31

struct QPt {
  QM31 x, y;
};
SB_HD QPt qp_add(QPt p, QPt q) {
  return {q_sub(q_mul(p.x, q.x), q_mul(p.y, q.y)), q_add(q_mul(p.x, q.y), q_mul(p.y, q.x))};
}
SB_HD QPt qp_from(Pt p) { return {q_fromm(p.x), q_fromm(p.y)}; }

SB_HD uint32_t bitrev32(uint32_t i, uint32_t log) {
#if defined(__CUDA_ARCH__)
  return log ? (__brev(i) >> (32 - log)) : 0;
#else
  uint32_t r = 0;
  for (uint32_t k = 0; k < log; k++) r |= ((i >> k) & 1u) << (log - 1 - k);
  return r;
#endif
}

}  // namespace sb
