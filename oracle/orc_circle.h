// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_field.h header).
//
// Circle group, cosets, canonic domains, twiddle tree and the circle FFT as the CpuBackend of
// stwo-prover 0.1.1 @ 31e8dbc defines them (upstream crates/prover/src/core/{circle,fft,utils}.rs,
// core/poly/circle/{canonic,domain,evaluation,poly}.rs, core/poly/twiddles.rs,
// core/backend/cpu/circle.rs) — restated from SURVEY.md Appendix A.2-A.4; source absent here.
// PARITY UNPINNED (no golden vector in the reference).  Reference call sites:
//   crates/brainfuck_prover/src/brainfuck_air/mod.rs:480-484  precompute_twiddles(CanonicCoset(27)…half_coset)
//   crates/brainfuck_prover/src/brainfuck_air/mod.rs:497,550-562,690-702  extend_evals -> interpolate_columns
//   crates/brainfuck_prover/src/brainfuck_air/mod.rs:500,583,723  commit -> evaluate_polynomials
#pragma once
#include "orc_field.h"
#include <vector>

namespace orc {

struct Pt { uint32_t x, y; };
static const Pt CIRCLE_GEN = {2u, 1268011823u};  // order 2^31
static const uint32_t LOG_ORDER = 31;

static inline Pt padd(Pt p, Pt q) {
  return {msub(mmul(p.x, q.x), mmul(p.y, q.y)), madd(mmul(p.x, q.y), mmul(p.y, q.x))};
}
static inline Pt pconj(Pt p) { return {p.x, mneg(p.y)}; }
static inline uint32_t double_x(uint32_t x) { return msub(mmul(2, mmul(x, x)), 1); }

// G^idx, idx taken mod 2^31: the plain square-and-multiply ladder (the definition) ...
static inline Pt point_at_index_ladder(uint32_t idx) {
  idx &= 0x7fffffffu;
  Pt r = {1, 0}, b = CIRCLE_GEN;
  while (idx) { if (idx & 1) r = padd(r, b); b = padd(b, b); idx >>= 1; }
  return r;
}
// ... and the same element from four 256-entry tables of G^(d * 256^k) (three group additions instead of up to 61): the
// row loops below call this once per domain row, as upstream iterates its domains incrementally instead of exponentiating.
struct PointTables {
  Pt t[4][256];
  PointTables() {
    for (int k = 0; k < 4; k++) {
      Pt base = point_at_index_ladder(1u << (8 * k));
      t[k][0] = {1, 0};
      for (int d = 1; d < 256; d++) t[k][d] = padd(t[k][d - 1], base);
    }
  }
};
static inline Pt point_at_index(uint32_t idx) {
  static const PointTables T;  // thread-safe static initialisation
  idx &= 0x7fffffffu;
  return padd(padd(T.t[0][idx & 255], T.t[1][(idx >> 8) & 255]), padd(T.t[2][(idx >> 16) & 255], T.t[3][idx >> 24]));
}

struct Coset {
  uint32_t initial, step, log;  // indices into <G>, mod 2^31
  uint32_t index_at(uint32_t i) const { return (initial + (uint32_t)((uint64_t)step * i)) & 0x7fffffffu; }
  Pt at(uint32_t i) const { return point_at_index(index_at(i)); }
  Coset dbl() const { return {(initial * 2) & 0x7fffffffu, (step * 2) & 0x7fffffffu, log ? log - 1 : 0}; }
  uint32_t size() const { return 1u << log; }
};
static inline uint32_t subgroup_gen_index(uint32_t log) { return log == 0 ? 0 : (1u << (LOG_ORDER - log)); }
static inline Coset coset_odds(uint32_t log) { return {subgroup_gen_index(log + 1), subgroup_gen_index(log), log}; }
static inline Coset coset_half_odds(uint32_t log) { return {subgroup_gen_index(log + 2), subgroup_gen_index(log), log}; }

// CanonicCoset::new(log).circle_domain(): half_coset = half_odds(log-1);
// at(i) = half.at(i) for i < 2^(log-1), else conj(half.at(i - 2^(log-1))).
struct CircleDomain {
  Coset half;
  uint32_t log_size() const { return half.log + 1; }
  uint32_t index_at(uint32_t i) const {
    uint32_t h = half.size();
    if (i < h) return half.index_at(i);
    return (0x80000000u - half.index_at(i - h)) & 0x7fffffffu;
  }
  Pt at(uint32_t i) const { return point_at_index(index_at(i)); }
};
static inline CircleDomain canonic_domain(uint32_t log) { return {coset_half_odds(log - 1)}; }

// precompute_twiddles(coset): per level, bit_reverse(x of first half of the coset); coset doubles; pad with 1.
static inline void precompute_twiddles(Coset coset, std::vector<uint32_t>& tw, std::vector<uint32_t>& itw) {
  size_t total = (size_t)1 << coset.log;
  tw.assign(total, 1);
  size_t off = 0;
  Coset c = coset;
  for (uint32_t lvl = 0; lvl < coset.log; lvl++) {
    uint32_t half = c.size() / 2, hl = c.log - 1;
    Pt p = point_at_index(c.initial), s = point_at_index(c.step);
    for (uint32_t i = 0; i < half; i++) {
      tw[off + bit_reverse(i, hl)] = p.x;
      p = padd(p, s);
    }
    off += half;
    c = c.dbl();
  }
  // element-wise inverse (batch inverse upstream; same values)
  itw.resize(total);
  // Montgomery batch trick in chunks to keep this fast.
  std::vector<uint32_t> pre(total);
  uint32_t acc = 1;
  for (size_t i = 0; i < total; i++) { pre[i] = acc; acc = mmul(acc, tw[i]); }
  uint32_t inv = minv(acc);
  for (size_t i = total; i-- > 0;) { itw[i] = mmul(inv, pre[i]); inv = mmul(inv, tw[i]); }
}

// Line-layer twiddles of a circle domain taken from a (bigger) tree: layer i (FFT layer i+1) is
// buf[len - 2^(i'+1) .. len - 2^i'] with i' = half.log - 1 - i, largest first.
static inline std::vector<const uint32_t*> domain_line_twiddles(const std::vector<uint32_t>& buf, uint32_t half_log) {
  std::vector<const uint32_t*> out;
  size_t len = buf.size();
  for (uint32_t i = 0; i < half_log; i++) {
    uint32_t l = half_log - 1 - i;
    out.push_back(buf.data() + (len - ((size_t)2 << l)));
  }
  return out;
}

static inline void fft_layer(uint32_t* v, uint32_t layer, size_t h, uint32_t t, bool inverse) {
  size_t span = (size_t)1 << layer;
  for (size_t l = 0; l < span; l++) {
    size_t i0 = (h << (layer + 1)) + l, i1 = i0 + span;
    uint32_t a = v[i0], b = v[i1];
    if (inverse) { v[i0] = madd(a, b); v[i1] = mmul(msub(a, b), t); }
    else { uint32_t m = mmul(b, t); v[i0] = madd(a, m); v[i1] = msub(a, m); }
  }
}
// One whole layer: butterfly k pairs (h 2^(layer+1) + l, + 2^layer) with h = k >> layer, l = k mod 2^layer, twiddle tw(h).
// Large layers are split over the host threads when the caller is not already inside a parallel region (a transform of
// few long columns, e.g. the four composition coordinates).
template <class TW>
static inline void fft_whole_layer(uint32_t* v, uint32_t log, uint32_t layer, bool inverse, TW tw) {
  const size_t half = (size_t)1 << (log - 1), span = (size_t)1 << layer;
#pragma omp parallel for schedule(static) if (half >= 32768)
  for (size_t k = 0; k < half; k++) {
    size_t h = k >> layer, l = k & (span - 1);
    size_t i0 = (h << (layer + 1)) + l, i1 = i0 + span;
    uint32_t a = v[i0], b = v[i1], t = tw(h);
    if (inverse) { v[i0] = madd(a, b); v[i1] = mmul(msub(a, b), t); }
    else { uint32_t m = mmul(b, t); v[i0] = madd(a, m); v[i1] = msub(a, m); }
  }
}

// Circle-layer twiddles from the first line layer: [x, y] -> [y, -y, -x, x].
static inline uint32_t circle_twiddle(const uint32_t* line0, size_t h) {
  uint32_t x = line0[(h >> 2) * 2], y = line0[(h >> 2) * 2 + 1];
  switch (h & 3) { case 0: return y; case 1: return mneg(y); case 2: return mneg(x); default: return x; }
}

// interpolate: bit-reversed evaluations on canonic_domain(log) -> coefficients, in place.
static inline void interpolate(uint32_t* v, uint32_t log, const std::vector<uint32_t>& itw) {
  size_t n = (size_t)1 << log;
  if (log == 0) return;
  if (log == 1) {  // direct definition: f = c0 + c1*y on {p, conj p}
    Pt p = canonic_domain(1).at(0);
    uint32_t a = v[0], b = v[1], i2 = minv(2);
    v[0] = mmul(madd(a, b), i2);
    v[1] = mmul(mmul(msub(a, b), i2), minv(p.y));
    return;
  }
  auto lines = domain_line_twiddles(itw, log - 1);
  if (log == 2) {  // one line twiddle only: circle twiddles are [y, -y] of the half-coset's initial point
    uint32_t iy = minv(canonic_domain(2).at(0).y);
    fft_layer(v, 0, 0, iy, true); fft_layer(v, 0, 1, mneg(iy), true);
  } else
  fft_whole_layer(v, log, 0, true, [&](size_t h) { return circle_twiddle(lines[0], h); });
  for (uint32_t layer = 1; layer < log; layer++) {
    const uint32_t* t = lines[layer - 1];
    fft_whole_layer(v, log, layer, true, [&](size_t h) { return t[h]; });
  }
  uint32_t ninv = minv(mpow(2, log));
#pragma omp parallel for schedule(static) if (n >= 65536)
  for (size_t i = 0; i < n; i++) v[i] = mmul(v[i], ninv);
}

// evaluate: coefficients (length 2^log) -> bit-reversed evaluations on canonic_domain(log), in place.
static inline void evaluate(uint32_t* v, uint32_t log, const std::vector<uint32_t>& tw) {
  if (log == 0) return;
  if (log == 1) {
    Pt p = canonic_domain(1).at(0);
    uint32_t c0 = v[0], c1 = mmul(v[1], p.y);
    v[0] = madd(c0, c1); v[1] = msub(c0, c1);
    return;
  }
  auto lines = domain_line_twiddles(tw, log - 1);
  for (uint32_t layer = log - 1; layer >= 1; layer--) {
    const uint32_t* t = lines[layer - 1];
    fft_whole_layer(v, log, layer, false, [&](size_t h) { return t[h]; });
  }
  if (log == 2) {
    uint32_t y = canonic_domain(2).at(0).y;
    fft_layer(v, 0, 0, y, false); fft_layer(v, 0, 1, mneg(y), false);
    return;
  }
  fft_whole_layer(v, log, 0, false, [&](size_t h) { return circle_twiddle(lines[0], h); });
}

// Secure-field circle point.
struct QPt { QM31 x, y; };
static inline QPt qpadd(QPt p, QPt q) {
  return {qsub(qmul(p.x, q.x), qmul(p.y, q.y)), qadd(qmul(p.x, q.y), qmul(p.y, q.x))};
}
static inline QPt qp_from(Pt p) { return {qfromm(p.x), qfromm(p.y)}; }
static inline QM31 qdouble_x(QM31 x) { return qsub(qmulm(qmul(x, x), 2), qfromm(1)); }

// eval_at_point: fold(coeffs, [y, x, pi(x), ...] reversed); top half of the array pairs with the first factor.
static inline QM31 eval_at_point(const uint32_t* c, uint32_t log, QPt p) {
  if (log == 0) return qfromm(c[0]);
  std::vector<QM31> map;  // map[k] <-> coefficient-index bit k
  map.push_back(p.y);
  QM31 x = p.x;
  for (uint32_t k = 1; k < log; k++) { map.push_back(x); x = qdouble_x(x); }
  size_t n = (size_t)1 << log;
  std::vector<QM31> cur(n / 2);
  for (size_t i = 0; i < n / 2; i++) cur[i] = qadd(qfromm(c[2 * i]), qmulm(map[0], c[2 * i + 1]));
  for (uint32_t k = 1; k < log; k++) {
    size_t m = n >> (k + 1);
    for (size_t i = 0; i < m; i++) cur[i] = qadd(cur[2 * i], qmul(cur[2 * i + 1], map[k]));
  }
  return cur[0];
}

}  // namespace orc
