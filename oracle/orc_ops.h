// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_field.h header).
//
// Row-local Backend ops as stwo-prover 0.1.1 @ 31e8dbc's CpuBackend defines them (upstream
// core/backend/cpu/{fri,quotients,accumulation,grind}.rs, core/utils.rs, core/backend/simd/prefix_sum.rs,
// core/pcs/quotients.rs — absent here; restated from SURVEY.md Appendix A.9-A.11).  PARITY UNPINNED.
// Reference entry: crates/brainfuck_prover/src/brainfuck_air/mod.rs:732 (prover::prove) and the
// `logup_gen.finalize_last()` calls, e.g. crates/brainfuck_prover/src/components/processor/table.rs:530.
#pragma once
#include "orc_blake2s.h"
#include "orc_circle.h"
#include <algorithm>

namespace orc {

static inline void bit_reverse_column(uint32_t* v, uint32_t log) {
  size_t n = (size_t)1 << log;
  for (size_t i = 0; i < n; i++) {
    size_t j = bit_reverse((uint32_t)i, log);
    if (i < j) std::swap(v[i], v[j]);
  }
}

static inline QM31 qat(const uint32_t* const* c, size_t i) { return qfrom(c[0][i], c[1][i], c[2][i], c[3][i]); }
static inline void qset(uint32_t* const* c, size_t i, QM31 v) { c[0][i] = v.a.a; c[1][i] = v.a.b; c[2][i] = v.b.a; c[3][i] = v.b.b; }

// fold_line: src on LineDomain(half_odds(log)) bit-reversed, 2^log values -> 2^(log-1) values.
//   x = domain.at(bit_reverse(2i, log)).x ; (f0,f1) = ibutterfly(e[2i], e[2i+1], 1/x) ; out[i] = f0 + alpha f1
static inline void fold_line(const uint32_t* const* src, uint32_t log, QM31 alpha, uint32_t* const* dst) {
  Coset dom = coset_half_odds(log);
  size_t n = (size_t)1 << log;
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n / 2; i++) {
    uint32_t x = dom.at(bit_reverse((uint32_t)(2 * i), log)).x;
    QM31 a = qat(src, 2 * i), b = qat(src, 2 * i + 1);
    QM31 f0 = qadd(a, b), f1 = qmulm(qsub(a, b), minv(x));
    qset(dst, i, qadd(f0, qmul(alpha, f1)));
  }
}

// fold_circle_into_line: src on canonic_domain(log) bit-reversed; dst (2^(log-1)) accumulates:
//   p = domain.at(bit_reverse(2i, log)); (f0,f1) = ibutterfly(src[2i], src[2i+1], 1/p.y); dst[i] = dst[i] alpha^2 + f0 + alpha f1
static inline void fold_circle_into_line(const uint32_t* const* src, uint32_t log, QM31 alpha, uint32_t* const* dst) {
  CircleDomain dom = canonic_domain(log);
  size_t n = (size_t)1 << log;
  QM31 a2 = qmul(alpha, alpha);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n / 2; i++) {
    Pt p = dom.at(bit_reverse((uint32_t)(2 * i), log));
    QM31 a = qat(src, 2 * i), b = qat(src, 2 * i + 1);
    QM31 f0 = qadd(a, b), f1 = qmulm(qsub(a, b), minv(p.y));
    QM31 d = qat((const uint32_t* const*)dst, i);
    qset(dst, i, qadd(qmul(d, a2), qadd(f0, qmul(alpha, f1))));
  }
}

// coset index -> circle-domain index (upstream core/utils.rs coset_index_to_circle_domain_index)
static inline size_t coset_to_domain_index(size_t i, uint32_t log) {
  return (i & 1) == 0 ? i / 2 : (((size_t)2 << log) - i) / 2;
}

// Inclusive prefix sum in trace-coset order over a column stored in bit-reversed circle-domain order.
static inline void prefix_sum_bitrev(uint32_t* v, uint32_t log) {
  size_t n = (size_t)1 << log;
  uint32_t acc = 0;
  for (size_t i = 0; i < n; i++) {
    size_t s = bit_reverse((uint32_t)coset_to_domain_index(i, log), log);
    acc = madd(acc, v[s]);
    v[s] = acc;
  }
}

// accumulate_quotients (CpuBackend::accumulate_quotients + pcs/quotients.rs helpers).
static inline void accumulate_quotients(uint32_t log, const uint32_t* const* cols, uint32_t ncols, QM31 alpha,
                                        const uint32_t* bpts, const uint32_t* bsizes, const uint32_t* ecols,
                                        const uint32_t* evals, uint32_t nb, uint32_t* const* out) {
  (void)ncols;
  struct LC { QM31 a, b, c; };
  std::vector<std::vector<LC>> lcs(nb);
  std::vector<QM31> bcoef(nb);
  std::vector<QPt> pts(nb);
  size_t e = 0;
  for (uint32_t b = 0; b < nb; b++) {
    const uint32_t* q = bpts + 8 * b;
    pts[b] = {qfrom(q[0], q[1], q[2], q[3]), qfrom(q[4], q[5], q[6], q[7])};
    QM31 al = qfromm(1);
    for (uint32_t j = 0; j < bsizes[b]; j++, e++) {
      al = qmul(al, alpha);
      QM31 v = qfrom(evals[4 * e], evals[4 * e + 1], evals[4 * e + 2], evals[4 * e + 3]);
      QM31 a = qsub(qconj(v), v);
      QM31 c = qsub(qconj(pts[b].y), pts[b].y);
      QM31 bb = qsub(qmul(v, c), qmul(a, pts[b].y));
      lcs[b].push_back({qmul(al, a), qmul(al, bb), qmul(al, c)});
    }
    bcoef[b] = qpow(alpha, bsizes[b]);
  }
  CircleDomain dom = canonic_domain(log);
  size_t n = (size_t)1 << log;
#pragma omp parallel for schedule(static)
  for (size_t row = 0; row < n; row++) {
    Pt p = dom.at(bit_reverse((uint32_t)row, log));
    QM31 acc = qfromm(0);
    size_t ee = 0;
    for (uint32_t b = 0; b < nb; b++) {
      CM31 prx = pts[b].x.a, pry = pts[b].y.a, pix = pts[b].x.b, piy = pts[b].y.b;
      CM31 den = csub(cmul(csub(prx, CM31{p.x, 0}), piy), cmul(csub(pry, CM31{p.y, 0}), pix));
      QM31 num = qfromm(0);
      for (uint32_t j = 0; j < bsizes[b]; j++, ee++) {
        const LC& l = lcs[b][j];
        QM31 value = qmulm(l.c, cols[ecols[ee]][row]);
        QM31 lin = qadd(qmulm(l.a, p.y), l.b);
        num = qadd(num, qsub(value, lin));
      }
      acc = qadd(qmul(acc, bcoef[b]), qmulc(num, cinv(den)));
    }
    qset(out, row, acc);
  }
}

// GrindOps: smallest nonce with trailing_zeros(mix_u64(nonce)) >= pow_bits; mix_u64 = raw compress(digest, [lo,hi,0..]).
static inline uint32_t trailing_zeros_128(const uint32_t d[8]) {
  for (int w = 0; w < 4; w++)
    if (d[w]) return 32 * w + (uint32_t)__builtin_ctz(d[w]);
  return 128;
}
static inline uint64_t grind(const uint32_t digest[8], uint32_t pow_bits) {
  for (uint64_t nonce = 0;; nonce++) {
    uint32_t h[8], m[16] = {0};
    memcpy(h, digest, 32);
    m[0] = (uint32_t)nonce; m[1] = (uint32_t)(nonce >> 32);
    b2s_compress(h, m, 0, 0, 0, 0);
    if (trailing_zeros_128(h) >= pow_bits) return nonce;
  }
}

}  // namespace orc
