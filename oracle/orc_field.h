// ORACLE — TEST INFRASTRUCTURE ONLY.  Not shipped, not linked into the product library.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// CPU restatement of the field tower used by the proving hot path of kkrt-labs/stwo-brainfuck.
// The arithmetic lives in the un-vendored dependency stwo-prover 0.1.1 (git starkware-libs/stwo
// rev 31e8dbcc4752240b596774743946c561ab5b9cd1, /root/reference/Cargo.toml:41, Cargo.lock:881-883),
// upstream files crates/prover/src/core/fields/{m31,cm31,qm31}.rs.  That source is absent from this
// container, so this file restates the published definitions.
// PARITY UNPINNED at the Backend boundary: the reference holds no golden vector for these ops.
// Reference call sites that depend on these definitions:
//   crates/brainfuck_prover/src/components/processor/table.rs:481-496 (PackedSecureField math)
//   crates/brainfuck_prover/src/components/mod.rs:122 (SECURE_EXTENSION_DEGREE = 4)
#pragma once
#include <cstdint>
#include <cstddef>

namespace orc {

static const uint32_t P = 2147483647u;  // 2^31 - 1

// Deliberately the most naive form (u64 %), so the oracle shares no tricks with the kernels.
static inline uint32_t madd(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a + b) % P); }
static inline uint32_t msub(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a + P - b) % P); }
static inline uint32_t mneg(uint32_t a) { return a == 0 ? 0 : P - a; }
static inline uint32_t mmul(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) % P); }
static inline uint32_t mpow(uint32_t a, uint64_t e) {
  uint32_t r = 1;
  while (e) { if (e & 1) r = mmul(r, a); a = mmul(a, a); e >>= 1; }
  return r;
}
static inline uint32_t minv(uint32_t a) { return mpow(a, P - 2); }

struct CM31 { uint32_t a, b; };  // a + b i, i^2 = -1
static inline CM31 cadd(CM31 x, CM31 y) { return {madd(x.a, y.a), madd(x.b, y.b)}; }
static inline CM31 csub(CM31 x, CM31 y) { return {msub(x.a, y.a), msub(x.b, y.b)}; }
static inline CM31 cneg(CM31 x) { return {mneg(x.a), mneg(x.b)}; }
static inline CM31 cmul(CM31 x, CM31 y) {
  return {msub(mmul(x.a, y.a), mmul(x.b, y.b)), madd(mmul(x.a, y.b), mmul(x.b, y.a))};
}
static inline CM31 cmulm(CM31 x, uint32_t m) { return {mmul(x.a, m), mmul(x.b, m)}; }
static inline CM31 cinv(CM31 x) {
  uint32_t n = minv(madd(mmul(x.a, x.a), mmul(x.b, x.b)));
  return {mmul(x.a, n), mmul(mneg(x.b), n)};
}

struct QM31 { CM31 a, b; };  // a + b u, u^2 = 2 + i
static inline QM31 qfrom(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return {{a, b}, {c, d}}; }
static inline QM31 qfromm(uint32_t a) { return {{a, 0}, {0, 0}}; }
static inline QM31 qadd(QM31 x, QM31 y) { return {cadd(x.a, y.a), cadd(x.b, y.b)}; }
static inline QM31 qsub(QM31 x, QM31 y) { return {csub(x.a, y.a), csub(x.b, y.b)}; }
static inline QM31 qneg(QM31 x) { return {cneg(x.a), cneg(x.b)}; }
static inline CM31 cmulR(CM31 x) { return cmul(x, CM31{2, 1}); }
static inline QM31 qmul(QM31 x, QM31 y) {
  return {cadd(cmul(x.a, y.a), cmulR(cmul(x.b, y.b))), cadd(cmul(x.a, y.b), cmul(x.b, y.a))};
}
static inline QM31 qmulm(QM31 x, uint32_t m) { return {cmulm(x.a, m), cmulm(x.b, m)}; }
static inline QM31 qmulc(QM31 x, CM31 c) { return {cmul(x.a, c), cmul(x.b, c)}; }
static inline QM31 qinv(QM31 x) {
  CM31 d = cinv(csub(cmul(x.a, x.a), cmulR(cmul(x.b, x.b))));
  return {cmul(x.a, d), cneg(cmul(x.b, d))};
}
static inline bool qeq(QM31 x, QM31 y) { return x.a.a == y.a.a && x.a.b == y.a.b && x.b.a == y.b.a && x.b.b == y.b.b; }
static inline QM31 qconj(QM31 x) { return {x.a, cneg(x.b)}; }  // u -> -u ("complex_conjugate" upstream)
static inline QM31 qpow(QM31 x, uint64_t e) {
  QM31 r = qfromm(1);
  while (e) { if (e & 1) r = qmul(r, x); x = qmul(x, x); e >>= 1; }
  return r;
}

static inline uint32_t bit_reverse(uint32_t i, uint32_t log) {
  if (log == 0) return 0;
  uint32_t r = 0;
  for (uint32_t k = 0; k < log; k++) r |= ((i >> k) & 1u) << (log - 1 - k);
  return r;
}

}  // namespace orc
