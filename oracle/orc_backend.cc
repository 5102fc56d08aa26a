// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_field.h header).
//
// CPU `Backend` for the protocol driver: every device op of the product replaced by the scalar restatements in this
// directory (own field arithmetic, own FFT/Merkle/quotient/fold loops, own row evaluators).  The driver itself
// (stwo-brainfuck_b200/csrc/host/{prover,prover_sharded,verifier}.hpp) and the AIR definition (host/air.hpp — the
// restatement of the reference's components/<name>/component.rs `evaluate` bodies) are shared with the product on purpose:
// they are host logic in the reference too; what this oracle checks is that the CUDA kernels compute the same field elements.
// Proof parity test: tests/test_prover_gpu.py compares the JSON of both provers byte for byte.
// The multi-rank entry point (orc_prove_sharded_json) runs the sharded driver on `world` threads with an in-process
// communicator, so that the N > 1 host logic is testable without GPUs.
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include "../stwo-brainfuck_b200/csrc/host/prover_sharded.hpp"
#include "../stwo-brainfuck_b200/csrc/host/verifier.hpp"
#include <omp.h>
#include "orc_ops.h"

using namespace sbf;

namespace {

struct HCol {  // host column: owns `store`, or is a view into another column's memory
  uint32_t* d = nullptr;
  size_t n = 0;
  std::vector<uint32_t> store;
  explicit HCol(size_t len) : n(len), store(len, 0) { d = store.data(); }
  HCol(const uint32_t* v, size_t len) : n(len), store(v, v + len) { d = store.data(); }
  HCol(uint32_t* view, size_t len, int) : d(view), n(len) {}
};
inline HCol* H(Col c) { return (HCol*)c; }
inline orc::QM31 oq(const sb::QM31& q) { return orc::qfrom(q.a.a, q.a.b, q.b.a, q.b.b); }
inline sb::QM31 sq(const orc::QM31& q) { return sb::q_make(q.a.a, q.a.b, q.b.a, q.b.b); }

// oracle-side field value types for the shared AIR templates
struct OFm { uint32_t v; };
struct OFq { orc::QM31 v; };
inline OFm operator+(OFm a, OFm b) { return {orc::madd(a.v, b.v)}; }
inline OFm operator-(OFm a, OFm b) { return {orc::msub(a.v, b.v)}; }
inline OFm operator*(OFm a, OFm b) { return {orc::mmul(a.v, b.v)}; }
inline OFq operator+(OFq a, OFq b) { return {orc::qadd(a.v, b.v)}; }
inline OFq operator-(OFq a, OFq b) { return {orc::qsub(a.v, b.v)}; }
inline OFq operator*(OFq a, OFq b) { return {orc::qmul(a.v, b.v)}; }
inline OFq operator*(OFq a, OFm b) { return {orc::qmulm(a.v, b.v)}; }

inline OFq ocombine(const LookupElements& le, const OFm* vals, int n) {
  orc::QM31 acc = orc::qfromm(0);
  for (int i = 0; i < n; i++) acc = orc::qadd(acc, orc::qmulm(oq(le.alpha_pow[i]), vals[i].v));
  return {orc::qsub(acc, oq(le.z))};
}

struct OrcLogupGen {
  typedef OFm F;
  typedef OFq EF;
  const std::vector<Col>* main;
  std::vector<HCol*>* out;  // null entries are not materialised
  const InteractionElements* el;
  size_t row, shift;
  int col = 0, batch = 0;
  orc::QM31 cum = orc::qfromm(0);
  F next() { return {H((*main)[col++])->d[row >> shift]}; }
  F is_first() { return {0}; }
  F cst(uint32_t c) { return {c}; }
  EF ef(F x) { return {orc::qfromm(x.v)}; }
  EF ef_neg_one() { return {orc::qfromm(orc::P - 1)}; }
  void add(F) {}
  void add(EF) {}
  void relation(int rel, EF num, const F* vals, int n) {
    OFq den = ocombine(el->rel[rel], vals, n);
    cum = orc::qadd(cum, orc::qmul(num.v, orc::qinv(den.v)));
    const uint32_t w[4] = {cum.a.a, cum.a.b, cum.b.a, cum.b.b};
    for (int k = 0; k < 4; k++) if ((*out)[4 * batch + k]) (*out)[4 * batch + k]->d[row] = w[k];
    batch++;
  }
  void finalize_logup() {}
};

struct OrcDomainEval {
  typedef OFm F;
  typedef OFq EF;
  const std::vector<Col>*main, *inter;
  const std::array<Col, 4>* prev = nullptr;  // pre-shifted last LogUp column (row-range evaluation), else index prev_row
  const uint32_t* is_first_col;
  const InteractionElements* el;
  const std::vector<sb::QM31>* coeff;
  orc::QM31 total;
  size_t row, prev_row;  // row indexes the (possibly partial) columns
  int col = 0, k = 0;
  orc::QM31 row_res = orc::qfromm(0);
  LogupState<OrcDomainEval> lg;
  F next() { return {H((*main)[col++])->d[row]}; }
  F is_first() { return {is_first_col[row]}; }
  F cst(uint32_t c) { return {c}; }
  EF ef(F x) { return {orc::qfromm(x.v)}; }
  EF ef_zero() { return {orc::qfromm(0)}; }
  EF ef_neg_one() { return {orc::qfromm(orc::P - 1)}; }
  EF total_sum() { return {total}; }
  void add(F c) { row_res = orc::qadd(row_res, orc::qmulm(oq((*coeff)[k++]), c.v)); }
  void add(EF c) { row_res = orc::qadd(row_res, orc::qmul(oq((*coeff)[k++]), c.v)); }
  void relation(int rel, EF num, const F* vals, int n) { lg.push(num, ocombine(el->rel[rel], vals, n)); }
  EF ext_at(int b, size_t r) {
    return {orc::qfrom(H((*inter)[4 * b])->d[r], H((*inter)[4 * b + 1])->d[r], H((*inter)[4 * b + 2])->d[r], H((*inter)[4 * b + 3])->d[r])};
  }
  EF ext_mask_cur(int b) { return ext_at(b, row); }
  void ext_mask_last(EF& pv, EF& cur) {
    if (prev) pv = {orc::qfrom(H((*prev)[0])->d[row], H((*prev)[1])->d[row], H((*prev)[2])->d[row], H((*prev)[3])->d[row])};
    else pv = ext_at(lg.n - 1, prev_row);
    cur = ext_at(lg.n - 1, row);
  }
  void finalize_logup() { lg.finalize(*this); }
};

// ---- in-process communicator for the multi-rank tests: `world` threads, one OrcBackend each
struct ThreadComm {
  int world;
  std::mutex mu;
  std::condition_variable cv;
  int waiting = 0;
  long generation = 0;
  std::vector<const uint32_t*> ptr;
  std::vector<const std::vector<size_t>*> counts;
  explicit ThreadComm(int w) : world(w), ptr(w, nullptr), counts(w, nullptr) {}
  void barrier() {
    std::unique_lock<std::mutex> lk(mu);
    long gen = generation;
    if (++waiting == world) { waiting = 0; generation++; cv.notify_all(); }
    else cv.wait(lk, [&] { return generation != gen; });
  }
};

struct OrcBackend : Backend {
  std::vector<uint32_t> tw, itw;
  ThreadComm* comm = nullptr;
  int my_rank = 0;
  const char* name() const override { return "oracle"; }
  static uint32_t lg2(size_t n) { uint32_t l = 0; while (((size_t)1 << l) < n) l++; return l; }

  Col from_host(const uint32_t* v, size_t n) override { return new HCol(v, n); }
  Col broadcast16(Col c) override {
    HCol* o = new HCol(H(c)->n * 16);
    for (size_t i = 0; i < o->n; i++) o->d[i] = H(c)->d[i >> 4];
    return o;
  }
  Col zeros(size_t n) override { return new HCol(n); }
  size_t len(Col c) override { return H(c)->n; }
  void read(Col c, size_t off, size_t n, uint32_t* out) override { memcpy(out, H(c)->d + off, n * 4); }
  void free_col(Col c) override { delete H(c); }
  std::vector<uint32_t> gather(const std::vector<Col>& cols, const std::vector<size_t>& offsets, uint32_t words) override {
    std::vector<uint32_t> out(cols.size() * words);
    for (size_t i = 0; i < cols.size(); i++) memcpy(&out[i * words], H(cols[i])->d + offsets[i], words * 4);
    return out;
  }

  // Per-column work over columns of very different lengths: long columns one after the other (their transforms split each
  // layer over the host threads, orc_circle.h), the short ones in parallel over the columns.
  template <class F>
  static void for_columns(const std::vector<Col>& v, uint32_t extra_log, F body, bool split = true) {
    const size_t big = split ? (size_t)1 << 18 : ~(size_t)0;
    for (size_t i = 0; i < v.size(); i++) if ((H(v[i])->n << extra_log) >= big) body(i);
#pragma omp parallel for schedule(dynamic)
    for (size_t i = 0; i < v.size(); i++) if ((H(v[i])->n << extra_log) < big) body(i);
  }
  void precompute_twiddles(uint32_t root_log) override { orc::precompute_twiddles(orc::coset_half_odds(root_log), tw, itw); }
  void interpolate(const std::vector<Col>& cols) override {
    for_columns(cols, 0, [&](size_t i) { orc::interpolate(H(cols[i])->d, lg2(H(cols[i])->n), itw); });
  }
  std::vector<Col> evaluate(const std::vector<Col>& coeffs, uint32_t log_blowup) override {
    std::vector<Col> out(coeffs.size());
    for_columns(coeffs, log_blowup, [&](size_t i) {
      HCol* o = new HCol(H(coeffs[i])->n << log_blowup);  // extend(): zero-pad at the end
      memcpy(o->d, H(coeffs[i])->d, H(coeffs[i])->n * 4);
      orc::evaluate(o->d, lg2(o->n), tw);
      out[i] = o;
    });
    return out;
  }
  std::vector<sb::QM31> eval_at_point(const std::vector<Col>& polys, const std::vector<QPoint>& pts) override {
    std::vector<sb::QM31> out(polys.size());
#pragma omp parallel for schedule(dynamic)
    for (size_t i = 0; i < polys.size(); i++)
      out[i] = sq(orc::eval_at_point(H(polys[i])->d, lg2(H(polys[i])->n), orc::QPt{oq(pts[i].x), oq(pts[i].y)}));
    return out;
  }
  // Lane-repeated columns the slow, obvious way: expand, run the plain transform, check the claimed structure, compact.
  static HCol* expand_values(const HCol* c, uint32_t rep) {
    HCol* o = new HCol(c->n << rep);
    for (size_t i = 0; i < o->n; i++) o->d[i] = c->d[i >> rep];
    return o;
  }
  static HCol* expand_coeffs(const HCol* c, uint32_t rep, uint32_t log_blowup) {
    HCol* o = new HCol((c->n << rep) << log_blowup);  // HCol(n) is zero-filled
    for (size_t i = 0; i < c->n; i++) o->d[i << rep] = c->d[i];
    return o;
  }
  std::vector<Col> interpolate_repeated(const std::vector<Col>& cols, uint32_t rep) override {
    bool bad = false;
    std::vector<Col> out(cols.size());
    for_columns(cols, rep, [&](size_t i) {
      HCol* full = expand_values(H(cols[i]), rep);
      orc::interpolate(full->d, lg2(full->n), itw);
      HCol* o = new HCol(H(cols[i])->n);
      for (size_t j = 0; j < full->n; j++) {
        if ((j & (((size_t)1 << rep) - 1)) == 0) o->d[j >> rep] = full->d[j];
        else if (full->d[j] != 0) bad = true;
      }
      delete full;
      out[i] = o;
    });
    if (bad) throw std::runtime_error("oracle: a repeated column has a non-zero coefficient off the 2^rep grid");
    return out;
  }
  std::vector<Col> evaluate_repeated(const std::vector<Col>& coeffs, uint32_t rep, uint32_t log_blowup) override {
    std::vector<Col> out(coeffs.size());
    for_columns(coeffs, rep + log_blowup, [&](size_t i) {
      HCol* o = expand_coeffs(H(coeffs[i]), rep, log_blowup);
      orc::evaluate(o->d, lg2(o->n), tw);
      out[i] = o;
    });
    return out;
  }
  std::vector<Col> evaluate_repeated_range(const std::vector<Col>& coeffs, uint32_t rep, uint32_t log_blowup, const std::vector<size_t>& offs,
                                           const std::vector<size_t>& cnts) override {
    std::vector<Col> full = evaluate_repeated(coeffs, rep, log_blowup), out(coeffs.size());
    for (size_t i = 0; i < coeffs.size(); i++) { out[i] = new HCol(H(full[i])->d + offs[i], cnts[i]); delete H(full[i]); }
    return out;
  }
  std::vector<sb::QM31> eval_at_point_repeated(const std::vector<Col>& polys, const std::vector<uint32_t>& reps, const std::vector<QPoint>& pts) override {
    std::vector<sb::QM31> out(polys.size());
    for_columns(polys, 0, [&](size_t i) {
      HCol* full = expand_coeffs(H(polys[i]), reps[i], 0);
      out[i] = sq(orc::eval_at_point(full->d, lg2(full->n), orc::QPt{oq(pts[i].x), oq(pts[i].y)}));
      delete full;
    }, false);
    return out;
  }
  std::vector<Col> merkle_commit_repeated(const std::vector<Col>& cols, uint32_t, Hash* root) override { return merkle_commit(cols, root); }
  std::vector<Col> merkle_commit(const std::vector<Col>& cols, Hash* root) override {
    // MerkleProver::commit (orc_blake2s.h merkle_commit), layer by layer straight into the columns that are kept
    uint32_t max_log = 0;
    for (Col c : cols) max_log = std::max(max_log, lg2(H(c)->n));
    std::vector<Col> out(max_log + 1);
    for (int lg = (int)max_log; lg >= 0; lg--) {
      std::vector<const uint32_t*> lc;
      for (Col c : cols) if (lg2(H(c)->n) == (uint32_t)lg) lc.push_back(H(c)->d);  // stable order
      HCol* o = new HCol((size_t)8 << lg);
      orc::commit_on_layer(lg, lg == (int)max_log ? nullptr : H(out[lg + 1])->d, lc.data(), lc.size(), o->d);
      out[lg] = o;
    }
    if (root) memcpy(root->data(), H(out[0])->d, 32);
    return out;
  }
  Col commit_layer(uint32_t log, Col prev, const std::vector<Col>& cols) override {
    std::vector<const uint32_t*> p;
    for (Col c : cols) p.push_back(H(c)->d);
    HCol* o = new HCol((size_t)8 << log);
    orc::commit_on_layer(log, prev ? H(prev)->d : nullptr, p.data(), p.size(), o->d);
    return o;
  }
  std::array<Col, 4> fold_line(const std::array<Col, 4>& src, uint32_t log, sb::QM31 alpha) override {
    return fold_line_range(src, log, 0, (size_t)1 << (log - 1), alpha);
  }
  void fold_circle_into_line(const std::array<Col, 4>& dst, const std::array<Col, 4>& src, uint32_t log, sb::QM31 alpha) override {
    fold_circle_into_line_range(dst, src, log, 0, (size_t)1 << (log - 1), alpha);
  }
  std::array<Col, 4> accumulate_quotients(uint32_t log, const std::vector<Col>& cols, sb::QM31 rc, const SampleBatchesFlat& b) override {
    return accumulate_quotients_range(log, 0, (size_t)1 << log, cols, rc, b);
  }
  void accumulate(const std::array<Col, 4>& dst, const std::array<Col, 4>& src) override {
    for (int k = 0; k < 4; k++) accumulate_col(dst[k], src[k]);
  }
  uint64_t grind(const Hash& digest, uint32_t pow_bits) override { return orc::grind(digest.data(), pow_bits); }
  Col gen_is_first(uint32_t log_size) override { HCol* c = new HCol((size_t)1 << log_size); c->d[0] = 1; return c; }

  std::vector<Col> logup_generate(int comp, const std::vector<Col>& main, const InteractionElements& el, sb::QM31& claimed) override {
    std::vector<uint8_t> want(4 * N_LOGUP_COLS[comp], 1);
    std::vector<Col> out = logup_generate_sel(comp, main, el, want);
    for (int k = 0; k < 4; k++) prefix_sum(out[out.size() - 4 + k]);
    size_t b = out.size() - 4;
    claimed = sb::q_make(H(out[b])->d[1], H(out[b + 1])->d[1], H(out[b + 2])->d[1], H(out[b + 3])->d[1]);
    return out;
  }
  void eval_constraints(int comp, uint32_t log_size, const std::vector<Col>& m, const std::vector<Col>& it, Col is_first,
                        const InteractionElements& el, sb::QM31 total, const std::vector<sb::QM31>& coeffs,
                        const std::array<Col, 4>& acc) override {
    eval_rows(comp, log_size, 0, (size_t)2 << log_size, m, it, nullptr, is_first, el, total, coeffs, acc);
  }

  // ---- multi-rank extension
  int rank() const override { return my_rank; }
  int world() const override { return comm ? comm->world : 1; }
  Col alloc(size_t n) override { return new HCol(n); }
  Col view(Col c, size_t off, size_t n) override { return new HCol(H(c)->d + off, n, 0); }
  void copy(Col dst, size_t dst_off, Col src, size_t src_off, size_t n) override { memcpy(H(dst)->d + dst_off, H(src)->d + src_off, n * 4); }
  void all_to_all(Col send, const std::vector<size_t>& sc, Col recv, const std::vector<size_t>& rc) override {
    if (!comm) { memcpy(H(recv)->d, H(send)->d, sc[0] * 4); return; }
    comm->ptr[my_rank] = H(send)->d;
    comm->counts[my_rank] = &sc;
    comm->barrier();
    size_t ro = 0;
    for (int s = 0; s < comm->world; s++) {
      size_t so = 0;
      for (int d = 0; d < my_rank; d++) so += (*comm->counts[s])[d];
      memcpy(H(recv)->d + ro, comm->ptr[s] + so, rc[s] * 4);
      ro += rc[s];
    }
    comm->barrier();
  }
  void all_gather(Col send, Col recv, size_t n) override {
    if (!comm) { memcpy(H(recv)->d, H(send)->d, n * 4); return; }
    comm->ptr[my_rank] = H(send)->d;
    comm->barrier();
    for (int s = 0; s < comm->world; s++) memcpy(H(recv)->d + s * n, comm->ptr[s], n * 4);
    comm->barrier();
  }
  void allreduce_host(uint32_t* buf, size_t n) override {
    if (!comm) return;
    comm->ptr[my_rank] = buf;
    comm->barrier();
    std::vector<uint32_t> sum(n, 0);
    for (int s = 0; s < comm->world; s++) for (size_t i = 0; i < n; i++) sum[i] += comm->ptr[s][i];
    comm->barrier();
    memcpy(buf, sum.data(), n * 4);
    comm->barrier();
  }
  // out[i] for i in [off, off+n): fold of inputs 2i, 2i+1 (src holds exactly those 2n inputs)
  std::array<Col, 4> fold_line_range(const std::array<Col, 4>& src, uint32_t log, size_t off, size_t n_out, sb::QM31 alpha) override {
    std::array<Col, 4> out;
    for (int k = 0; k < 4; k++) out[k] = new HCol(n_out);
    orc::Coset dom = orc::coset_half_odds(log);
    orc::QM31 al = oq(alpha);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n_out; i++) {
      uint32_t x = dom.at(orc::bit_reverse((uint32_t)(2 * (off + i)), log)).x;
      orc::QM31 a = orc::qfrom(H(src[0])->d[2 * i], H(src[1])->d[2 * i], H(src[2])->d[2 * i], H(src[3])->d[2 * i]);
      orc::QM31 b = orc::qfrom(H(src[0])->d[2 * i + 1], H(src[1])->d[2 * i + 1], H(src[2])->d[2 * i + 1], H(src[3])->d[2 * i + 1]);
      orc::QM31 r = orc::qadd(orc::qadd(a, b), orc::qmul(al, orc::qmulm(orc::qsub(a, b), orc::minv(x))));
      H(out[0])->d[i] = r.a.a; H(out[1])->d[i] = r.a.b; H(out[2])->d[i] = r.b.a; H(out[3])->d[i] = r.b.b;
    }
    return out;
  }
  void fold_circle_into_line_range(const std::array<Col, 4>& dst, const std::array<Col, 4>& src, uint32_t log, size_t off, size_t n_out,
                                   sb::QM31 alpha) override {
    orc::CircleDomain dom = orc::canonic_domain(log);
    orc::QM31 al = oq(alpha), a2 = orc::qmul(al, al);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n_out; i++) {
      orc::Pt p = dom.at(orc::bit_reverse((uint32_t)(2 * (off + i)), log));
      orc::QM31 a = orc::qfrom(H(src[0])->d[2 * i], H(src[1])->d[2 * i], H(src[2])->d[2 * i], H(src[3])->d[2 * i]);
      orc::QM31 b = orc::qfrom(H(src[0])->d[2 * i + 1], H(src[1])->d[2 * i + 1], H(src[2])->d[2 * i + 1], H(src[3])->d[2 * i + 1]);
      orc::QM31 f = orc::qadd(orc::qadd(a, b), orc::qmul(al, orc::qmulm(orc::qsub(a, b), orc::minv(p.y))));
      orc::QM31 d = orc::qfrom(H(dst[0])->d[i], H(dst[1])->d[i], H(dst[2])->d[i], H(dst[3])->d[i]);
      orc::QM31 r = orc::qadd(orc::qmul(d, a2), f);
      H(dst[0])->d[i] = r.a.a; H(dst[1])->d[i] = r.a.b; H(dst[2])->d[i] = r.b.a; H(dst[3])->d[i] = r.b.b;
    }
  }
  std::array<Col, 4> accumulate_quotients_range(uint32_t log, size_t row_off, size_t n_rows, const std::vector<Col>& cols, sb::QM31 rc,
                                                const SampleBatchesFlat& f) override {
    std::array<Col, 4> out;
    for (int k = 0; k < 4; k++) out[k] = new HCol(n_rows);
    // quotient_constants (pcs/quotients.rs): per sample the line coefficients already multiplied by their power of alpha,
    // per batch alpha^len — computed once per call, as upstream does, with the oracle's own arithmetic
    struct LC { orc::QM31 a, b, c; uint32_t col; };
    struct Batch { orc::CM31 prx, pry, pix, piy; orc::QM31 coef; std::vector<LC> lc; };
    std::vector<Batch> batches(f.sizes.size());
    orc::QM31 alpha = oq(rc);
    size_t e = 0;
    for (size_t b = 0; b < f.sizes.size(); b++) {
      const uint32_t* q = &f.points[8 * b];
      orc::QM31 sx = orc::qfrom(q[0], q[1], q[2], q[3]), sy = orc::qfrom(q[4], q[5], q[6], q[7]);
      Batch& B = batches[b];
      B.prx = sx.a; B.pix = sx.b; B.pry = sy.a; B.piy = sy.b;
      orc::QM31 al = orc::qfromm(1), c = orc::qsub(orc::qconj(sy), sy);
      for (uint32_t j = 0; j < f.sizes[b]; j++, e++) {
        al = orc::qmul(al, alpha);
        orc::QM31 v = orc::qfrom(f.entry_vals[4 * e], f.entry_vals[4 * e + 1], f.entry_vals[4 * e + 2], f.entry_vals[4 * e + 3]);
        orc::QM31 a = orc::qsub(orc::qconj(v), v);
        orc::QM31 bb = orc::qsub(orc::qmul(v, c), orc::qmul(a, sy));
        B.lc.push_back({orc::qmul(al, a), orc::qmul(al, bb), orc::qmul(al, c), f.entry_cols[e]});
      }
      B.coef = orc::qpow(alpha, f.sizes[b]);
    }
    const orc::CircleDomain dom = orc::canonic_domain(log);
#pragma omp parallel for schedule(static)
    for (size_t r = 0; r < n_rows; r++) {
      orc::Pt p = dom.at(orc::bit_reverse((uint32_t)(row_off + r), log));
      orc::QM31 acc = orc::qfromm(0);
      for (const Batch& B : batches) {
        orc::CM31 den = orc::csub(orc::cmul(orc::csub(B.prx, orc::CM31{p.x, 0}), B.piy), orc::cmul(orc::csub(B.pry, orc::CM31{p.y, 0}), B.pix));
        orc::QM31 num = orc::qfromm(0);
        for (const LC& l : B.lc) {
          orc::QM31 value = orc::qmulm(l.c, H(cols[l.col])->d[r]);
          orc::QM31 lin = orc::qadd(orc::qmulm(l.a, p.y), l.b);
          num = orc::qadd(num, orc::qsub(value, lin));
        }
        acc = orc::qadd(orc::qmul(acc, B.coef), orc::qmulc(num, orc::cinv(den)));
      }
      H(out[0])->d[r] = acc.a.a; H(out[1])->d[r] = acc.a.b; H(out[2])->d[r] = acc.b.a; H(out[3])->d[r] = acc.b.b;
    }
    return out;
  }
  Col shift_prev(Col c, uint32_t trace_log) override {
    uint32_t e = trace_log + 1;
    size_t n = (size_t)1 << e, half = n / 2;
    HCol* o = new HCol(n);
    for (size_t row = 0; row < n; row++) {
      size_t idx = orc::bit_reverse((uint32_t)row, e);
      size_t pidx = idx < half ? (idx + half - 1) % half : ((idx - half + 1) % half) + half;
      o->d[row] = H(c)->d[orc::bit_reverse((uint32_t)pidx, e)];
    }
    return o;
  }
  void accumulate_col(Col dst, Col src) override { for (size_t i = 0; i < H(dst)->n; i++) H(dst)->d[i] = orc::madd(H(dst)->d[i], H(src)->d[i]); }
  void prefix_sum(Col c) override { orc::prefix_sum_bitrev(H(c)->d, lg2(H(c)->n)); }
  std::vector<Col> logup_generate_sel(int comp, const std::vector<Col>& main, const InteractionElements& el,
                                      const std::vector<uint8_t>& want) override {
    size_t n = H(main[0])->n << LOG_N_LANES;
    std::vector<HCol*> out(4 * N_LOGUP_COLS[comp], nullptr);
    for (size_t i = 0; i < out.size(); i++) if (want[i]) out[i] = new HCol(n);
#pragma omp parallel for schedule(static)
    for (size_t row = 0; row < n; row++) {
      OrcLogupGen e;
      e.main = &main; e.out = &out; e.el = &el; e.row = row; e.shift = LOG_N_LANES;
      eval_component(comp, e);
    }
    return std::vector<Col>(out.begin(), out.end());
  }
  void eval_constraints_range(int comp, uint32_t log_size, size_t row_off, size_t n_rows, const std::vector<Col>& m, const std::vector<Col>& it,
                              const std::array<Col, 4>& prev, Col is_first, const InteractionElements& el, sb::QM31 total,
                              const std::vector<sb::QM31>& coeffs, const std::array<Col, 4>& acc) override {
    eval_rows(comp, log_size, row_off, n_rows, m, it, &prev, is_first, el, total, coeffs, acc);
  }
  void eval_rows(int comp, uint32_t log_size, size_t row_off, size_t n_rows, const std::vector<Col>& m, const std::vector<Col>& it,
                 const std::array<Col, 4>* prev, Col is_first, const InteractionElements& el, sb::QM31 total,
                 const std::vector<sb::QM31>& coeffs, const std::array<Col, 4>& acc) {
    uint32_t e = log_size + 1;
    size_t half = (size_t)1 << log_size;
    // 1 / coset_vanishing(CanonicCoset(log_size).coset, eval_domain.at(i)), i = 0,1 (the translation cancels for canonic cosets)
    uint32_t dinv[2];
    for (uint32_t i = 0; i < 2; i++) {
      uint32_t x = orc::canonic_domain(e).at(i).x;
      for (uint32_t k = 1; k < log_size; k++) x = orc::double_x(x);
      dinv[i] = orc::minv(x);
    }
#pragma omp parallel for schedule(static)
    for (size_t row = 0; row < n_rows; row++) {
      size_t grow = row + row_off;
      size_t idx = orc::bit_reverse((uint32_t)grow, e);
      size_t pidx = idx < half ? (idx + half - 1) % half : ((idx - half + 1) % half) + half;
      OrcDomainEval ev;
      ev.main = &m; ev.inter = &it; ev.prev = prev; ev.is_first_col = H(is_first)->d; ev.el = &el; ev.coeff = &coeffs; ev.total = oq(total);
      ev.row = row; ev.prev_row = orc::bit_reverse((uint32_t)pidx, e);
      eval_component(comp, ev);
      orc::QM31 r = orc::qmulm(ev.row_res, dinv[grow >> log_size]);
      H(acc[0])->d[row] = orc::madd(H(acc[0])->d[row], r.a.a); H(acc[1])->d[row] = orc::madd(H(acc[1])->d[row], r.a.b);
      H(acc[2])->d[row] = orc::madd(H(acc[2])->d[row], r.b.a); H(acc[3])->d[row] = orc::madd(H(acc[3])->d[row], r.b.b);
    }
  }
};

// AssertEvaluator: every constraint vanishes on every row of the trace domain (upstream constraint_framework/assert.rs;
// the reference's 13 `test_*_constraints` tests, e.g. components/processor/component.rs:178-236).
struct OrcAssertEval : OrcDomainEval {
  std::string* err;
  // message: "constraint K row R left (a + bi) + (c + di)u" — the value as QM31's Display prints it upstream (pinned by the
  // should_panic strings of the reference's memory/component.rs tests)
  static std::string show(orc::QM31 v) {
    return "(" + std::to_string(v.a.a) + " + " + std::to_string(v.a.b) + "i) + (" + std::to_string(v.b.a) + " + " + std::to_string(v.b.b) + "i)u";
  }
  size_t report_row = 0;
  void fail(orc::QM31 v) { if (err->empty()) *err = "constraint " + std::to_string(k) + " row " + std::to_string(report_row) + " left " + show(v); }
  void add(F c) { if (c.v != 0) fail(orc::qfromm(c.v)); k++; }
  void add(EF c) { if (!orc::qeq(c.v, orc::qfromm(0))) fail(c.v); k++; }
  LogupState<OrcAssertEval> lg2s;
  void relation(int rel, EF num, const F* vals, int n) { lg2s.push(num, ocombine(el->rel[rel], vals, n)); }
  void ext_mask_last(EF& pv, EF& cur) { pv = ext_at(lg2s.n - 1, prev_row); cur = ext_at(lg2s.n - 1, row); }
  void finalize_logup() { lg2s.finalize(*this); }
};

// Same walk, but the value of every constraint at one row is recorded instead of asserted (tests/air_model.py compares
// them with a Python transcription of the reference's evaluate() bodies on tables that do NOT satisfy the AIR).
struct OrcRecordEval : OrcDomainEval {
  std::vector<orc::QM31>* rec;
  void add(F c) { rec->push_back(orc::qfromm(c.v)); k++; }
  void add(EF c) { rec->push_back(c.v); k++; }
  LogupState<OrcRecordEval> lg2s;
  void relation(int rel, EF num, const F* vals, int n) { lg2s.push(num, ocombine(el->rel[rel], vals, n)); }
  void ext_mask_last(EF& pv, EF& cur) { pv = ext_at(lg2s.n - 1, prev_row); cur = ext_at(lg2s.n - 1, row); }
  void finalize_logup() { lg2s.finalize(*this); }
};

// constraint_framework::assert_constraints for one component: every constraint must vanish on every row of the trace domain.
// Upstream evaluates the polynomials back on the trace domain, bit-reverses to NATURAL order and walks `row` over that, so
// the first failure it reports is the first in natural order; the same walk here ("row" in the message is that index).
static std::string assert_component(OrcBackend& B, int c, const Table& table, const InteractionElements& el) {
  std::vector<Col> compact, full;
  for (auto& col : table.cols) { compact.push_back(B.from_host(col.data(), col.size())); full.push_back(B.broadcast16(compact.back())); }
  sb::QM31 claimed;
  std::vector<Col> inter = B.logup_generate(c, compact, el, claimed);
  uint32_t ls = table.log_size;
  size_t n = (size_t)1 << ls;
  Col isf = B.gen_is_first(ls);
  std::vector<sb::QM31> coeffs(N_CONSTRAINTS[c], sb::q_fromm(1));
  std::string err;
  for (size_t nat = 0; nat < n && err.empty(); nat++) {
    size_t row = orc::bit_reverse((uint32_t)nat, ls);   // storage index of natural row `nat`
    size_t idx = nat;
    size_t half = n / 2;
    size_t ci = idx < half ? 2 * idx : 2 * n - 1 - 2 * idx;   // circle-domain index -> coset index
    size_t pc = (ci + n - 1) % n;                              // coset-order predecessor
    size_t pidx = orc::coset_to_domain_index(pc, ls);
    OrcAssertEval ev;
    ev.main = &full; ev.inter = &inter; ev.is_first_col = H(isf)->d; ev.el = &el; ev.coeff = &coeffs; ev.total = oq(claimed);
    ev.row = row; ev.prev_row = orc::bit_reverse((uint32_t)pidx, ls); ev.err = &err; ev.report_row = nat;
    eval_component(c, ev);
  }
  for (Col x : compact) B.free_col(x);
  for (Col x : full) B.free_col(x);
  for (Col x : inter) B.free_col(x);
  B.free_col(isf);
  return err;
}

char* dupstr(const std::string& s) { char* p = (char*)malloc(s.size() + 1); memcpy(p, s.c_str(), s.size() + 1); return p; }
thread_local std::string g_err;
std::string g_err_shared;

}  // namespace

extern "C" {

const char* orc_last_error() { return g_err.empty() ? g_err_shared.c_str() : g_err.c_str(); }

// Full CPU proof with the oracle backend; returns the proof JSON (malloc'd) or NULL.  verify != 0 also runs the verifier.
char* orc_prove_json(const char* code, const uint8_t* input, size_t input_len, uint32_t log_max_rows, int verify) {
  try {
    std::vector<uint32_t> program = compile(code);
    Machine vm(program, std::vector<uint8_t>(input, input + input_len));
    vm.execute();
    OrcBackend B;
    ProverConfig cfg;
    cfg.log_max_rows = log_max_rows;
    ProveResult r = prove_brainfuck(B, program, vm.trace, cfg);
    if (getenv("ORC_STAGES")) for (auto& kv : r.times.ms) fprintf(stderr, "orc stage %-28s %10.1f ms\n", kv.first.c_str(), kv.second);
    if (verify) verify_brainfuck(r.proof, cfg);
    return dupstr(proof_to_json(r.proof));
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}

// The same driver with a PreprocessedCache kept across `n` proofs on one backend (prover.hpp: the preprocessed tree is
// program-independent).  Returns the proofs' JSON, one per line, then a last line "fills hits"; NULL on failure.
char* orc_prove_sequence_cached_json(int n, const char* const* codes, const uint8_t* const* inputs, const size_t* input_lens,
                                     const uint32_t* log_max_rows, int verify) {
  try {
    OrcBackend B;
    PreprocessedCache cache;
    std::string out;
    for (int i = 0; i < n; i++) {
      std::vector<uint32_t> program = compile(codes[i]);
      Machine vm(program, std::vector<uint8_t>(inputs[i], inputs[i] + input_lens[i]));
      vm.execute();
      ProverConfig cfg;
      cfg.log_max_rows = log_max_rows[i];
      ProveResult r = prove_brainfuck(B, program, vm.trace, cfg, nullptr, &cache);
      if (verify) verify_brainfuck(r.proof, cfg);
      out += proof_to_json(r.proof) + "\n";
    }
    out += std::to_string(cache.fills) + " " + std::to_string(cache.hits);
    cache.release(B);
    return dupstr(out);
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}

// Test hook: the column -> owner-rank assignment of the sharded driver (prover_sharded.hpp assign_owners), so that its balance
// and determinism can be pinned from Python (tests/test_sharded_prover.py).
void orc_assign_owners(const uint32_t* logs, size_t n, int world, int32_t* out) {
  std::vector<int> o = assign_owners(std::vector<uint32_t>(logs, logs + n), world);
  for (size_t i = 0; i < n; i++) out[i] = o[i];
}

// The sharded driver on `world` in-process ranks (threads).  Every rank must produce the same proof; returns rank 0's JSON,
// or NULL if any rank failed or the ranks disagree.
char* orc_prove_sharded_json(const char* code, const uint8_t* input, size_t input_len, uint32_t log_max_rows, int world, int verify) {
  try {
    std::vector<uint32_t> program = compile(code);
    Machine vm(program, std::vector<uint8_t>(input, input + input_len));
    vm.execute();
    ProverConfig cfg;
    cfg.log_max_rows = log_max_rows;
    if (getenv("ORC_SHARD_MIN_LOG")) cfg.shard_min_log = (uint32_t)atoi(getenv("ORC_SHARD_MIN_LOG"));   // tests: replicate small columns
    ThreadComm comm(world);
    std::vector<std::string> json(world), err(world);
    std::vector<std::thread> th;
    for (int r = 0; r < world; r++)
      th.emplace_back([&, r] {
        try {
          OrcBackend B;
          B.my_rank = r;
          if (world > 1) B.comm = &comm;
          ProveResult res = prove_brainfuck_sharded(B, program, TraceSource([&] {
            TraceInput in;
            in.regs = vm.trace.data(); in.n = vm.trace.size(); in.stats = trace_stats(in.regs, in.n);   // recomputed, not the VM's own
            return in;
          }), cfg);
          if (verify) verify_brainfuck(res.proof, cfg);
          json[r] = proof_to_json(res.proof);
        } catch (const std::exception& e) {
          err[r] = e.what();
          // a failed rank would dead-lock the others at the next barrier: make every later barrier a no-op
          std::lock_guard<std::mutex> lk(comm.mu);
          comm.world = 1 << 30;
          comm.generation++;
          comm.cv.notify_all();
        }
      });
    for (auto& t : th) t.join();
    for (int r = 0; r < world; r++) if (!err[r].empty()) { g_err_shared = "rank " + std::to_string(r) + ": " + err[r]; g_err.clear(); return nullptr; }
    for (int r = 1; r < world; r++) if (json[r] != json[0]) { g_err_shared = "ranks disagree on the proof"; g_err.clear(); return nullptr; }
    return dupstr(json[0]);
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
void orc_free(char* p) { free(p); }

// VM + tables as text for the host-logic tests: "steps;output bytes;log_sizes;program;first memory cells"
char* orc_vm_summary(const char* code, const uint8_t* input, size_t input_len) {
  try {
    std::vector<uint32_t> program = compile(code);
    Machine vm(program, std::vector<uint8_t>(input, input + input_len));
    vm.execute();
    auto tables = build_tables(vm.trace, program);
    std::ostringstream o;
    o << vm.trace.size() << ";";
    for (uint8_t b : vm.output) o << (int)b << ",";
    o << ";";
    for (auto& t : tables) o << t.log_size << ",";
    o << ";";
    for (uint32_t w : program) o << w << ",";
    o << ";";
    for (int i = 0; i < 5 && i < (int)vm.ram.size(); i++) o << vm.ram[i] << ",";
    return dupstr(o.str());
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}

// The VM's register rows (n x 7 words, registers.rs order) and two statements of their statistics: the counters the VM keeps
// while it runs and trace_stats() recomputed from the finished trace (16 words each, layout of sc_trace_stats_host).
// out == NULL: returns the number of rows only.
size_t orc_vm_registers(const char* code, const uint8_t* input, size_t input_len, uint32_t* out, uint64_t* stats_vm, uint64_t* stats_recomputed,
                        uint32_t* program_out, size_t* program_len) {
  try {
    std::vector<uint32_t> program = compile(code);
    Machine vm(program, std::vector<uint8_t>(input, input + input_len));
    vm.execute();
    if (program_len) *program_len = program.size();
    if (program_out) memcpy(program_out, program.data(), program.size() * 4);
    auto pack = [](const TraceStats& st, uint64_t* w) {
      memset(w, 0, 16 * sizeof(uint64_t));
      w[0] = st.steps; w[1] = st.memory_rows;
      for (int k = 0; k < 8; k++) w[2 + k] = st.op_count[k];
      w[10] = st.zero_ci; w[11] = st.zero_ci_index; w[12] = st.max_mp; w[13] = st.max_ip;
    };
    if (stats_vm) pack(vm.stats, stats_vm);
    if (stats_recomputed) pack(trace_stats(vm.trace.data(), vm.trace.size()), stats_recomputed);
    if (out) memcpy(out, vm.trace.data(), vm.trace.size() * sizeof(Registers));
    return vm.trace.size();
  } catch (const std::exception& e) {
    g_err = e.what();
    return 0;
  }
}

// Dumps one component table: rows x cols, row-major, into out (caller sized via a first call with out == NULL).
size_t orc_table_dump(const char* code, const uint8_t* input, size_t input_len, int comp, uint32_t* out, uint32_t* n_cols) {
  std::vector<uint32_t> program = compile(code);
  Machine vm(program, std::vector<uint8_t>(input, input + input_len));
  vm.execute();
  auto tables = build_tables(vm.trace, program);
  const Table& t = tables[comp];
  *n_cols = (uint32_t)t.cols.size();
  if (out)
    for (size_t r = 0; r < t.rows(); r++)
      for (size_t c = 0; c < t.cols.size(); c++) out[r * t.cols.size() + c] = t.cols[c][r];
  return t.rows();
}

// One component table built from an explicit register list (n rows of clk, ip, ci, ni, mp, mv, mvi, in clk order) and program
// words: the shape of the reference's table.rs unit tests, which start from hand-written `Registers`.  Returns the number of
// rows (0 and a message in orc_last_error on failure).
size_t orc_table_from_registers(const uint32_t* regs, size_t n, const uint32_t* program, size_t n_program, int comp, uint32_t* out,
                                uint32_t* n_cols) {
  try {
    std::vector<Registers> tr(n);
    for (size_t i = 0; i < n; i++) { const uint32_t* r = regs + 7 * i; tr[i].clk = r[0]; tr[i].ip = r[1]; tr[i].ci = r[2]; tr[i].ni = r[3]; tr[i].mp = r[4]; tr[i].mv = r[5]; tr[i].mvi = r[6]; }
    std::vector<uint32_t> prog(program, program + n_program);
    Table t = build_table(comp, tr, prog);
    *n_cols = (uint32_t)t.cols.size();
    if (out)
      for (size_t r = 0; r < t.rows(); r++)
        for (size_t c = 0; c < t.cols.size(); c++) out[r * t.cols.size() + c] = t.cols[c][r];
    return t.rows();
  } catch (const std::exception& e) {
    g_err = e.what();
    return 0;
  }
}

// assert_constraints for every component of a program, with shifted-dummy or channel-drawn lookup elements.
// Returns NULL on success, else a malloc'd message.
char* orc_assert_constraints(const char* code, const uint8_t* input, size_t input_len, int dummy_elements) {
  try {
    std::vector<uint32_t> program = compile(code);
    Machine vm(program, std::vector<uint8_t>(input, input + input_len));
    vm.execute();
    auto tables = build_tables(vm.trace, program);
    InteractionElements el;
    Channel ch;
    if (dummy_elements) {
      // LookupElements::dummy() (z = alpha^i = 1) makes some denominators vanish (memory/component.rs:225-227); keep the
      // all-ones alpha powers but move z away
      for (int r = 0; r < 3; r++) { el.rel[r].z = sb::q_make(5, 6, 7, 8); for (int i = 0; i < 7; i++) el.rel[r].alpha_pow[i] = sb::q_fromm(1); }
    } else {
      el = draw_elements(ch);
    }
    OrcBackend B;
    B.precompute_twiddles(8);
    for (int c = 0; c < N_COMPONENTS; c++) {
      std::string err = assert_component(B, c, tables[c], el);
      if (!err.empty()) return dupstr(std::string(COMPONENT_NAMES[c]) + ": " + err);
    }
    return nullptr;
  } catch (const std::exception& e) {
    return dupstr(std::string("exception: ") + e.what());
  }
}

// LogupTraceGenerator output for ONE component given its table (rows x cols, row-major) and explicit lookup elements
// (96 words: 3 x {z[4], alpha_powers[7][4]}, the C ABI's layout): out receives 4 * N_LOGUP_COLS columns of 16 * n_rows
// words each, concatenated; claimed[4] the claimed sum.  Returns the number of columns written, 0 on error.
size_t orc_logup_table(int comp, const uint32_t* rows, size_t n_rows, size_t n_cols, const uint32_t* elements, uint32_t* out,
                       uint32_t* claimed) {
  try {
    if (comp < 0 || comp >= N_COMPONENTS || (int)n_cols != N_MAIN_COLS[comp]) throw std::runtime_error("bad component / column count");
    InteractionElements el;
    static_assert(sizeof(InteractionElements) == 96 * 4, "elements layout");
    memcpy(&el, elements, sizeof(el));
    OrcBackend B;
    std::vector<Col> compact;
    for (size_t c = 0; c < n_cols; c++) {
      std::vector<uint32_t> col(n_rows);
      for (size_t r = 0; r < n_rows; r++) col[r] = rows[r * n_cols + c];
      compact.push_back(B.from_host(col.data(), col.size()));
    }
    sb::QM31 sum;
    std::vector<Col> inter = B.logup_generate(comp, compact, el, sum);
    const size_t len = 16 * n_rows;
    for (size_t k = 0; k < inter.size(); k++) B.read(inter[k], 0, len, out + k * len);
    claimed[0] = sum.a.a; claimed[1] = sum.a.b; claimed[2] = sum.b.a; claimed[3] = sum.b.b;
    for (Col x : compact) B.free_col(x);
    for (Col x : inter) B.free_col(x);
    return inter.size();
  } catch (const std::exception& e) {
    g_err = e.what();
    return 0;
  }
}

// ComponentProver::evaluate_constraint_quotients_on_domain for ONE component given its table: interpolate the main trace,
// generate and interpolate the LogUp columns, extend everything (and IsFirst) to CanonicCoset(log_size + 1), evaluate the
// constraints with the given per-constraint coefficients (n_constraints x 4 words) into a zeroed accumulator.
// out: 4 coordinate columns of 32 * n_rows words.  Returns the claimed sum through claimed[4]; 0 on error, else 4.
size_t orc_constraints_table(int comp, const uint32_t* rows, size_t n_rows, size_t n_cols, const uint32_t* elements,
                             const uint32_t* coeffs, uint32_t* out, uint32_t* claimed) {
  try {
    if (comp < 0 || comp >= N_COMPONENTS || (int)n_cols != N_MAIN_COLS[comp]) throw std::runtime_error("bad component / column count");
    InteractionElements el;
    memcpy(&el, elements, sizeof(el));
    OrcBackend B;
    const uint32_t ls = OrcBackend::lg2(n_rows) + LOG_N_LANES;
    B.precompute_twiddles(ls + 2);
    std::vector<Col> compact, full;
    for (size_t c = 0; c < n_cols; c++) {
      std::vector<uint32_t> col(n_rows);
      for (size_t r = 0; r < n_rows; r++) col[r] = rows[r * n_cols + c];
      compact.push_back(B.from_host(col.data(), col.size()));
      full.push_back(B.broadcast16(compact.back()));
    }
    sb::QM31 sum;
    std::vector<Col> inter = B.logup_generate(comp, compact, el, sum);
    Col isf = B.gen_is_first(ls);
    B.interpolate(full);
    B.interpolate(inter);
    B.interpolate({isf});
    std::vector<Col> main_lde = B.evaluate(full, 1), inter_lde = B.evaluate(inter, 1), isf_lde = B.evaluate({isf}, 1);
    std::vector<sb::QM31> cf(N_CONSTRAINTS[comp]);
    for (int k = 0; k < N_CONSTRAINTS[comp]; k++) cf[k] = sb::q_make(coeffs[4 * k], coeffs[4 * k + 1], coeffs[4 * k + 2], coeffs[4 * k + 3]);
    const size_t len = (size_t)2 << ls;
    std::array<Col, 4> acc;
    for (int k = 0; k < 4; k++) acc[k] = B.zeros(len);
    B.eval_constraints(comp, ls, main_lde, inter_lde, isf_lde[0], el, sum, cf, acc);
    for (int k = 0; k < 4; k++) B.read(acc[k], 0, len, out + k * len);
    claimed[0] = sum.a.a; claimed[1] = sum.a.b; claimed[2] = sum.b.a; claimed[3] = sum.b.b;
    for (auto* v : {&compact, &full, &inter, &main_lde, &inter_lde, &isf_lde}) for (Col x : *v) B.free_col(x);
    B.free_col(isf);
    for (Col x : acc) B.free_col(x);
    return 4;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 0;
  }
}

// The Fiat–Shamir channel (csrc/host/channel.hpp, host logic shared by the product and the oracle) driven by a script, so
// that a test can pin it against an independent model.  Ops: 'R' + 32 bytes mix_root; 'F' + u32 n + 16 n bytes mix_felts;
// 'U' + 8 bytes mix_u64; 'D' draw_felt (emits 16 bytes); 'S' + u32 n draw_felts(n) (emits 16 n bytes); 'B' draw_random_bytes
// (emits 32 bytes); 'Z' trailing_zeros (emits 4 bytes).  The final digest (32 bytes) is appended.  Returns bytes written.
size_t orc_channel_script(const uint8_t* script, size_t len, uint8_t* out, size_t cap) {
  Channel ch;
  size_t w = 0;
  auto emit = [&](const void* p, size_t n) { if (w + n <= cap) memcpy(out + w, p, n); w += n; };
  auto rd32 = [&](size_t at) { uint32_t v; memcpy(&v, script + at, 4); return v; };
  size_t i = 0;
  while (i < len) {
    const uint8_t op = script[i++];
    if (op == 'R') { Hash h; memcpy(h.data(), script + i, 32); i += 32; ch.mix_root(h); }
    else if (op == 'F') {
      uint32_t n = rd32(i); i += 4;
      std::vector<sb::QM31> f(n);
      for (uint32_t k = 0; k < n; k++, i += 16) f[k] = sb::q_make(rd32(i), rd32(i + 4), rd32(i + 8), rd32(i + 12));
      ch.mix_felts(f);
    }
    else if (op == 'U') { uint64_t v; memcpy(&v, script + i, 8); i += 8; ch.mix_u64(v); }
    else if (op == 'D') { sb::QM31 q = ch.draw_felt(); uint32_t v[4] = {q.a.a, q.a.b, q.b.a, q.b.b}; emit(v, 16); }
    else if (op == 'S') {
      uint32_t n = rd32(i); i += 4;
      for (sb::QM31 q : ch.draw_felts(n)) { uint32_t v[4] = {q.a.a, q.a.b, q.b.a, q.b.b}; emit(v, 16); }
    }
    else if (op == 'B') { Hash h = ch.draw_random_bytes(); emit(h.data(), 32); }
    else if (op == 'Z') { uint32_t z = ch.trailing_zeros(); emit(&z, 4); }
    else return 0;
  }
  emit(ch.digest.data(), 32);
  return w;
}

// Values of all constraints of ONE component at natural trace-domain row `nat`, for an explicit table and explicit lookup
// elements; the LogUp columns and the claimed sum are generated from the table.  out: N_CONSTRAINTS x 4 words.  Returns the
// number of constraints, 0 on error.
size_t orc_constraint_values(int comp, const uint32_t* rows, size_t n_rows, size_t n_cols, const uint32_t* elements, size_t nat,
                             uint32_t* out) {
  try {
    if (comp < 0 || comp >= N_COMPONENTS || (int)n_cols != N_MAIN_COLS[comp]) throw std::runtime_error("bad component / column count");
    InteractionElements el;
    memcpy(&el, elements, sizeof(el));
    OrcBackend B;
    std::vector<Col> compact, full;
    for (size_t c = 0; c < n_cols; c++) {
      std::vector<uint32_t> col(n_rows);
      for (size_t r = 0; r < n_rows; r++) col[r] = rows[r * n_cols + c];
      compact.push_back(B.from_host(col.data(), col.size()));
      full.push_back(B.broadcast16(compact.back()));
    }
    sb::QM31 claimed;
    std::vector<Col> inter = B.logup_generate(comp, compact, el, claimed);
    const uint32_t ls = OrcBackend::lg2(n_rows) + LOG_N_LANES;
    const size_t n = (size_t)1 << ls, half = n / 2;
    if (nat >= n) throw std::runtime_error("row out of range");
    Col isf = B.gen_is_first(ls);
    std::vector<sb::QM31> coeffs(N_CONSTRAINTS[comp], sb::q_fromm(1));
    std::vector<orc::QM31> rec;
    const size_t ci = nat < half ? 2 * nat : 2 * n - 1 - 2 * nat;   // circle-domain index -> coset index
    const size_t pidx = orc::coset_to_domain_index((ci + n - 1) % n, ls);
    OrcRecordEval ev;
    ev.main = &full; ev.inter = &inter; ev.is_first_col = H(isf)->d; ev.el = &el; ev.coeff = &coeffs; ev.total = oq(claimed);
    ev.row = orc::bit_reverse((uint32_t)nat, ls); ev.prev_row = orc::bit_reverse((uint32_t)pidx, ls); ev.rec = &rec;
    eval_component(comp, ev);
    for (size_t k = 0; k < rec.size(); k++) { out[4 * k] = rec[k].a.a; out[4 * k + 1] = rec[k].a.b; out[4 * k + 2] = rec[k].b.a; out[4 * k + 3] = rec[k].b.b; }
    for (auto* v : {&compact, &full, &inter}) for (Col x : *v) B.free_col(x);
    B.free_col(isf);
    return rec.size();
  } catch (const std::exception& e) {
    g_err = e.what();
    return 0;
  }
}

// assert_constraints on ONE component given its table explicitly (rows x cols, row-major): the shape of the reference's
// negative component tests, which corrupt a table by hand.  elements: 0 = drawn from a fresh channel (MemoryElements::draw
// on Blake2sChannel::default()), 2 = LookupElements::dummy() (z = 1, every alpha power = 1).  NULL = all constraints hold.
char* orc_assert_table(int comp, const uint32_t* rows, size_t n_rows, size_t n_cols, int elements) {
  try {
    if (comp < 0 || comp >= N_COMPONENTS || (int)n_cols != N_MAIN_COLS[comp]) return dupstr("bad component / column count");
    Table t;
    t.cols.assign(n_cols, ColVec(n_rows));
    for (size_t r = 0; r < n_rows; r++) for (size_t c = 0; c < n_cols; c++) t.cols[c][r] = rows[r * n_cols + c];
    t.log_size = OrcBackend::lg2(n_rows) + LOG_N_LANES;
    InteractionElements el;
    if (elements == 2) {
      for (int r = 0; r < 3; r++) { el.rel[r].z = sb::q_fromm(1); for (int i = 0; i < 7; i++) el.rel[r].alpha_pow[i] = sb::q_fromm(1); }
    } else {
      Channel ch;
      el = draw_elements(ch);
    }
    OrcBackend B;
    B.precompute_twiddles(8);
    std::string err = assert_component(B, comp, t, el);
    return err.empty() ? nullptr : dupstr(err);
  } catch (const std::exception& e) {
    return dupstr(std::string("exception: ") + e.what());
  }
}

}  // extern "C"
