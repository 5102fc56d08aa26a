// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_field.h header).
//
// CPU `Backend` for the protocol driver: every device op of the product replaced by the scalar restatements in this
// directory (own field arithmetic, own FFT/Merkle/quotient/fold loops, own row evaluators).  The driver itself
// (stwo-brainfuck_b200/csrc/host/{prover,verifier}.hpp) and the AIR definition (host/air.hpp — the restatement of the
// reference's components/<name>/component.rs `evaluate` bodies) are shared with the product on purpose: they are host
// logic in the reference too; what this oracle checks is that the CUDA kernels compute the same field elements.
// Proof parity test: tests/test_prover_gpu.py compares the JSON of both provers byte for byte.
#include <memory>
#include "../stwo-brainfuck_b200/csrc/host/verifier.hpp"
#include "orc_ops.h"

using namespace sbf;

namespace {

struct HCol { std::vector<uint32_t> v; };
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
  std::vector<HCol*>* out;
  const InteractionElements* el;
  size_t row, shift;
  int col = 0, batch = 0;
  orc::QM31 cum = orc::qfromm(0);
  F next() { return {H((*main)[col++])->v[row >> shift]}; }
  F is_first() { return {0}; }
  F cst(uint32_t c) { return {c}; }
  EF ef(F x) { return {orc::qfromm(x.v)}; }
  EF ef_neg_one() { return {orc::qfromm(orc::P - 1)}; }
  void add(F) {}
  void add(EF) {}
  void relation(int rel, EF num, const F* vals, int n) {
    OFq den = ocombine(el->rel[rel], vals, n);
    cum = orc::qadd(cum, orc::qmul(num.v, orc::qinv(den.v)));
    (*out)[4 * batch]->v[row] = cum.a.a; (*out)[4 * batch + 1]->v[row] = cum.a.b;
    (*out)[4 * batch + 2]->v[row] = cum.b.a; (*out)[4 * batch + 3]->v[row] = cum.b.b;
    batch++;
  }
  void finalize_logup() {}
};

struct OrcDomainEval {
  typedef OFm F;
  typedef OFq EF;
  const std::vector<Col>*main, *inter;
  const uint32_t* is_first_col;
  const InteractionElements* el;
  const std::vector<sb::QM31>* coeff;
  orc::QM31 total;
  size_t row, prev_row;
  int col = 0, k = 0;
  orc::QM31 row_res = orc::qfromm(0);
  LogupState<OrcDomainEval> lg;
  F next() { return {H((*main)[col++])->v[row]}; }
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
    return {orc::qfrom(H((*inter)[4 * b])->v[r], H((*inter)[4 * b + 1])->v[r], H((*inter)[4 * b + 2])->v[r], H((*inter)[4 * b + 3])->v[r])};
  }
  EF ext_mask_cur(int b) { return ext_at(b, row); }
  void ext_mask_last(EF& prev, EF& cur) { prev = ext_at(lg.n - 1, prev_row); cur = ext_at(lg.n - 1, row); }
  void finalize_logup() { lg.finalize(*this); }
};

struct OrcBackend : Backend {
  std::vector<uint32_t> tw, itw;
  const char* name() const override { return "oracle"; }
  static uint32_t lg2(size_t n) { uint32_t l = 0; while (((size_t)1 << l) < n) l++; return l; }

  Col from_host(const uint32_t* v, size_t n) override { return new HCol{std::vector<uint32_t>(v, v + n)}; }
  Col broadcast16(Col c) override {
    HCol* o = new HCol;
    o->v.resize(H(c)->v.size() * 16);
    for (size_t i = 0; i < o->v.size(); i++) o->v[i] = H(c)->v[i >> 4];
    return o;
  }
  Col zeros(size_t n) override { return new HCol{std::vector<uint32_t>(n, 0)}; }
  size_t len(Col c) override { return H(c)->v.size(); }
  void read(Col c, size_t off, size_t n, uint32_t* out) override { memcpy(out, H(c)->v.data() + off, n * 4); }
  void free_col(Col c) override { delete H(c); }
  std::vector<uint32_t> gather(const std::vector<Col>& cols, const std::vector<size_t>& offsets, uint32_t words) override {
    std::vector<uint32_t> out(cols.size() * words);
    for (size_t i = 0; i < cols.size(); i++) memcpy(&out[i * words], H(cols[i])->v.data() + offsets[i], words * 4);
    return out;
  }

  void precompute_twiddles(uint32_t root_log) override { orc::precompute_twiddles(orc::coset_half_odds(root_log), tw, itw); }
  void interpolate(const std::vector<Col>& cols) override {
#pragma omp parallel for schedule(dynamic)
    for (size_t i = 0; i < cols.size(); i++) orc::interpolate(H(cols[i])->v.data(), lg2(H(cols[i])->v.size()), itw);
  }
  std::vector<Col> evaluate(const std::vector<Col>& coeffs, uint32_t log_blowup) override {
    std::vector<Col> out(coeffs.size());
#pragma omp parallel for schedule(dynamic)
    for (size_t i = 0; i < coeffs.size(); i++) {
      HCol* o = new HCol;
      o->v = H(coeffs[i])->v;
      o->v.resize(o->v.size() << log_blowup, 0);  // extend(): zero-pad at the end
      orc::evaluate(o->v.data(), lg2(o->v.size()), tw);
      out[i] = o;
    }
    return out;
  }
  std::vector<sb::QM31> eval_at_point(const std::vector<Col>& polys, const std::vector<QPoint>& pts) override {
    std::vector<sb::QM31> out(polys.size());
#pragma omp parallel for schedule(dynamic)
    for (size_t i = 0; i < polys.size(); i++)
      out[i] = sq(orc::eval_at_point(H(polys[i])->v.data(), lg2(H(polys[i])->v.size()), orc::QPt{oq(pts[i].x), oq(pts[i].y)}));
    return out;
  }
  std::vector<Col> merkle_commit(const std::vector<Col>& cols, Hash* root) override {
    std::vector<const uint32_t*> p;
    std::vector<uint32_t> logs;
    for (Col c : cols) { p.push_back(H(c)->v.data()); logs.push_back(lg2(H(c)->v.size())); }
    std::vector<std::vector<uint32_t>> layers;
    orc::merkle_commit(p.data(), logs.data(), cols.size(), layers);
    std::vector<Col> out;
    for (auto& l : layers) out.push_back(new HCol{l});
    if (root) memcpy(root->data(), layers[0].data(), 32);
    return out;
  }
  std::array<Col, 4> fold_line(const std::array<Col, 4>& src, uint32_t log, sb::QM31 alpha) override {
    std::array<Col, 4> out;
    const uint32_t* s[4]; uint32_t* d[4];
    for (int k = 0; k < 4; k++) { out[k] = zeros((size_t)1 << (log - 1)); s[k] = H(src[k])->v.data(); d[k] = H(out[k])->v.data(); }
    orc::fold_line(s, log, oq(alpha), d);
    return out;
  }
  void fold_circle_into_line(const std::array<Col, 4>& dst, const std::array<Col, 4>& src, uint32_t log, sb::QM31 alpha) override {
    const uint32_t* s[4]; uint32_t* d[4];
    for (int k = 0; k < 4; k++) { s[k] = H(src[k])->v.data(); d[k] = H(dst[k])->v.data(); }
    orc::fold_circle_into_line(s, log, oq(alpha), d);
  }
  std::array<Col, 4> accumulate_quotients(uint32_t log, const std::vector<Col>& cols, sb::QM31 rc, const SampleBatchesFlat& b) override {
    std::array<Col, 4> out;
    uint32_t* d[4];
    for (int k = 0; k < 4; k++) { out[k] = zeros((size_t)1 << log); d[k] = H(out[k])->v.data(); }
    std::vector<const uint32_t*> p;
    for (Col c : cols) p.push_back(H(c)->v.data());
    orc::accumulate_quotients(log, p.data(), (uint32_t)p.size(), oq(rc), b.points.data(), b.sizes.data(), b.entry_cols.data(),
                              b.entry_vals.data(), (uint32_t)b.sizes.size(), d);
    return out;
  }
  void accumulate(const std::array<Col, 4>& dst, const std::array<Col, 4>& src) override {
    for (int k = 0; k < 4; k++)
      for (size_t i = 0; i < H(dst[k])->v.size(); i++) H(dst[k])->v[i] = orc::madd(H(dst[k])->v[i], H(src[k])->v[i]);
  }
  uint64_t grind(const Hash& digest, uint32_t pow_bits) override { return orc::grind(digest.data(), pow_bits); }
  Col gen_is_first(uint32_t log_size) override { HCol* c = (HCol*)zeros((size_t)1 << log_size); c->v[0] = 1; return c; }

  std::vector<Col> logup_generate(int comp, const std::vector<Col>& main, const InteractionElements& el, sb::QM31& claimed) override {
    size_t n = H(main[0])->v.size() << LOG_N_LANES;
    uint32_t log = lg2(n);
    std::vector<HCol*> out(4 * N_LOGUP_COLS[comp]);
    for (auto& o : out) o = (HCol*)zeros(n);
#pragma omp parallel for schedule(static)
    for (size_t row = 0; row < n; row++) {
      OrcLogupGen e;
      e.main = &main; e.out = &out; e.el = &el; e.row = row; e.shift = LOG_N_LANES;
      eval_component(comp, e);
    }
    for (int k = 0; k < 4; k++) orc::prefix_sum_bitrev(out[out.size() - 4 + k]->v.data(), log);
    size_t b = out.size() - 4;
    claimed = sb::q_make(out[b]->v[1], out[b + 1]->v[1], out[b + 2]->v[1], out[b + 3]->v[1]);
    return std::vector<Col>(out.begin(), out.end());
  }
  void eval_constraints(int comp, uint32_t log_size, const std::vector<Col>& m, const std::vector<Col>& it, Col is_first,
                        const InteractionElements& el, sb::QM31 total, const std::vector<sb::QM31>& coeffs,
                        const std::array<Col, 4>& acc) override {
    uint32_t e = log_size + 1;
    size_t n = (size_t)1 << e, half = n / 2;
    // 1 / coset_vanishing(CanonicCoset(log_size).coset, eval_domain.at(i)), i = 0,1 (the translation cancels for canonic cosets)
    uint32_t dinv[2];
    for (uint32_t i = 0; i < 2; i++) {
      uint32_t x = orc::canonic_domain(e).at(i).x;
      for (uint32_t k = 1; k < log_size; k++) x = orc::double_x(x);
      dinv[i] = orc::minv(x);
    }
#pragma omp parallel for schedule(static)
    for (size_t row = 0; row < n; row++) {
      size_t idx = orc::bit_reverse((uint32_t)row, e);
      size_t pidx = idx < half ? (idx + half - 1) % half : ((idx - half + 1) % half) + half;
      OrcDomainEval ev;
      ev.main = &m; ev.inter = &it; ev.is_first_col = H(is_first)->v.data(); ev.el = &el; ev.coeff = &coeffs; ev.total = oq(total);
      ev.row = row; ev.prev_row = orc::bit_reverse((uint32_t)pidx, e);
      eval_component(comp, ev);
      orc::QM31 r = orc::qmulm(ev.row_res, dinv[row >> log_size]);
      H(acc[0])->v[row] = orc::madd(H(acc[0])->v[row], r.a.a); H(acc[1])->v[row] = orc::madd(H(acc[1])->v[row], r.a.b);
      H(acc[2])->v[row] = orc::madd(H(acc[2])->v[row], r.b.a); H(acc[3])->v[row] = orc::madd(H(acc[3])->v[row], r.b.b);
    }
  }
};

// AssertEvaluator: every constraint vanishes on every row of the trace domain (upstream constraint_framework/assert.rs;
// the reference's 13 `test_*_constraints` tests, e.g. components/processor/component.rs:178-236).
struct OrcAssertEval : OrcDomainEval {
  std::string* err;
  void add(F c) { if (c.v != 0 && err->empty()) *err = "constraint " + std::to_string(k) + " row " + std::to_string(row); k++; }
  void add(EF c) {
    if (!orc::qeq(c.v, orc::qfromm(0)) && err->empty()) *err = "constraint " + std::to_string(k) + " row " + std::to_string(row);
    k++;
  }
  LogupState<OrcAssertEval> lg2s;
  void relation(int rel, EF num, const F* vals, int n) { lg2s.push(num, ocombine(el->rel[rel], vals, n)); }
  void ext_mask_last(EF& prev, EF& cur) { prev = ext_at(lg2s.n - 1, prev_row); cur = ext_at(lg2s.n - 1, row); }
  void finalize_logup() { lg2s.finalize(*this); }
};

char* dupstr(const std::string& s) { char* p = (char*)malloc(s.size() + 1); memcpy(p, s.c_str(), s.size() + 1); return p; }
thread_local std::string g_err;

}  // namespace

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

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
    if (verify) verify_brainfuck(r.proof, cfg);
    return dupstr(proof_to_json(r.proof));
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
void orc_free(char* p) { free(p); }

// VM + tables as JSON-ish text for the host-logic tests: "steps;output-hex;log_sizes;program"
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

// assert_constraints for every component of a program, with dummy (all-ones) or seeded lookup elements.
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
      for (int r = 0; r < 3; r++) { el.rel[r].z = sb::q_fromm(1); for (int i = 0; i < 7; i++) el.rel[r].alpha_pow[i] = sb::q_fromm(1); }
      // LookupElements::dummy() makes the (clk 0, ...) denominators vanish for some tables; use a shifted z instead
      for (int r = 0; r < 3; r++) el.rel[r].z = sb::q_make(5, 6, 7, 8);
    } else {
      el = draw_elements(ch);
    }
    OrcBackend B;
    B.precompute_twiddles(8);
    for (int c = 0; c < N_COMPONENTS; c++) {
      std::vector<Col> compact, full;
      for (auto& col : tables[c].cols) { compact.push_back(B.from_host(col.data(), col.size())); full.push_back(B.broadcast16(compact.back())); }
      sb::QM31 claimed;
      std::vector<Col> inter = B.logup_generate(c, compact, el, claimed);
      uint32_t ls = tables[c].log_size;
      size_t n = (size_t)1 << ls;
      Col isf = B.gen_is_first(ls);
      std::vector<sb::QM31> coeffs(N_CONSTRAINTS[c], sb::q_fromm(1));
      std::string err;
      for (size_t row = 0; row < n; row++) {
        // trace-domain offset -1: same index arithmetic with eval_log == log_size (step 2^-1 -> handled as half-domain walk)
        size_t idx = orc::bit_reverse((uint32_t)row, ls);
        // coset order neighbour: convert circle-domain index -> coset index, subtract one, convert back
        size_t half = n / 2;
        size_t ci = idx < half ? 2 * idx : 2 * n - 1 - 2 * idx;   // circle-domain -> coset index (inverse of coset_to_domain_index)
        size_t pc = (ci + n - 1) % n;
        size_t pidx = orc::coset_to_domain_index(pc, ls);
        OrcAssertEval ev;
        ev.main = &full; ev.inter = &inter; ev.is_first_col = H(isf)->v.data(); ev.el = &el; ev.coeff = &coeffs; ev.total = oq(claimed);
        ev.row = row; ev.prev_row = orc::bit_reverse((uint32_t)pidx, ls); ev.err = &err;
        eval_component(c, ev);
        if (!err.empty()) return dupstr(std::string(COMPONENT_NAMES[c]) + ": " + err);
      }
      for (Col x : compact) B.free_col(x);
      for (Col x : full) B.free_col(x);
      for (Col x : inter) B.free_col(x);
      B.free_col(isf);
    }
    return nullptr;
  } catch (const std::exception& e) {
    return dupstr(std::string("exception: ") + e.what());
  }
}

}  // extern "C"
