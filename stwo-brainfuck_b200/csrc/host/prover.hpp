// Protocol driver: prove_brainfuck + stwo's prover::prove / CommitmentSchemeProver / FriProver / MerkleProver::decommit,
// written against `Backend` (backend.hpp).  Follows crates/brainfuck_prover/src/brainfuck_air/mod.rs:471-735 for the phase
// order and stwo-prover 0.1.1 @ 31e8dbc core/{prover/mod.rs, pcs/prover.rs, pcs/quotients.rs, fri.rs, vcs/prover.rs,
// air/accumulation.rs} for everything inside `prove`.
#pragma once
#include <chrono>
#include <cstdio>
#include <functional>
#include "proof.hpp"
#include "tables.hpp"

namespace sbf {

struct StageTimes { std::vector<std::pair<std::string, double>> ms; };

struct CommitTree {
  std::vector<Col> polys;      // coefficient columns (compact when rep > 0: coefficient j stands for j << rep, backend.hpp)
  uint32_t rep = 0;            // every evaluation column repeats each value 2^rep times
  std::vector<uint32_t> logs;  // log size of each polynomial
  std::vector<Col> evals;      // LDE columns (log + blowup)
  std::vector<Col> layers;     // Merkle layers by log size
  Hash root;
  bool borrowed = false;       // the columns belong to a PreprocessedCache: the proof does not free them
};

// The preprocessed tree is program-independent: the IsFirst columns of log size LOG_MAX_ROWS .. 4, whatever the program
// (brainfuck_air/mod.rs:453-464,493-500).  A caller that proves many programs on one backend may keep the tree (its
// polynomials, their LDEs, the Merkle layers and the root) between proofs.  The transcript is
// unchanged: the cached root is mixed where the computed one would be.  Off unless a cache object is passed.
struct PreprocessedCache {
  bool valid = false;
  uint32_t log_max_rows = 0, log_blowup = 0;
  CommitTree tree;
  uint64_t fills = 0, hits = 0;  // (re)builds and reuses so far
  bool matches(const ProverConfig& cfg) const { return valid && log_max_rows == cfg.log_max_rows && log_blowup == cfg.log_blowup; }
  void release(Backend& B) {
    for (Col x : tree.polys) B.free_col(x);
    for (Col x : tree.evals) B.free_col(x);
    for (Col x : tree.layers) B.free_col(x);
    forget();
  }
  void forget() { tree = CommitTree(); valid = false; }  // the columns are already gone (e.g. released with a failed proof)
};

// Mask points of every committed column: tree -> column -> points (Components::mask_points + composition, prover/mod.rs)
struct MaskLayout {
  std::vector<std::vector<std::vector<QPoint>>> points;
};
inline MaskLayout mask_points(const ProverConfig& cfg, const uint32_t log_size[N_COMPONENTS], QPoint oods) {
  MaskLayout m;
  m.points.resize(4);
  uint32_t n_pre = cfg.log_max_rows - LOG_N_LANES + 1;
  m.points[0].assign(n_pre, {});
  for (int c = 0; c < N_COMPONENTS; c++) m.points[0][cfg.log_max_rows - log_size[c]] = {oods};  // IsFirst(log_size) is used
  for (int c = 0; c < N_COMPONENTS; c++) {
    for (int k = 0; k < N_MAIN_COLS[c]; k++) m.points[1].push_back({oods});
    int ni = 4 * N_LOGUP_COLS[c];
    // trace step of CanonicCoset(log_size): G^(2^(31-log_size)); offset -1 -> subtract it
    QPoint minus_step = to_qpoint(p_conj(point_at_index(1u << (31 - log_size[c]))));
    QPoint prev = qp_add(oods, minus_step);
    for (int k = 0; k < ni; k++) {
      if (k < ni - 4) m.points[2].push_back({oods});
      else m.points[2].push_back({prev, oods});
    }
  }
  m.points[3].assign(4, {oods});
  return m;
}

// Host point evaluator (PointEvaluator + PointEvaluationAccumulator): eval_composition_polynomial_at_point.
struct PointEval {
  typedef Fq F;
  typedef Fq EF;
  const std::vector<std::vector<QM31>>* main;   // this component's main columns: [col][sample]
  const std::vector<std::vector<QM31>>* inter;  // this component's interaction columns
  size_t main_off, inter_off;
  QM31 is_first_v, total, denom_inv, random_coeff;
  const InteractionElements* el;
  QM31* accumulation;
  size_t col = 0;
  LogupState<PointEval> lg;
  F next() { return {(*main)[main_off + col++][0]}; }
  F is_first() { return {is_first_v}; }
  F cst(uint32_t c) { return {q_fromm(c)}; }
  EF ef(F x) { return x; }
  EF ef_zero() { return {q_zero()}; }
  EF ef_neg_one() { return {q_fromm(P - 1)}; }
  EF total_sum() { return {total}; }
  void add(F c) { *accumulation = q_add(q_mul(*accumulation, random_coeff), q_mul(denom_inv, c.v)); }
  void relation(int rel, EF num, const F* vals, int n) { lg.push(num, combine_q(el->rel[rel], vals, n)); }
  EF ext(int b, int sample) {
    const auto& c = *inter;
    QM31 v0 = c[inter_off + 4 * b][sample], v1 = c[inter_off + 4 * b + 1][sample], v2 = c[inter_off + 4 * b + 2][sample],
         v3 = c[inter_off + 4 * b + 3][sample];
    // SecureField::from_partial_evals: v0 + v1*i + v2*u + v3*iu
    QM31 r = v0;
    r = q_add(r, q_mul(v1, q_make(0, 1, 0, 0)));
    r = q_add(r, q_mul(v2, q_make(0, 0, 1, 0)));
    r = q_add(r, q_mul(v3, q_make(0, 0, 0, 1)));
    return {r};
  }
  EF ext_mask_cur(int b) { return ext(b, 0); }
  void ext_mask_last(EF& prev, EF& cur) { prev = ext(lg.n - 1, 0); cur = ext(lg.n - 1, 1); }
  void finalize_logup() { lg.finalize(*this); }
};

inline QM31 coset_vanishing_canonic(uint32_t log_size, QM31 x) {  // pi^(log_size-1)(x), SURVEY.md A.7
  for (uint32_t k = 1; k < log_size; k++) x = q_sub(q_mulm(q_sqr(x), 2), q_fromm(1));
  return x;
}

inline QM31 eval_composition_at_point(const ProverConfig& cfg, const uint32_t log_size[N_COMPONENTS], const QM31 claimed[N_COMPONENTS],
                                      const InteractionElements& el, QPoint oods,
                                      const std::vector<std::vector<std::vector<QM31>>>& sampled, QM31 random_coeff) {
  QM31 acc = q_zero();
  size_t main_off = 0, inter_off = 0;
  for (int c = 0; c < N_COMPONENTS; c++) {
    PointEval e;
    e.main = &sampled[1]; e.inter = &sampled[2]; e.main_off = main_off; e.inter_off = inter_off;
    const auto& pre = sampled[0][cfg.log_max_rows - log_size[c]];
    if (pre.empty()) throw std::runtime_error("missing IsFirst sample");
    e.is_first_v = pre[0];
    e.total = claimed[c];
    e.denom_inv = q_inv(coset_vanishing_canonic(log_size[c], oods.x));
    e.random_coeff = random_coeff;
    e.el = &el;
    e.accumulation = &acc;
    eval_component(c, e);
    main_off += N_MAIN_COLS[c];
    inter_off += 4 * N_LOGUP_COLS[c];
  }
  return acc;
}

inline InteractionElements draw_elements(Channel& ch) {  // BrainfuckInteractionElements::draw (brainfuck_air/mod.rs:149-165)
  InteractionElements el;
  for (int r = 0; r < 3; r++) {
    auto za = ch.draw_felts(2);
    el.rel[r].z = za[0];
    QM31 cur = q_fromm(1);
    for (int i = 0; i < 7; i++) { el.rel[r].alpha_pow[i] = cur; cur = q_mul(cur, za[1]); }
  }
  return el;
}

// Every element the decommit phase reads (`Column::at` upstream) is known from the query positions alone, so the walks
// below only record what they need; one flush then fetches everything with two batched gathers (32-byte hashes, 4-byte
// values) and hands each walk its slice.  Output references passed to the walks must stay valid until flush().
struct GatherQueue {
  Backend& B;
  std::vector<Col> hcols, vcols;
  std::vector<size_t> hoff, voff;
  std::vector<std::function<void(const uint32_t*, const uint32_t*)>> done;
  explicit GatherQueue(Backend& b) : B(b) {}
  void flush() {
    std::vector<uint32_t> hw = B.gather(hcols, hoff, 8), vw = B.gather(vcols, voff, 1);
    for (auto& f : done) f(hw.data(), vw.data());
    hcols.clear(); vcols.clear(); hoff.clear(); voff.clear(); done.clear();
  }
};

// MerkleProver::decommit (core/vcs/prover.rs): values come back per column in the original column order.
inline void merkle_decommit(GatherQueue& G, const CommitTree& t, const std::vector<Col>& columns, const std::map<uint32_t, std::vector<size_t>>& queries,
                            std::vector<std::vector<uint32_t>>* queried_values, MerkleDecommitment& d) {
  Backend& B = G.B;
  if (queried_values) queried_values->assign(columns.size(), {});
  struct VReq { size_t col; bool queried; };
  std::vector<VReq> vreq;
  const size_t hbase = G.hcols.size(), vbase = G.vcols.size();
  std::vector<size_t> last_queries;
  int n_layers = (int)t.layers.size();
  for (int lg = n_layers - 1; lg >= 0; lg--) {
    std::vector<size_t> lcols;
    for (size_t c = 0; c < columns.size(); c++) if (B.len(columns[c]) == ((size_t)1 << lg)) lcols.push_back(c);
    static const std::vector<size_t> none;
    auto it = queries.find((uint32_t)lg);
    const std::vector<size_t>& colq = it == queries.end() ? none : it->second;
    size_t pi = 0, ci = 0;
    std::vector<size_t> total;
    while (pi < last_queries.size() || ci < colq.size()) {
      size_t node;
      if (pi < last_queries.size() && ci < colq.size()) node = std::min(last_queries[pi] / 2, colq[ci]);
      else if (pi < last_queries.size()) node = last_queries[pi] / 2;
      else node = colq[ci];
      if (lg + 1 < n_layers) {
        for (size_t child = 2 * node; child <= 2 * node + 1; child++) {
          if (pi < last_queries.size() && last_queries[pi] == child) pi++;
          else { G.hcols.push_back(t.layers[lg + 1]); G.hoff.push_back(8 * child); }
        }
      }
      bool queried = ci < colq.size() && colq[ci] == node;
      if (queried) ci++;
      for (size_t c : lcols) { G.vcols.push_back(columns[c]); G.voff.push_back(node); vreq.push_back({c, queried}); }
      total.push_back(node);
    }
    last_queries = total;
  }
  const size_t nh = G.hcols.size() - hbase;
  G.done.push_back([hbase, vbase, nh, vreq = std::move(vreq), queried_values, &d](const uint32_t* hw, const uint32_t* vw) {
    for (size_t i = 0; i < nh; i++) { Hash h; memcpy(h.data(), hw + 8 * (hbase + i), 32); d.hash_witness.push_back(h); }
    for (size_t i = 0; i < vreq.size(); i++) {
      uint32_t v = vw[vbase + i];
      if (!vreq[i].queried) d.column_witness.push_back(v);
      else if (queried_values) (*queried_values)[vreq[i].col].push_back(v);
    }
  });
}

// FRI witness evaluations of one layer: the positions of each fold coset that are not themselves queried.
inline void fri_witness(GatherQueue& G, const std::array<Col, 4>& eval, const std::vector<size_t>& queries, const std::vector<size_t>& pos,
                        std::vector<QM31>& out) {
  const size_t vbase = G.vcols.size();
  size_t k = 0;
  for (size_t p : pos) {
    while (k < queries.size() && queries[k] < p) k++;
    if (k < queries.size() && queries[k] == p) continue;
    for (int c = 0; c < 4; c++) { G.vcols.push_back(eval[c]); G.voff.push_back(p); }
  }
  const size_t n = G.vcols.size() - vbase;
  G.done.push_back([vbase, n, &out](const uint32_t*, const uint32_t* vw) {
    const uint32_t* w = vw + vbase;
    for (size_t i = 0; i + 3 < n; i += 4) out.push_back(q_make(w[i], w[i + 1], w[i + 2], w[i + 3]));
  });
}

struct ProveResult {
  BrainfuckProof proof;
  StageTimes times;
};

// `run_vm` yields the execution trace.  With cfg.overlap_host it is called only after the (program-independent)
// preprocessed phase has been enqueued, so the VM run as well as the table building hide behind that device work.
typedef std::function<TraceInput()> TraceSource;
inline ProveResult prove_brainfuck(Backend& B, const std::vector<uint32_t>& code, const TraceSource& run_vm,
                                   const ProverConfig& cfg, const std::function<void()>& sync = nullptr,
                                   PreprocessedCache* pp_cache = nullptr) {
  ProveResult R;
  BrainfuckProof& proof = R.proof;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* name) {
    if (sync) sync();
    auto now = std::chrono::steady_clock::now();
    R.times.ms.push_back({name, std::chrono::duration<double, std::milli>(now - t_last).count()});
    t_last = now;
  };

  // ---- setup (mod.rs:479-487)
  B.precompute_twiddles(cfg.log_max_rows + cfg.log_blowup + 1);
  Channel ch;
  std::vector<CommitTree> trees;
  auto commit_tree = [&](CommitTree& t) {  // TreeBuilder::commit -> CommitmentTreeProver::new
    if (t.rep) {
      t.evals = B.evaluate_repeated(t.polys, t.rep, cfg.log_blowup);
      t.layers = B.merkle_commit_repeated(t.evals, t.rep, &t.root);
    } else {
      t.evals = B.evaluate(t.polys, cfg.log_blowup);
      t.layers = B.merkle_commit(t.evals, &t.root);
    }
    ch.mix_root(t.root);
  };
  lap("twiddles");

  // ---- phase 0: preprocessed trace (mod.rs:493-500).  With cfg.overlap_host the device work of this phase is only
  // enqueued here; the host builds the 13 tables meanwhile and the root is read back (and mixed) afterwards — the
  // transcript order is unchanged.
  std::vector<std::vector<Col>> compact(N_COMPONENTS);
  // The 13 tables in lane-compact form (one word per table row).  The CUDA backend uploads the 7-word register rows and
  // builds every table on the device (csrc/tables.cu); the copy runs beside the preprocessed phase on the copy stream.
  auto make_tables = [&] {
    TraceInput in = run_vm();
    B.trace_tables(in, code, cfg.log_max_rows, compact, proof.log_size);
  };
  if (!cfg.overlap_host) { make_tables(); lap("tables(host)"); }
  {
    CommitTree t;
    const bool hit = pp_cache && pp_cache->matches(cfg);
    if (hit) {
      t = pp_cache->tree;  // handles only; `borrowed` is set
      pp_cache->hits++;
    } else {
      if (pp_cache && pp_cache->valid) pp_cache->release(B);  // built for another LOG_MAX_ROWS / blowup
      // the polynomials (closed-form coefficients) are kept for the OODS samples; the extension is written directly
      // (Backend::is_first_lde: the polynomial is a rank-one product, no transform on CUDA)
      for (uint32_t lg = cfg.log_max_rows; lg >= LOG_N_LANES; lg--) {
        t.polys.push_back(B.is_first_poly(lg)); t.logs.push_back(lg);
        t.evals.push_back(B.is_first_lde(lg, cfg.log_blowup, 0, (size_t)1 << (lg + cfg.log_blowup)));
      }
      t.layers = B.merkle_commit(t.evals, nullptr);
    }
    if (cfg.overlap_host) make_tables();  // the VM runs on this thread while the device is busy with the phase queued above
    if (cfg.overlap_host) { R.times.ms.push_back({"tables(host)", 0}); lap("tables(host)+preprocessed"); }
    if (!hit) B.read(t.layers[0], 0, 8, t.root.data());
    ch.mix_root(t.root);
    if (pp_cache && !hit) {
      t.borrowed = true;
      pp_cache->tree = t;
      pp_cache->log_max_rows = cfg.log_max_rows;
      pp_cache->log_blowup = cfg.log_blowup;
      pp_cache->valid = true;
      pp_cache->fills++;
    }
    trees.push_back(std::move(t));
  }
  lap("preprocessed");

  // ---- phase 1: main trace (mod.rs:506-583)
  {
    CommitTree t;
    std::vector<Col> values;
    for (int c = 0; c < N_COMPONENTS; c++)
      for (Col cc : compact[c]) { values.push_back(cc); t.logs.push_back(proof.log_size[c]); }
    // a table row fills all 16 lanes of its column (table.rs trace_evaluation): interpolate / extend / hash the distinct values only
    t.rep = LOG_N_LANES;
    t.polys = B.interpolate_repeated(values, t.rep);   // the values themselves are kept for the LogUp generation below
    for (int c = 0; c < N_COMPONENTS; c++) ch.mix_u64(proof.log_size[c]);
    commit_tree(t);
    trees.push_back(std::move(t));
  }
  lap("main_trace");

  // ---- phase 2: interaction trace (mod.rs:589-723)
  InteractionElements el = draw_elements(ch);
  {
    CommitTree t;
    std::vector<Col> sum_cols;
    for (int c = 0; c < N_COMPONENTS; c++) {
      std::vector<Col> cols = B.logup_generate_deferred(c, compact[c], el);
      for (Col cc : compact[c]) B.free_col(cc);
      for (Col x : cols) { t.polys.push_back(x); t.logs.push_back(proof.log_size[c]); }
      sum_cols.insert(sum_cols.end(), cols.end() - 4, cols.end());
    }
    {  // claimed sums (LogupTraceGenerator::finalize_last: the cumulative column at index 1), one read-back for all components
      std::vector<uint32_t> w = B.gather(sum_cols, std::vector<size_t>(sum_cols.size(), 1), 1);
      for (int c = 0; c < N_COMPONENTS; c++) proof.claimed_sum[c] = q_make(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
      B.check_tables();  // the device has just been waited for: collect what the table kernels flagged, if anything
    }
    B.interpolate(t.polys);
    for (int c = 0; c < N_COMPONENTS; c++) ch.mix_felts({proof.claimed_sum[c]});
    commit_tree(t);
    trees.push_back(std::move(t));
  }
  lap("interaction_trace");

  // ---- prover::prove
  QM31 random_coeff = ch.draw_felt();
  int total_constraints = 0;
  for (int c = 0; c < N_COMPONENTS; c++) total_constraints += N_CONSTRAINTS[c];
  std::vector<QM31> powers(total_constraints);
  { QM31 a = q_fromm(1); for (auto& p : powers) { p = a; a = q_mul(a, random_coeff); } }
  std::map<uint32_t, std::array<Col, 4>> sub;  // DomainEvaluationAccumulator::sub_accumulations by eval log size
  {
    size_t main_off = 0, inter_off = 0;
    int g = 0;
    for (int c = 0; c < N_COMPONENTS; c++) {
      uint32_t ls = proof.log_size[c], elog = ls + 1;
      if (!sub.count(elog)) sub[elog] = {B.zeros((size_t)1 << elog), B.zeros((size_t)1 << elog), B.zeros((size_t)1 << elog), B.zeros((size_t)1 << elog)};
      std::vector<QM31> coeffs(N_CONSTRAINTS[c]);
      for (int k = 0; k < N_CONSTRAINTS[c]; k++) coeffs[k] = powers[total_constraints - 1 - (g + k)];
      g += N_CONSTRAINTS[c];
      std::vector<Col> m(trees[1].evals.begin() + main_off, trees[1].evals.begin() + main_off + N_MAIN_COLS[c]);
      std::vector<Col> it(trees[2].evals.begin() + inter_off, trees[2].evals.begin() + inter_off + 4 * N_LOGUP_COLS[c]);
      B.eval_constraints(c, ls, m, it, trees[0].evals[cfg.log_max_rows - ls], el, proof.claimed_sum[c], coeffs, sub[elog]);
      main_off += N_MAIN_COLS[c];
      inter_off += 4 * N_LOGUP_COLS[c];
    }
  }
  lap("constraints");
  // DomainEvaluationAccumulator::finalize.  Upstream walks the sizes upwards: evaluate the running polynomial on the next
  // domain, add the values, interpolate.  interpolate(values + evaluate(p)) = interpolate(values) + p exactly (p has lower
  // degree than the domain), so the same coefficients come from interpolating every size once — one batched call — and
  // adding each running polynomial into the low coefficients of the next: no lifting transforms at all.
  CommitTree comp_tree;
  {
    std::vector<Col> all;
    for (auto& kv : sub) for (Col x : kv.second) all.push_back(x);
    B.interpolate(all);
    std::vector<Col> cur;
    uint32_t cur_log = 0;
    for (auto& kv : sub) {
      std::array<Col, 4> vals = kv.second;
      if (!cur.empty()) {
        std::array<Col, 4> low;
        for (int k = 0; k < 4; k++) low[k] = B.view(vals[k], 0, (size_t)1 << cur_log);
        B.accumulate(low, {cur[0], cur[1], cur[2], cur[3]});
        for (Col x : low) B.free_col(x);
        for (Col x : cur) B.free_col(x);
      }
      cur = {vals[0], vals[1], vals[2], vals[3]};
      cur_log = kv.first;
    }
    comp_tree.polys = cur;
    comp_tree.logs.assign(4, cur_log);
    commit_tree(comp_tree);
    trees.push_back(std::move(comp_tree));
  }
  lap("composition");

  // ---- OODS sampling (prove_values)
  QPoint oods = random_point(ch);
  MaskLayout mask = mask_points(cfg, proof.log_size, oods);
  CommitmentSchemeProof& P = proof.proof;
  {
    std::vector<Col> polys;
    std::vector<QPoint> pts;
    std::vector<uint32_t> reps;
    for (size_t t = 0; t < trees.size(); t++)
      for (size_t c = 0; c < trees[t].polys.size(); c++)
        for (auto& p : mask.points[t][c]) { polys.push_back(trees[t].polys[c]); pts.push_back(p); reps.push_back(trees[t].rep); }
    std::vector<QM31> vals = B.eval_at_point_repeated(polys, reps, pts);
    size_t k = 0;
    P.sampled_values.resize(trees.size());
    std::vector<QM31> flat;
    for (size_t t = 0; t < trees.size(); t++) {
      P.sampled_values[t].resize(trees[t].polys.size());
      for (size_t c = 0; c < trees[t].polys.size(); c++)
        for (size_t s = 0; s < mask.points[t][c].size(); s++) { P.sampled_values[t][c].push_back(vals[k]); flat.push_back(vals[k]); k++; }
    }
    ch.mix_felts(flat);
  }
  lap("oods_eval");

  // ---- DEEP quotients (compute_fri_quotients): columns of all trees, grouped by LDE size, descending
  QM31 quot_coeff = ch.draw_felt();
  struct FlatCol { Col eval; uint32_t lde_log; std::vector<PointSample> samples; };
  std::vector<FlatCol> flat_cols;
  for (size_t t = 0; t < trees.size(); t++)
    for (size_t c = 0; c < trees[t].evals.size(); c++) {
      FlatCol f{trees[t].evals[c], trees[t].logs[c] + cfg.log_blowup, {}};
      for (size_t s = 0; s < mask.points[t][c].size(); s++) f.samples.push_back({mask.points[t][c][s], P.sampled_values[t][c][s]});
      flat_cols.push_back(std::move(f));
    }
  std::map<uint32_t, std::vector<const FlatCol*>, std::greater<uint32_t>> groups;
  for (auto& f : flat_cols) groups[f.lde_log].push_back(&f);
  std::vector<std::pair<uint32_t, std::array<Col, 4>>> quotients;  // descending log size
  for (auto& kv : groups) {
    std::vector<Col> cols;
    std::vector<const std::vector<PointSample>*> samples;
    for (auto* f : kv.second) { cols.push_back(f->eval); samples.push_back(&f->samples); }
    quotients.push_back({kv.first, B.accumulate_quotients(kv.first, cols, quot_coeff, batch_samples(samples))});
  }
  // ---- sanity check (ProvingError::ConstraintsNotSatisfied).  Host arithmetic on the sampled values: done here, while the
  // device works through the quotient kernels queued above, instead of at the end of the proof where it is pure latency
  {
    const auto& cs = P.sampled_values[3];
    QM31 comp = cs[0][0];
    comp = q_add(comp, q_mul(cs[1][0], q_make(0, 1, 0, 0)));
    comp = q_add(comp, q_mul(cs[2][0], q_make(0, 0, 1, 0)));
    comp = q_add(comp, q_mul(cs[3][0], q_make(0, 0, 0, 1)));
    QM31 want = eval_composition_at_point(cfg, proof.log_size, proof.claimed_sum, el, oods, P.sampled_values, random_coeff);
    if (!q_eq(comp, want)) throw std::runtime_error("ConstraintsNotSatisfied");
  }
  lap("quotients");

  // ---- FRI commit (FriProver::commit)
  CommitTree fri_first;
  std::vector<Col> first_cols;
  for (auto& q : quotients) for (Col x : q.second) first_cols.push_back(x);
  struct InnerLayer { std::array<Col, 4> eval; uint32_t log; CommitTree tree; };
  std::vector<InnerLayer> inner;
  const uint32_t last_log = cfg.log_last_layer_degree_bound + cfg.log_blowup;
  if (cfg.log_last_layer_degree_bound != 0) throw std::runtime_error("only log_last_layer_degree_bound = 0 is supported");
  std::vector<QM31> last_values;   // the last layer's evaluation
  std::array<Col, 4> layer = {nullptr, nullptr, nullptr, nullptr};
  Backend::FriCommitResult fused;
  if (B.fri_commit(quotients, ch.digest, last_log, fused)) {
    // The whole phase ran on the device with its own copy of the channel; replay the transcript from the roots it returned.
    fri_first.layers = fused.first_layers;
    fri_first.root = fused.first_root;
    ch.mix_root(fri_first.root);
    ch.draw_felt();                      // the circle fold's coefficient
    for (auto& L : fused.inner) {
      InnerLayer I{L.eval, L.log, {}};
      I.tree.layers = L.layers;
      I.tree.root = L.root;
      ch.mix_root(L.root);
      ch.draw_felt();                    // this layer's folding coefficient
      inner.push_back(std::move(I));
    }
    last_values = fused.last_layer;
  } else {
    fri_first.layers = B.merkle_commit(first_cols, &fri_first.root);
    ch.mix_root(fri_first.root);
    QM31 circle_alpha = ch.draw_felt();
    uint32_t line_log = quotients[0].first - 1;
    layer = {B.zeros((size_t)1 << line_log), B.zeros((size_t)1 << line_log), B.zeros((size_t)1 << line_log), B.zeros((size_t)1 << line_log)};
    size_t qi = 0;
    while (line_log > last_log) {
      while (qi < quotients.size() && quotients[qi].first - 1 == line_log) { B.fold_circle_into_line(layer, quotients[qi].second, quotients[qi].first, circle_alpha); qi++; }
      InnerLayer L{layer, line_log, {}};
      L.tree.layers = B.merkle_commit({layer[0], layer[1], layer[2], layer[3]}, &L.tree.root);
      ch.mix_root(L.tree.root);
      QM31 alpha = ch.draw_felt();
      layer = B.fold_line(layer, line_log, alpha);
      line_log--;
      inner.push_back(std::move(L));
    }
    if (qi != quotients.size()) throw std::runtime_error("FRI: not all columns consumed");
    size_t n = (size_t)1 << line_log;
    std::vector<std::vector<uint32_t>> cv(4, std::vector<uint32_t>(n));
    for (int k = 0; k < 4; k++) B.read(layer[k], 0, n, cv[k].data());
    for (size_t i = 0; i < n; i++) last_values.push_back(q_make(cv[0][i], cv[1][i], cv[2][i], cv[3][i]));
  }
  // last layer: interpolate on the host (LineEvaluation::interpolate), degree bound 2^log_last_layer_degree_bound = 1
  {
    QM31 v0 = last_values[0];
    for (size_t i = 1; i < last_values.size(); i++)
      if (!q_eq(v0, last_values[i])) throw std::runtime_error("FRI: invalid degree (last layer not constant)");
    P.fri_proof.last_layer_poly = {v0};
    ch.mix_felts(P.fri_proof.last_layer_poly);
  }
  lap("fri_commit");

  // ---- proof of work
  P.proof_of_work = B.grind(ch.digest, cfg.pow_bits);
  ch.mix_u64(P.proof_of_work);
  lap("grind");

  // ---- FRI decommit
  uint32_t max_log = quotients[0].first;
  Queries queries = Queries::generate(ch, max_log, cfg.n_queries);
  std::map<uint32_t, std::vector<size_t>> positions_by_log;
  GatherQueue G(B);
  P.fri_proof.inner_layers.resize(inner.size());  // the queued walks keep references into these
  {
    std::map<uint32_t, std::vector<size_t>> fri_pos;
    for (auto& q : quotients) {
      Queries cq = queries.fold(max_log - q.first);
      positions_by_log[q.first] = cq.positions;
      std::vector<size_t> pos = decommitment_positions(cq.positions, 1);
      fri_pos[q.first] = pos;
      fri_witness(G, q.second, cq.positions, pos, P.fri_proof.first_layer.fri_witness);
    }
    merkle_decommit(G, fri_first, first_cols, fri_pos, nullptr, P.fri_proof.first_layer.decommitment);
    P.fri_proof.first_layer.commitment = fri_first.root;
    Queries lq = queries.fold(1);
    for (size_t li = 0; li < inner.size(); li++) {
      auto& L = inner[li];
      FriLayerProof& lp = P.fri_proof.inner_layers[li];
      std::vector<size_t> pos = decommitment_positions(lq.positions, 1);
      fri_witness(G, L.eval, lq.positions, pos, lp.fri_witness);
      std::map<uint32_t, std::vector<size_t>> m{{L.log, pos}};
      merkle_decommit(G, L.tree, {L.eval[0], L.eval[1], L.eval[2], L.eval[3]}, m, nullptr, lp.decommitment);
      lp.commitment = L.tree.root;
      lq = lq.fold(1);
    }
  }
  // ---- decommit the four trees on the query positions
  P.queried_values.resize(trees.size());
  P.decommitments.resize(trees.size());
  for (size_t t = 0; t < trees.size(); t++) {
    P.commitments.push_back(trees[t].root);
    merkle_decommit(G, trees[t], trees[t].evals, positions_by_log, &P.queried_values[t], P.decommitments[t]);
  }
  G.flush();
  lap("decommit");

  // ---- release device memory
  for (auto& t : trees) {
    if (t.borrowed) continue;  // the preprocessed tree stays with its cache
    for (Col x : t.polys) B.free_col(x);
    for (Col x : t.evals) B.free_col(x);
    for (Col x : t.layers) B.free_col(x);
  }
  for (auto& q : quotients) for (Col x : q.second) B.free_col(x);
  for (Col x : fri_first.layers) B.free_col(x);
  for (auto& L : inner) { for (Col x : L.eval) B.free_col(x); for (Col x : L.tree.layers) B.free_col(x); }
  for (Col x : layer) if (x) B.free_col(x);
  lap("check+free");
  return R;
}

inline ProveResult prove_brainfuck(Backend& B, const std::vector<uint32_t>& code, const std::vector<Registers>& vm_trace,
                                   const ProverConfig& cfg, const std::function<void()>& sync = nullptr,
                                   PreprocessedCache* pp_cache = nullptr) {
  return prove_brainfuck(B, code, TraceSource([&] {
    TraceInput in;
    in.regs = vm_trace.data(); in.n = vm_trace.size(); in.stats = trace_stats(in.regs, in.n);
    return in;
  }), cfg, sync, pp_cache);
}

}  // namespace sbf
