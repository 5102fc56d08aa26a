// verify_brainfuck + stwo's verify / CommitmentSchemeVerifier / FriVerifier / MerkleVerifier — pure host code, as in the
// reference (crates/brainfuck_prover/src/brainfuck_air/mod.rs:738-797; "stays on CPU").  It is the
// acceptance test for a device-made proof.  Restates stwo-prover 0.1.1 @ 31e8dbc core/{prover/mod.rs (verify),
// pcs/verifier.rs, pcs/quotients.rs (fri_answers), fri.rs (FriVerifier), vcs/verifier.rs}.
#pragma once
#include "prover.hpp"

namespace sbf {

struct VerifyError : std::runtime_error { using std::runtime_error::runtime_error; };

// MerkleVerifier::verify.  column_logs: log size per column (tree order); queried: per column values in query order.
inline void merkle_verify(const Hash& root, const std::vector<uint32_t>& column_logs, const std::map<uint32_t, std::vector<size_t>>& queries,
                          const std::vector<std::vector<uint32_t>>& queried, const MerkleDecommitment& d) {
  if (column_logs.empty()) return;
  uint32_t max_log = *std::max_element(column_logs.begin(), column_logs.end());
  std::vector<size_t> qpos(column_logs.size(), 0);
  size_t hw = 0, cw = 0;
  std::vector<std::pair<size_t, Hash>> last;
  bool have_last = false;
  for (int lg = (int)max_log; lg >= 0; lg--) {
    std::vector<size_t> lcols;
    for (size_t c = 0; c < column_logs.size(); c++) if (column_logs[c] == (uint32_t)lg) lcols.push_back(c);
    static const std::vector<size_t> none;
    auto it = queries.find((uint32_t)lg);
    const std::vector<size_t>& colq = it == queries.end() ? none : it->second;
    size_t pi = 0, ci = 0;
    std::vector<std::pair<size_t, Hash>> total;
    while (pi < last.size() || ci < colq.size()) {
      size_t node;
      if (pi < last.size() && ci < colq.size()) node = std::min(last[pi].first / 2, colq[ci]);
      else if (pi < last.size()) node = last[pi].first / 2;
      else node = colq[ci];
      uint32_t children[16];
      if (have_last) {
        for (size_t k = 0; k < 2; k++) {
          size_t child = 2 * node + k;
          Hash h;
          if (pi < last.size() && last[pi].first == child) h = last[pi++].second;
          else { if (hw >= d.hash_witness.size()) throw VerifyError("Merkle: WitnessTooShort"); h = d.hash_witness[hw++]; }
          memcpy(children + 8 * k, h.data(), 32);
        }
      }
      bool q = ci < colq.size() && colq[ci] == node;
      if (q) ci++;
      std::vector<uint32_t> vals;
      for (size_t c : lcols) {
        if (q) { if (qpos[c] >= queried[c].size()) throw VerifyError("Merkle: ColumnValuesTooShort"); vals.push_back(queried[c][qpos[c]++]); }
        else { if (cw >= d.column_witness.size()) throw VerifyError("Merkle: WitnessTooShort"); vals.push_back(d.column_witness[cw++]); }
      }
      total.push_back({node, b2s::hash_node(have_last ? children : nullptr, vals.data(), vals.size())});
    }
    last = total;
    have_last = true;
  }
  if (hw != d.hash_witness.size() || cw != d.column_witness.size()) throw VerifyError("Merkle: WitnessTooLong");
  for (size_t c = 0; c < column_logs.size(); c++) if (qpos[c] != queried[c].size()) throw VerifyError("Merkle: ColumnValuesTooLong");
  if (last.size() != 1 || last[0].second != root) throw VerifyError("Merkle: RootMismatch");
}

struct SparseEval { std::vector<std::vector<QM31>> subsets; std::vector<size_t> subset_start; };

// compute_decommitment_positions_and_rebuild_evals (fold_step = 1)
inline SparseEval rebuild_evals(const std::vector<size_t>& queries, const std::vector<QM31>& query_evals, const std::vector<QM31>& witness,
                                size_t& wpos, std::vector<size_t>& positions) {
  SparseEval s;
  size_t i = 0;
  while (i < queries.size()) {
    size_t start = (queries[i] >> 1) << 1;
    std::vector<QM31> sub;
    for (size_t p = start; p < start + 2; p++) {
      positions.push_back(p);
      if (i < queries.size() && queries[i] == p) { sub.push_back(query_evals[i]); i++; }
      else { if (wpos >= witness.size()) throw VerifyError("FRI: InsufficientWitness"); sub.push_back(witness[wpos++]); }
    }
    s.subsets.push_back(sub);
    s.subset_start.push_back(start);
  }
  return s;
}

inline void verify_brainfuck(const BrainfuckProof& proof, const ProverConfig& cfg) {
  const CommitmentSchemeProof& P = proof.proof;
  if (P.commitments.size() != 4 || P.sampled_values.size() != 4 || P.queried_values.size() != 4 || P.decommitments.size() != 4)
    throw VerifyError("InvalidStructure");
  Channel ch;
  // column log sizes per tree (BrainfuckClaim::log_sizes, preprocessed overwritten by IS_FIRST_LOG_SIZES)
  std::vector<std::vector<uint32_t>> logs(4);
  for (uint32_t lg = cfg.log_max_rows; lg >= LOG_N_LANES; lg--) logs[0].push_back(lg);
  uint32_t max_ls = 0;
  for (int c = 0; c < N_COMPONENTS; c++) {
    uint32_t ls = proof.log_size[c];
    if (ls < LOG_N_LANES || ls > cfg.log_max_rows) throw VerifyError("InvalidStructure: log_size");
    max_ls = std::max(max_ls, ls);
    for (int k = 0; k < N_MAIN_COLS[c]; k++) logs[1].push_back(ls);
    for (int k = 0; k < 4 * N_LOGUP_COLS[c]; k++) logs[2].push_back(ls);
  }
  logs[3].assign(4, max_ls + 1);
  ch.mix_root(P.commitments[0]);
  for (int c = 0; c < N_COMPONENTS; c++) ch.mix_u64(proof.log_size[c]);
  ch.mix_root(P.commitments[1]);
  InteractionElements el = draw_elements(ch);
  QM31 sum = q_zero();
  for (int c = 0; c < N_COMPONENTS; c++) sum = q_add(sum, proof.claimed_sum[c]);
  if (!q_eq(sum, q_zero())) throw VerifyError("InvalidLookup: Invalid LogUp sum");
  for (int c = 0; c < N_COMPONENTS; c++) ch.mix_felts({proof.claimed_sum[c]});
  ch.mix_root(P.commitments[2]);
  // stwo verify()
  QM31 random_coeff = ch.draw_felt();
  ch.mix_root(P.commitments[3]);
  QPoint oods = random_point(ch);
  MaskLayout mask = mask_points(cfg, proof.log_size, oods);
  for (int t = 0; t < 4; t++) {
    if (P.sampled_values[t].size() != logs[t].size() || P.queried_values[t].size() != logs[t].size()) throw VerifyError("InvalidStructure");
    for (size_t c = 0; c < logs[t].size(); c++)
      if (P.sampled_values[t][c].size() != mask.points[t][c].size()) throw VerifyError("InvalidStructure: sampled_values");
  }
  const auto& cs = P.sampled_values[3];
  QM31 comp = cs[0][0];
  comp = q_add(comp, q_mul(cs[1][0], q_make(0, 1, 0, 0)));
  comp = q_add(comp, q_mul(cs[2][0], q_make(0, 0, 1, 0)));
  comp = q_add(comp, q_mul(cs[3][0], q_make(0, 0, 0, 1)));
  if (!q_eq(comp, eval_composition_at_point(cfg, proof.log_size, proof.claimed_sum, el, oods, P.sampled_values, random_coeff)))
    throw VerifyError("OodsNotMatching");

  // verify_values
  std::vector<QM31> flat;
  for (int t = 0; t < 4; t++) for (auto& c : P.sampled_values[t]) for (auto& v : c) flat.push_back(v);
  ch.mix_felts(flat);
  QM31 quot_coeff = ch.draw_felt();
  std::set<uint32_t, std::greater<uint32_t>> lde_logs;
  for (int t = 0; t < 4; t++) for (uint32_t l : logs[t]) lde_logs.insert(l + cfg.log_blowup);
  std::vector<uint32_t> col_logs(lde_logs.begin(), lde_logs.end());  // FRI column sizes, descending
  // FriVerifier::commit
  const FriProof& F = P.fri_proof;
  ch.mix_root(F.first_layer.commitment);
  QM31 circle_alpha = ch.draw_felt();
  uint32_t max_log = col_logs[0];
  std::vector<QM31> layer_alphas;
  uint32_t line_log = max_log - 1;
  // the last-layer check below compares evaluations with coefficient 0 only: a non-constant last layer is not supported
  // (the prover refuses it too), so it must not be accepted as if it were constant
  if (cfg.log_last_layer_degree_bound != 0) throw VerifyError("only log_last_layer_degree_bound = 0 is supported");
  const uint32_t last_log = cfg.log_last_layer_degree_bound + cfg.log_blowup;
  if (line_log < last_log || F.inner_layers.size() != line_log - last_log) throw VerifyError("FRI: InvalidNumFriLayers");
  for (auto& L : F.inner_layers) { ch.mix_root(L.commitment); layer_alphas.push_back(ch.draw_felt()); }
  if (F.last_layer_poly.size() > ((size_t)1 << cfg.log_last_layer_degree_bound)) throw VerifyError("FRI: LastLayerDegreeInvalid");
  ch.mix_felts(F.last_layer_poly);
  // proof of work
  ch.mix_u64(P.proof_of_work);
  if (ch.trailing_zeros() < cfg.pow_bits) throw VerifyError("ProofOfWork");
  // queries
  Queries queries = Queries::generate(ch, max_log, cfg.n_queries);
  std::map<uint32_t, std::vector<size_t>> positions_by_log;
  for (uint32_t l : col_logs) positions_by_log[l] = queries.fold(max_log - l).positions;
  // Merkle decommitments of the four trees
  for (int t = 0; t < 4; t++) {
    std::vector<uint32_t> ext;
    for (uint32_t l : logs[t]) ext.push_back(l + cfg.log_blowup);
    merkle_verify(P.commitments[t], ext, positions_by_log, P.queried_values[t], P.decommitments[t]);
  }
  // fri_answers: per size group (descending), quotient value at every query position
  struct FC { uint32_t lde; int t; size_t c; };
  std::vector<FC> flat_cols;
  for (int t = 0; t < 4; t++) for (size_t c = 0; c < logs[t].size(); c++) flat_cols.push_back({logs[t][c] + cfg.log_blowup, t, c});
  std::vector<std::vector<QM31>> answers;  // per FRI column
  for (uint32_t l : col_logs) {
    std::vector<const FC*> grp;
    for (auto& f : flat_cols) if (f.lde == l) grp.push_back(&f);
    std::vector<std::vector<PointSample>> samples(grp.size());
    std::vector<const std::vector<PointSample>*> sp;
    for (size_t g = 0; g < grp.size(); g++) {
      for (size_t s = 0; s < mask.points[grp[g]->t][grp[g]->c].size(); s++)
        samples[g].push_back({mask.points[grp[g]->t][grp[g]->c][s], P.sampled_values[grp[g]->t][grp[g]->c][s]});
    }
    for (auto& s : samples) sp.push_back(&s);
    SampleBatchesFlat bf = batch_samples(sp);
    const std::vector<size_t>& qp = positions_by_log[l];
    std::vector<QM31> ans;
    for (size_t qi = 0; qi < qp.size(); qi++) {
      std::vector<uint32_t> row;
      for (auto* f : grp) {
        const auto& qv = P.queried_values[f->t][f->c];
        if (qv.size() != qp.size()) throw VerifyError("InvalidStructure: queried_values");
        row.push_back(qv[qi]);
      }
      Pt dp = canonic_domain_at(l, bitrev32((uint32_t)qp[qi], l));
      ans.push_back(row_quotient(bf, row, quot_coeff, dp));
    }
    answers.push_back(ans);
  }
  // FriVerifier::decommit — first layer
  std::vector<SparseEval> sparse(col_logs.size());
  {
    size_t wpos = 0;
    std::map<uint32_t, std::vector<size_t>> dpos;
    std::vector<std::vector<uint32_t>> dvals;
    std::vector<uint32_t> dlogs;
    for (size_t k = 0; k < col_logs.size(); k++) {
      std::vector<size_t> pos;
      sparse[k] = rebuild_evals(positions_by_log[col_logs[k]], answers[k], F.first_layer.fri_witness, wpos, pos);
      dpos[col_logs[k]] = pos;
      for (int c = 0; c < 4; c++) {
        std::vector<uint32_t> v;
        for (auto& sub : sparse[k].subsets) for (auto& q : sub) { uint32_t w[4] = {q.a.a, q.a.b, q.b.a, q.b.b}; v.push_back(w[c]); }
        dvals.push_back(v);
        dlogs.push_back(col_logs[k]);
      }
    }
    if (wpos != F.first_layer.fri_witness.size()) throw VerifyError("FRI: FirstLayerEvaluationsInvalid");
    merkle_verify(F.first_layer.commitment, dlogs, dpos, dvals, F.first_layer.decommitment);
  }
  // fold the first-layer subsets into the line, then through the inner layers
  auto fold_circle_pair = [&](uint32_t log, size_t start, const std::vector<QM31>& sub, QM31 alpha) {
    Pt p = canonic_domain_at(log, bitrev32((uint32_t)start, log));
    QM31 f0 = q_add(sub[0], sub[1]), f1 = q_mulm(q_sub(sub[0], sub[1]), m_inv(p.y));
    return q_add(f0, q_mul(alpha, f1));
  };
  auto fold_line_pair = [&](uint32_t log, size_t start, const std::vector<QM31>& sub, QM31 alpha) {
    // LineDomain(half_odds(log)).at(bitrev(start)) x-coordinate
    uint32_t i = bitrev32((uint32_t)start, log);
    Pt p = point_at_index((1u << (29 - log)) + (uint32_t)(((uint64_t)i << (31 - log)) & 0x7fffffffu));
    QM31 f0 = q_add(sub[0], sub[1]), f1 = q_mulm(q_sub(sub[0], sub[1]), m_inv(p.x));
    return q_add(f0, q_mul(alpha, f1));
  };
  Queries lq = queries.fold(1);
  std::vector<QM31> layer_evals(lq.positions.size(), q_zero());
  size_t col_i = 0;
  QM31 alpha_sq = q_mul(circle_alpha, circle_alpha);
  for (size_t li = 0; li <= F.inner_layers.size(); li++) {
    // fold in the circle columns whose folded size matches this layer
    while (col_i < col_logs.size() && col_logs[col_i] - 1 == line_log) {
      const SparseEval& s = sparse[col_i];
      // subsets of this column are indexed by its own folded queries; map onto this layer's query list
      Queries cq = queries.fold(max_log - col_logs[col_i]).fold(1);
      if (cq.positions != lq.positions || s.subsets.size() != lq.positions.size()) throw VerifyError("FRI: query mismatch");
      for (size_t k = 0; k < lq.positions.size(); k++) {
        QM31 folded = fold_circle_pair(col_logs[col_i], s.subset_start[k], s.subsets[k], circle_alpha);
        layer_evals[k] = q_add(q_mul(layer_evals[k], alpha_sq), folded);
      }
      col_i++;
    }
    if (li == F.inner_layers.size()) break;
    const FriLayerProof& L = F.inner_layers[li];
    size_t wpos = 0;
    std::vector<size_t> pos;
    SparseEval s = rebuild_evals(lq.positions, layer_evals, L.fri_witness, wpos, pos);
    if (wpos != L.fri_witness.size()) throw VerifyError("FRI: InnerLayerEvaluationsInvalid");
    std::vector<std::vector<uint32_t>> dvals(4);
    for (auto& sub : s.subsets) for (auto& q : sub) { dvals[0].push_back(q.a.a); dvals[1].push_back(q.a.b); dvals[2].push_back(q.b.a); dvals[3].push_back(q.b.b); }
    merkle_verify(L.commitment, std::vector<uint32_t>(4, line_log), {{line_log, pos}}, dvals, L.decommitment);
    Queries nq = lq.fold(1);
    std::vector<QM31> next(nq.positions.size());
    if (s.subsets.size() != nq.positions.size()) throw VerifyError("FRI: fold mismatch");   // before anything is written through k
    for (size_t k = 0; k < s.subsets.size(); k++) next[k] = fold_line_pair(line_log, s.subset_start[k], s.subsets[k], layer_alphas[li]);
    lq = nq;
    layer_evals = next;
    line_log--;
  }
  if (col_i != col_logs.size()) throw VerifyError("FRI: columns not consumed");
  // last layer: constant polynomial
  for (auto& v : layer_evals) if (!q_eq(v, F.last_layer_poly.empty() ? q_zero() : F.last_layer_poly[0])) throw VerifyError("FRI: LastLayerEvaluationsInvalid");
}

}  // namespace sbf
