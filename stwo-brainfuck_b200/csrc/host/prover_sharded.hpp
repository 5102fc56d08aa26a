// One proof split over N = 2^w ranks (one GPU each) — the multi-GPU form of prover.hpp, same transcript, same proof bytes.
//
// Partitioning (BASELINE.json north_star):
//   * polynomials are COLUMN-sharded: each committed column has one owner rank that interpolates it, extends it and
//     evaluates it at the OODS points (columns are independent: no communication);
//   * every low-degree extension is then RE-SHARDED BY ROWS with one all-to-all per commitment tree: rank r ends up with
//     rows [r*R/N, (r+1)*R/N) (bit-reversed order) of every column of that size; Merkle hashing, constraint evaluation,
//     DEEP quotients and FRI folds are local to such a range (children 2i, 2i+1 and fold partners are adjacent);
//   * per tree / FRI layer the N sub-roots are all-gathered (N x 32 bytes) and the top log2 N layers are hashed by every rank;
//   * tiny things (columns below 32 rows per rank, the FRI tail, the channel, the grind) are replicated;
//   * the only non-local mask, the LogUp cumulative column at coset offset -1, travels as an extra pre-shifted column in
//     the interaction tree's all-to-all;
//   * composition accumulators are row-sharded; their four coordinate columns are gathered to four owners for the
//     interpolate/lift chain (FFTs need whole columns), then committed like any other tree.
// With world = 1 this driver degenerates to the single-GPU prover and produces the identical proof (tested).
#pragma once
#include <memory>
#include "prover.hpp"

namespace sbf {

struct ShardLayout {
  int w = 0, rank = 0, world = 1;
  // Columns of fewer than 2^min_log rows are replicated instead of sharded (ProverConfig::shard_min_log): below ~2^16 rows a
  // rank's share is a handful of thread blocks, and every sharded layer costs an all-gather of sub-roots plus three launches
  // for the replicated top of its tree — more than hashing the whole small tree on every rank.
  uint32_t min_log = 0;
  bool circle_sharded(uint32_t L) const { return L >= std::max<uint32_t>((uint32_t)w + 5, min_log + 1); }  // LDE / circle-domain columns of log L
  bool line_sharded(uint32_t l) const { return l >= std::max<uint32_t>((uint32_t)w + 4, min_log); }        // FRI line layers of log l
};

// A column as this rank sees it after re-sharding: its rows [rank*seg, (rank+1)*seg) if sharded, else the whole column.
struct RowCol {
  Col rows = nullptr;
  uint32_t L = 0;
  bool sharded = false;
  size_t seg() const { return rows_len; }
  size_t rows_len = 0;
};
struct Loc { Col col; size_t off; bool mine; };
inline Loc locate(const RowCol& c, size_t row, const ShardLayout& sl) {
  if (c.sharded) return {c.rows, row % c.rows_len, (int)(row / c.rows_len) == sl.rank};
  return {c.rows, row, sl.rank == 0};
}
// Values that live on different ranks, delivered to every rank: each rank gathers what it owns, the rest arrives by all-reduce.
inline std::vector<uint32_t> fetch(Backend& B, const std::vector<Loc>& req, uint32_t words) {
  std::vector<uint32_t> buf(req.size() * words, 0);
  std::vector<Col> cols;
  std::vector<size_t> offs, slot;
  for (size_t i = 0; i < req.size(); i++)
    if (req[i].mine) { cols.push_back(req[i].col); offs.push_back(req[i].off); slot.push_back(i); }
  std::vector<uint32_t> got = B.gather(cols, offs, words);
  for (size_t j = 0; j < slot.size(); j++) memcpy(&buf[slot[j] * words], &got[j * words], words * 4);
  if (B.world() > 1) B.allreduce_host(buf.data(), buf.size());
  return buf;
}

// Batched form: requests are registered while walking the trees, resolved with ONE gather + ONE all-reduce per word size,
// then the registered finishers distribute the values (MerkleProver::decommit touches ~100 scattered nodes per tree).
struct FetchBatch {
  std::vector<Loc> req[2];             // [0]: 1-word values, [1]: 8-word hashes
  std::vector<uint32_t> data[2];
  std::vector<std::function<void()>> finish;
  size_t add_value(const Loc& l) { req[0].push_back(l); return req[0].size() - 1; }
  size_t add_hash(const Loc& l) { req[1].push_back(l); return req[1].size() - 1; }
  void run(Backend& B) {
    data[0] = fetch(B, req[0], 1);
    data[1] = fetch(B, req[1], 8);
    for (auto& f : finish) f();
  }
};

// Owner of every column of a tree: longest-processing-time-first over the transform cost (2^log words): columns in
// descending size, each to the rank with the least work so far (ties: the lowest rank), so that four 2^24-word columns and
// sixteen 2^22-word columns on eight ranks come out as 4 x (one large) + 4 x (four small) instead of the 6 : 3 split a plain
// round-robin over the sorted list gives.  Every rank computes the same assignment.
inline std::vector<int> assign_owners(const std::vector<uint32_t>& logs, int world) {
  std::vector<size_t> order(logs.size());
  for (size_t i = 0; i < order.size(); i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return logs[a] > logs[b]; });
  std::vector<int> owner(logs.size());
  std::vector<uint64_t> load((size_t)world, 0);
  for (size_t k = 0; k < order.size(); k++) {
    int best = 0;
    for (int r = 1; r < world; r++) if (load[(size_t)r] < load[(size_t)best]) best = r;
    owner[order[k]] = best;
    load[(size_t)best] += (uint64_t)1 << logs[order[k]];
  }
  return owner;
}

struct SMerkle {
  std::vector<Col> layers;  // [k]: k > w: this rank's sub-layer (2^(k-w) nodes); k <= w: the whole layer (replicated)
  Hash root;
  bool whole = false;       // no column of the tree is sharded: every rank hashed the whole tree, every layer is complete
};
// Row-sharded MerkleProver::commit over columns that are either row ranges (sharded) or whole (replicated).  The node
// function is local to a row, so the layers with more than `world` nodes are an ordinary Merkle tree over this rank's row
// ranges (log sizes L - w): one backend call, which fuses the small layers and stages one pointer table.  The `world`
// sub-roots are all-gathered and the top w layers hashed by every rank.
// rep: every column repeats each value 2^rep times (main trace); the deepest `rep` layers then hash one node per group.
// before_root_read: called when all device work is queued, just before the blocking read of the root.
inline SMerkle merkle_sharded(Backend& B, const ShardLayout& sl, const std::vector<RowCol>& cols, uint32_t rep = 0,
                              const std::function<void()>& before_root_read = nullptr, bool read_root = true) {
  SMerkle m;
  uint32_t maxL = 0;
  for (auto& c : cols) maxL = std::max(maxL, c.L);
  m.layers.assign(maxL + 1, nullptr);
  const uint32_t w = (uint32_t)sl.w;
  bool any_sharded = false;
  for (auto& c : cols) any_sharded |= c.sharded;
  if (!any_sharded) {
    // Every column is replicated (a small tree, ShardLayout::min_log): hash the whole tree here, like every other rank —
    // one fused backend call, no sub-root all-gather and no separate launches for the top layers.
    std::vector<Col> all;
    for (auto& c : cols) all.push_back(c.rows);
    m.layers = rep ? B.merkle_commit_repeated(all, rep, nullptr) : B.merkle_commit(all, nullptr);
    m.whole = true;
    if (before_root_read) before_root_read();
    if (read_root) B.read(m.layers[0], 0, 8, m.root.data());
    return m;
  }
  Col prev = nullptr;
  int k = (int)maxL;
  if (maxL >= w) {
    std::vector<Col> local, tmp;
    for (auto& c : cols) {
      if (c.L < w) continue;
      size_t n = (size_t)1 << (c.L - w);
      if (c.sharded) local.push_back(c.rows);
      else { Col v = B.view(c.rows, (size_t)sl.rank * n, n); tmp.push_back(v); local.push_back(v); }
    }
    std::vector<Col> ll = rep ? B.merkle_commit_repeated(local, rep, nullptr) : B.merkle_commit(local, nullptr);
    for (Col v : tmp) B.free_col(v);
    for (size_t j = 0; j < ll.size(); j++) m.layers[j + w] = ll[j];
    Col full = B.alloc((size_t)8 << w);   // N sub-roots -> every rank
    B.all_gather(ll[0], full, 8);
    B.free_col(ll[0]);
    m.layers[w] = full;
    prev = full;
    k = (int)w - 1;
  }
  for (; k >= 0; k--) {
    std::vector<Col> lc;
    for (auto& c : cols) if (c.L == (uint32_t)k) lc.push_back(c.rows);  // always replicated at these sizes
    Col layer = B.commit_layer((uint32_t)k, prev, lc);
    m.layers[k] = layer;
    prev = layer;
  }
  if (before_root_read) before_root_read();
  if (read_root) B.read(m.layers[0], 0, 8, m.root.data());   // otherwise the caller mixes it on the device (Backend::dchan_*)
  return m;
}

struct STree {
  std::vector<uint32_t> logs;  // polynomial log size per column
  std::vector<int> owner;
  std::vector<Col> polys;      // coefficient column on the owner, nullptr elsewhere (every rank when `rep`)
  uint32_t rep = 0;            // main trace: compact coefficients of lane-repeated columns (backend.hpp), held by every rank
  std::vector<RowCol> rows;    // LDE as seen by this rank
  std::vector<Col> bufs;       // buffers behind `rows`
  SMerkle merkle;
};
struct ExtraCol { Col full; uint32_t L; int owner; };  // non-committed column riding in a tree's all-to-all

// Column-shard -> row-shard exchange of the LDEs (`lde[c]` on owners) + extras, then the row-sharded Merkle commit.
inline void exchange_and_commit(Backend& B, const ShardLayout& sl, uint32_t log_blowup, STree& t, std::vector<Col>& lde,
                                const std::vector<ExtraCol>& extras, std::vector<RowCol>* extra_rows,
                                const std::function<void()>& before_root_read = nullptr) {
  const int N = sl.world, me = sl.rank;
  struct E { uint32_t L; int owner; bool sharded; size_t seg; Col full; };
  std::vector<E> es;
  for (size_t c = 0; c < t.logs.size(); c++) {
    uint32_t L = t.logs[c] + log_blowup;
    bool sh = sl.circle_sharded(L);
    es.push_back({L, t.owner[c], sh, sh ? ((size_t)1 << (L - sl.w)) : ((size_t)1 << L), lde[c]});
  }
  for (auto& x : extras) {
    bool sh = sl.circle_sharded(x.L);
    es.push_back({x.L, x.owner, sh, sh ? ((size_t)1 << (x.L - sl.w)) : ((size_t)1 << x.L), x.full});
  }
  std::vector<RowCol> out(es.size());
  if (N == 1) {  // nothing to exchange: the LDE columns are the row ranges
    for (size_t i = 0; i < es.size(); i++) { out[i] = {es[i].full, es[i].L, es[i].sharded, es[i].seg}; t.bufs.push_back(es[i].full); }
  } else {
    std::vector<size_t> scount(N, 0), rcount(N, 0);
    for (auto& e : es) { rcount[e.owner] += e.seg; if (e.owner == me) for (int d = 0; d < N; d++) scount[d] += e.seg; }
    size_t stot = 0, rtot = 0;
    for (int d = 0; d < N; d++) { stot += scount[d]; rtot += rcount[d]; }
    std::vector<Col> pc;
    std::vector<size_t> ps;
    std::vector<uint8_t> psh;
    for (auto& e : es) if (e.owner == me) { pc.push_back(e.full); ps.push_back(e.seg); psh.push_back(e.sharded ? 1 : 0); }
    // CUDA: one kernel writes every destination's block into that rank's receive window over NVLink (Backend::exchange_push);
    // otherwise (no peer mapping, window still growing, CPU backends) pack a send buffer and hand it to the all-to-all
    Col recv = B.exchange_push(pc, ps, psh, rcount);
    if (!recv) {
      Col send = B.alloc(std::max<size_t>(stot, 4));
      recv = B.alloc(std::max<size_t>(rtot, 4));
      B.pack_exchange(send, pc, ps, psh);   // one launch on CUDA instead of world x columns copies
      B.all_to_all(send, scount, recv, rcount);
      B.free_col(send);
    }
    for (auto& e : es) if (e.owner == me) B.free_col(e.full);
    std::vector<size_t> roff(N, 0);
    { size_t o = 0; for (int s = 0; s < N; s++) { roff[s] = o; o += rcount[s]; } }
    for (size_t i = 0; i < es.size(); i++) {
      out[i] = {B.view(recv, roff[es[i].owner], es[i].seg), es[i].L, es[i].sharded, es[i].seg};
      roff[es[i].owner] += es[i].seg;
      t.bufs.push_back(out[i].rows);
    }
    t.bufs.push_back(recv);
  }
  t.rows.assign(out.begin(), out.begin() + t.logs.size());
  if (extra_rows) extra_rows->assign(out.begin() + t.logs.size(), out.end());
  t.merkle = merkle_sharded(B, sl, t.rows, 0, before_root_read);
}

// MerkleProver::decommit over a sharded tree: same walk as merkle_decommit; the reads are registered in `fb` and the outputs
// are filled by a finisher once the batch has run.
inline void merkle_decommit_sharded(FetchBatch& fb, const ShardLayout& sl, const SMerkle& m, const std::vector<RowCol>& columns,
                                    const std::map<uint32_t, std::vector<size_t>>& queries,
                                    std::vector<std::vector<uint32_t>>* queried_values, MerkleDecommitment* d) {
  if (queried_values) queried_values->assign(columns.size(), {});
  struct VReq { size_t slot, col; bool queried; };
  auto hslots = std::make_shared<std::vector<size_t>>();
  auto vinfo = std::make_shared<std::vector<VReq>>();
  std::vector<size_t> last_queries;
  int n_layers = (int)m.layers.size();
  for (int lg = n_layers - 1; lg >= 0; lg--) {
    std::vector<size_t> lcols;
    for (size_t c = 0; c < columns.size(); c++) if (columns[c].L == (uint32_t)lg) lcols.push_back(c);
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
        int k = lg + 1;
        for (size_t child = 2 * node; child <= 2 * node + 1; child++) {
          if (pi < last_queries.size() && last_queries[pi] == child) { pi++; continue; }
          if (k > sl.w && !m.whole) {
            size_t n = (size_t)1 << (k - sl.w);
            hslots->push_back(fb.add_hash({m.layers[k], 8 * (child % n), (int)(child / n) == sl.rank}));
          } else {
            hslots->push_back(fb.add_hash({m.layers[k], 8 * child, sl.rank == 0}));
          }
        }
      }
      bool queried = ci < colq.size() && colq[ci] == node;
      if (queried) ci++;
      for (size_t c : lcols) vinfo->push_back({fb.add_value(locate(columns[c], node, sl)), c, queried});
      total.push_back(node);
    }
    last_queries = total;
  }
  FetchBatch* pfb = &fb;
  fb.finish.push_back([pfb, hslots, vinfo, queried_values, d] {
    for (size_t s : *hslots) { Hash h; memcpy(h.data(), &pfb->data[1][8 * s], 32); d->hash_witness.push_back(h); }
    for (auto& v : *vinfo) {
      uint32_t x = pfb->data[0][v.slot];
      if (v.queried) { if (queried_values) (*queried_values)[v.col].push_back(x); } else d->column_witness.push_back(x);
    }
  });
}
inline void fri_witness_sharded(FetchBatch& fb, const ShardLayout& sl, const std::array<RowCol, 4>& eval, const std::vector<size_t>& queries,
                                const std::vector<size_t>& pos, std::vector<QM31>* out) {
  auto slots = std::make_shared<std::vector<size_t>>();
  size_t k = 0;
  for (size_t p : pos) {
    while (k < queries.size() && queries[k] < p) k++;
    if (k < queries.size() && queries[k] == p) continue;
    for (int c = 0; c < 4; c++) slots->push_back(fb.add_value(locate(eval[c], p, sl)));
  }
  FetchBatch* pfb = &fb;
  fb.finish.push_back([pfb, slots, out] {
    const auto& w = pfb->data[0];
    for (size_t i = 0; i + 3 < slots->size(); i += 4)
      out->push_back(q_make(w[(*slots)[i]], w[(*slots)[i + 1]], w[(*slots)[i + 2]], w[(*slots)[i + 3]]));
  });
}

inline ProveResult prove_brainfuck_sharded(Backend& B, const std::vector<uint32_t>& code, const TraceSource& run_vm,
                                           const ProverConfig& cfg, const std::function<void()>& sync = nullptr) {
  ProveResult R;
  BrainfuckProof& proof = R.proof;
  ShardLayout sl;
  sl.rank = B.rank(); sl.world = B.world(); sl.min_log = cfg.shard_min_log;
  while ((1 << sl.w) < sl.world) sl.w++;
  if ((1 << sl.w) != sl.world) throw std::runtime_error("world size must be a power of two");
  const int N = sl.world, me = sl.rank;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* name) {
    if (sync) sync();
    auto now = std::chrono::steady_clock::now();
    R.times.ms.push_back({name, std::chrono::duration<double, std::milli>(now - t_last).count()});
    t_last = now;
  };
  auto row_off = [&](const RowCol& c) { return c.sharded ? (size_t)me * c.rows_len : (size_t)0; };

  B.precompute_twiddles(cfg.log_max_rows + cfg.log_blowup + 1);
  Channel ch;
  std::vector<STree> trees;
  lap("twiddles");

  // owned polynomials of a tree: interpolate in place, extend, exchange, commit
  auto lde_owned = [&](STree& t) {
    std::vector<Col> owned;
    std::vector<size_t> idx;
    for (size_t c = 0; c < t.polys.size(); c++) if (t.polys[c]) { owned.push_back(t.polys[c]); idx.push_back(c); }
    std::vector<Col> ev = B.evaluate(owned, cfg.log_blowup);
    std::vector<Col> lde(t.polys.size(), nullptr);
    for (size_t j = 0; j < idx.size(); j++) lde[idx[j]] = ev[j];
    return lde;
  };
  auto interpolate_owned = [&](STree& t) {
    std::vector<Col> owned;
    for (Col p : t.polys) if (p) owned.push_back(p);
    B.interpolate(owned);
  };

  // Every rank runs the VM itself (it is deterministic and there is a host core per GPU) on a host thread beside the
  // program-independent preprocessed phase below, uploads the 7-word register rows and builds all 13 tables on its own
  // device (Backend::trace_tables; csrc/tables.cu on CUDA): no table crosses NVLink.
  std::vector<std::vector<Col>> compact(N_COMPONENTS);
  TraceInput trace_in;
  std::string host_err;
  double host_wait_ms = 0;
  struct Joiner { std::thread t; ~Joiner() { if (t.joinable()) t.join(); } } host;
  if (cfg.overlap_host) {
    host.t = std::thread([&] {
      try { trace_in = run_vm(); } catch (const std::exception& e) { host_err = e.what(); }
    });
  } else {
    // VM, upload and table building in front of every other device operation of the proof (bench.py's device-timed `value`:
    // the backend marks "register rows resident" behind the upload, so the mark then precedes the whole proof)
    trace_in = run_vm();
    B.trace_tables(trace_in, code, cfg.log_max_rows, compact, proof.log_size);
  }

  // ---- phase 0: preprocessed trace
  {
    if (N > 1) B.exchange_begin();   // collective: the receive windows of the direct column->row exchange (Backend::exchange_push)
    STree t;
    for (uint32_t lg = cfg.log_max_rows; lg >= LOG_N_LANES; lg--) t.logs.push_back(lg);
    t.owner = assign_owners(t.logs, N);
    for (size_t c = 0; c < t.logs.size(); c++) t.polys.push_back(t.owner[c] == me ? B.is_first_poly(t.logs[c]) : nullptr);
    // The extension of an IsFirst column is a row-local closed form (Backend::is_first_lde): every rank writes its own row
    // range of every column, so this tree needs neither transforms nor a column->row exchange; the polynomials stay with
    // their owners for the OODS samples.
    for (size_t c = 0; c < t.logs.size(); c++) {
      const uint32_t L = t.logs[c] + cfg.log_blowup;
      const bool sh = sl.circle_sharded(L);
      const size_t seg = sh ? ((size_t)1 << (L - sl.w)) : ((size_t)1 << L);
      Col r = B.is_first_lde(t.logs[c], cfg.log_blowup, sh ? (size_t)me * seg : 0, seg);
      t.rows.push_back({r, L, sh, seg});
      t.bufs.push_back(r);
    }
    // With every kernel of this phase queued, wait for the host thread and queue the uploads of this rank's tables: they
    // run on the copy stream beside the tail of the phase instead of in front of the main-trace exchange.
    void *ev_queued = nullptr, *ev_host = nullptr;
    t.merkle = merkle_sharded(B, sl, t.rows, 0, [&] {
      if (!cfg.overlap_host) return;
      ev_queued = B.mark();   // end of this phase's kernels
      host.t.join();
      ev_host = B.mark();     // first moment the device could be given the next phase: the gap is what the host cost it
      if (!host_err.empty()) return;
      try { B.trace_tables(trace_in, code, cfg.log_max_rows, compact, proof.log_size); } catch (const std::exception& e) { host_err = e.what(); }
    });
    if (ev_queued && ev_host) host_wait_ms = B.gap_ms(ev_queued, ev_host);
    if (!host_err.empty()) throw std::runtime_error(host_err);
    ch.mix_root(t.merkle.root);
    trees.push_back(std::move(t));
  }
  lap("preprocessed");
  // the part of the VM + table time the DEVICE waited for (the rest hid behind the preprocessed phase)
  R.times.ms.push_back({"tables(host)", host_wait_ms});

  // ---- phase 1: main trace.  The compact columns (one word per table row, 63 MB in total for fib19) were built on every
  // rank's own device above, because LogUp generation and the transforms below need them everywhere.
  {
    STree t;
    for (int c = 0; c < N_COMPONENTS; c++) {
      if (proof.log_size[c] < LOG_N_LANES) throw std::runtime_error("bad table size");
      for (int j = 0; j < N_MAIN_COLS[c]; j++) t.logs.push_back(proof.log_size[c]);
    }
    t.owner = assign_owners(t.logs, N);
    // Every table row fills 16 lanes, so a column's transforms run on its distinct values at 1/16 of the cost; the values
    // are on every rank already, so each rank transforms ALL main columns and expands only its own row range of the LDE:
    // the main tree needs no column->row exchange at all.
    std::vector<Col> values;
    for (int c = 0; c < N_COMPONENTS; c++) for (Col cc : compact[c]) values.push_back(cc);
    t.rep = LOG_N_LANES;
    t.polys = B.interpolate_repeated(values, t.rep);
    for (int c = 0; c < N_COMPONENTS; c++) ch.mix_u64(proof.log_size[c]);
    {
      std::vector<size_t> offs, cnts;
      for (size_t c = 0; c < t.logs.size(); c++) {
        uint32_t L = t.logs[c] + cfg.log_blowup;
        bool sh = sl.circle_sharded(L);
        size_t seg = sh ? ((size_t)1 << (L - sl.w)) : ((size_t)1 << L);
        offs.push_back(sh ? (size_t)me * seg : 0);
        cnts.push_back(seg);
      }
      std::vector<Col> rows = B.evaluate_repeated_range(t.polys, t.rep, cfg.log_blowup, offs, cnts);
      for (size_t c = 0; c < t.logs.size(); c++) {
        uint32_t L = t.logs[c] + cfg.log_blowup;
        t.rows.push_back({rows[c], L, sl.circle_sharded(L), cnts[c]});
        t.bufs.push_back(rows[c]);
      }
      t.merkle = merkle_sharded(B, sl, t.rows, t.rep);
    }
    ch.mix_root(t.merkle.root);
    trees.push_back(std::move(t));
  }
  lap("main_trace");

  // ---- phase 2: interaction trace.  Ownership is per coordinate column; an owner recomputes the (cheap) fractions itself.
  InteractionElements el = draw_elements(ch);
  std::vector<std::array<RowCol, 4>> prev_rows(N_COMPONENTS);  // last LogUp column at coset offset -1, row-sharded
  {
    STree t;
    std::vector<int> col_comp;
    for (int c = 0; c < N_COMPONENTS; c++)
      for (int j = 0; j < 4 * N_LOGUP_COLS[c]; j++) { t.logs.push_back(proof.log_size[c]); col_comp.push_back(c); }
    t.owner = assign_owners(t.logs, N);
    t.polys.assign(t.logs.size(), nullptr);
    std::vector<uint32_t> claimed(4 * N_COMPONENTS, 0);
    std::vector<Col> sum_cols;
    std::vector<size_t> sum_slots;
    size_t base = 0;
    for (int c = 0; c < N_COMPONENTS; c++) {
      int nout = 4 * N_LOGUP_COLS[c];
      std::vector<uint8_t> want(nout, 0);
      bool any = false;
      for (int j = 0; j < nout; j++) if (t.owner[base + j] == me) { want[j] = 1; any = true; }
      if (any) {
        std::vector<Col> outs = B.logup_generate_sel(c, compact[c], el, want);
        for (int j = 0; j < nout; j++) {
          if (!outs[j]) continue;
          if (j >= nout - 4) {  // LogupTraceGenerator::finalize_last: coset-order prefix sum, claimed_sum = col.at(1)
            B.prefix_sum(outs[j]);
            sum_cols.push_back(outs[j]);
            sum_slots.push_back(4 * c + (j - (nout - 4)));
          }
          t.polys[base + j] = outs[j];
        }
      }
      for (Col cc : compact[c]) B.free_col(cc);
      base += nout;
    }
    if (!sum_cols.empty()) {  // one read-back for every cumulative column this rank owns
      std::vector<uint32_t> wv = B.gather(sum_cols, std::vector<size_t>(sum_cols.size(), 1), 1);
      for (size_t i = 0; i < sum_slots.size(); i++) claimed[sum_slots[i]] = wv[i];
    }
    if (N > 1) B.allreduce_host(claimed.data(), claimed.size());
    for (int c = 0; c < N_COMPONENTS; c++) proof.claimed_sum[c] = q_make(claimed[4 * c], claimed[4 * c + 1], claimed[4 * c + 2], claimed[4 * c + 3]);
    B.check_tables();
    interpolate_owned(t);
    for (int c = 0; c < N_COMPONENTS; c++) ch.mix_felts({proof.claimed_sum[c]});
    std::vector<Col> lde = lde_owned(t);
    // extras: the last LogUp column of every component, shifted to coset offset -1, one column per coordinate
    std::vector<ExtraCol> extras;
    base = 0;
    for (int c = 0; c < N_COMPONENTS; c++) {
      int nout = 4 * N_LOGUP_COLS[c];
      for (int k = 0; k < 4; k++) {
        size_t ci = base + nout - 4 + k;
        Col sh = t.owner[ci] == me ? B.shift_prev(lde[ci], proof.log_size[c]) : nullptr;
        extras.push_back({sh, proof.log_size[c] + cfg.log_blowup, t.owner[ci]});
      }
      base += nout;
    }
    std::vector<RowCol> extra_rows;
    exchange_and_commit(B, sl, cfg.log_blowup, t, lde, extras, &extra_rows);
    for (int c = 0; c < N_COMPONENTS; c++) for (int k = 0; k < 4; k++) prev_rows[c][k] = extra_rows[4 * c + k];
    ch.mix_root(t.merkle.root);
    trees.push_back(std::move(t));
  }
  lap("interaction_trace");

  // ---- composition: constraint quotients on this rank's rows
  QM31 random_coeff = ch.draw_felt();
  int total_constraints = 0;
  for (int c = 0; c < N_COMPONENTS; c++) total_constraints += N_CONSTRAINTS[c];
  std::vector<QM31> powers(total_constraints);
  { QM31 a = q_fromm(1); for (auto& p : powers) { p = a; a = q_mul(a, random_coeff); } }
  struct Sub { std::array<Col, 4> cols; bool sharded; size_t len; };
  std::map<uint32_t, Sub> sub;
  {
    size_t main_off = 0, inter_off = 0;
    int g = 0;
    for (int c = 0; c < N_COMPONENTS; c++) {
      uint32_t ls = proof.log_size[c], L = ls + 1;
      bool sh = sl.circle_sharded(L);
      size_t len = sh ? ((size_t)1 << (L - sl.w)) : ((size_t)1 << L);
      if (!sub.count(L)) sub[L] = {{B.zeros(len), B.zeros(len), B.zeros(len), B.zeros(len)}, sh, len};
      std::vector<QM31> coeffs(N_CONSTRAINTS[c]);
      for (int k = 0; k < N_CONSTRAINTS[c]; k++) coeffs[k] = powers[total_constraints - 1 - (g + k)];
      g += N_CONSTRAINTS[c];
      std::vector<Col> m, it;
      for (int j = 0; j < N_MAIN_COLS[c]; j++) m.push_back(trees[1].rows[main_off + j].rows);
      for (int j = 0; j < 4 * N_LOGUP_COLS[c]; j++) it.push_back(trees[2].rows[inter_off + j].rows);
      Col isf = trees[0].rows[cfg.log_max_rows - ls].rows;
      if (sh) {
        std::array<Col, 4> pv = {prev_rows[c][0].rows, prev_rows[c][1].rows, prev_rows[c][2].rows, prev_rows[c][3].rows};
        B.eval_constraints_range(c, ls, (size_t)me * len, len, m, it, pv, isf, el, proof.claimed_sum[c], coeffs, sub[L].cols);
      } else {
        B.eval_constraints(c, ls, m, it, isf, el, proof.claimed_sum[c], coeffs, sub[L].cols);
      }
      main_off += N_MAIN_COLS[c];
      inter_off += 4 * N_LOGUP_COLS[c];
    }
  }
  for (auto& pr : prev_rows) for (auto& rc : pr) if (N > 1 && rc.rows) { /* views into the tree's recv buffer: freed with it */ }
  lap("constraints");

  // ---- accumulator finalize: coordinate k of every size goes to rank k % N (one all-to-all), which runs the FFT chain
  STree comp_tree;
  {
    auto owner_of = [&](int k) { return k % N; };
    std::map<uint32_t, std::array<Col, 4>> full;  // whole accumulator columns on their coordinate owner
    if (N == 1) {
      for (auto& kv : sub) full[kv.first] = kv.second.cols;
    } else {
      // Direct path (CUDA): every rank writes its row range of coordinate k straight into the whole column inside owner k % N's
      // receive window (Backend::exchange_scatter) — no send buffer, no all-to-all, no reassembly copies.  The layout of an
      // owner's region: its coordinates of every sharded size, ascending, each a whole column of len * N words.
      for (auto& kv : sub) full[kv.first] = {nullptr, nullptr, nullptr, nullptr};
      Col region = nullptr;
      {
        std::vector<size_t> tot(N, 0);
        std::map<std::pair<uint32_t, int>, size_t> off;   // (size, coordinate) -> offset in its owner's region
        for (auto& kv : sub)
          if (kv.second.sharded)
            for (int k = 0; k < 4; k++) { off[{kv.first, k}] = tot[owner_of(k)]; tot[owner_of(k)] += kv.second.len * (size_t)N; }
        size_t region_words = 4;
        for (int d = 0; d < N; d++) region_words = std::max(region_words, tot[d]);
        std::vector<Col> pieces;
        std::vector<uint32_t> dest;
        std::vector<size_t> doff;
        for (auto& kv : sub)
          if (kv.second.sharded)
            for (int k = 0; k < 4; k++) {
              pieces.push_back(kv.second.cols[k]); dest.push_back((uint32_t)owner_of(k));
              doff.push_back(off[{kv.first, k}] + (size_t)me * kv.second.len);
            }
        region = B.exchange_scatter(pieces, dest, doff, region_words);
        if (region) {
          for (auto& kv : sub)
            if (kv.second.sharded)
              for (int k = 0; k < 4; k++)
                if (owner_of(k) == me) full[kv.first][k] = B.view(region, off[{kv.first, k}], kv.second.len * (size_t)N);
          B.free_col(region);   // a handle on window memory: the views stay valid for the proof
        }
      }
      std::vector<size_t> scount(N, 0), rcount(N, 0);
      if (!region)
      for (auto& kv : sub) {
        if (!kv.second.sharded) continue;
        for (int k = 0; k < 4; k++) { scount[owner_of(k)] += kv.second.len; if (owner_of(k) == me) for (int s = 0; s < N; s++) rcount[s] += kv.second.len; }
      }
      if (!region) {   // packed send buffer -> all-to-all -> reassembly
        size_t stot = 0, rtot = 0;
        for (int d = 0; d < N; d++) { stot += scount[d]; rtot += rcount[d]; }
        Col send = B.alloc(std::max<size_t>(stot, 4)), recv = B.alloc(std::max<size_t>(rtot, 4));
        size_t so = 0;
        for (int d = 0; d < N; d++)
          for (auto& kv : sub)
            if (kv.second.sharded)
              for (int k = 0; k < 4; k++)
                if (owner_of(k) == d) { B.copy(send, so, kv.second.cols[k], 0, kv.second.len); so += kv.second.len; }
        B.all_to_all(send, scount, recv, rcount);
        B.free_col(send);
        size_t ro = 0;
        for (int s = 0; s < N; s++)
          for (auto& kv : sub)
            if (kv.second.sharded)
              for (int k = 0; k < 4; k++)
                if (owner_of(k) == me) {
                  Col& dst = full[kv.first][k];
                  if (!dst) dst = B.alloc(kv.second.len * N);
                  B.copy(dst, (size_t)s * kv.second.len, recv, ro, kv.second.len);
                  ro += kv.second.len;
                }
        B.free_col(recv);
      }
      for (auto& kv : sub)
        for (int k = 0; k < 4; k++) {
          if (kv.second.sharded) B.free_col(kv.second.cols[k]);
          else if (owner_of(k) == me) full[kv.first][k] = kv.second.cols[k];   // replicated: the owner keeps its copy
          else B.free_col(kv.second.cols[k]);
        }
    }
    uint32_t comp_log = sub.rbegin()->first;
    comp_tree.logs.assign(4, comp_log);
    for (int k = 0; k < 4; k++) comp_tree.owner.push_back(owner_of(k));
    comp_tree.polys.assign(4, nullptr);
    {  // finalize in coefficient space (prover.hpp): interpolate every size once, add each running polynomial into the low
       // coefficients of the next size
      std::vector<Col> mine;
      for (auto& kv : full) for (int k = 0; k < 4; k++) if (owner_of(k) == me) mine.push_back(kv.second[k]);
      B.interpolate(mine);
    }
    for (int k = 0; k < 4; k++) {
      if (owner_of(k) != me) continue;
      Col cur = nullptr;
      uint32_t cur_log = 0;
      for (auto& kv : full) {
        Col vals = kv.second[k];
        if (cur) {
          Col low = B.view(vals, 0, (size_t)1 << cur_log);
          B.accumulate_col(low, cur);
          B.free_col(low);
          B.free_col(cur);
        }
        cur = vals;
        cur_log = kv.first;
      }
      comp_tree.polys[k] = cur;
    }
    std::vector<Col> lde = lde_owned(comp_tree);
    exchange_and_commit(B, sl, cfg.log_blowup, comp_tree, lde, {}, nullptr);
    ch.mix_root(comp_tree.merkle.root);
    trees.push_back(std::move(comp_tree));
  }
  lap("composition");

  // ---- OODS sampling: owners evaluate their polynomials, the table is completed by all-reduce
  QPoint oods = random_point(ch);
  MaskLayout mask = mask_points(cfg, proof.log_size, oods);
  CommitmentSchemeProof& P = proof.proof;
  {
    std::vector<Col> polys;
    std::vector<QPoint> pts;
    std::vector<size_t> slots;
    std::vector<uint32_t> reps;
    size_t n_slots = 0;
    for (size_t t = 0; t < trees.size(); t++)
      for (size_t c = 0; c < trees[t].polys.size(); c++)
        for (auto& p : mask.points[t][c]) {
          if (trees[t].polys[c] && (!trees[t].rep || trees[t].owner[c] == me)) {
            polys.push_back(trees[t].polys[c]); pts.push_back(p); reps.push_back(trees[t].rep); slots.push_back(n_slots);
          }
          n_slots++;
        }
    std::vector<QM31> vals = B.eval_at_point_repeated(polys, reps, pts);
    std::vector<uint32_t> table(4 * n_slots, 0);
    for (size_t j = 0; j < slots.size(); j++) memcpy(&table[4 * slots[j]], &vals[j], 16);
    if (N > 1) B.allreduce_host(table.data(), table.size());
    size_t k = 0;
    P.sampled_values.resize(trees.size());
    std::vector<QM31> flat;
    for (size_t t = 0; t < trees.size(); t++) {
      P.sampled_values[t].resize(trees[t].polys.size());
      for (size_t c = 0; c < trees[t].polys.size(); c++)
        for (size_t s = 0; s < mask.points[t][c].size(); s++) {
          QM31 v = q_make(table[4 * k], table[4 * k + 1], table[4 * k + 2], table[4 * k + 3]);
          P.sampled_values[t][c].push_back(v);
          flat.push_back(v);
          k++;
        }
    }
    ch.mix_felts(flat);
  }
  lap("oods_eval");

  // ---- DEEP quotients on this rank's rows
  QM31 quot_coeff = ch.draw_felt();
  struct FlatCol { RowCol rc; std::vector<PointSample> samples; };
  std::vector<FlatCol> flat_cols;
  for (size_t t = 0; t < trees.size(); t++)
    for (size_t c = 0; c < trees[t].rows.size(); c++) {
      FlatCol f{trees[t].rows[c], {}};
      for (size_t s = 0; s < mask.points[t][c].size(); s++) f.samples.push_back({mask.points[t][c][s], P.sampled_values[t][c][s]});
      flat_cols.push_back(std::move(f));
    }
  std::map<uint32_t, std::vector<const FlatCol*>, std::greater<uint32_t>> groups;
  for (auto& f : flat_cols) groups[f.rc.L].push_back(&f);
  std::vector<std::pair<uint32_t, std::array<RowCol, 4>>> quotients;  // descending log size
  for (auto& kv : groups) {
    std::vector<Col> cols;
    std::vector<const std::vector<PointSample>*> samples;
    for (auto* f : kv.second) { cols.push_back(f->rc.rows); samples.push_back(&f->samples); }
    const RowCol& r0 = kv.second[0]->rc;
    std::array<Col, 4> q = r0.sharded ? B.accumulate_quotients_range(kv.first, row_off(r0), r0.rows_len, cols, quot_coeff, batch_samples(samples))
                                      : B.accumulate_quotients(kv.first, cols, quot_coeff, batch_samples(samples));
    std::array<RowCol, 4> rq;
    for (int k = 0; k < 4; k++) rq[k] = {q[k], kv.first, r0.sharded, r0.rows_len};
    quotients.push_back({kv.first, rq});
  }
  // ---- sanity check (ProvingError::ConstraintsNotSatisfied): host arithmetic on the sampled values, done while the device works
  // through the quotient kernels queued above instead of at the end of the proof
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

  // ---- FRI commit
  std::vector<RowCol> first_cols;
  for (auto& q : quotients) for (auto& x : q.second) first_cols.push_back(x);
  // The transcript of this phase runs on the device when the backend has a device channel (Backend::dchan_*): every rank
  // mixes the (replicated) root of each layer and draws the folding coefficient in a one-thread kernel, the folds read it
  // from device memory, and no layer waits for a read-back; the host replays its channel from the roots afterwards.
  const uint32_t last_log = cfg.log_last_layer_degree_bound + cfg.log_blowup;
  const uint32_t n_mixes = quotients[0].first - last_log;   // first layer + one per inner layer
  void* dc = B.dchan_begin(ch.digest, n_mixes);
  uint32_t mixes = 0;
  SMerkle fri_first = merkle_sharded(B, sl, first_cols, 0, nullptr, dc == nullptr);
  QM31 circle_alpha = q_zero();
  if (dc) { B.dchan_mix_root_draw(dc, fri_first.layers[0]); mixes++; }
  else { ch.mix_root(fri_first.root); circle_alpha = ch.draw_felt(); }
  struct InnerLayer { std::array<RowCol, 4> eval; uint32_t log; SMerkle tree; };
  std::vector<InnerLayer> inner;
  uint32_t line_log = quotients[0].first - 1;
  auto line_cols = [&](uint32_t l, std::array<Col, 4> c) {
    bool sh = sl.line_sharded(l);
    size_t len = sh ? ((size_t)1 << (l - sl.w)) : ((size_t)1 << l);
    std::array<RowCol, 4> r;
    for (int k = 0; k < 4; k++) r[k] = {c[k], l, sh, len};
    return r;
  };
  std::array<RowCol, 4> layer;
  {
    bool sh = sl.line_sharded(line_log);
    size_t len = sh ? ((size_t)1 << (line_log - sl.w)) : ((size_t)1 << line_log);
    layer = line_cols(line_log, {B.zeros(len), B.zeros(len), B.zeros(len), B.zeros(len)});
  }
  size_t qi = 0;
  auto cols_of = [](const std::array<RowCol, 4>& r) { return std::array<Col, 4>{r[0].rows, r[1].rows, r[2].rows, r[3].rows}; };
  while (line_log > last_log) {
    if (dc && !layer[0].sharded && line_log <= B.fri_tail_max_log()) {
      // The line evaluation is replicated and small: every remaining layer (circle fold, commit, channel, fold_line) in one
      // backend call on the device transcript (Backend::fri_tail_dc; csrc/fri.cu fri_tail_kernel) instead of ~10 launches each.
      std::vector<std::array<Col, 4>> tq;
      size_t qj = qi;
      for (uint32_t lg = line_log; lg > last_log; lg--) {
        if (qj < quotients.size() && quotients[qj].first - 1 == lg) { tq.push_back(cols_of(quotients[qj].second)); qj++; }
        else tq.push_back({nullptr, nullptr, nullptr, nullptr});
      }
      bool whole = true;   // the quotient columns of these sizes are replicated too (below ShardLayout::min_log)
      for (size_t j = qi; j < qj; j++) whole &= !quotients[j].second[0].sharded;
      Backend::FriTailResult tr;
      if (whole && B.fri_tail_dc(dc, cols_of(layer), line_log, last_log, tq, tr)) {
        for (int k = 0; k < 4; k++) B.free_col(layer[k].rows);
        for (size_t t = 0; t < tr.evals.size(); t++) {
          InnerLayer Lr{line_cols(line_log, tr.evals[t]), line_log, {}};
          Lr.tree.layers = tr.trees[t];
          Lr.tree.whole = true;
          inner.push_back(std::move(Lr));
          mixes++;
          line_log--;
        }
        layer = line_cols(line_log, tr.last);
        qi = qj;
        break;
      }
    }
    while (qi < quotients.size() && quotients[qi].first - 1 == line_log) {
      const size_t off = layer[0].sharded ? (size_t)me * layer[0].rows_len : 0;
      if (dc) B.fold_circle_into_line_range_dc(cols_of(layer), cols_of(quotients[qi].second), quotients[qi].first, off, layer[0].rows_len, dc, 0);
      else if (layer[0].sharded) B.fold_circle_into_line_range(cols_of(layer), cols_of(quotients[qi].second), quotients[qi].first, off,
                                                               layer[0].rows_len, circle_alpha);
      else B.fold_circle_into_line(cols_of(layer), cols_of(quotients[qi].second), quotients[qi].first, circle_alpha);
      qi++;
    }
    InnerLayer Lr{layer, line_log, {}};
    Lr.tree = merkle_sharded(B, sl, std::vector<RowCol>(layer.begin(), layer.end()), 0, nullptr, dc == nullptr);
    QM31 alpha = q_zero();
    if (dc) { B.dchan_mix_root_draw(dc, Lr.tree.layers[0]); mixes++; }
    else { ch.mix_root(Lr.tree.root); alpha = ch.draw_felt(); }
    std::array<Col, 4> next;
    if (layer[0].sharded) {
      size_t n_out = layer[0].rows_len / 2;
      next = dc ? B.fold_line_range_dc(cols_of(layer), line_log, (size_t)me * n_out, n_out, dc, mixes - 1)
                : B.fold_line_range(cols_of(layer), line_log, (size_t)me * n_out, n_out, alpha);
      if (!sl.line_sharded(line_log - 1)) {  // the layer becomes too small to shard: replicate it
        for (int k = 0; k < 4; k++) {
          Col fullc = B.alloc((size_t)1 << (line_log - 1));
          B.all_gather(next[k], fullc, n_out);
          B.free_col(next[k]);
          next[k] = fullc;
        }
      }
    } else {
      next = dc ? B.fold_line_range_dc(cols_of(layer), line_log, 0, (size_t)1 << (line_log - 1), dc, mixes - 1)
                : B.fold_line(cols_of(layer), line_log, alpha);
    }
    line_log--;
    layer = line_cols(line_log, next);
    inner.push_back(std::move(Lr));
  }
  if (dc) {   // one read-back for the whole phase, then the host transcript catches up
    std::vector<Hash> roots = B.dchan_finish(dc, mixes);
    fri_first.root = roots[0];
    ch.mix_root(roots[0]);
    ch.draw_felt();
    for (size_t i = 0; i < inner.size(); i++) {
      inner[i].tree.root = roots[i + 1];
      ch.mix_root(roots[i + 1]);
      ch.draw_felt();
    }
  }
  if (qi != quotients.size()) throw std::runtime_error("FRI: not all columns consumed");
  {
    size_t n = (size_t)1 << line_log;
    std::vector<std::vector<uint32_t>> cv(4, std::vector<uint32_t>(n));
    for (int k = 0; k < 4; k++) B.read(layer[k].rows, 0, n, cv[k].data());
    if (cfg.log_last_layer_degree_bound != 0) throw std::runtime_error("only log_last_layer_degree_bound = 0 is supported");
    QM31 v0 = q_make(cv[0][0], cv[1][0], cv[2][0], cv[3][0]);
    for (size_t i = 1; i < n; i++)
      if (!q_eq(v0, q_make(cv[0][i], cv[1][i], cv[2][i], cv[3][i]))) throw std::runtime_error("FRI: invalid degree (last layer not constant)");
    P.fri_proof.last_layer_poly = {v0};
    ch.mix_felts(P.fri_proof.last_layer_poly);
  }
  lap("fri_commit");

  P.proof_of_work = B.grind(ch.digest, cfg.pow_bits);
  ch.mix_u64(P.proof_of_work);
  lap("grind");

  // ---- decommit: every read of every tree and FRI layer goes into one batch (one gather + one all-reduce per word size)
  uint32_t max_log = quotients[0].first;
  Queries queries = Queries::generate(ch, max_log, cfg.n_queries);
  std::map<uint32_t, std::vector<size_t>> positions_by_log;
  {
    FetchBatch fb;
    std::map<uint32_t, std::vector<size_t>> fri_pos;
    for (auto& q : quotients) {
      Queries cq = queries.fold(max_log - q.first);
      positions_by_log[q.first] = cq.positions;
      std::vector<size_t> pos = decommitment_positions(cq.positions, 1);
      fri_pos[q.first] = pos;
      fri_witness_sharded(fb, sl, q.second, cq.positions, pos, &P.fri_proof.first_layer.fri_witness);
    }
    merkle_decommit_sharded(fb, sl, fri_first, first_cols, fri_pos, nullptr, &P.fri_proof.first_layer.decommitment);
    P.fri_proof.first_layer.commitment = fri_first.root;
    P.fri_proof.inner_layers.resize(inner.size());
    Queries lq = queries.fold(1);
    std::vector<std::vector<RowCol>> inner_cols(inner.size());
    for (size_t i = 0; i < inner.size(); i++) {
      InnerLayer& L = inner[i];
      FriLayerProof& lp = P.fri_proof.inner_layers[i];
      std::vector<size_t> pos = decommitment_positions(lq.positions, 1);
      fri_witness_sharded(fb, sl, L.eval, lq.positions, pos, &lp.fri_witness);
      std::map<uint32_t, std::vector<size_t>> m{{L.log, pos}};
      inner_cols[i].assign(L.eval.begin(), L.eval.end());
      merkle_decommit_sharded(fb, sl, L.tree, inner_cols[i], m, nullptr, &lp.decommitment);
      lp.commitment = L.tree.root;
      lq = lq.fold(1);
    }
    P.queried_values.resize(trees.size());
    P.decommitments.resize(trees.size());
    for (size_t t = 0; t < trees.size(); t++) {
      P.commitments.push_back(trees[t].merkle.root);
      merkle_decommit_sharded(fb, sl, trees[t].merkle, trees[t].rows, positions_by_log, &P.queried_values[t], &P.decommitments[t]);
    }
    lap("decommit:walk");
    fb.run(B);
  }
  lap("decommit");

  // ---- release
  for (auto& t : trees) {
    for (Col x : t.polys) if (x) B.free_col(x);
    for (Col x : t.bufs) B.free_col(x);
    for (Col x : t.merkle.layers) if (x) B.free_col(x);
  }
  for (auto& q : quotients) for (auto& x : q.second) B.free_col(x.rows);
  for (Col x : fri_first.layers) if (x) B.free_col(x);
  for (auto& L : inner) { for (auto& x : L.eval) B.free_col(x.rows); for (Col x : L.tree.layers) if (x) B.free_col(x); }
  for (auto& x : layer) B.free_col(x.rows);
  lap("check+free");
  return R;
}

}  // namespace sbf
