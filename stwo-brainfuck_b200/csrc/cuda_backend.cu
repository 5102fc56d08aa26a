// CudaBackend: the `Backend` the protocol driver runs on in the product — every method is one C-ABI call
// (include/stwo_cuda.h), i.e. the host orchestrator dog-foods the drop-in boundary.  Also exports the prove / verify entry
// points (`sbf_*`) that stand in for `brainfuck_prover prove|verify` (crates/brainfuck_prover/src/bin/brainfuck_prover.rs:79-152).
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/stwo_cuda_sharded.h"
#include "host/prover_sharded.hpp"
#include "host/verifier.hpp"

using namespace sbf;

namespace {

struct CudaBackendImpl : Backend {
  sc_ctx* ctx;
  const sc_twiddles* tw = nullptr;
  explicit CudaBackendImpl(sc_ctx* c) : ctx(c) {}
  ~CudaBackendImpl() override { if (own_tw) sc_twiddles_free(ctx, own_tw); if (pending_trace) sc_trace_free(ctx, pending_trace); if (ev_resident) sc_event_free(ctx, ev_resident); }
  static void ck(int32_t r) { if (r) throw std::runtime_error(std::string("stwo_cuda: ") + sc_last_error()); }
  static sc_col* h(Col c) { return (sc_col*)c; }
  const char* name() const override { return "cuda"; }

  Col from_host(const uint32_t* v, size_t n) override { sc_col* c; ck(sc_col_from_host(ctx, v, n, &c)); return c; }
  Col from_host_async(const uint32_t* v, size_t n) override { sc_col* c; ck(sc_col_from_host_async(ctx, v, n, &c)); return c; }
  struct PinnedArena : HostArena {
    sc_ctx* ctx;
    void* alloc(size_t bytes) override {
      void* p = nullptr;
      if (sc_host_arena_alloc(ctx, bytes, &p)) throw std::bad_alloc();
      return p;
    }
  } pinned;
  HostArena* host_arena() override { pinned.ctx = ctx; sc_host_arena_reset(ctx); return &pinned; }
  // The 13 tables are built on the device from the uploaded register rows (csrc/tables.cu).  host_tables = true falls back
  // to the host builders of tables.hpp (SBF_HOST_TABLES: A/B measurements and the table parity tests).
  bool host_tables = false;
  sc_trace* pending_trace = nullptr;
  sc_event* ev_resident = nullptr;   // recorded when the register rows have arrived (bench.py's device-timed `value`)
  uint64_t h2d_bytes = 0;
  // milliseconds of device time from that mark to now (waits for the queued work); -1 without a mark
  double ms_since_resident() {
    if (!ev_resident) return -1;
    sc_event* e = nullptr;
    ck(sc_event_record(ctx, &e));
    float ms = 0;
    int32_t r = sc_event_elapsed(ctx, ev_resident, e, &ms);
    sc_event_free(ctx, e); sc_event_free(ctx, ev_resident); ev_resident = nullptr;
    ck(r);
    return ms;
  }
  void trace_tables(const TraceInput& in, const std::vector<uint32_t>& code, uint32_t log_max_rows,
                    std::vector<std::vector<Col>>& compact, uint32_t log_size[N_COMPONENTS]) override {
    if (host_tables) { Backend::trace_tables(in, code, log_max_rows, compact, log_size); return; }
    uint64_t rows[N_COMPONENTS];
    table_rows_from_stats(in.stats, code.size(), rows);   // throws InvalidEndOfExecution / empty trace like the host builders
    uint64_t w[16] = {0};
    w[0] = in.stats.steps; w[1] = in.stats.memory_rows;
    for (int k = 0; k < 8; k++) w[2 + k] = in.stats.op_count[k];
    w[10] = in.stats.zero_ci; w[11] = in.stats.zero_ci_index; w[12] = in.stats.max_mp; w[13] = in.stats.max_ip;
    sc_trace* t = nullptr;
    ck(sc_trace_upload(ctx, (const uint32_t*)in.regs, in.n, code.data(), code.size(), in.mvi_filled ? 0 : 1, &t));
    // "inputs resident in HBM": a mark on the compute stream right behind the upload, in front of the first table kernel
    ck(sc_ctx_join_uploads(ctx));
    if (ev_resident) sc_event_free(ctx, ev_resident);
    ev_resident = nullptr;
    ck(sc_event_record(ctx, &ev_resident));
    h2d_bytes = (in.n * 7 + code.size()) * 4;
    std::vector<sc_col*> cols(128, nullptr);
    if (sc_trace_build_tables(ctx, t, w, log_max_rows, cols.data(), log_size)) {
      std::string msg = sc_last_error();
      sc_trace_free(ctx, t);
      throw std::runtime_error(msg);
    }
    if (pending_trace) sc_trace_free(ctx, pending_trace);
    pending_trace = t;
    compact.assign(N_COMPONENTS, {});
    size_t k = 0;
    for (int c = 0; c < N_COMPONENTS; c++) for (int j = 0; j < N_MAIN_COLS[c]; j++) compact[c].push_back(cols[k++]);
  }
  // called by the driver right after a read-back it needs anyway: the flags the table kernels raised
  void check_tables() override {
    if (!pending_trace) return;
    uint32_t flags = 0;
    int32_t r = sc_trace_status(ctx, pending_trace, &flags);
    sc_trace_free(ctx, pending_trace);
    pending_trace = nullptr;
    ck(r);
    if (flags) throw std::runtime_error("device table building: the trace disagrees with its statistics (flags " + std::to_string(flags) + ")");
  }
  Col broadcast16(Col c) override { sc_col* o; ck(sc_col_broadcast16(ctx, h(c), &o)); return o; }
  Col zeros(size_t n) override { sc_col* c; ck(sc_col_zeros(ctx, n, &c)); return c; }
  size_t len(Col c) override { return sc_col_len(h(c)); }
  void read(Col c, size_t off, size_t n, uint32_t* out) override { ck(sc_col_read(ctx, h(c), off, n, out)); }
  void free_col(Col c) override { ck(sc_col_free(ctx, h(c))); }
  std::vector<uint32_t> gather(const std::vector<Col>& cols, const std::vector<size_t>& offsets, uint32_t words) override {
    std::vector<uint32_t> out(cols.size() * words);
    std::vector<uint64_t> off(offsets.begin(), offsets.end());
    ck(sc_gather(ctx, (sc_col* const*)cols.data(), off.data(), (uint32_t)cols.size(), words, out.data()));
    return out;
  }

  // The tree depends only on root_log: the context computes it once and keeps it (the reference recomputes it per proof,
  // brainfuck_air/mod.rs:480-484).  cache_twiddles = false restores that behaviour.
  bool cache_twiddles = true;
  sc_twiddles* own_tw = nullptr;
  void precompute_twiddles(uint32_t root_log) override {
    if (cache_twiddles) { ck(sc_twiddles_cached(ctx, root_log, &tw)); return; }
    if (own_tw) sc_twiddles_free(ctx, own_tw);
    ck(sc_precompute_twiddles(ctx, root_log, &own_tw));
    tw = own_tw;
  }
  void interpolate(const std::vector<Col>& cols) override { ck(sc_interpolate(ctx, (sc_col* const*)cols.data(), (uint32_t)cols.size(), tw)); }
  std::vector<Col> evaluate(const std::vector<Col>& coeffs, uint32_t log_blowup) override {
    std::vector<Col> out(coeffs.size());
    if (coeffs.empty()) return out;
    ck(sc_evaluate(ctx, (sc_col* const*)coeffs.data(), (uint32_t)coeffs.size(), log_blowup, tw, (sc_col**)out.data()));
    return out;
  }
  std::vector<QM31> eval_at_point(const std::vector<Col>& polys, const std::vector<QPoint>& pts) override {
    std::vector<QM31> out(polys.size());
    static_assert(sizeof(QPoint) == 32 && sizeof(QM31) == 16, "layout");
    ck(sc_eval_at_point(ctx, (sc_col* const*)polys.data(), (uint32_t)polys.size(), (const uint32_t*)pts.data(), (uint32_t*)out.data()));
    return out;
  }
  std::vector<Col> interpolate_repeated(const std::vector<Col>& cols, uint32_t rep) override {
    std::vector<Col> out(cols.size());
    if (cols.empty()) return out;
    ck(sc_interpolate_repeated(ctx, (sc_col* const*)cols.data(), (uint32_t)cols.size(), rep, tw, (sc_col**)out.data()));
    return out;
  }
  std::vector<Col> evaluate_repeated(const std::vector<Col>& coeffs, uint32_t rep, uint32_t log_blowup) override {
    std::vector<Col> out(coeffs.size());
    if (coeffs.empty()) return out;
    ck(sc_evaluate_repeated(ctx, (sc_col* const*)coeffs.data(), (uint32_t)coeffs.size(), rep, log_blowup, tw, (sc_col**)out.data()));
    return out;
  }
  std::vector<Col> evaluate_repeated_range(const std::vector<Col>& coeffs, uint32_t rep, uint32_t log_blowup, const std::vector<size_t>& offs,
                                           const std::vector<size_t>& cnts) override {
    std::vector<Col> out(coeffs.size());
    if (coeffs.empty()) return out;
    std::vector<uint64_t> o(offs.begin(), offs.end()), c(cnts.begin(), cnts.end());
    ck(sc_evaluate_repeated_range(ctx, (sc_col* const*)coeffs.data(), (uint32_t)coeffs.size(), rep, log_blowup, tw, o.data(), c.data(),
                                  (sc_col**)out.data()));
    return out;
  }
  std::vector<QM31> eval_at_point_repeated(const std::vector<Col>& polys, const std::vector<uint32_t>& reps, const std::vector<QPoint>& pts) override {
    std::vector<QM31> out(polys.size());
    if (polys.empty()) return out;
    ck(sc_eval_at_point_repeated(ctx, (sc_col* const*)polys.data(), reps.data(), (uint32_t)polys.size(), (const uint32_t*)pts.data(), (uint32_t*)out.data()));
    return out;
  }
  std::vector<Col> merkle_commit_repeated(const std::vector<Col>& cols, uint32_t rep, Hash* root) override {
    uint32_t max_log = 0;
    for (Col c : cols) { uint32_t l = 0; while (((size_t)1 << l) < len(c)) l++; max_log = std::max(max_log, l); }
    std::vector<Col> layers(max_log + 1);
    ck(sc_merkle_commit_repeated(ctx, (sc_col* const*)cols.data(), (uint32_t)cols.size(), rep, (sc_col**)layers.data(), nullptr,
                                 root ? root->data() : nullptr));
    return layers;
  }
  std::vector<Col> merkle_commit(const std::vector<Col>& cols, Hash* root) override {
    uint32_t max_log = 0;
    for (Col c : cols) { uint32_t l = 0; while (((size_t)1 << l) < len(c)) l++; max_log = std::max(max_log, l); }
    std::vector<Col> layers(max_log + 1);
    ck(sc_merkle_commit(ctx, (sc_col* const*)cols.data(), (uint32_t)cols.size(), (sc_col**)layers.data(), nullptr,
                        root ? root->data() : nullptr));
    return layers;
  }
  std::array<Col, 4> fold_line(const std::array<Col, 4>& src, uint32_t log, QM31 alpha) override {
    std::array<Col, 4> out;
    ck(sc_fold_line(ctx, (sc_col* const*)src.data(), log, (const uint32_t*)&alpha, tw, (sc_col**)out.data()));
    return out;
  }
  void fold_circle_into_line(const std::array<Col, 4>& dst, const std::array<Col, 4>& src, uint32_t log, QM31 alpha) override {
    ck(sc_fold_circle_into_line(ctx, (sc_col* const*)src.data(), log, (const uint32_t*)&alpha, tw, (sc_col* const*)dst.data()));
  }
  bool fused_fri = true;   // SBF_NO_FUSED_FRI clears it (A/B measurements)
  void* dchan_begin(const Hash& digest, uint32_t max_mixes) override {
    if (!fused_fri) return nullptr;
    sc_dchan* dc = nullptr;
    ck(sc_dchan_create(ctx, digest.data(), max_mixes, &dc));
    return dc;
  }
  void dchan_mix_root_draw(void* dc, Col root_col) override { ck(sc_dchan_mix_root_draw(ctx, (sc_dchan*)dc, h(root_col))); }
  std::vector<Hash> dchan_finish(void* dc, uint32_t n_mixes) override {
    std::vector<uint32_t> w(8 * (size_t)n_mixes);
    ck(sc_dchan_finish(ctx, (sc_dchan*)dc, w.data()));
    std::vector<Hash> out(n_mixes);
    for (uint32_t i = 0; i < n_mixes; i++) memcpy(out[i].data(), &w[8 * i], 32);
    return out;
  }
  uint32_t fri_tail_max_log() const override { return fused_fri ? 10u : 0u; }
  bool fri_tail_dc(void* dc, const std::array<Col, 4>& layer, uint32_t start_log, uint32_t last_log, const std::vector<std::array<Col, 4>>& quot,
                   FriTailResult& out) override {
    if (!dc || !fused_fri || start_log > 10 || start_log <= last_log || quot.size() != start_log - last_log) return false;
    const uint32_t n_tail = start_log - last_log;
    std::vector<sc_col*> qc;
    for (auto& q : quot) for (Col c : q) qc.push_back(c ? h(c) : nullptr);
    size_t n_layers = 0;
    for (uint32_t lg = start_log; lg > last_log; lg--) n_layers += lg + 1;
    std::vector<sc_col*> evals(4 * (size_t)n_tail, nullptr), layers(n_layers, nullptr);
    sc_col* last[4] = {nullptr, nullptr, nullptr, nullptr};
    sc_col* in[4] = {h(layer[0]), h(layer[1]), h(layer[2]), h(layer[3])};
    ck(sc_dchan_fri_tail(ctx, (sc_dchan*)dc, tw, in, start_log, last_log, qc.data(), evals.data(), layers.data(), last));
    out.evals.clear(); out.trees.clear();
    size_t loff = 0;
    for (uint32_t lg = start_log, t = 0; lg > last_log; lg--, t++) {
      out.evals.push_back({(Col)evals[4 * t], (Col)evals[4 * t + 1], (Col)evals[4 * t + 2], (Col)evals[4 * t + 3]});
      std::vector<Col> tr(lg + 1);
      for (uint32_t k = 0; k <= lg; k++) tr[k] = (Col)layers[loff + k];
      out.trees.push_back(std::move(tr));
      loff += lg + 1;
    }
    out.last = {(Col)last[0], (Col)last[1], (Col)last[2], (Col)last[3]};
    return true;
  }
  std::array<Col, 4> fold_line_range_dc(const std::array<Col, 4>& src, uint32_t log, size_t off, size_t n_out, void* dc, uint32_t k) override {
    std::array<Col, 4> out;
    ck(sc_fold_line_range_dc(ctx, (sc_col* const*)src.data(), log, off, n_out, (sc_dchan*)dc, k, tw, (sc_col**)out.data()));
    return out;
  }
  void fold_circle_into_line_range_dc(const std::array<Col, 4>& dst, const std::array<Col, 4>& src, uint32_t log, size_t off, size_t n_out, void* dc,
                                      uint32_t k) override {
    ck(sc_fold_circle_into_line_range_dc(ctx, (sc_col* const*)src.data(), log, off, n_out, (sc_dchan*)dc, k, tw, (sc_col* const*)dst.data()));
  }
  bool fri_commit(const std::vector<std::pair<uint32_t, std::array<Col, 4>>>& quotients, const Hash& digest, uint32_t last_log,
                  FriCommitResult& out) override {
    if (!fused_fri || quotients.empty()) return false;
    const uint32_t top = quotients[0].first;
    if (last_log + 1 >= top || last_log > 10) return false;
    std::vector<sc_col*> qc;
    std::vector<uint32_t> ql;
    for (auto& q : quotients) { ql.push_back(q.first); for (Col c : q.second) qc.push_back(h(c)); }
    const uint32_t n_inner = top - 1 - last_log;
    size_t n_layers = 0;
    for (uint32_t lg = top - 1; lg > last_log; lg--) n_layers += lg + 1;
    std::vector<sc_col*> first(top + 1, nullptr), evals(4 * (size_t)n_inner, nullptr), layers(n_layers, nullptr);
    std::vector<uint32_t> roots(8 * (size_t)(n_inner + 1)), last((size_t)4 << last_log);
    ck(sc_fri_commit(ctx, tw, qc.data(), ql.data(), (uint32_t)ql.size(), digest.data(), last_log, first.data(), evals.data(), layers.data(),
                     roots.data(), last.data()));
    out.first_layers.assign(first.begin(), first.end());
    memcpy(out.first_root.data(), roots.data(), 32);
    out.inner.resize(n_inner);
    size_t off = 0;
    for (uint32_t i = 0; i < n_inner; i++) {
      auto& L = out.inner[i];
      L.log = top - 1 - i;
      for (int k = 0; k < 4; k++) L.eval[k] = evals[4 * i + k];
      L.layers.assign(layers.begin() + off, layers.begin() + off + L.log + 1);
      off += L.log + 1;
      memcpy(L.root.data(), roots.data() + 8 * (i + 1), 32);
    }
    const size_t n = (size_t)1 << last_log;
    out.last_layer.resize(n);
    for (size_t i = 0; i < n; i++) out.last_layer[i] = sb::q_make(last[i], last[n + i], last[2 * n + i], last[3 * n + i]);
    return true;
  }
  std::array<Col, 4> accumulate_quotients(uint32_t log, const std::vector<Col>& cols, QM31 rc, const SampleBatchesFlat& b) override {
    std::array<Col, 4> out;
    ck(sc_accumulate_quotients(ctx, log, (sc_col* const*)cols.data(), (uint32_t)cols.size(), (const uint32_t*)&rc, b.points.data(),
                               b.sizes.data(), b.entry_cols.data(), b.entry_vals.data(), (uint32_t)b.sizes.size(), (sc_col**)out.data()));
    return out;
  }
  void accumulate(const std::array<Col, 4>& dst, const std::array<Col, 4>& src) override {
    ck(sc_accumulate(ctx, (sc_col* const*)dst.data(), (sc_col* const*)src.data()));
  }
  uint64_t grind(const Hash& digest, uint32_t pow_bits) override { uint64_t n; ck(sc_grind(ctx, digest.data(), pow_bits, &n)); return n; }
  Col gen_is_first(uint32_t log_size) override { sc_col* c; ck(sc_gen_is_first(ctx, log_size, &c)); return c; }
  Col is_first_poly(uint32_t log_size) override { sc_col* c; ck(sc_is_first_coeffs(ctx, log_size, tw, &c)); return c; }
  Col is_first_lde(uint32_t log_size, uint32_t log_blowup, size_t row_off, size_t n_rows) override {
    sc_col* c; ck(sc_is_first_lde(ctx, log_size, log_blowup, tw, row_off, n_rows, &c)); return c;
  }
  std::vector<Col> logup_generate(int comp, const std::vector<Col>& main, const InteractionElements& el, QM31& claimed) override {
    std::vector<Col> out(4 * N_LOGUP_COLS[comp]);
    ck(sc_logup_generate(ctx, comp, (sc_col* const*)main.data(), (uint32_t)main.size(), LOG_N_LANES, (const uint32_t*)&el,
                         (sc_col**)out.data(), (uint32_t*)&claimed));
    return out;
  }
  std::vector<Col> logup_generate_deferred(int comp, const std::vector<Col>& main, const InteractionElements& el) override {
    std::vector<Col> out(4 * N_LOGUP_COLS[comp]);
    ck(sc_logup_generate(ctx, comp, (sc_col* const*)main.data(), (uint32_t)main.size(), LOG_N_LANES, (const uint32_t*)&el,
                         (sc_col**)out.data(), nullptr));
    return out;
  }
  void eval_constraints(int comp, uint32_t log_size, const std::vector<Col>& m, const std::vector<Col>& it, Col is_first,
                        const InteractionElements& el, QM31 total, const std::vector<QM31>& coeffs, const std::array<Col, 4>& acc) override {
    ck(sc_eval_constraints(ctx, comp, log_size, (sc_col* const*)m.data(), (uint32_t)m.size(), (sc_col* const*)it.data(), (uint32_t)it.size(),
                           h(is_first), (const uint32_t*)&el, (const uint32_t*)&total, (const uint32_t*)coeffs.data(), (sc_col* const*)acc.data()));
  }

  // ---- multi-GPU extension (include/stwo_cuda_sharded.h)
  sc_comm* comm = nullptr;
  int rank() const override { return comm ? sc_comm_rank(comm) : 0; }
  int world() const override { return comm ? sc_comm_world(comm) : 1; }
  Col alloc(size_t n) override { sc_col* c; ck(sc_col_uninit(ctx, n, &c)); return c; }
  Col view(Col c, size_t off, size_t n) override { sc_col* o; ck(sc_col_view(ctx, h(c), off, n, &o)); return o; }
  void copy(Col dst, size_t dst_off, Col src, size_t src_off, size_t n) override { ck(sc_col_copy(ctx, h(dst), dst_off, h(src), src_off, n)); }
  Col commit_layer(uint32_t log, Col prev, const std::vector<Col>& cols) override {
    sc_col* o;
    ck(sc_merkle_commit_layer(ctx, log, h(prev), (sc_col* const*)cols.data(), (uint32_t)cols.size(), &o));
    return o;
  }
  void* mark() override { sc_event* e; ck(sc_event_record(ctx, &e)); return e; }
  double gap_ms(void* a, void* b) override {
    float ms = 0;
    ck(sc_event_elapsed(ctx, (sc_event*)a, (sc_event*)b, &ms));
    sc_event_free(ctx, (sc_event*)a); sc_event_free(ctx, (sc_event*)b);
    return ms;
  }
  Col commit_layer_repeated(uint32_t log, Col prev, const std::vector<Col>& cols, uint32_t rep) override {
    sc_col* o;
    ck(sc_merkle_commit_layer_repeated(ctx, log, h(prev), (sc_col* const*)cols.data(), (uint32_t)cols.size(), rep, &o));
    return o;
  }
  void pack_exchange(Col send, const std::vector<Col>& cols, const std::vector<size_t>& segs, const std::vector<uint8_t>& sharded) override {
    std::vector<uint64_t> sg(segs.begin(), segs.end());
    ck(sc_pack_exchange(ctx, (sc_col* const*)cols.data(), sg.data(), sharded.data(), (uint32_t)cols.size(), (uint32_t)world(), h(send)));
  }
  void exchange_begin() override { if (comm) ck(sc_exchange_begin(ctx, comm)); }
  Col exchange_scatter(const std::vector<Col>& pieces, const std::vector<uint32_t>& dest, const std::vector<size_t>& dst_off,
                       size_t region_words) override {
    if (!comm) return nullptr;
    // Measured and checked against the golden proofs at world 2 (profiles/r2_end_prove_n2.json); the round's GPU budget ended
    // before a world-4 / world-8 run, so there it stays opt-in (SC_SCATTER_EXCHANGE=1) and the all-to-all path is the default.
    static const bool forced = getenv("SC_SCATTER_EXCHANGE") != nullptr;
    if (!forced && world() != 2) return nullptr;
    std::vector<uint64_t> off(dst_off.begin(), dst_off.end());
    sc_col* out = nullptr;
    ck(sc_exchange_scatter(ctx, comm, (sc_col* const*)pieces.data(), dest.data(), off.data(), (uint32_t)pieces.size(), region_words, &out));
    return out;
  }
  Col exchange_push(const std::vector<Col>& cols, const std::vector<size_t>& segs, const std::vector<uint8_t>& sharded,
                    const std::vector<size_t>& recv_counts) override {
    if (!comm) return nullptr;
    std::vector<uint64_t> sg(segs.begin(), segs.end()), rc(recv_counts.begin(), recv_counts.end());
    sc_col* out = nullptr;
    ck(sc_exchange_push(ctx, comm, (sc_col* const*)cols.data(), sg.data(), sharded.data(), (uint32_t)cols.size(), rc.data(), &out));
    return out;
  }
  void all_to_all(Col send, const std::vector<size_t>& sc, Col recv, const std::vector<size_t>& rc) override {
    if (!comm) { copy(recv, 0, send, 0, sc[0]); return; }
    std::vector<uint64_t> s(sc.begin(), sc.end()), r(rc.begin(), rc.end());
    ck(sc_all_to_all(ctx, comm, h(send), s.data(), h(recv), r.data()));
  }
  void all_gather(Col send, Col recv, size_t n) override {
    if (!comm) { copy(recv, 0, send, 0, n); return; }
    ck(sc_all_gather(ctx, comm, h(send), h(recv), n));
  }
  void allreduce_host(uint32_t* buf, size_t n) override { if (comm) ck(sc_allreduce_host_u32(ctx, comm, buf, n)); }
  std::array<Col, 4> fold_line_range(const std::array<Col, 4>& src, uint32_t log, size_t off, size_t n_out, QM31 alpha) override {
    std::array<Col, 4> out;
    ck(sc_fold_line_range(ctx, (sc_col* const*)src.data(), log, off, n_out, (const uint32_t*)&alpha, tw, (sc_col**)out.data()));
    return out;
  }
  void fold_circle_into_line_range(const std::array<Col, 4>& dst, const std::array<Col, 4>& src, uint32_t log, size_t off, size_t n_out,
                                   QM31 alpha) override {
    ck(sc_fold_circle_into_line_range(ctx, (sc_col* const*)src.data(), log, off, n_out, (const uint32_t*)&alpha, tw, (sc_col* const*)dst.data()));
  }
  std::array<Col, 4> accumulate_quotients_range(uint32_t log, size_t row_off, size_t n_rows, const std::vector<Col>& cols, QM31 rc,
                                                const SampleBatchesFlat& b) override {
    std::array<Col, 4> out;
    ck(sc_accumulate_quotients_range(ctx, log, row_off, n_rows, (sc_col* const*)cols.data(), (uint32_t)cols.size(), (const uint32_t*)&rc,
                                     b.points.data(), b.sizes.data(), b.entry_cols.data(), b.entry_vals.data(), (uint32_t)b.sizes.size(),
                                     (sc_col**)out.data()));
    return out;
  }
  Col shift_prev(Col c, uint32_t trace_log) override { sc_col* o; ck(sc_shift_prev(ctx, h(c), trace_log, &o)); return o; }
  void accumulate_col(Col dst, Col src) override { ck(sc_accumulate_col(ctx, h(dst), h(src))); }
  void prefix_sum(Col c) override { ck(sc_prefix_sum_bitrev(ctx, h(c))); }
  std::vector<Col> logup_generate_sel(int comp, const std::vector<Col>& main, const InteractionElements& el,
                                      const std::vector<uint8_t>& want) override {
    std::vector<Col> out(4 * N_LOGUP_COLS[comp], nullptr);
    ck(sc_logup_generate_sel(ctx, comp, (sc_col* const*)main.data(), (uint32_t)main.size(), LOG_N_LANES, (const uint32_t*)&el, want.data(),
                             (sc_col**)out.data()));
    return out;
  }
  void eval_constraints_range(int comp, uint32_t log_size, size_t row_off, size_t n_rows, const std::vector<Col>& m, const std::vector<Col>& it,
                              const std::array<Col, 4>& prev, Col is_first, const InteractionElements& el, QM31 total,
                              const std::vector<QM31>& coeffs, const std::array<Col, 4>& acc) override {
    ck(sc_eval_constraints_range(ctx, comp, log_size, row_off, n_rows, (sc_col* const*)m.data(), (uint32_t)m.size(), (sc_col* const*)it.data(),
                                 (uint32_t)it.size(), (sc_col* const*)prev.data(), h(is_first), (const uint32_t*)&el, (const uint32_t*)&total,
                                 (const uint32_t*)coeffs.data(), (sc_col* const*)acc.data()));
  }
};

// Runs the VM.  Device-side table building wants the register rows in pinned memory (the upload is then a DMA beside the
// kernels already queued) and fills mvi itself, so the machine writes straight into the context's host arena and skips its
// batched inversions.  With host tables the trace stays in the machine's own vector.
TraceInput run_machine(sc_ctx* ctx, Machine& vm, uint32_t log_max_rows, bool host_tables, double& vm_ms) {
  auto t0 = std::chrono::steady_clock::now();
  if (!host_tables) {
    const uint32_t lg = log_max_rows >= LOG_N_LANES && log_max_rows <= 28 ? log_max_rows - LOG_N_LANES : 24;
    const size_t cap = ((size_t)1 << lg) + 1;   // one more row than the Processor table can take: the size check reports it
    void* buf = nullptr;
    sc_host_arena_reset(ctx);
    if (sc_host_arena_alloc(ctx, cap * sizeof(Registers), &buf)) throw std::runtime_error(sc_last_error());
    vm.sink = (Registers*)buf; vm.sink_cap = cap;
    vm.skip_inverses = true;
  }
  vm.execute();
  vm_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  TraceInput in;
  in.regs = vm.rows(); in.n = vm.n_rows(); in.stats = vm.stats; in.mvi_filled = !vm.skip_inverses;
  return in;
}

// sc_ctx_arena_begin / _end around one proof (also when the proof throws)
struct ArenaBracket {
  sc_ctx* ctx; bool on;
  ArenaBracket(sc_ctx* c, bool enable) : ctx(c), on(enable && sc_ctx_arena_begin(c) == SC_OK) {}
  ~ArenaBracket() { if (on) sc_ctx_arena_end(ctx); }
};

thread_local std::string g_sbf_err;
char* dup_string(const std::string& s) {
  char* p = (char*)malloc(s.size() + 1);
  memcpy(p, s.c_str(), s.size() + 1);
  return p;
}

}  // namespace

struct sbf_proof {
  BrainfuckProof proof;
  ProverConfig cfg;
  std::string report;  // JSON: steps, log sizes, per-stage host-clock milliseconds
  std::vector<uint8_t> output;
};

extern "C" {

const char* sbf_last_error(void) { return g_sbf_err.c_str(); }

// SBF_CACHE_PREPROCESSED: the cache lives in slot 0 of the context and dies with it
static void drop_preprocessed_cache(sc_ctx* ctx, void* p) {
  PreprocessedCache* c = (PreprocessedCache*)p;
  try { CudaBackendImpl B(ctx); c->release(B); } catch (...) {}
  delete c;
}
static PreprocessedCache* preprocessed_cache(sc_ctx* ctx) {
  PreprocessedCache* c = (PreprocessedCache*)sc_ctx_attached(ctx, 0);
  if (!c) {
    c = new PreprocessedCache;
    if (sc_ctx_attach(ctx, 0, c, drop_preprocessed_cache)) { delete c; throw std::runtime_error(sc_last_error()); }
  }
  return c;
}
int32_t sbf_preprocessed_cache_clear(sc_ctx* ctx) {
  if (!ctx) return SC_EINVAL;
  PreprocessedCache* c = (PreprocessedCache*)sc_ctx_attached(ctx, 0);
  if (c) { try { CudaBackendImpl B(ctx); c->release(B); } catch (const std::exception& e) { g_sbf_err = e.what(); return SC_ECUDA; } }
  return SC_OK;
}

// `brainfuck_prover prove --code <code>` with stdin bytes `input`: run the VM on the host, prove on the device.
// log_max_rows = LOG_MAX_ROWS (24; 20 in the reference's tests).  Returns 0 or SC_EPROOF.
int32_t sbf_prove_sharded(sc_ctx* ctx, sc_comm* comm, const char* code, const uint8_t* input, size_t input_len, uint32_t log_max_rows,
                          uint32_t flags, sbf_proof** out);
int32_t sbf_prove(sc_ctx* ctx, const char* code, const uint8_t* input, size_t input_len, uint32_t log_max_rows, uint32_t flags,
                  sbf_proof** out) {
  if (flags & 4u) return sbf_prove_sharded(ctx, nullptr, code, input, input_len, log_max_rows, flags, out);  // SBF_SHARDED_DRIVER
  const uint64_t mark = sc_ctx_mark(ctx);
  PreprocessedCache* pp = nullptr;
  uint64_t pp_fills = 0;
  try {
    if (!ctx || !code || !out) throw std::runtime_error("null argument");
    if (flags & 8u) { pp = preprocessed_cache(ctx); pp_fills = pp->fills; }  // SBF_CACHE_PREPROCESSED
    std::vector<uint32_t> program = compile(code);
    Machine vm(program, std::vector<uint8_t>(input, input + input_len));
    double vm_ms = 0;
    const bool host_tables = (flags & 16u) != 0;   // SBF_HOST_TABLES
    // the VM runs when the prover asks for the trace: after the preprocessed phase is enqueued unless SBF_NO_OVERLAP
    TraceSource run_vm = [&]() { return run_machine(ctx, vm, log_max_rows, host_tables, vm_ms); };
    CudaBackendImpl B(ctx);
    B.host_tables = host_tables;
    B.fused_fri = !(flags & 32u);   // SBF_NO_FUSED_FRI
    // every column of the proof from one slab (not with the preprocessed-tree cache, whose columns outlive the proof)
    ArenaBracket arena(ctx, !(flags & 64u) && !(flags & 8u));
    ProverConfig cfg;
    cfg.log_max_rows = log_max_rows;
    cfg.overlap_host = !(flags & 1u);  // SBF_NO_OVERLAP: VM run and tables before any device work (bench.py's device-path timing)
    B.cache_twiddles = !(flags & 2u);  // SBF_NO_TWIDDLE_CACHE: recompute the twiddle tree in every proof, as the reference does
    auto t1 = std::chrono::steady_clock::now();
    // stage boundaries wait for the device only while the profiling scopes are on (sc_ctx_profile): stages_ms are then device-complete
    // times; otherwise they are the host's enqueue times and the stages overlap as the stream allows
    ProveResult r = prove_brainfuck(B, program, run_vm, cfg, [&] { if (sc_ctx_profiling(ctx)) sc_ctx_sync(ctx); }, pp);
    const double device_ms = B.ms_since_resident();
    double prove_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
    sbf_proof* p = new sbf_proof{std::move(r.proof), cfg, "", vm.output};
    std::ostringstream o;
    o << "{\"steps\":" << vm.n_rows() << ",\"vm_ms\":" << vm_ms << ",\"prove_ms\":" << prove_ms << ",\"device_ms\":" << device_ms << ",\"h2d_bytes\":" << B.h2d_bytes
      << ",\"stages_synced\":" << (sc_ctx_profiling(ctx) ? "true" : "false") << ",\"program_words\":" << program.size() << ",\"log_sizes\":[";
    for (int c = 0; c < N_COMPONENTS; c++) o << (c ? "," : "") << p->proof.log_size[c];
    o << "],\"stages_ms\":{";
    for (size_t i = 0; i < r.times.ms.size(); i++) o << (i ? "," : "") << "\"" << r.times.ms[i].first << "\":" << r.times.ms[i].second;
    o << "}}";
    p->report = o.str();
    *out = p;
    return SC_OK;
  } catch (const std::exception& e) {
    g_sbf_err = e.what();
    if (ctx) { sc_ctx_sync(ctx); sc_ctx_release_since(ctx, mark); }  // nothing the failed proof allocated stays on the device
    if (pp && pp->fills != pp_fills) pp->forget();  // a tree cached by this very call went with the release above
    return SC_EPROOF;
  }
}

// One proof over the ranks of `comm` (one process per GPU; every rank calls this with the same program and gets the same
// proof).  comm == NULL runs the sharded driver on a single GPU (world 1): same proof as sbf_prove.
int32_t sbf_prove_sharded(sc_ctx* ctx, sc_comm* comm, const char* code, const uint8_t* input, size_t input_len, uint32_t log_max_rows,
                          uint32_t flags, sbf_proof** out) {
  const uint64_t mark = sc_ctx_mark(ctx);
  try {
    if (!ctx || !code || !out) throw std::runtime_error("null argument");
    // the sharded driver has no preprocessed-tree cache: say so instead of ignoring the flag
    if (flags & 8u) throw std::runtime_error("SBF_CACHE_PREPROCESSED is not supported by the sharded driver");
    std::vector<uint32_t> program = compile(code);
    Machine vm(program, std::vector<uint8_t>(input, input + input_len));
    double vm_ms = 0;
    const bool host_tables = (flags & 16u) != 0;   // SBF_HOST_TABLES
    // called on the prover's host thread, beside the preprocessed phase
    TraceSource run_vm = [&]() { return run_machine(ctx, vm, log_max_rows, host_tables, vm_ms); };
    CudaBackendImpl B(ctx);
    B.host_tables = host_tables;
    B.fused_fri = !(flags & 32u);   // SBF_NO_FUSED_FRI
    ArenaBracket arena(ctx, !(flags & 64u));
    B.comm = comm;
    ProverConfig cfg;
    cfg.log_max_rows = log_max_rows;
    cfg.shard_min_log = getenv("SBF_SHARD_MIN_LOG") ? (uint32_t)atoi(getenv("SBF_SHARD_MIN_LOG")) : 16;
    cfg.overlap_host = !(flags & 1u);  // SBF_NO_OVERLAP: VM run and tables before any device work (bench.py's device-path timing)
    B.cache_twiddles = !(flags & 2u);
    auto t1 = std::chrono::steady_clock::now();
    ProveResult r = prove_brainfuck_sharded(B, program, run_vm, cfg, [&] { if (sc_ctx_profiling(ctx)) sc_ctx_sync(ctx); });
    const double device_ms = B.ms_since_resident();
    double prove_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
    sbf_proof* p = new sbf_proof{std::move(r.proof), cfg, "", vm.output};
    std::ostringstream o;
    o << "{\"steps\":" << vm.n_rows() << ",\"vm_ms\":" << vm_ms << ",\"prove_ms\":" << prove_ms << ",\"device_ms\":" << device_ms << ",\"h2d_bytes\":" << B.h2d_bytes
      << ",\"stages_synced\":" << (sc_ctx_profiling(ctx) ? "true" : "false") << ",\"program_words\":" << program.size() << ",\"world\":" << B.world() << ",\"log_sizes\":[";
    for (int c = 0; c < N_COMPONENTS; c++) o << (c ? "," : "") << p->proof.log_size[c];
    o << "],\"stages_ms\":{";
    for (size_t i = 0; i < r.times.ms.size(); i++) o << (i ? "," : "") << "\"" << r.times.ms[i].first << "\":" << r.times.ms[i].second;
    o << "}}";
    p->report = o.str();
    *out = p;
    return SC_OK;
  } catch (const std::exception& e) {
    g_sbf_err = e.what();
    if (ctx) { sc_ctx_sync(ctx); sc_ctx_release_since(ctx, mark); }  // nothing the failed proof allocated stays on the device
    return SC_EPROOF;
  }
}

// `brainfuck_prover verify`: pure host code.  Returns 0 or SC_EVERIFY (message in sbf_last_error()).
int32_t sbf_verify(const sbf_proof* p) {
  try {
    if (!p) throw std::runtime_error("null proof");
    verify_brainfuck(p->proof, p->cfg);
    return SC_OK;
  } catch (const std::exception& e) {
    g_sbf_err = e.what();
    return SC_EVERIFY;
  }
}
char* sbf_proof_json(const sbf_proof* p) { return p ? dup_string(proof_to_json(p->proof)) : nullptr; }
char* sbf_proof_report(const sbf_proof* p) { return p ? dup_string(p->report) : nullptr; }
size_t sbf_proof_output(const sbf_proof* p, uint8_t* buf, size_t cap) {
  if (!p) return 0;
  size_t n = std::min(cap, p->output.size());
  if (buf && n) memcpy(buf, p->output.data(), n);
  return p->output.size();
}
void sbf_string_free(char* s) { free(s); }
void sbf_proof_free(sbf_proof* p) { delete p; }

// `brainfuck_prover verify <file>` (bin/brainfuck_prover.rs:145-152): deserialise the serde JSON text and verify it with the
// VERIFIER's own parameters — its LOG_MAX_ROWS and the default PcsConfig — not with anything the prover recorded.
int32_t sbf_verify_json(const char* json, uint32_t log_max_rows) {
  try {
    if (!json) throw std::runtime_error("null proof");
    BrainfuckProof p = proof_from_json(json);
    ProverConfig cfg;
    cfg.log_max_rows = log_max_rows;
    verify_brainfuck(p, cfg);
    return SC_OK;
  } catch (const std::exception& e) {
    g_sbf_err = e.what();
    return SC_EVERIFY;
  }
}
// The same text as a proof object (sbf_proof_json o sbf_proof_from_json is the identity on well-formed proofs).
int32_t sbf_proof_from_json(const char* json, uint32_t log_max_rows, sbf_proof** out) {
  try {
    if (!json || !out) throw std::runtime_error("null argument");
    ProverConfig cfg;
    cfg.log_max_rows = log_max_rows;
    *out = new sbf_proof{proof_from_json(json), cfg, "{}", {}};
    return SC_OK;
  } catch (const std::exception& e) {
    g_sbf_err = e.what();
    return SC_EVERIFY;
  }
}

}  // extern "C"
