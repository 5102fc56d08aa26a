// The backend surface the protocol driver (prover.hpp) is written against — Stwo's `Backend` trait family plus the two
// SimdBackend-only pieces (LogUp generation, constraint evaluation), SURVEY.md §8b.  The product implements it over the C
// ABI (cuda_backend.cu); the CPU oracle implements it with scalar loops (oracle/orc_backend.cc, tests only).
#pragma once
#include <array>
#include <mutex>
#include <vector>
#include "air.hpp"
#include "channel.hpp"
#include "tables.hpp"

namespace sbf {

typedef void* Col;  // opaque column handle owned by the backend
struct QPoint { QM31 x, y; };

struct SampleBatchesFlat {   // ColumnSampleBatch list for one LDE size, flattened as the C ABI wants it
  std::vector<uint32_t> points;      // nb x 8
  std::vector<uint32_t> sizes;       // nb
  std::vector<uint32_t> entry_cols;  // column index inside the size group
  std::vector<uint32_t> entry_vals;  // 4 words per entry
};

// The VM's output as the table builders see it (crates/brainfuck_vm/src/registers.rs:5-21: seven words per step, in clk
// order) together with the sizes derived from it (vm.hpp TraceStats).
struct TraceInput {
  const Registers* regs = nullptr;
  size_t n = 0;
  TraceStats stats;
  bool mvi_filled = true;   // false: the mvi column is still zero and the table builder computes mv^-1 itself
};
// Rows of every table (a power of two) from the trace statistics — what the reference learns by building the tables
// (memory/table.rs:259-318, instruction/table.rs:250-281, program/table.rs:36-46, processor/table.rs:195-207,
// instructions/table.rs:293-328, jump/table.rs:264-297, end_of_execution/table.rs:71-77).
inline void table_rows_from_stats(const TraceStats& st, size_t program_len, uint64_t rows[N_COMPONENTS]) {
  auto p2 = [](uint64_t n) { uint64_t p = 1; while (p < n) p <<= 1; return p; };
  if (st.steps == 0) throw std::runtime_error("empty trace");
  if (program_len == 0) throw std::runtime_error("empty program");
  if (st.zero_ci != 1) throw std::runtime_error("InvalidEndOfExecution");
  rows[MEMORY] = p2(st.memory_rows);
  rows[INSTRUCTION] = p2(program_len + st.steps);
  rows[PROGRAM] = p2(program_len);
  rows[PROCESSOR] = p2(st.steps);
  for (int k = 0; k < 8; k++) rows[JNZ + k] = st.op_count[k] ? p2(2ull * st.op_count[k]) / 2 : 1;
  rows[EOE] = 1;
}

struct Backend {
  virtual ~Backend() {}
  virtual const char* name() const = 0;
  // The 13 `trace_evaluation`s (brainfuck_air/mod.rs:511-547) in lane-compact form: compact[c] receives N_MAIN_COLS[c] columns of
  // one word per table row and log_size[c] the column log size (rows x 16 lanes).  Default: the host builders of tables.hpp
  // and one upload per column; the CUDA backend builds the tables on the device from the uploaded registers (csrc/tables.cu).
  // Throws "component too large: <name>" before anything is allocated when a table exceeds 2^(log_max_rows - 4) rows.
  virtual void trace_tables(const TraceInput& in, const std::vector<uint32_t>& code, uint32_t log_max_rows,
                            std::vector<std::vector<Col>>& compact, uint32_t log_size[N_COMPONENTS]) {
    std::vector<Registers> regs(in.regs, in.regs + in.n);
    if (!in.mvi_filled) for (auto& r : regs) r.mvi = r.mv ? sb::m_inv(r.mv) : 0;
    static std::mutex arena_mu;   // current_arena() is process-wide: concurrent proofs take turns building host tables
    std::lock_guard<std::mutex> lk(arena_mu);
    struct ArenaGuard {  // tables are built into the backend's host arena (pinned memory on CUDA) and die before it is released
      HostArena* prev;
      explicit ArenaGuard(HostArena* a) : prev(current_arena()) { current_arena() = a; }
      ~ArenaGuard() { current_arena() = prev; }
    } guard(host_arena());
    std::vector<Table> tables = build_tables(regs, code, log_max_rows);
    compact.assign(N_COMPONENTS, {});
    for (int c = 0; c < N_COMPONENTS; c++) {
      log_size[c] = tables[c].log_size;
      if (tables[c].log_size > log_max_rows) throw std::runtime_error(std::string("component too large: ") + COMPONENT_NAMES[c]);
    }
    for (int c = 0; c < N_COMPONENTS; c++)
      for (auto& col : tables[c].cols) compact[c].push_back(from_host(col.data(), col.size()));   // synchronous: the tables die here
  }
  // Column<T>
  virtual Col from_host(const uint32_t* v, size_t n) = 0;
  // Optional fast upload path: a host arena whose memory the backend can DMA from directly, and an upload that does not
  // wait (the caller keeps the memory alive until the next synchronising call).
  virtual HostArena* host_arena() { return nullptr; }
  virtual Col from_host_async(const uint32_t* v, size_t n) { return from_host(v, n); }
  virtual Col broadcast16(Col c) = 0;
  virtual Col zeros(size_t n) = 0;
  virtual size_t len(Col c) = 0;
  virtual void read(Col c, size_t off, size_t n, uint32_t* out) = 0;
  virtual void free_col(Col c) = 0;
  // batched Column::at: out[i*words .. +words) = cols[i][offsets[i] .. +words)
  virtual std::vector<uint32_t> gather(const std::vector<Col>& cols, const std::vector<size_t>& offsets, uint32_t words) = 0;
  // PolyOps
  virtual void precompute_twiddles(uint32_t root_log) = 0;
  virtual void interpolate(const std::vector<Col>& cols) = 0;                                  // in place
  virtual std::vector<Col> evaluate(const std::vector<Col>& coeffs, uint32_t log_blowup) = 0;
  virtual std::vector<QM31> eval_at_point(const std::vector<Col>& polys, const std::vector<QPoint>& pts) = 0;
  // ---- lane-repeated columns: a stored value stands for 2^rep consecutive rows of the evaluation (rep = LOG_N_LANES for
  // the main trace, whose table rows are written into all 16 SIMD lanes: components/<name>/table.rs trace_evaluation).
  // The polynomial of such a column has one non-zero coefficient in 2^rep; `interpolate_repeated` leaves those in the
  // column (coefficient j = coefficient j << rep of the full vector) and the two evaluators take that compact form.
  virtual std::vector<Col> interpolate_repeated(const std::vector<Col>& values, uint32_t rep) = 0;                       // new columns; inputs kept
  virtual std::vector<Col> evaluate_repeated(const std::vector<Col>& coeffs, uint32_t rep, uint32_t log_blowup) = 0;     // full-length evaluations
  // rows [offs[i], offs[i] + cnts[i]) of evaluate_repeated's column i (multiples of 2^rep): a rank's share in the sharded prover
  virtual std::vector<Col> evaluate_repeated_range(const std::vector<Col>& coeffs, uint32_t rep, uint32_t log_blowup,
                                                   const std::vector<size_t>& offs, const std::vector<size_t>& cnts) = 0;
  virtual std::vector<QM31> eval_at_point_repeated(const std::vector<Col>& polys, const std::vector<uint32_t>& reps,
                                                   const std::vector<QPoint>& pts) = 0;
  // merkle_commit of full-length columns that all repeat each value 2^rep times (the deepest `rep` layers then repeat too)
  virtual std::vector<Col> merkle_commit_repeated(const std::vector<Col>& cols, uint32_t rep, Hash* root) = 0;
  // MerkleOps: layers[k] = layer of log size k
  // root == nullptr: enqueue only (the caller reads layers[0] later), so the host can overlap other work
  virtual std::vector<Col> merkle_commit(const std::vector<Col>& cols, Hash* root) = 0;
  // MerkleOps::commit_on_layer on 2^log rows (also used on row ranges: the node function is local to a row)
  virtual Col commit_layer(uint32_t log, Col prev, const std::vector<Col>& cols) = 0;
  // same when every column (hence every node) of these 2^log rows repeats 2^rep times: backends may hash one node per group
  virtual Col commit_layer_repeated(uint32_t log, Col prev, const std::vector<Col>& cols, uint32_t rep) { (void)rep; return commit_layer(log, prev, cols); }
  // FriOps
  virtual std::array<Col, 4> fold_line(const std::array<Col, 4>& src, uint32_t log, QM31 alpha) = 0;
  virtual void fold_circle_into_line(const std::array<Col, 4>& dst, const std::array<Col, 4>& src, uint32_t log, QM31 alpha) = 0;
  // FriProver::commit as one backend call with the channel on the device (optional; csrc/fri.cu).  Returns false when the
  // backend has no fused path — the driver then runs the layer-by-layer loop over fold_* / merkle_commit with its own channel.
  struct FriCommitResult {
    std::vector<Col> first_layers; Hash first_root;
    struct Inner { std::array<Col, 4> eval; uint32_t log; std::vector<Col> layers; Hash root; };
    std::vector<Inner> inner;
    std::vector<QM31> last_layer;   // 2^last_log values
  };
  virtual bool fri_commit(const std::vector<std::pair<uint32_t, std::array<Col, 4>>>& quotients, const Hash& channel_digest, uint32_t last_log,
                          FriCommitResult& out) { (void)quotients; (void)channel_digest; (void)last_log; (void)out; return false; }
  // The same idea for drivers that walk the layers themselves (prover_sharded.hpp): a device-resident transcript.  dchan_begin
  // returns NULL when the backend has none.  Coefficient #k is the one drawn after the k-th mix (k = 0: the circle fold's).
  virtual void* dchan_begin(const Hash& channel_digest, uint32_t max_mixes) { (void)channel_digest; (void)max_mixes; return nullptr; }
  virtual void dchan_mix_root_draw(void* dc, Col root_col) { (void)dc; (void)root_col; }
  virtual std::vector<Hash> dchan_finish(void* dc, uint32_t n_mixes) { (void)dc; (void)n_mixes; return {}; }   // waits; frees dc
  // All FRI layers from `start_log` down to last_log + 1 on the device transcript in one call (optional; false = not
  // available, the caller walks the layers).  `layer`: the replicated line evaluation of log start_log BEFORE the quotient of
  // that size is folded in; quot[t]: the quotient columns of log start_log - t + 1 (nullptr entries when there are none).
  // Out, per layer t: the committed evaluation, its tree (layers[k] = 2^k nodes); `last`: the 2^last_log values left.
  struct FriTailResult { std::vector<std::array<Col, 4>> evals; std::vector<std::vector<Col>> trees; std::array<Col, 4> last; };
  virtual uint32_t fri_tail_max_log() const { return 0; }
  virtual bool fri_tail_dc(void* dc, const std::array<Col, 4>& layer, uint32_t start_log, uint32_t last_log,
                           const std::vector<std::array<Col, 4>>& quot, FriTailResult& out) {
    (void)dc; (void)layer; (void)start_log; (void)last_log; (void)quot; (void)out; return false;
  }
  virtual std::array<Col, 4> fold_line_range_dc(const std::array<Col, 4>& src, uint32_t log, size_t out_off, size_t n_out, void* dc, uint32_t k) {
    (void)src; (void)log; (void)out_off; (void)n_out; (void)dc; (void)k; throw std::runtime_error("no device channel");
  }
  virtual void fold_circle_into_line_range_dc(const std::array<Col, 4>& dst, const std::array<Col, 4>& src, uint32_t log, size_t out_off, size_t n_out,
                                              void* dc, uint32_t k) {
    (void)dst; (void)src; (void)log; (void)out_off; (void)n_out; (void)dc; (void)k; throw std::runtime_error("no device channel");
  }
  // QuotientOps
  virtual std::array<Col, 4> accumulate_quotients(uint32_t log, const std::vector<Col>& cols, QM31 random_coeff,
                                                  const SampleBatchesFlat& b) = 0;
  // AccumulationOps
  virtual void accumulate(const std::array<Col, 4>& dst, const std::array<Col, 4>& src) = 0;
  // GrindOps
  virtual uint64_t grind(const Hash& digest, uint32_t pow_bits) = 0;
  // constraint_framework
  virtual Col gen_is_first(uint32_t log_size) = 0;
  // the interpolated IsFirst column (backends may have a closed form)
  virtual Col is_first_poly(uint32_t log_size) { Col c = gen_is_first(log_size); interpolate({c}); return c; }
  // rows [row_off, row_off + n_rows) of the evaluation of that polynomial on the domain of log_size + log_blowup.  CUDA writes
  // them in closed form (csrc/quotients.cu is_first_lde_kernel); the default transforms the polynomial and slices.
  virtual Col is_first_lde(uint32_t log_size, uint32_t log_blowup, size_t row_off, size_t n_rows) {
    Col p = is_first_poly(log_size);
    std::vector<Col> e = evaluate({p}, log_blowup);
    free_col(p);
    if (row_off == 0 && n_rows == ((size_t)1 << (log_size + log_blowup))) return e[0];
    Col out = alloc(n_rows);
    copy(out, 0, e[0], row_off, n_rows);
    free_col(e[0]);
    return out;
  }
  virtual std::vector<Col> logup_generate(int comp, const std::vector<Col>& main, const InteractionElements& el, QM31& claimed_sum) = 0;
  // same without the read-back: the claimed sum is element 1 of each of the last four returned columns (fetch them with one gather)
  virtual std::vector<Col> logup_generate_deferred(int comp, const std::vector<Col>& main, const InteractionElements& el) {
    QM31 s;
    return logup_generate(comp, main, el, s);
  }
  virtual void eval_constraints(int comp, uint32_t log_size, const std::vector<Col>& main_lde, const std::vector<Col>& inter_lde,
                                Col is_first_lde, const InteractionElements& el, QM31 total_sum, const std::vector<QM31>& coeffs,
                                const std::array<Col, 4>& accum) = 0;

  // trace_tables may defer its consistency checks to the next point where the driver waits for the device anyway
  virtual void check_tables() {}

  // device-side stopwatch: mark() a point in the queued work; gap_ms(a, b) = device time between two marks (consumes both)
  virtual void* mark() { return nullptr; }
  virtual double gap_ms(void* a, void* b) { (void)a; (void)b; return 0; }

  // ---- multi-GPU extension (prover_sharded.hpp): one rank of `world`; row-range variants and collectives
  virtual int rank() const { return 0; }
  virtual int world() const { return 1; }
  virtual Col alloc(size_t n) = 0;                                            // uninitialised
  virtual Col view(Col c, size_t off, size_t n) = 0;                          // non-owning slice (free_col drops the handle)
  virtual void copy(Col dst, size_t dst_off, Col src, size_t src_off, size_t n) = 0;
  // send[d * per_dest + off_j ..) = rows [d * seg_j, (d+1) * seg_j) of cols[j] (the whole column when !sharded[j]), per_dest = sum of segs
  virtual void pack_exchange(Col send, const std::vector<Col>& cols, const std::vector<size_t>& segs, const std::vector<uint8_t>& sharded) {
    size_t per = 0;
    for (size_t x : segs) per += x;
    for (int d = 0; d < world(); d++) {
      size_t o = 0;
      for (size_t j = 0; j < cols.size(); j++) { copy(send, (size_t)d * per + o, cols[j], sharded[j] ? (size_t)d * segs[j] : 0, segs[j]); o += segs[j]; }
    }
  }
  virtual void all_to_all(Col send, const std::vector<size_t>& send_counts, Col recv, const std::vector<size_t>& recv_counts) = 0;
  // Optional direct exchange (CUDA: peer stores into the destination's receive window, csrc/sharded.cu): exchange_begin once per
  // proof (collective); exchange_push = pack_exchange + all_to_all in one step, returning the receive buffer, or nullptr when
  // the backend has no such path (the caller then packs and calls all_to_all).  Every rank gets the same answer.
  // the same windows, pieces -> arbitrary (rank, offset): piece j goes whole to word dst_off[j] of rank dest[j]'s region of
  // region_words words (the same number on every rank); returns this rank's region or nullptr (no direct path)
  virtual Col exchange_scatter(const std::vector<Col>& pieces, const std::vector<uint32_t>& dest, const std::vector<size_t>& dst_off,
                               size_t region_words) {
    (void)pieces; (void)dest; (void)dst_off; (void)region_words; return nullptr;
  }
  virtual void exchange_begin() {}
  virtual Col exchange_push(const std::vector<Col>& cols, const std::vector<size_t>& segs, const std::vector<uint8_t>& sharded,
                            const std::vector<size_t>& recv_counts) {
    (void)cols; (void)segs; (void)sharded; (void)recv_counts; return nullptr;
  }
  virtual void all_gather(Col send, Col recv, size_t n) = 0;
  virtual void allreduce_host(uint32_t* buf, size_t n) = 0;                   // sum; exactly one contributor per slot
  virtual std::array<Col, 4> fold_line_range(const std::array<Col, 4>& src, uint32_t log, size_t out_off, size_t n_out, QM31 alpha) = 0;
  virtual void fold_circle_into_line_range(const std::array<Col, 4>& dst, const std::array<Col, 4>& src, uint32_t log, size_t out_off,
                                           size_t n_out, QM31 alpha) = 0;
  virtual std::array<Col, 4> accumulate_quotients_range(uint32_t log, size_t row_off, size_t n_rows, const std::vector<Col>& cols,
                                                        QM31 random_coeff, const SampleBatchesFlat& b) = 0;
  virtual Col shift_prev(Col lde_col, uint32_t trace_log) = 0;
  virtual void accumulate_col(Col dst, Col src) = 0;
  virtual void prefix_sum(Col c) = 0;
  virtual std::vector<Col> logup_generate_sel(int comp, const std::vector<Col>& main, const InteractionElements& el,
                                              const std::vector<uint8_t>& want) = 0;   // null for unwanted outputs
  virtual void eval_constraints_range(int comp, uint32_t log_size, size_t row_off, size_t n_rows, const std::vector<Col>& main_lde,
                                      const std::vector<Col>& inter_lde, const std::array<Col, 4>& prev, Col is_first_lde,
                                      const InteractionElements& el, QM31 total_sum, const std::vector<QM31>& coeffs,
                                      const std::array<Col, 4>& accum) = 0;
};

}  // namespace sbf
