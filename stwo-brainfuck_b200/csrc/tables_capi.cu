// C ABI of the device-side table builder (kernels in tables.cu): the 13 `trace_evaluation`s of
// crates/brainfuck_prover/src/brainfuck_air/mod.rs:511-547 from the uploaded register rows.
#include "capi_internal.cuh"
#include "host/air_ids.hpp"
#include "host/vm.hpp"

using namespace sbf;

struct sc_trace {
  sc_col* raw = nullptr;     // n x 7 words as uploaded (freed once unpacked)
  sc_col* code = nullptr;    // the compiled program
  sc_col* status = nullptr;  // 4 words; word 0 collects the consistency flags of the kernels
  uint64_t n = 0, np = 0;
  bool fill_mvi = false, built = false;
};

static void stats_to_words(const TraceStats& st, uint64_t w[16]) {
  memset(w, 0, 16 * sizeof(uint64_t));
  w[0] = st.steps; w[1] = st.memory_rows;
  for (int k = 0; k < 8; k++) w[2 + k] = st.op_count[k];
  w[10] = st.zero_ci; w[11] = st.zero_ci_index; w[12] = st.max_mp; w[13] = st.max_ip;
}
static TraceStats stats_from_words(const uint64_t w[16]) {
  TraceStats st;
  st.steps = w[0]; st.memory_rows = w[1];
  for (int k = 0; k < 8; k++) st.op_count[k] = (uint32_t)w[2 + k];
  st.zero_ci = (uint32_t)w[10]; st.zero_ci_index = w[11]; st.max_mp = (uint32_t)w[12]; st.max_ip = (uint32_t)w[13];
  return st;
}
static uint32_t bit_length(uint64_t v) { uint32_t b = 0; while (v) { b++; v >>= 1; } return b; }

extern "C" {

int32_t sc_trace_stats_host(const uint32_t* regs, uint64_t n_steps, uint64_t stats_out[16]) {
  if (!regs || !stats_out) return fail(SC_EINVAL, "null argument");
  static_assert(sizeof(Registers) == 28, "Registers is seven words");
  stats_to_words(trace_stats(reinterpret_cast<const Registers*>(regs), n_steps), stats_out);
  return SC_OK;
}

int32_t sc_trace_upload(sc_ctx* ctx, const uint32_t* regs, uint64_t n_steps, const uint32_t* program, uint64_t program_len,
                        int32_t fill_mvi, sc_trace** out) {
  ENTER_NOJOIN();
  if (!regs || !program || !out || !n_steps || !program_len) return fail(SC_EINVAL, "trace_upload: null or empty argument");
  if (n_steps >= (1ull << 28) || program_len >= (1ull << 28)) return fail(SC_EINVAL, "trace_upload: trace too long");
  sc_trace* t = new sc_trace;
  t->n = n_steps; t->np = program_len; t->fill_mvi = fill_mvi != 0;
  int32_t r = sc_col_from_host_async(ctx, regs, n_steps * 7, &t->raw);
  if (!r) r = sc_col_from_host_async(ctx, program, program_len, &t->code);
  if (r) { sc_col_free(ctx, t->raw); sc_col_free(ctx, t->code); delete t; return r; }
  *out = t;
  return SC_OK;
}

int32_t sc_trace_free(sc_ctx* ctx, sc_trace* t) {
  if (!t) return SC_OK;
  if (!ctx) return fail(SC_EINVAL, "null context");
  sc_col_free(ctx, t->raw); sc_col_free(ctx, t->code); sc_col_free(ctx, t->status);
  delete t;
  return SC_OK;
}

// Waits for everything queued and returns the consistency flags the table kernels raised: 1 clk not increasing, 2 the
// EndOfExecution row has ci != 0, 4 an opcode count differs from `stats`, 8 the Memory row count differs from `stats`.
int32_t sc_trace_status(sc_ctx* ctx, const sc_trace* t, uint32_t* flags) {
  ENTER();
  if (!t || !flags) return fail(SC_EINVAL, "null argument");
  *flags = 0;
  if (!t->status) return SC_OK;
  CK(cudaMemcpyAsync(flags, t->status->d, 4, cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return SC_OK;
}

int32_t sc_trace_build_tables(sc_ctx* ctx, sc_trace* t, const uint64_t stats_words[16], uint32_t log_max_rows, sc_col** cols_out,
                              uint32_t log_sizes_out[13]) {
  ENTER();   // joins the upload
  if (!t || !stats_words || !cols_out || !log_sizes_out) return fail(SC_EINVAL, "null argument");
  if (t->built || !t->raw) return fail(SC_EINVAL, "trace_build_tables: the trace was already consumed");
  const TraceStats st = stats_from_words(stats_words);
  if (st.steps != t->n) return fail(SC_EINVAL, "trace_build_tables: stats are for another trace");
  if (st.zero_ci != 1 || st.zero_ci_index >= t->n) return fail(SC_EINVAL, "InvalidEndOfExecution");
  // ---- sizes (the reference learns them by building the tables)
  auto p2 = [](uint64_t n) { uint64_t p = 1; while (p < n) p <<= 1; return p; };
  uint64_t rows[N_COMPONENTS];
  rows[MEMORY] = p2(st.memory_rows); rows[INSTRUCTION] = p2(t->np + t->n); rows[PROGRAM] = p2(t->np); rows[PROCESSOR] = p2(t->n);
  for (int k = 0; k < 8; k++) rows[JNZ + k] = st.op_count[k] ? p2(2ull * st.op_count[k]) / 2 : 1;
  rows[EOE] = 1;
  const uint32_t lmr = log_max_rows < 32 ? log_max_rows : 31;
  for (int c = 0; c < N_COMPONENTS; c++) {
    uint32_t lg = 0;
    while ((1ull << lg) < rows[c]) lg++;
    log_sizes_out[c] = lg + LOG_N_LANES;
    if (lg + LOG_N_LANES > lmr || lg > 27)
      return fail(SC_EINVAL, std::string("component too large: ") + COMPONENT_NAMES[c] + " (" + std::to_string(c == MEMORY ? st.memory_rows : rows[c]) +
                  (c == MEMORY ? " rows after filling the clk gaps)" : " rows)"));
  }
  ProfScope ps(ctx, "build_tables");
  const uint32_t m = (uint32_t)t->n, np = (uint32_t)t->np, total = np + m;
  // ---- outputs
  std::vector<sc_col*> made;
  auto bail = [&](int32_t r) { for (sc_col* c : made) sc_col_free(ctx, c); return r; };
  int off[N_COMPONENTS + 1];
  off[0] = 0;
  for (int c = 0; c < N_COMPONENTS; c++) off[c + 1] = off[c] + N_MAIN_COLS[c];
  for (int c = 0; c < N_COMPONENTS; c++)
    for (int j = 0; j < N_MAIN_COLS[c]; j++) {
      sc_col* col = nullptr;
      int32_t r = new_col(ctx, rows[c], &col);
      if (r) return bail(r);
      made.push_back(col);
      cols_out[off[c] + j] = col;
    }
  auto ptrs = [&](int c) { ColPtrs p{}; for (int j = 0; j < N_MAIN_COLS[c]; j++) p.p[j] = cols_out[off[c] + j]->d; return p; };
  // ---- scratch: one allocation, carved up
  uint64_t n_steps_idx = 0;
  for (int k = 0; k < 8; k++) n_steps_idx += std::max<uint32_t>(st.op_count[k], 1);
  const size_t big = std::max<size_t>(total, m);
  const size_t words = 7 * (size_t)m + 4 + n_steps_idx + tb_opcode_scratch_words(m) + 5 * big + rs_scratch_words((uint32_t)big) + m +
                       scan_scratch_words(m) + 64;
  sc_col* scratch = nullptr;
  { int32_t r = new_col(ctx, words, &scratch); if (r) return bail(r); }
  if (!t->status) { int32_t r = new_col(ctx, 4, &t->status); if (r) { sc_col_free(ctx, scratch); return bail(r); } }
  cudaError_t ce = cudaMemsetAsync(t->status->d, 0, 16, ctx->st);
  if (ce != cudaSuccess) { sc_col_free(ctx, scratch); bail(0); CK(ce); }
  uint32_t* w = scratch->d;
  auto take = [&](size_t n) { uint32_t* p = w; w += (n + 3) & ~(size_t)3; return p; };
  TraceSoA soa;
  soa.clk = take(m); soa.ip = take(m); soa.ci = take(m); soa.ni = take(m); soa.mp = take(m); soa.mv = take(m); soa.mvi = take(m);
  OpSteps steps; OpCounts cnt; OpTables tabs;
  for (int k = 0; k < 8; k++) {
    steps.p[k] = take(std::max<uint32_t>(st.op_count[k], 1));
    cnt.n[k] = st.op_count[k];
    tabs.rows[k] = (uint32_t)rows[JNZ + k];
    for (int j = 0; j < 13; j++) tabs.cols[k][j] = j < N_MAIN_COLS[JNZ + k] ? cols_out[off[JNZ + k] + j]->d : nullptr;
  }
  uint32_t* d_cnt = take(tb_opcode_scratch_words(m));
  uint32_t* kbuf[2] = {take(big), take(big)};
  uint32_t* vbuf[2] = {take(big), take(big)};
  uint32_t* keys_in = take(big);
  uint32_t* d_hist = take(rs_scratch_words((uint32_t)big));
  uint32_t* d_delta = take(m);
  uint32_t* d_sums = take(scan_scratch_words(m));
  uint32_t* d_status = t->status->d;
  cudaStream_t s = ctx->st;
  int e = 0;
  auto run = [&]() -> int {
    if ((e = launch_tb_unpack(t->raw->d, m, t->fill_mvi, soa, d_status, s))) return e;
    if ((e = launch_tb_processor(soa, m, (uint32_t)rows[PROCESSOR], ptrs(PROCESSOR), s))) return e;
    if ((e = launch_tb_program(t->code->d, np, (uint32_t)rows[PROGRAM], ptrs(PROGRAM), s))) return e;
    if ((e = launch_tb_eoe(soa, (uint32_t)st.zero_ci_index, ptrs(EOE), d_status, s))) return e;
    if ((e = launch_tb_opcodes(soa, m, cnt, steps, tabs, d_cnt, d_status, s))) return e;
    uint32_t* ord = nullptr;
    if ((e = launch_radix_sort_index(soa.mp, kbuf, vbuf, m, bit_length(st.max_mp), d_hist, &ord, s))) return e;
    if ((e = launch_tb_memory(soa, ord, m, (uint32_t)st.memory_rows, (uint32_t)rows[MEMORY], ptrs(MEMORY), d_delta, d_sums, d_status, s))) return e;
    if ((e = launch_tb_ins_keys(soa.ip, np, total, keys_in, s))) return e;
    if ((e = launch_radix_sort_index(keys_in, kbuf, vbuf, total, bit_length(std::max<uint64_t>(np ? np - 1 : 0, st.max_ip)), d_hist, &ord, s))) return e;
    if ((e = launch_tb_instruction(soa, t->code->d, ord, np, total, (uint32_t)rows[INSTRUCTION], ptrs(INSTRUCTION), s))) return e;
    return 0;
  };
  e = run();
  sc_col_free(ctx, scratch);                       // stream-ordered: the kernels above finish first
  sc_col_free(ctx, t->raw); t->raw = nullptr;      // the register rows are consumed
  sc_col_free(ctx, t->code); t->code = nullptr;
  t->built = true;
  if (e) { bail(0); CKL(e); }
  return SC_OK;
}

}  // extern "C"
