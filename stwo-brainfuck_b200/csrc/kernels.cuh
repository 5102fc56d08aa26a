// Internal launch interface between the C ABI (capi.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>
#include "m31.cuh"

namespace sb {

// Kernel-launch counter of the context the calling thread entered last (ENTER in capi_internal.cuh points it at
// sc_ctx::launches): bench.py's gpu_launches is per context, not per process.
extern thread_local unsigned long long* g_launch_counter;
#define g_launch_count (*::sb::g_launch_counter)

// fft.cu
int launch_twiddle_tree(uint32_t* tw, uint32_t* itw, uint32_t R, cudaStream_t st);
int launch_interpolate(uint32_t* const* cols, uint32_t ncols, uint32_t n, const uint32_t* itw_end, cudaStream_t st);
int launch_interpolate_repeated(const uint32_t* const* src, uint32_t* const* cols, uint32_t ncols, uint32_t n, const uint32_t* itw_end,
                                cudaStream_t st);
int launch_evaluate_repeated(const uint32_t* const* coeffs, uint32_t* const* out, uint32_t ncols, uint32_t src_log, uint32_t n,
                             const uint32_t* tw_end, cudaStream_t st);
int launch_evaluate(const uint32_t* const* coeffs, uint32_t* const* out, uint32_t ncols, uint32_t src_log, uint32_t n,
                    const uint32_t* tw_end, cudaStream_t st);

// merkle.cu
constexpr uint32_t MERKLE_TOP_LOG = 9;  // layers of <= 2^9 nodes are hashed by one CTA in a single launch (merkle.cu)
constexpr uint32_t MERKLE_SUB_MAX = 9;   // a CTA of the sub-tree kernel owns up to 2^9 nodes of its first layer (16 KB + 8 KB of digests)
constexpr uint32_t MERKLE_SUB_FROM = 19; // layers of more than 2^19 nodes keep one launch each: they fill the machine on their own
int launch_commit_subtree(uint32_t L, uint32_t S, const uint32_t* prev, const uint32_t* const* cols, const uint32_t* col_off,
                          uint32_t* const* out, cudaStream_t st);
int launch_commit_top(uint32_t top_log, const uint32_t* prev, const uint32_t* const* cols, const uint32_t* col_off,
                      uint32_t* const* out, cudaStream_t st);
int launch_commit_layer(uint32_t log_size, const uint32_t* prev, const uint32_t* const* cols, uint32_t ncols,
                        uint32_t* out, cudaStream_t st, uint32_t rep_log = 0);
int launch_grind(const uint32_t digest[8], uint32_t pow_bits, unsigned long long* d_result, cudaStream_t st);

// microbench.cu (measurement only)
int launch_int_pipe_bench(int kind, uint32_t iters, int n_sm, void* d_scratch, double out[4], cudaStream_t st);

// ops.cu
int launch_bit_reverse(uint32_t* v, uint32_t log, cudaStream_t st);
int launch_batch_inverse_m31(const uint32_t* src, uint32_t* dst, size_t n, cudaStream_t st);
int launch_batch_inverse_qm31(const uint32_t* const src[4], uint32_t* const dst[4], size_t n, cudaStream_t st);
int launch_fold_line(const uint32_t* const src[4], uint32_t log, QM31 alpha, uint32_t* const dst[4], const uint32_t* itw_end,
                     cudaStream_t st);
int launch_fold_circle_into_line(const uint32_t* const src[4], uint32_t log, QM31 alpha, uint32_t* const dst[4],
                                 const uint32_t* itw_end, cudaStream_t st);
int launch_accumulate(uint32_t* const dst[4], const uint32_t* const src[4], size_t n, cudaStream_t st);
int launch_fill(uint32_t* v, size_t n, uint32_t value, cudaStream_t st);
int launch_gen_is_first(uint32_t* v, uint32_t log, cudaStream_t st);
int launch_is_first_coeffs(uint32_t* out, uint32_t log, const uint32_t* itw_plain_end, cudaStream_t st);
int launch_prefix_sum_bitrev(uint32_t* v, uint32_t log, uint32_t* scratch, cudaStream_t st);
size_t prefix_sum_tiled_words(uint32_t log);
int launch_prefix_sum_bitrev_tiled(uint32_t* const* v, uint32_t ncols, uint32_t log, uint32_t* scratch, cudaStream_t st);
int launch_prefix_sum_bitrev4(uint32_t* const v[4], uint32_t log, uint32_t* scratch, size_t words, cudaStream_t st);
struct EvalTaskHost {  // mirrors ops.cu EvalTask
  const uint32_t* coeffs;
  uint32_t log;
  uint32_t first_block;
  QM31 f[28];
};
int launch_eval_at_point_tasks(const void* d_tasks, uint32_t ntasks, uint32_t total_blocks, QM31* d_partials, QM31* d_work,
                               QM31* d_out, cudaStream_t st);
int launch_gather(const uint32_t* const* d_src, uint32_t n, uint32_t words, uint32_t* d_out, cudaStream_t st);
int launch_broadcast16(const uint32_t* src, uint32_t* dst, size_t src_len, cudaStream_t st);
// dst[c][i] = src[c][i >> rep_log] for ncols columns of src_len values (device pointer arrays), 2 <= rep_log <= 8
int launch_broadcast_cols(const uint32_t* const* src, uint32_t* const* dst, uint32_t ncols, size_t src_len, uint32_t rep_log, cudaStream_t st);

// fri.cu — FRI commit phase with the channel on the device
constexpr uint32_t FRI_TAIL_LOG = 10;   // line evaluations of <= 2^10 values are folded, hashed and mixed by one persistent CTA
struct FriTailArgs {
  uint32_t start_log, last_log, one;
  const uint32_t* layer_in[4];       // evaluation of line log start_log before its circle fold (all NULL: zeros)
  const uint32_t* itw_end;           // end of the plain inverse twiddle buffer
  uint32_t* digest;                  // channel digest, 8 words, in/out
  const uint32_t* circle_alpha;      // the first layer's folding coefficient, 4 words
  uint32_t* const* eval_out;         // [layer * 4 + k]: the committed evaluation of each layer (2^lg words per coordinate)
  uint32_t* const* tree_out;         // per layer lg: its Merkle layers of log lg, lg-1, ..., 0, concatenated over the layers
  const uint32_t* const* quot;       // [layer * 4 + k]: the quotient column of log lg + 1 that folds into layer lg, or NULL
  uint32_t* roots_out;               // 8 words per layer
  uint32_t* last_out[4];             // the last layer's evaluation (2^last_log words per coordinate)
};
int launch_fri_channel(uint32_t* d_digest, const uint32_t* d_root, uint32_t* d_alpha_out, uint32_t* d_root_copy, cudaStream_t st);
int launch_fold_line_dev(const uint32_t* const src[4], uint32_t log, const uint32_t* d_alpha, uint32_t* const dst[4], const uint32_t* itw_end,
                         cudaStream_t st);
int launch_fold_circle_dev(const uint32_t* const src[4], uint32_t log, const uint32_t* d_alpha, uint32_t* const dst[4], const uint32_t* itw_end,
                           bool first, cudaStream_t st);
int launch_fri_tail(const FriTailArgs& a, cudaStream_t st);

// tables.cu — device-side table building
struct TraceSoA { uint32_t *clk, *ip, *ci, *ni, *mp, *mv, *mvi; };   // the register rows as seven arrays
struct ColPtrs { uint32_t* p[13]; };
struct OpSteps { uint32_t* p[8]; };                                   // step indices per opcode slot: ] [ , < - . + >
struct OpCounts { uint32_t n[8]; };
struct OpTables { uint32_t* cols[8][13]; uint32_t rows[8]; };
int launch_tb_unpack(const uint32_t* d_regs, uint32_t n, bool fill_mvi, const TraceSoA& t, uint32_t* d_status, cudaStream_t st);
int launch_tb_processor(const TraceSoA& t, uint32_t m, uint32_t n, const ColPtrs& c, cudaStream_t st);
int launch_tb_program(const uint32_t* d_code, uint32_t np, uint32_t n, const ColPtrs& c, cudaStream_t st);
int launch_tb_eoe(const TraceSoA& t, uint32_t idx, const ColPtrs& c, uint32_t* d_status, cudaStream_t st);
size_t tb_opcode_scratch_words(uint32_t m);
int launch_tb_opcodes(const TraceSoA& t, uint32_t m, const OpCounts& cnt, const OpSteps& steps, const OpTables& tabs, uint32_t* d_cnt,
                      uint32_t* d_status, cudaStream_t st);
size_t rs_scratch_words(uint32_t n);
int launch_radix_sort_index(const uint32_t* keys_in, uint32_t* const kbuf[2], uint32_t* const vbuf[2], uint32_t n, uint32_t key_bits,
                            uint32_t* d_hist, uint32_t** ord_out, cudaStream_t st);
size_t scan_scratch_words(uint32_t n);
int launch_inclusive_scan(uint32_t* v, uint32_t n, uint32_t* d_sums, cudaStream_t st);
int launch_tb_memory(const TraceSoA& t, const uint32_t* d_ord, uint32_t m, uint32_t rows, uint32_t n, const ColPtrs& c, uint32_t* d_delta,
                     uint32_t* d_sums, uint32_t* d_status, cudaStream_t st);
int launch_tb_ins_keys(const uint32_t* d_ip, uint32_t np, uint32_t total, uint32_t* d_keys, cudaStream_t st);
int launch_tb_instruction(const TraceSoA& t, const uint32_t* d_code, const uint32_t* d_ord, uint32_t np, uint32_t total, uint32_t n,
                          const ColPtrs& c, cudaStream_t st);

// quotients.cu
struct QuotEntry { uint32_t col; uint32_t c[4]; };
struct QuotBatch { CM31 prx, pry, pix, piy, c0; QM31 suma, sumb, coeff; uint32_t first, count; };  // c0 = prx*piy - pry*pix
int launch_accumulate_quotients(uint32_t log, uint64_t row_off, uint64_t nrows, const uint32_t* const* d_cols,
                                const QuotBatch* d_batches, uint32_t nb, const QuotEntry* d_entries, uint32_t* const out[4],
                                cudaStream_t st, uint32_t* d_scratch = nullptr);
size_t quotients_scratch_words(uint32_t log, uint64_t row_off, uint64_t nrows);
int launch_is_first_lde(uint32_t* out, uint32_t L, uint32_t dom_log, uint64_t row_off, uint64_t nrows, const uint32_t* itw_plain_end,
                        cudaStream_t st, uint32_t* d_scratch);

}  // namespace sb
