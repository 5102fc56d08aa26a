/* stwo_cuda_sharded.h — multi-GPU part of the C ABI of libstwo_cuda.so (one process per GPU, one sc_ctx per process).
 *
 * The reference is single-process (no collectives, no multi-GPU); these entry points are what the sharded
 * prover (stwo-brainfuck_b200/csrc/host/prover_sharded.hpp) needs on top of stwo_cuda.h to split one proof over N GPUs the
 * way BASELINE.json's north_star prescribes: columns for interpolation/extension, one all-to-all per commitment tree into
 * bit-reversed row ranges, then row-local hashing, constraint evaluation, quotients and FRI folds.
 * "Range" variants compute rows [row_off, row_off + n) of a domain; column handles passed to them cover exactly that range. */
#ifndef STWO_CUDA_SHARDED_H
#define STWO_CUDA_SHARDED_H
#include "stwo_cuda.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct sc_comm sc_comm; /* NCCL communicator (resolved with dlopen("libnccl.so.2")) */

/* rank 0 creates the id, the launcher broadcasts its 128 bytes (e.g. torch.distributed), every rank calls sc_comm_init */
int32_t sc_comm_unique_id(uint8_t out[128]);
int32_t sc_comm_init(sc_ctx* ctx, int32_t rank, int32_t world, const uint8_t id[128], sc_comm** out);
int32_t sc_comm_destroy(sc_ctx* ctx, sc_comm* comm);
int32_t sc_comm_rank(const sc_comm* comm);
int32_t sc_comm_world(const sc_comm* comm);
/* column-shard -> row-shard exchange: per-peer counts in words, blocks laid out in rank order in both buffers */
int32_t sc_all_to_all(sc_ctx* ctx, sc_comm* comm, const sc_col* send, const uint64_t* send_counts, sc_col* recv, const uint64_t* recv_counts);
/* Direct column->row exchange: instead of packing a send buffer and handing it to NCCL's send/recv all-to-all, every rank
 * owns a receive WINDOW (cudaMalloc + CUDA IPC, mapped by all peers) and one kernel of a sending rank writes every
 * destination's block straight into that rank's window over NVLink; a one-word all-reduce behind it is the completion
 * barrier.  sc_exchange_begin: collective, once per proof (sizes / maps the windows, one small all-reduce in steady state).
 * sc_exchange_push: arguments as sc_pack_exchange + the per-source receive counts of sc_all_to_all; *recv_out = the receive
 * buffer (a region of the window), or NULL when the direct path is unavailable on this box / the window is still too small:
 * the caller then falls back to sc_pack_exchange + sc_all_to_all (every rank takes the same branch).  SC_NO_PUSH_EXCHANGE=1
 * disables it (A/B measurements). */
int32_t sc_exchange_begin(sc_ctx* ctx, sc_comm* comm);
int32_t sc_exchange_push(sc_ctx* ctx, sc_comm* comm, sc_col* const* cols, const uint64_t* segs, const uint8_t* sharded, uint32_t n,
                         const uint64_t* recv_counts, sc_col** recv_out);
/* The same windows for the opposite direction (row ranges -> whole columns on an owner: the composition accumulators before
 * the accumulator-finalize transforms): piece j of this rank lands, whole, at word dst_offs[j] of rank dest_ranks[j]'s region.
 * Every rank passes the same region_words; *recv_out = this rank's region or NULL (fall back to sc_all_to_all). */
int32_t sc_exchange_scatter(sc_ctx* ctx, sc_comm* comm, sc_col* const* cols, const uint32_t* dest_ranks, const uint64_t* dst_offs, uint32_t n,
                            uint64_t region_words, sc_col** recv_out);
int32_t sc_all_gather(sc_ctx* ctx, sc_comm* comm, const sc_col* send, sc_col* recv, uint64_t n);   /* Merkle sub-roots */
int32_t sc_allreduce_host_u32(sc_ctx* ctx, sc_comm* comm, uint32_t* buf, uint64_t n);              /* tiny host tables */

/* the send buffer of sc_all_to_all in one launch: for every destination d, block d = the rows [d*seg_j, (d+1)*seg_j) of each
 * owned column j in order (the whole column when sharded[j] == 0) */
int32_t sc_pack_exchange(sc_ctx* ctx, sc_col* const* cols, const uint64_t* segs, const uint8_t* sharded, uint32_t n, uint32_t world, sc_col* send);
int32_t sc_col_copy(sc_ctx* ctx, sc_col* dst, uint64_t dst_off, const sc_col* src, uint64_t src_off, uint64_t n);
int32_t sc_col_view(sc_ctx* ctx, sc_col* col, uint64_t off, uint64_t n, sc_col** out);   /* non-owning slice, off % 4 == 0 */

/* FriOps on a row range: outputs [out_off, out_off + n_out) of the folded layer; src covers inputs [2*out_off, 2*(out_off+n_out)) */
int32_t sc_fold_line_range(sc_ctx* ctx, sc_col* const src[4], uint32_t log, uint64_t out_off, uint64_t n_out, const uint32_t alpha[4],
                           const sc_twiddles* tw, sc_col* dst_out[4]);
int32_t sc_fold_circle_into_line_range(sc_ctx* ctx, sc_col* const src[4], uint32_t log, uint64_t out_off, uint64_t n_out,
                                       const uint32_t alpha[4], const sc_twiddles* tw, sc_col* const dst[4]);
/* The FRI transcript as a device object, for callers that drive the commit loop layer by layer (the sharded prover): the
 * channel digest, one folding coefficient per mixed root and a copy of every root live in device memory, so a layer never
 * waits for a host round trip; the host reads the roots once (sc_dchan_finish, which also frees the object) and replays its
 * own channel.  Same arithmetic as Blake2sMerkleChannel::mix_root + Blake2sChannel::draw_felt (upstream core/vcs/
 * blake2_merkle.rs, core/channel/blake2s.rs).  The *_dc folds take coefficient #k of the object. */
typedef struct sc_dchan sc_dchan;
int32_t sc_dchan_create(sc_ctx* ctx, const uint32_t digest[8], uint32_t max_mixes, sc_dchan** out);
int32_t sc_dchan_mix_root_draw(sc_ctx* ctx, sc_dchan* dc, const sc_col* root_col);
int32_t sc_dchan_finish(sc_ctx* ctx, sc_dchan* dc, uint32_t* roots_out);
const uint32_t* sc_dchan_coeff_ptr(const sc_dchan* dc, uint32_t k);   /* device address of coefficient #k (4 words), NULL if not drawn yet */
int32_t sc_fold_line_range_dc(sc_ctx* ctx, sc_col* const src[4], uint32_t log, uint64_t out_off, uint64_t n_out, const sc_dchan* dc, uint32_t k,
                              const sc_twiddles* tw, sc_col* dst_out[4]);
int32_t sc_fold_circle_into_line_range_dc(sc_ctx* ctx, sc_col* const src[4], uint32_t log, uint64_t out_off, uint64_t n_out, const sc_dchan* dc,
                                          uint32_t k, const sc_twiddles* tw, sc_col* const dst[4]);
/* Every remaining FRI layer in one launch once the line evaluation is replicated and has <= 2^10 values (the sharded
 * driver's counterpart of sc_fri_commit's tail): for lg = start_log .. last_log + 1 fold quot_cols[4t..4t+3] (the quotient
 * column of log lg + 1; four NULLs when there is none) into the evaluation with coefficient #0, commit it (evals_out[4t..],
 * layers_out: lg + 1 layers per tree, root at index 0 of each tree's block), mix the root, draw, fold_line (FriProver::
 * commit_inner_layers, stwo core/fri.rs).  last_out: the 2^last_log values left.  dc advances by start_log - last_log mixes. */
int32_t sc_dchan_fri_tail(sc_ctx* ctx, sc_dchan* dc, const sc_twiddles* tw, sc_col* const layer_in[4], uint32_t start_log, uint32_t last_log,
                          sc_col* const* quot_cols, sc_col** evals_out, sc_col** layers_out, sc_col* last_out[4]);
/* QuotientOps on a row range (row_off, n_rows multiples of 4) */
int32_t sc_accumulate_quotients_range(sc_ctx* ctx, uint32_t log, uint64_t row_off, uint64_t n_rows, sc_col* const* cols, uint32_t n,
                                      const uint32_t random_coeff[4], const uint32_t* batch_points, const uint32_t* batch_sizes,
                                      const uint32_t* entry_cols, const uint32_t* entry_vals, uint32_t nb, sc_col* out[4]);
/* the LDE of a LogUp cumulative column read at coset offset -1, as a column (so it can be re-sharded like the others) */
int32_t sc_shift_prev(sc_ctx* ctx, const sc_col* col, uint32_t trace_log, sc_col** out);
int32_t sc_accumulate_col(sc_ctx* ctx, sc_col* dst, const sc_col* src);
/* sc_evaluate_repeated (stwo_cuda.h) restricted to rows [row_off[i], row_off[i]+row_cnt[i]) of column i, both multiples of
 * 2^log_repeat.  The main-trace values are replicated on every rank and their transforms cost 1/16 of a full column, so a
 * rank computes all of them and keeps only its row range: the main tree needs no LDE exchange. */
int32_t sc_evaluate_repeated_range(sc_ctx* ctx, sc_col* const* coeffs, uint32_t n, uint32_t log_repeat, uint32_t log_blowup,
                                   const sc_twiddles* tw, const uint64_t* row_off, const uint64_t* row_cnt, sc_col** out);

/* LogUp generation materialising only the wanted coordinate columns (no prefix sum; use sc_prefix_sum_bitrev on the owner) */
int32_t sc_logup_generate_sel(sc_ctx* ctx, int32_t component, sc_col* const* main_cols, uint32_t n_main, uint32_t log_repeat,
                              const uint32_t* elements, const uint8_t* want, sc_col** out);
int32_t sc_eval_constraints_range(sc_ctx* ctx, int32_t component, uint32_t log_size, uint64_t row_off, uint64_t n_rows,
                                  sc_col* const* main_lde, uint32_t n_main, sc_col* const* inter_lde, uint32_t n_inter,
                                  sc_col* const prev[4], const sc_col* is_first_lde, const uint32_t* elements,
                                  const uint32_t total_sum[4], const uint32_t* coeffs, sc_col* const accum[4]);

#ifdef __cplusplus
}
#endif
#endif
