/* stwo_cuda.h — C ABI of libstwo_cuda.so, the B200 (sm_100a) backend for stwo-brainfuck's proving hot path.
 *
 * Every entry point replaces one method of Stwo's `Backend` trait family as the reference instantiates it at
 * `SimdBackend` (stwo-prover 0.1.1 @ 31e8dbc, an un-vendored git dependency: /root/reference/Cargo.toml:41).  The
 * reference has no FFI of its own (it selects the backend by type parameter, crates/brainfuck_prover/src/
 * brainfuck_air/mod.rs:486-487,732), so each declaration cites the reference call site that reaches the trait method
 * and the upstream trait it stands in for.  The Rust binding a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *  - all functions return 0 on success or a negative sc_status; sc_last_error() gives a thread-local message;
 *  - no exceptions/panics cross the ABI; a sticky CUDA error makes the context unusable (SC_ECUDA from then on);
 *  - handles are owned by the caller and freed exactly once; the library never keeps host pointers past a call;
 *  - a column (`sc_col`) is a device buffer of 32-bit words: M31 values in [0, 2^31-1), or 8 words per Blake2s digest;
 *  - circle evaluations are in bit-reversed circle-domain order on CanonicCoset(log).circle_domain();
 *  - QM31 values cross the ABI as 4 words (a + bi) + (c + di)u -> {a,b,c,d}; secure columns are 4 coordinate columns;
 *  - calls on one context are serialised on its stream and are asynchronous unless they return host data.
 */
#ifndef STWO_CUDA_H
#define STWO_CUDA_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { SC_OK = 0, SC_EINVAL = -1, SC_ECUDA = -2, SC_ENOMEM = -3, SC_EPROOF = -4, SC_EVERIFY = -5 } sc_status;

typedef struct sc_ctx sc_ctx;             /* device + stream + scratch */
typedef struct sc_col sc_col;             /* Backend::Column (BaseColumn / Col<B, Blake2sHash>) */
typedef struct sc_twiddles sc_twiddles;   /* TwiddleTree<B> */

const char* sc_last_error(void);
int32_t sc_version(void);

/* `stream` is a cudaStream_t to launch on (e.g. torch.cuda.current_stream().cuda_stream) or NULL for a private one. */
int32_t sc_ctx_create(int32_t device, void* stream, sc_ctx** out);
int32_t sc_ctx_destroy(sc_ctx* ctx);
int32_t sc_ctx_sync(sc_ctx* ctx);
/* Proof arena: between begin and end every new column is a slice of one persistent device slab (bump allocation; sc_col_free
 * of such a column only drops the handle).  The caller promises that every column made inside the bracket is dead before the
 * next sc_ctx_arena_begin.  The slab is sized from the previous bracket's total and reallocated, if it has to grow, in begin;
 * what does not fit falls back to the stream-ordered pool.  sbf_prove brackets every proof with it. */
int32_t sc_ctx_arena_begin(sc_ctx* ctx);
int32_t sc_ctx_arena_end(sc_ctx* ctx);
/* The compute stream waits for the asynchronous uploads issued so far (every other entry point does this first, too). */
int32_t sc_ctx_join_uploads(sc_ctx* ctx);
/* Number of kernels launched through this context so far (bench.py's gpu_launches). */
uint64_t sc_ctx_launch_count(const sc_ctx* ctx);
/* Per-kernel-class device timing: when enabled every entry point brackets its launches with CUDA events on the launch
 * stream; the report is "tag:milliseconds:count;..." (sum per tag since the last report) and clears the records. */
int32_t sc_ctx_profile(sc_ctx* ctx, int32_t enable);
int32_t sc_ctx_profiling(const sc_ctx* ctx);   /* 1 while enabled (sbf_prove then also waits at its stage boundaries, so that stages_ms are device-complete) */
/* Device-side stopwatch on the compute stream: sc_event_elapsed waits for `b` and returns the milliseconds between the marks. */
typedef struct sc_event sc_event;
int32_t sc_event_record(sc_ctx* ctx, sc_event** out);
int32_t sc_event_elapsed(sc_ctx* ctx, const sc_event* a, const sc_event* b, float* ms);
int32_t sc_event_free(sc_ctx* ctx, sc_event* e);
/* sc_ctx_mark: a token; sc_ctx_release_since frees every column created on the context after the token that is still alive
 * (a Rust caller unwinding from a panic inside `prove` has no other way to find them). */
uint64_t sc_ctx_mark(sc_ctx* ctx);
int32_t sc_ctx_release_since(sc_ctx* ctx, uint64_t mark);
uint64_t sc_ctx_live_columns(sc_ctx* ctx);   /* number of column handles currently alive on the context */
/* Ties a caller-owned object to the context: `dtor(ctx, p)` runs inside sc_ctx_destroy while the context can still free
 * columns (and when the slot is overwritten).  slot < 4; slot 0 is used by sbf_prove for its preprocessed-tree cache. */
typedef void (*sc_attach_dtor)(sc_ctx* ctx, void* p);
int32_t sc_ctx_attach(sc_ctx* ctx, uint32_t slot, void* p, sc_attach_dtor dtor);
void* sc_ctx_attached(sc_ctx* ctx, uint32_t slot);
size_t sc_ctx_profile_report(sc_ctx* ctx, char* buf, size_t cap);
size_t sc_ctx_profile_timeline(sc_ctx* ctx, char* buf, size_t cap);  /* "tag:start_ms:dur_ms;" per scope, not cleared */

/* ---- Column<T> (upstream core/backend/mod.rs `Column`: zeros, uninitialized, to_cpu, len, at, set, clone) ---- */
int32_t sc_col_zeros(sc_ctx* ctx, uint64_t len, sc_col** out);
int32_t sc_col_uninit(sc_ctx* ctx, uint64_t len, sc_col** out);
int32_t sc_col_from_host(sc_ctx* ctx, const uint32_t* host, uint64_t len, sc_col** out);   /* FromIterator */
/* Upload without waiting: `host` must stay valid until the next synchronising call; pinned memory (sc_host_arena_alloc)
 * makes it a true asynchronous DMA.  The reference fills its columns in ordinary host memory (e.g. processor/table.rs:
 * 86-100); filling them in this arena instead removes the staged pageable copy from the upload.  The copy runs on a
 * second stream owned by the context, beside kernels already queued; every other entry point first makes the compute
 * stream wait for the uploads issued so far. */
int32_t sc_col_from_host_async(sc_ctx* ctx, const uint32_t* host, uint64_t len, sc_col** out);
int32_t sc_host_arena_alloc(sc_ctx* ctx, uint64_t bytes, void** out);   /* thread-safe bump allocation, 64-byte aligned */
int32_t sc_host_arena_reset(sc_ctx* ctx);                               /* releases every allocation, keeps the blocks */
int32_t sc_col_to_host(sc_ctx* ctx, const sc_col* col, uint32_t* host);                    /* to_cpu */
int32_t sc_col_read(sc_ctx* ctx, const sc_col* col, uint64_t offset, uint64_t n, uint32_t* host);   /* at */
int32_t sc_col_write(sc_ctx* ctx, sc_col* col, uint64_t offset, uint64_t n, const uint32_t* host);  /* set */
int32_t sc_col_clone(sc_ctx* ctx, const sc_col* col, sc_col** out);
int32_t sc_col_free(sc_ctx* ctx, sc_col* col);
uint64_t sc_col_len(const sc_col* col);
void* sc_col_device_ptr(sc_col* col);
/* Non-owning column over caller-owned device memory (a torch tensor, an NCCL receive buffer); sc_col_free only drops
 * the handle.  Used by the multi-GPU path to hash row ranges received by all-to-all without another copy. */
int32_t sc_col_wrap(sc_ctx* ctx, void* device_ptr, uint64_t len, sc_col** out);
/* Expands `src` (len L) to len 16*L with every value repeated 16x — the reference's PackedBaseField broadcast
 * (crates/brainfuck_prover/src/components/processor/table.rs:86-100), so only 1/16 of a trace column crosses PCIe. */
int32_t sc_col_broadcast16(sc_ctx* ctx, const sc_col* src, sc_col** out);

/* ---- ColumnOps::bit_reverse_column (upstream core/backend/simd/bit_reverse.rs) ---- */
int32_t sc_bit_reverse(sc_ctx* ctx, sc_col* col);

/* ---- FieldOps::batch_inverse for M31 and QM31 (LogupColGenerator::finalize_col, reached from e.g.
 *      crates/brainfuck_prover/src/components/memory/table.rs:513) ---- */
int32_t sc_batch_inverse_m31(sc_ctx* ctx, const sc_col* src, sc_col* dst);
int32_t sc_batch_inverse_qm31(sc_ctx* ctx, sc_col* const src[4], sc_col* const dst[4]);

/* ---- PolyOps (upstream core/poly/circle/ops.rs) ---- */
/* precompute_twiddles(half_odds(root_log)) — crates/brainfuck_prover/src/brainfuck_air/mod.rs:480-484 (root_log 26). */
int32_t sc_precompute_twiddles(sc_ctx* ctx, uint32_t root_log, sc_twiddles** out);
int32_t sc_twiddles_free(sc_ctx* ctx, sc_twiddles* tw);
/* The same tree, computed once per context and returned as a borrowed handle (never pass it to sc_twiddles_free). */
int32_t sc_twiddles_cached(sc_ctx* ctx, uint32_t root_log, const sc_twiddles** out);
int32_t sc_twiddles_to_host(sc_ctx* ctx, const sc_twiddles* tw, uint32_t* twiddles, uint32_t* itwiddles);
/* interpolate_columns: in place, evaluations -> coefficients; columns may have mixed sizes (log >= 3).
 * brainfuck_air/mod.rs:497,550-562,690-702 (tree_builder.extend_evals). */
int32_t sc_interpolate(sc_ctx* ctx, sc_col* const* cols, uint32_t n, const sc_twiddles* tw);
/* evaluate_polynomials: coefficients -> evaluations on CanonicCoset(log + log_blowup); allocates out[i].
 * brainfuck_air/mod.rs:500,583,723 (tree_builder.commit). */
int32_t sc_evaluate(sc_ctx* ctx, sc_col* const* coeffs, uint32_t n, uint32_t log_blowup, const sc_twiddles* tw, sc_col** out);
/* eval_at_point for n (polynomial, point) pairs; points: n x 8 words {x[4], y[4]}; out: n x 4 words.
 * Reached from prove_values inside prover::prove (brainfuck_air/mod.rs:732). */
int32_t sc_eval_at_point(sc_ctx* ctx, sc_col* const* polys, uint32_t n, const uint32_t* points, uint32_t* out);

/* ---- MerkleOps<Blake2sMerkleHasher>::commit_on_layer (upstream core/vcs/ops.rs, core/backend/simd/blake2s.rs) ----
 * prev may be NULL; out receives a new 8*2^log_size-word column.  brainfuck_air/mod.rs:500,583,723 and every FRI layer. */
int32_t sc_merkle_commit_layer(sc_ctx* ctx, uint32_t log_size, const sc_col* prev, sc_col* const* cols, uint32_t n, sc_col** out);
/* MerkleProver::commit over mixed-size columns (stable by size, descending): layers_out[k] = layer of log size k,
 * k = 0..max_log (caller provides max_log+1 slots); root (8 words) copied to root_out if not NULL. */
int32_t sc_merkle_commit(sc_ctx* ctx, sc_col* const* cols, uint32_t n, sc_col** layers_out, uint32_t* max_log_out, uint32_t root_out[8]);

/* ---- FriOps (upstream core/backend/simd/fri.rs) — inside prover::prove, brainfuck_air/mod.rs:732 ---- */
/* src: 4 coordinate columns of a LineEvaluation on half_odds(log); dst: 4 new columns of length 2^(log-1). */
int32_t sc_fold_line(sc_ctx* ctx, sc_col* const src[4], uint32_t log, const uint32_t alpha[4], const sc_twiddles* tw, sc_col* dst_out[4]);
/* dst (length 2^(log-1), 4 coords) <- dst*alpha^2 + fold(src on CanonicCoset(log)). */
int32_t sc_fold_circle_into_line(sc_ctx* ctx, sc_col* const src[4], uint32_t log, const uint32_t alpha[4], const sc_twiddles* tw, sc_col* const dst[4]);

/* FriProver::commit (upstream core/fri.rs) as ONE call: first-layer tree over all quotient columns, then per line layer
 * fold_circle_into_line of the columns that join it, Merkle tree, channel.mix_root + channel.draw_felt, fold_line — with the
 * Blake2s channel state in device memory, so no layer waits for a host round trip (upstream reads every root back before it
 * can draw the next folding coefficient).  The caller replays its own channel from roots_out afterwards: mix_root(roots[0]),
 * draw_felt() (the circle fold's coefficient), then mix_root / draw_felt per inner layer — the digests and coefficients are
 * the same, so the transcript and the proof are unchanged.
 * quot_cols: nq secure columns (4 coordinate columns each) of strictly descending log sizes quot_logs; channel_digest: the
 * channel's digest before the phase; last_log = log_last_layer_degree_bound + log_blowup.  With top = quot_logs[0] and
 * n_inner = top - 1 - last_log:  first_layers_out: top + 1 slots (layer k = log size k);  inner_evals_out: 4 * n_inner
 * columns (the committed evaluation of line log top-1, top-2, ...);  inner_layers_out: for each inner layer of line log lg,
 * lg + 1 slots (layer k at slot k), concatenated in layer order;  roots_out: 8 * (1 + n_inner) words;  last_values_out:
 * 4 * 2^last_log words, coordinate-major. */
int32_t sc_fri_commit(sc_ctx* ctx, const sc_twiddles* tw, sc_col* const* quot_cols, const uint32_t* quot_logs, uint32_t nq,
                      const uint32_t channel_digest[8], uint32_t last_log, sc_col** first_layers_out, sc_col** inner_evals_out,
                      sc_col** inner_layers_out, uint32_t* roots_out, uint32_t* last_values_out);

/* ---- QuotientOps::accumulate_quotients (upstream core/backend/simd/quotients.rs, core/pcs/quotients.rs) ----
 * cols: the n columns of one LDE size 2^log; batches: nb sample batches; batch b has point batch_points[8b..8b+8) and
 * batch_sizes[b] entries; entries (column index, value[4]) are concatenated in entry_cols / entry_vals.
 * out: 4 new coordinate columns. */
int32_t sc_accumulate_quotients(sc_ctx* ctx, uint32_t log, sc_col* const* cols, uint32_t n, const uint32_t random_coeff[4],
                                const uint32_t* batch_points, const uint32_t* batch_sizes, const uint32_t* entry_cols,
                                const uint32_t* entry_vals, uint32_t nb, sc_col* out[4]);

/* ---- AccumulationOps (upstream core/backend/simd/accumulation.rs) ---- */
int32_t sc_accumulate(sc_ctx* ctx, sc_col* const dst[4], sc_col* const src[4]);
int32_t sc_secure_powers(const uint32_t felt[4], uint32_t n, uint32_t* out /* n x 4 */);

/* ---- GrindOps<Blake2sChannel>::grind: smallest nonce with >= pow_bits trailing zeros (upstream simd/grind.rs) ---- */
int32_t sc_grind(sc_ctx* ctx, const uint32_t digest[8], uint32_t pow_bits, uint64_t* nonce_out);

/* ---- Device-side table building: the 13 `trace_evaluation`s of crates/brainfuck_prover/src/brainfuck_air/
 * mod.rs:511-547 (memory/table.rs:85-117,249-318; instruction/table.rs:85-110,250-281; program/table.rs:36-46; processor/
 * table.rs:117-142,195-207; processor/instructions/table.rs:293-328; .../jump/table.rs:264-297; .../end_of_execution/table.rs:
 * 71-77) from the VM's register rows, in the lane-compact form the *_repeated entry points take (one word per table row).
 * regs: n_steps x 7 words {clk, ip, ci, ni, mp, mv, mvi} (crates/brainfuck_vm/src/registers.rs:5-21), in clk order; program:
 * the compiled program words.  The upload does not wait (host memory must stay valid until the next synchronising call;
 * pinned memory from sc_host_arena_alloc makes it a DMA beside queued kernels).  fill_mvi != 0: mvi is computed on the
 * device and the input's seventh word is ignored.
 * stats (16 words): [0] steps, [1] Memory rows before padding, [2..9] steps per opcode ] [ , < - . + > (with a successor),
 * [10] rows with ci = 0, [11] index of the first, [12] max mp, [13] max ip — kept by the VM while it runs, or computed by
 * sc_trace_stats_host.  They fix every table size, so no device round trip is needed before the columns are allocated;
 * the kernels cross-check them (sc_trace_status).
 * sc_trace_build_tables: cols_out receives the 128 main-trace columns in (component, column) order, log_sizes_out the 13
 * column log sizes (rows x 16 lanes).  SC_EINVAL "component too large: <name>" when a table exceeds 2^(log_max_rows-4) rows. ---- */
typedef struct sc_trace sc_trace;
int32_t sc_trace_stats_host(const uint32_t* regs, uint64_t n_steps, uint64_t stats_out[16]);
int32_t sc_trace_upload(sc_ctx* ctx, const uint32_t* regs, uint64_t n_steps, const uint32_t* program, uint64_t program_len,
                        int32_t fill_mvi, sc_trace** out);
int32_t sc_trace_build_tables(sc_ctx* ctx, sc_trace* trace, const uint64_t stats[16], uint32_t log_max_rows, sc_col** cols_out,
                              uint32_t log_sizes_out[13]);
int32_t sc_trace_status(sc_ctx* ctx, const sc_trace* trace, uint32_t* flags);
int32_t sc_trace_free(sc_ctx* ctx, sc_trace* trace);

/* ---- measurement only: integer-pipe micro-benchmark, the measured peak of the Blake2s roofline ("INT32 ALU
 * pipe ... measure").  kind 0 LOP3, 1 SHF, 2 PRMT, 3 IADD3, 4 IMAD, 5 LOP3+IMAD interleaved, 6 the Blake2s G mix.
 * out = {ALU-pipe lane-ops/clk/SM, FMA-pipe lane-ops/clk/SM, kernel ms, #SMs}. ---- */
int32_t sc_microbench_int(sc_ctx* ctx, int32_t kind, uint32_t iters, double out[4]);

/* ---- constraint_framework pieces the reference uses concretely at SimdBackend ---- */
/* gen_is_first::<B>(log_size) — brainfuck_air/mod.rs:497. */
int32_t sc_gen_is_first(sc_ctx* ctx, uint32_t log_size, sc_col** out);
/* coefficients of that column's polynomial in closed form (= sc_gen_is_first + sc_interpolate, without the transform) */
int32_t sc_is_first_coeffs(sc_ctx* ctx, uint32_t log_size, const sc_twiddles* tw, sc_col** out);
/* The low-degree extension of that column — rows [row_off, row_off + n_rows) (multiples of 4) of its evaluation on
 * CanonicCoset(log_size + log_blowup).circle_domain(), bit-reversed — written in closed form: the polynomial is the rank-one
 * product 2^-L (1 + t_0 y)(1 + t_1 x) prod_l (1 + t_l pi^(l-1)(x)), so the extension of the preprocessed trace
 * (brainfuck_air/mod.rs:493-500: gen_is_first -> interpolate -> commit's evaluate) needs no transform and, row range by row
 * range, no exchange between GPUs.  Same values as sc_evaluate(sc_is_first_coeffs(..)). */
int32_t sc_is_first_lde(sc_ctx* ctx, uint32_t log_size, uint32_t log_blowup, const sc_twiddles* tw, uint64_t row_off, uint64_t n_rows, sc_col** out);
/* simd/prefix_sum.rs inclusive_prefix_sum: in-place inclusive prefix sum in trace-coset order of a bit-reversed
 * column (LogupTraceGenerator::finalize_last, e.g. crates/brainfuck_prover/src/components/processor/table.rs:530). */
int32_t sc_prefix_sum_bitrev(sc_ctx* ctx, sc_col* col);

/* ---- batched `Column::at` for decommitment (MerkleProver::decommit / FriProver::decommit read single elements):
 * out_host[i*words .. +words) = cols[i][offsets[i] .. +words). ---- */
int32_t sc_gather(sc_ctx* ctx, sc_col* const* cols, const uint64_t* offsets, uint32_t n, uint32_t words, uint32_t* out_host);

/* ---- Lane-repeated columns.  The reference writes every table row into all 16 SIMD lanes of a trace column
 * (components/processor/table.rs:86-100 and the six other table.rs files: `data[vec_row] = value.into()`), so a main-trace
 * evaluation of log size m+4 holds 2^m distinct values, each filling 16 consecutive (bit-reversed-order) rows.  Such a
 * column's polynomial has one non-zero coefficient in 16 and the first four FFT layers only scale / replicate, so
 * PolyOps::interpolate_columns / evaluate_polynomials / eval_at_point and MerkleOps::commit_on_layer can work on the 2^m
 * values.  A column passed here stores the distinct values (or compact coefficients: coefficient j = coefficient j<<r of the
 * full vector).  sc_evaluate_repeated returns ordinary full-length columns, bit-identical to sc_evaluate of the expanded
 * input.  The *_repeated Merkle calls take ordinary full columns and rely on the caller's promise that they repeat. ---- */
/* out == NULL: in place.  Otherwise the inputs are left untouched and out[i] receives a new column with the coefficients. */
int32_t sc_interpolate_repeated(sc_ctx* ctx, sc_col* const* cols, uint32_t n, uint32_t log_repeat, const sc_twiddles* tw,
                                sc_col** out);
int32_t sc_evaluate_repeated(sc_ctx* ctx, sc_col* const* coeffs, uint32_t n, uint32_t log_repeat, uint32_t log_blowup,
                             const sc_twiddles* tw, sc_col** out);
int32_t sc_eval_at_point_repeated(sc_ctx* ctx, sc_col* const* polys, const uint32_t* log_repeats, uint32_t n,
                                  const uint32_t* points, uint32_t* out);
int32_t sc_merkle_commit_layer_repeated(sc_ctx* ctx, uint32_t log_size, const sc_col* prev, sc_col* const* cols, uint32_t n,
                                        uint32_t log_repeat, sc_col** out);
int32_t sc_merkle_commit_repeated(sc_ctx* ctx, sc_col* const* cols, uint32_t n, uint32_t log_repeat, sc_col** layers_out,
                                  uint32_t* max_log_out, uint32_t root_out[8]);

/* ---- LogupTraceGenerator for one component: write_frac / finalize_col per relation entry, finalize_last.
 * Stands in for the reference's interaction_trace_evaluation (e.g. crates/brainfuck_prover/src/components/processor/
 * table.rs:456-533; memory/table.rs:485-518).  component: 0 memory, 1 instruction, 2 program, 3 processor, 4 `]`, 5 `[`,
 * 6 `,`, 7 `<`, 8 `-`, 9 `.`, 10 `+`, 11 `>`, 12 end_of_execution (BrainfuckClaim order, brainfuck_air/mod.rs:79-93).
 * main_cols: the component's main-trace evaluations, one value per 2^log_repeat rows (4 = the lane-compact form, 0 =
 * full columns); elements: 3 x {z[4], alpha_powers[7][4]} for the memory,
 * instruction and processor relations (brainfuck_air/mod.rs:149-165).  out: 4 x (#LogUp columns) new columns.
 * claimed_sum may be NULL: it is element 1 of each of the last four columns and can be fetched later (sc_gather). ---- */
int32_t sc_logup_generate(sc_ctx* ctx, int32_t component, sc_col* const* main_cols, uint32_t n_main, uint32_t log_repeat,
                          const uint32_t* elements, sc_col** out, uint32_t claimed_sum[4]);
/* ---- ComponentProver::evaluate_constraint_quotients_on_domain for one component (upstream constraint_framework/
 * component.rs + simd_domain.rs; the `evaluate()` bodies are the reference's components/<name>/component.rs).
 * Columns are the LDEs on CanonicCoset(log_size+1); coeffs: n_constraints x 4 words, coeffs[k] multiplies constraint k;
 * accum (4 coordinate columns of the same length) += sum_k coeffs[k]*C_k / vanishing. ---- */
int32_t sc_eval_constraints(sc_ctx* ctx, int32_t component, uint32_t log_size, sc_col* const* main_lde, uint32_t n_main,
                            sc_col* const* inter_lde, uint32_t n_inter, const sc_col* is_first_lde, const uint32_t* elements,
                            const uint32_t total_sum[4], const uint32_t* coeffs, sc_col* const accum[4]);

#ifdef __cplusplus
}
#endif
#endif /* STWO_CUDA_H */
