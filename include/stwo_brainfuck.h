/* stwo_brainfuck.h — prove / verify entry points of libstwo_cuda.so: the host orchestrator that stands in for
 * `brainfuck_prover prove|verify` (crates/brainfuck_prover/src/bin/brainfuck_prover.rs:79-152) and for
 * prove_brainfuck / verify_brainfuck (crates/brainfuck_prover/src/brainfuck_air/mod.rs:471-797).
 * The VM runs on the host; the 13 tables are built on the device from the uploaded register rows (sc_trace_*, stwo_cuda.h);
 * every Backend operation goes through the C ABI of stwo_cuda.h. */
#ifndef STWO_BRAINFUCK_H
#define STWO_BRAINFUCK_H
#include "stwo_cuda.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct sbf_proof sbf_proof; /* BrainfuckProof { claim, interaction_claim, proof } */
const char* sbf_last_error(void);
/* prove --code <code> with stdin bytes; log_max_rows = LOG_MAX_ROWS (24; 20 under cfg(test)), brainfuck_air/mod.rs:427-433 */
#define SBF_NO_OVERLAP 1u /* flags: build the host tables before any device work instead of overlapping them with phase 0 */
#define SBF_NO_TWIDDLE_CACHE 2u /* flags: recompute the twiddle tree in every proof (the reference does, mod.rs:480-484) */
/* flags: keep the preprocessed tree (the IsFirst columns of log size log_max_rows..4, their LDEs and Merkle layers — the
 * same for every program, brainfuck_air/mod.rs:453-464,493-500) on the context between proofs.  The reference rebuilds it
 * in every proof; the proof bytes are the same either way.  sbf_preprocessed_cache_clear drops it (sc_ctx_destroy does too). */
#define SBF_CACHE_PREPROCESSED 8u
/* flags: build the 13 tables with the host builders (csrc/host/tables.hpp) and upload the finished columns, as round 1 did,
 * instead of building them on the device from the register rows; same proof bytes (A/B measurements, parity tests). */
#define SBF_HOST_TABLES 16u
/* flags: run the FRI commit phase layer by layer with the host channel (one root read-back per layer) instead of
 * sc_fri_commit's device-side channel; same proof bytes (A/B measurements). */
#define SBF_NO_FUSED_FRI 32u
/* flags: allocate the proof's columns from the stream-ordered pool one by one instead of the per-proof arena (sc_ctx_arena_*) */
#define SBF_NO_ARENA 64u
int32_t sbf_preprocessed_cache_clear(sc_ctx* ctx);
int32_t sbf_prove(sc_ctx* ctx, const char* code, const uint8_t* input, size_t input_len, uint32_t log_max_rows, uint32_t flags,
                  sbf_proof** out);
/* verify_brainfuck: host only; 0 or SC_EVERIFY */
/* The same proof split over the ranks of `comm` (include/stwo_cuda_sharded.h: one process per GPU, every rank calls this with
 * the same program and ends with the same proof bytes).  comm == NULL runs the sharded driver on one GPU. */
#define SBF_SHARDED_DRIVER 4u /* flags for sbf_prove: take the sharded driver with world size 1 */
struct sc_comm;
int32_t sbf_prove_sharded(sc_ctx* ctx, struct sc_comm* comm, const char* code, const uint8_t* input, size_t input_len,
                          uint32_t log_max_rows, uint32_t flags, sbf_proof** out);
int32_t sbf_verify(const sbf_proof* proof);   /* with the LOG_MAX_ROWS the proof object was made with */
char* sbf_proof_json(const sbf_proof* proof);    /* serde-shaped JSON of the proof; free with sbf_string_free */
char* sbf_proof_report(const sbf_proof* proof);  /* steps, log sizes, per-stage milliseconds */
size_t sbf_proof_output(const sbf_proof* proof, uint8_t* buf, size_t cap); /* program stdout */
void sbf_string_free(char* s);
void sbf_proof_free(sbf_proof* proof);
/* `brainfuck_prover verify <file>`: the serde JSON text of a proof (what sbf_proof_json / the reference's `prove --output`
 * write) checked with the verifier's own LOG_MAX_ROWS and the default PcsConfig; 0 or SC_EVERIFY.  Pure host code. */
int32_t sbf_verify_json(const char* json, uint32_t log_max_rows);
int32_t sbf_proof_from_json(const char* json, uint32_t log_max_rows, sbf_proof** out);
#ifdef __cplusplus
}
#endif
#endif
