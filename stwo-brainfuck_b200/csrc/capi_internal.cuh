// Internals shared by the translation units that implement the C ABI (capi.cu, sharded.cu): handle layouts, error
// plumbing, the small-table staging ring and the profiling scope.
#pragma once
#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/stwo_cuda.h"
#include "kernels.cuh"

using namespace sb;

struct sc_col {
  uint32_t* d;
  uint64_t len;
  bool owned = true;  // false: a view over caller-owned device memory (sc_col_wrap) or a slice of the proof arena
  bool slab = false;  // lives in the context's upload slab (sc_col_from_host_async)
  uint64_t id = 0;    // creation order within the context (sc_ctx_mark / sc_ctx_release_since)
};
struct sc_twiddles {
  uint32_t root_log;
  uint32_t* tw;   // 2^root_log words
  uint32_t* itw;  // 2^root_log words
};
struct sc_ctx {
  int device;
  cudaStream_t st;
  bool own_stream;
  bool poisoned;
  unsigned long long launches = 0;  // kernels launched through this context (sc_ctx_launch_count)
  // staging ring for small host->device tables (pointer arrays, task tables)
  uint8_t* h_ring;
  uint8_t* d_ring;
  size_t ring_size, ring_off;
  // optional per-kernel-class timing (CUDA events on the launch stream), see sc_ctx_profile
  bool profiling = false;
  struct ProfRec { const char* tag; cudaEvent_t a, b; };
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> ev_pool;
  // pinned host arena (sc_host_arena_*): blocks are kept for the life of the context and reused after a reset
  struct ArenaBlock { uint8_t* p; size_t size, used; };
  std::vector<ArenaBlock> arena;
  std::mutex arena_mu;
  std::map<uint32_t, sc_twiddles*> tw_cache;  // sc_twiddles_cached
  // every live column handle by creation id: lets a caller that failed half-way (an exception inside the prover) release
  // what it created since a mark instead of leaking gigabytes of device memory
  std::map<uint64_t, sc_col*> live;
  uint64_t next_id = 1;
  // sc_col_from_host_async copies on a second stream so that uploads overlap kernels already queued on `st`; the next
  // call of any other entry point makes `st` wait for them (join_uploads)
  cudaStream_t copy_st = nullptr;
  cudaEvent_t copy_ev = nullptr, slab_ev = nullptr;
  bool uploads_pending = false;
  // Upload targets come from a persistent slab, not from the stream-ordered pool: a pool allocation on the copy stream
  // either inherits a dependency on the compute stream or maps fresh memory, and both stall the kernels it should overlap.
  // Bump allocation; the slab rewinds when its last column is freed (the copy stream then waits for the compute stream's
  // position at that moment before it overwrites anything).
  uint8_t* slab = nullptr;
  size_t slab_cap = 0, slab_used = 0, slab_high = 0;
  int slab_live = 0;
  bool slab_fence = false;
  // Proof arena (sc_ctx_arena_begin/end): between the two calls new columns are bump-allocated from one persistent slab and
  // freeing them is a host-side no-op — a proof makes ~1200 columns, and at eight GPUs its kernels are short enough that the
  // 2400 cudaMallocAsync / cudaFreeAsync calls were a visible part of the (host-bound) critical path.  Nothing is reused inside
  // a proof, so there is no ordering hazard; the slab is sized from the previous proof's total and grows between proofs.
  uint8_t* parena = nullptr;
  size_t parena_cap = 0, parena_off = 0, parena_need = 0, parena_want = 0;
  bool parena_on = false;
  // sc_ctx_attach: caller-owned objects whose life ends with the context (e.g. the prover's preprocessed-tree cache)
  struct Attached { void* p = nullptr; void (*dtor)(sc_ctx*, void*) = nullptr; } attached[4];
};

// RAII: brackets the kernels launched in a scope with two events when profiling is on.
struct ProfScope {
  sc_ctx* c; cudaEvent_t a = nullptr, b = nullptr; const char* tag;
  static cudaEvent_t get(sc_ctx* c) {
    if (!c->ev_pool.empty()) { cudaEvent_t e = c->ev_pool.back(); c->ev_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
  ProfScope(sc_ctx* ctx, const char* t) : c(ctx), tag(t) {
    if (c && c->profiling) { a = get(c); b = get(c); cudaEventRecord(a, c->st); }
  }
  ~ProfScope() { if (a) { cudaEventRecord(b, c->st); c->prof.push_back({tag, a, b}); } }
};

extern thread_local std::string g_sc_err;
static inline int32_t fail(int32_t code, const std::string& m) { g_sc_err = m; return code; }
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      if (ctx) ctx->poisoned = true;                                                               \
      return fail(SC_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                   \
    }                                                                                              \
  } while (0)
#define CKL(expr)                                                                                  \
  do {                                                                                             \
    int e_ = (expr);                                                                               \
    if (e_ > 0) { ctx->poisoned = true; return fail(SC_ECUDA, std::string(#expr) + ": " + cudaGetErrorString((cudaError_t)e_)); } \
    if (e_ < 0) return fail(SC_EINVAL, std::string(#expr) + ": invalid argument");                 \
  } while (0)
#define ENTER_NOJOIN()                                                                             \
  if (!ctx) return fail(SC_EINVAL, "null context");                                                \
  if (ctx->poisoned) return fail(SC_ECUDA, "context unusable after an earlier CUDA error");        \
  ::sb::g_launch_counter = &ctx->launches;                                                         \
  CK(cudaSetDevice(ctx->device))
#define ENTER()                                                                                    \
  ENTER_NOJOIN();                                                                                  \
  if (ctx->uploads_pending) {                                                                      \
    CK(cudaEventRecord(ctx->copy_ev, ctx->copy_st));                                               \
    CK(cudaStreamWaitEvent(ctx->st, ctx->copy_ev, 0));                                             \
    ctx->uploads_pending = false;                                                                  \
  }

static inline bool is_pow2(uint64_t x) { return x && !(x & (x - 1)); }
static inline uint32_t ilog2(uint64_t x) { uint32_t l = 0; while ((1ull << l) < x) l++; return l; }

// Copies a small host table to the device through the pinned ring; returns the device address.
static inline int32_t stage(sc_ctx* ctx, const void* host, size_t bytes, void** dptr) {
  size_t need = (bytes + 255) & ~(size_t)255;
  if (need > ctx->ring_size) return fail(SC_ENOMEM, "staging table too large");
  if (ctx->ring_off + need > ctx->ring_size) {
    CK(cudaStreamSynchronize(ctx->st));
    ctx->ring_off = 0;
  }
  memcpy(ctx->h_ring + ctx->ring_off, host, bytes);
  CK(cudaMemcpyAsync(ctx->d_ring + ctx->ring_off, ctx->h_ring + ctx->ring_off, bytes, cudaMemcpyHostToDevice, ctx->st));
  *dptr = ctx->d_ring + ctx->ring_off;
  ctx->ring_off += need;
  return SC_OK;
}

static inline void track(sc_ctx* ctx, sc_col* c) { c->id = ctx->next_id++; ctx->live[c->id] = c; }
static inline int32_t new_col(sc_ctx* ctx, uint64_t len, sc_col** out) {
  if (ctx->parena_on) {
    const size_t bytes = ((size_t)std::max<uint64_t>(len, 4) * 4 + 255) & ~(size_t)255;
    ctx->parena_need += bytes;
    if (ctx->parena && ctx->parena_off + bytes <= ctx->parena_cap) {
      *out = new sc_col{reinterpret_cast<uint32_t*>(ctx->parena + ctx->parena_off), len};
      (*out)->owned = false;
      ctx->parena_off += bytes;
      track(ctx, *out);
      return SC_OK;
    }
  }
  uint32_t* d = nullptr;
  cudaError_t e = cudaMallocAsync((void**)&d, std::max<uint64_t>(len, 4) * 4, ctx->st);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(SC_ENOMEM, std::string("cudaMallocAsync: ") + cudaGetErrorString(e)); }
  *out = new sc_col{d, len};
  track(ctx, *out);
  return SC_OK;
}


// MerkleProver::commit over mixed-size columns (capi.cu): layers_out[k] = layer of log size k; root_out may be NULL (no read-back).
extern "C" int32_t merkle_commit_impl(sc_ctx* ctx, sc_col* const* cols, uint32_t n, uint32_t log_repeat, sc_col** layers_out,
                                        uint32_t* max_log_out, uint32_t root_out[8]);

// Device-resident transcript for the FRI commit loop (fri_capi.cu): digest, one coefficient per mix, a copy of every root.
struct sc_dchan {
  sc_col* buf = nullptr;     // 8 (digest) + 4 * max (coefficients) + 8 * max (roots) words
  uint32_t max = 0, n = 0;   // mixes allowed / done
};
extern "C" const uint32_t* sc_dchan_coeff_ptr(const sc_dchan* dc, uint32_t k);
