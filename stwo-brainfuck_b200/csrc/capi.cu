// C ABI of libstwo_cuda.so (declared in include/stwo_cuda.h): context, columns, and one entry per Backend trait method.
// Host-side glue only — argument checking, grouping by size, staging of small tables; all arithmetic is in the kernels.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <tuple>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/stwo_cuda.h"
#include "kernels.cuh"
#include "air_params.cuh"

using namespace sb;

namespace sb {
static unsigned long long g_launch_orphans = 0;  // launches made before any context was entered on the thread
thread_local unsigned long long* g_launch_counter = &g_launch_orphans;
}
thread_local std::string g_sc_err;

#include "capi_internal.cuh"

extern "C" {

const char* sc_last_error(void) { return g_sc_err.c_str(); }
int32_t sc_version(void) { return 1; }

int32_t sc_ctx_create(int32_t device, void* stream, sc_ctx** out) {
  sc_ctx* ctx = nullptr;
  if (!out) return fail(SC_EINVAL, "null out");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(SC_ECUDA, "no CUDA device available (this library has no CPU path)"); }
  if (device < 0 || device >= ndev) return fail(SC_EINVAL, "bad device index");
  CK(cudaSetDevice(device));
  sc_ctx* c = new sc_ctx();
  c->device = device; c->poisoned = false; c->ring_size = 4u << 20; c->ring_off = 0;
  if (stream) { c->st = (cudaStream_t)stream; c->own_stream = false; }
  else { CK(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking)); c->own_stream = true; }
  CK(cudaMallocHost((void**)&c->h_ring, c->ring_size));
  CK(cudaMalloc((void**)&c->d_ring, c->ring_size));
  // keep freed blocks in the stream-ordered pool: column churn must not hit the driver allocator
  cudaMemPool_t pool;
  CK(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t thr = ~0ull;
  CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
  *out = c;
  return SC_OK;
}
int32_t sc_ctx_destroy(sc_ctx* ctx) {
  if (!ctx) return SC_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->st);
  for (auto& a : ctx->attached) if (a.p && a.dtor) { a.dtor(ctx, a.p); a.p = nullptr; }  // may free columns: the context is still whole
  {  // columns the caller never freed die with their context (their handles are invalid from here on)
    std::vector<sc_col*> left;
    for (auto& kv : ctx->live) left.push_back(kv.second);
    for (sc_col* c : left) sc_col_free(ctx, c);
  }
  cudaStreamSynchronize(ctx->st);
  cudaFreeHost(ctx->h_ring);
  for (auto& b : ctx->arena) cudaFreeHost(b.p);
  for (auto& kv : ctx->tw_cache) { cudaFree(kv.second->tw); cudaFree(kv.second->itw); delete kv.second; }
  cudaFree(ctx->d_ring);
  if (ctx->copy_st) { cudaStreamSynchronize(ctx->copy_st); cudaStreamDestroy(ctx->copy_st); cudaEventDestroy(ctx->copy_ev); cudaEventDestroy(ctx->slab_ev); }
  if (ctx->slab) cudaFree(ctx->slab);
  if (ctx->parena) cudaFree(ctx->parena);
  if (ctx->own_stream) cudaStreamDestroy(ctx->st);
  delete ctx;
  return SC_OK;
}
int32_t sc_ctx_sync(sc_ctx* ctx) { ENTER(); CK(cudaStreamSynchronize(ctx->st)); return SC_OK; }
// Proof arena: see sc_ctx::parena.  begin: (re)size the slab to what the previous bracket asked for (a growing slab is
// reallocated here, while nothing of the arena is alive) and start bump allocation; end: stop and remember the total.
int32_t sc_ctx_arena_begin(sc_ctx* ctx) {
  ENTER();
  if (ctx->parena_on) return fail(SC_EINVAL, "arena_begin: already inside a bracket");
  if (ctx->parena_want > ctx->parena_cap) {
    CK(cudaStreamSynchronize(ctx->st));
    if (ctx->parena) CK(cudaFree(ctx->parena));
    ctx->parena = nullptr; ctx->parena_cap = 0;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);   // the pool's cached blocks make room
    const size_t want = ctx->parena_want + ctx->parena_want / 16 + ((size_t)64 << 20);
    if (cudaMalloc((void**)&ctx->parena, want) == cudaSuccess) ctx->parena_cap = want;
    else { cudaGetLastError(); ctx->parena = nullptr; }   // no room: stay on the pool
  }
  ctx->parena_off = 0; ctx->parena_need = 0; ctx->parena_on = true;
  return SC_OK;
}
int32_t sc_ctx_arena_end(sc_ctx* ctx) {
  if (!ctx) return fail(SC_EINVAL, "null context");
  if (ctx->parena_on) { ctx->parena_on = false; ctx->parena_want = std::max(ctx->parena_want, ctx->parena_need); }
  return SC_OK;
}
// Makes the compute stream wait for every upload issued so far (sc_col_from_host_async, sc_trace_upload) without blocking the host.
int32_t sc_ctx_join_uploads(sc_ctx* ctx) { ENTER(); return SC_OK; }
int32_t sc_ctx_attach(sc_ctx* ctx, uint32_t slot, void* p, sc_attach_dtor dtor) {
  if (!ctx || slot >= 4) return fail(SC_EINVAL, "ctx_attach: bad argument");
  auto& a = ctx->attached[slot];
  if (a.p && a.dtor && a.p != p) a.dtor(ctx, a.p);
  a.p = p;
  a.dtor = dtor;
  return SC_OK;
}
void* sc_ctx_attached(sc_ctx* ctx, uint32_t slot) { return ctx && slot < 4 ? ctx->attached[slot].p : nullptr; }
uint64_t sc_ctx_launch_count(const sc_ctx* ctx) { return ctx ? ctx->launches : 0; }
int32_t sc_ctx_profile(sc_ctx* ctx, int32_t enable) {
  ENTER();
  ctx->profiling = enable != 0;
  return SC_OK;
}
int32_t sc_ctx_profiling(const sc_ctx* ctx) { return ctx && ctx->profiling ? 1 : 0; }
// Sums the recorded scopes per tag into "tag:ms:count;..." and clears them.  Returns the needed length.
size_t sc_ctx_profile_report(sc_ctx* ctx, char* buf, size_t cap) {
  if (!ctx) return 0;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->st);
  std::map<std::string, std::pair<double, int>> acc;
  for (auto& r : ctx->prof) {
    float ms = 0;
    cudaEventElapsedTime(&ms, r.a, r.b);
    acc[r.tag].first += ms; acc[r.tag].second += 1;
    ctx->ev_pool.push_back(r.a); ctx->ev_pool.push_back(r.b);
  }
  ctx->prof.clear();
  std::string out;
  for (auto& kv : acc) { char t[160]; snprintf(t, sizeof t, "%s:%.6f:%d;", kv.first.c_str(), kv.second.first, kv.second.second); out += t; }
  if (buf && cap) { size_t n = std::min(cap - 1, out.size()); memcpy(buf, out.data(), n); buf[n] = 0; }
  return out.size() + 1;
}

// Per-scope GPU timeline of the records collected so far ("tag:start_ms:dur_ms;" relative to the first record); does not
// clear them, so call it before sc_ctx_profile_report.  Holes between consecutive records are time the GPU sat idle.
size_t sc_ctx_profile_timeline(sc_ctx* ctx, char* buf, size_t cap) {
  if (!ctx) return 0;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->st);
  std::string out;
  for (auto& r : ctx->prof) {
    float t0 = 0, ms = 0;
    cudaEventElapsedTime(&t0, ctx->prof.front().a, r.a);
    cudaEventElapsedTime(&ms, r.a, r.b);
    char t[160]; snprintf(t, sizeof t, "%s:%.6f:%.6f;", r.tag, t0, ms); out += t;
  }
  if (buf && cap) { size_t n = std::min(cap - 1, out.size()); memcpy(buf, out.data(), n); buf[n] = 0; }
  return out.size() + 1;
}

// ------------------------------------------------------------------ columns
int32_t sc_col_uninit(sc_ctx* ctx, uint64_t len, sc_col** out) { ENTER(); if (!out) return fail(SC_EINVAL, "null out"); return new_col(ctx, len, out); }
int32_t sc_col_zeros(sc_ctx* ctx, uint64_t len, sc_col** out) {
  ENTER();
  if (!out) return fail(SC_EINVAL, "null out");
  int32_t r = new_col(ctx, len, out);
  if (r) return r;
  CK(cudaMemsetAsync((*out)->d, 0, len * 4, ctx->st));
  return SC_OK;
}
int32_t sc_col_from_host(sc_ctx* ctx, const uint32_t* host, uint64_t len, sc_col** out) {
  ENTER();
  if (!out || (!host && len)) return fail(SC_EINVAL, "null argument");
  int32_t r = new_col(ctx, len, out);
  if (r) return r;
  CK(cudaMemcpyAsync((*out)->d, host, len * 4, cudaMemcpyHostToDevice, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));  // host buffer may be pageable and is not retained
  return SC_OK;
}
// Same as sc_col_from_host but does not wait: `host` must stay valid (and should be pinned, e.g. from sc_host_arena_alloc,
// for the copy to be a real asynchronous DMA) until the next synchronising call on this context.
int32_t sc_col_from_host_async(sc_ctx* ctx, const uint32_t* host, uint64_t len, sc_col** out) {
  ENTER_NOJOIN();
  if (!out || (!host && len)) return fail(SC_EINVAL, "null argument");
  if (!ctx->copy_st) {
    CK(cudaStreamCreateWithFlags(&ctx->copy_st, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ctx->copy_ev, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->slab_ev, cudaEventDisableTiming));
  }
  const size_t bytes = ((std::max<uint64_t>(len, 4) * 4) + 255) & ~(size_t)255;
  if (ctx->slab_live == 0 && ctx->slab_used == 0 && ctx->slab_high > ctx->slab_cap) {
    // the previous round of uploads did not fit: grow once, while nothing lives in the slab (cudaFree/cudaMalloc synchronise)
    if (ctx->slab) CK(cudaFree(ctx->slab));
    ctx->slab = nullptr; ctx->slab_cap = 0;
    size_t want = ctx->slab_high + (ctx->slab_high >> 2);
    CK(cudaMalloc((void**)&ctx->slab, want));
    ctx->slab_cap = want;
  }
  ctx->slab_high = std::max(ctx->slab_high, ctx->slab_used + bytes);
  if (ctx->slab_used + bytes > ctx->slab_cap) {
    // no room (first proof on this context, or a larger program): ordinary pool column, copied on the compute stream
    int32_t r = new_col(ctx, len, out);
    if (r) return r;          // nothing was counted yet: a failed allocation leaves the slab bookkeeping untouched
    ctx->slab_used += bytes;  // keep counting so that slab_high sees the whole round
    ctx->slab_live++;
    (*out)->slab = true;      // only for the live count; `owned` stays true, so the memory goes back to the pool
    CK(cudaMemcpyAsync((*out)->d, host, len * 4, cudaMemcpyHostToDevice, ctx->st));
    return SC_OK;
  }
  if (ctx->slab_fence) {  // the slab was rewound: wait for the compute stream's position at that moment (sc_col_free)
    CK(cudaStreamWaitEvent(ctx->copy_st, ctx->slab_ev, 0));
    ctx->slab_fence = false;
  }
  uint32_t* d = reinterpret_cast<uint32_t*>(ctx->slab + ctx->slab_used);
  ctx->slab_used += bytes;
  ctx->slab_live++;
  sc_col* c = new sc_col{d, len};
  c->owned = false; c->slab = true;
  track(ctx, c);
  *out = c;
  // In pieces: the small host->device copies of the compute stream (pointer tables, stage()) share the one H2D copy
  // engine with these uploads and would otherwise sit behind a whole 16 MB column, idling the kernels that wait for them.
  const size_t piece = (size_t)512 << 10;
  for (size_t off = 0; off < len * 4; off += piece)
    CK(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(d) + off, reinterpret_cast<const uint8_t*>(host) + off, std::min(piece, (size_t)(len * 4 - off)),
                       cudaMemcpyHostToDevice, ctx->copy_st));
  ctx->uploads_pending = true;
  return SC_OK;
}
// Pinned host arena owned by the context: thread-safe bump allocation (64-byte aligned); reset releases everything at once
// but keeps the pinned blocks for the next proof.
int32_t sc_host_arena_alloc(sc_ctx* ctx, uint64_t bytes, void** out) {
  if (!ctx || !out) return fail(SC_EINVAL, "null argument");
  std::lock_guard<std::mutex> lk(ctx->arena_mu);
  size_t need = (bytes + 63) & ~(size_t)63;
  for (auto& b : ctx->arena)
    if (b.used + need <= b.size) { *out = b.p + b.used; b.used += need; return SC_OK; }
  size_t sz = std::max<size_t>(need, (size_t)64 << 20);
  uint8_t* p = nullptr;
  cudaSetDevice(ctx->device);
  if (cudaMallocHost((void**)&p, sz) != cudaSuccess) { cudaGetLastError(); return fail(SC_ENOMEM, "cudaMallocHost failed"); }
  ctx->arena.push_back({p, sz, need});
  *out = p;
  return SC_OK;
}
int32_t sc_host_arena_reset(sc_ctx* ctx) {
  if (!ctx) return fail(SC_EINVAL, "null context");
  std::lock_guard<std::mutex> lk(ctx->arena_mu);
  for (auto& b : ctx->arena) b.used = 0;
  return SC_OK;
}
int32_t sc_col_to_host(sc_ctx* ctx, const sc_col* col, uint32_t* host) {
  ENTER();
  if (!col || !host) return fail(SC_EINVAL, "null argument");
  CK(cudaMemcpyAsync(host, col->d, col->len * 4, cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return SC_OK;
}
int32_t sc_col_read(sc_ctx* ctx, const sc_col* col, uint64_t offset, uint64_t n, uint32_t* host) {
  ENTER_NOJOIN();
  if (!col || !host || offset + n > col->len) return fail(SC_EINVAL, "read out of range");
  // a read of an ordinary column does not wait for uploads still in flight on the copy stream (a Merkle root can be read
  // back while the next phase's inputs are being copied); only a column that lives in the upload slab needs them
  if (col->slab && ctx->uploads_pending) {
    CK(cudaEventRecord(ctx->copy_ev, ctx->copy_st));
    CK(cudaStreamWaitEvent(ctx->st, ctx->copy_ev, 0));
    ctx->uploads_pending = false;
  }
  CK(cudaMemcpyAsync(host, col->d + offset, n * 4, cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return SC_OK;
}
int32_t sc_col_write(sc_ctx* ctx, sc_col* col, uint64_t offset, uint64_t n, const uint32_t* host) {
  ENTER();
  if (!col || !host || offset + n > col->len) return fail(SC_EINVAL, "write out of range");
  CK(cudaMemcpyAsync(col->d + offset, host, n * 4, cudaMemcpyHostToDevice, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return SC_OK;
}
int32_t sc_col_clone(sc_ctx* ctx, const sc_col* col, sc_col** out) {
  ENTER();
  if (!col || !out) return fail(SC_EINVAL, "null argument");
  int32_t r = new_col(ctx, col->len, out);
  if (r) return r;
  CK(cudaMemcpyAsync((*out)->d, col->d, col->len * 4, cudaMemcpyDeviceToDevice, ctx->st));
  return SC_OK;
}
int32_t sc_col_free(sc_ctx* ctx, sc_col* col) {
  if (!col) return SC_OK;
  if (!ctx) return fail(SC_EINVAL, "null context");
  cudaSetDevice(ctx->device);
  ctx->live.erase(col->id);
  if (col->owned) cudaFreeAsync(col->d, ctx->st);
  if (col->slab && --ctx->slab_live == 0) {
    // rewind; whatever is queued on `st` up to here may still read the slab, later uploads must not overtake it
    ctx->slab_used = 0;
    if (ctx->slab_ev) { cudaEventRecord(ctx->slab_ev, ctx->st); ctx->slab_fence = true; }
  }
  delete col;
  return SC_OK;
}
// GPU-side stopwatch for callers that overlap host work with queued kernels: sc_event_record marks a point on the compute
// stream; sc_event_elapsed(a, b) waits for b and returns the device time between the two marks (the idle gap when b was
// recorded after the host finished something the device had to wait for).
struct sc_event { cudaEvent_t ev; };
int32_t sc_event_record(sc_ctx* ctx, sc_event** out) {
  ENTER_NOJOIN();
  if (!out) return fail(SC_EINVAL, "null out");
  sc_event* e = new sc_event;
  CK(cudaEventCreate(&e->ev));
  CK(cudaEventRecord(e->ev, ctx->st));
  *out = e;
  return SC_OK;
}
int32_t sc_event_elapsed(sc_ctx* ctx, const sc_event* a, const sc_event* b, float* ms) {
  ENTER_NOJOIN();
  if (!a || !b || !ms) return fail(SC_EINVAL, "null argument");
  CK(cudaEventSynchronize(b->ev));
  CK(cudaEventElapsedTime(ms, a->ev, b->ev));
  return SC_OK;
}
int32_t sc_event_free(sc_ctx* ctx, sc_event* e) {
  if (!e) return SC_OK;
  if (ctx) cudaSetDevice(ctx->device);
  cudaEventDestroy(e->ev);
  delete e;
  return SC_OK;
}
// sc_ctx_mark returns a token; sc_ctx_release_since frees every column created on this context after that token that is
// still alive (views included).  For callers that unwind from a failure and no longer know what they allocated.
uint64_t sc_ctx_mark(sc_ctx* ctx) { return ctx ? ctx->next_id : 0; }
uint64_t sc_ctx_live_columns(sc_ctx* ctx) { return ctx ? ctx->live.size() : 0; }
int32_t sc_ctx_release_since(sc_ctx* ctx, uint64_t mark) {
  if (!ctx) return fail(SC_EINVAL, "null context");
  std::vector<sc_col*> victims;
  for (auto it = ctx->live.lower_bound(mark); it != ctx->live.end(); ++it) victims.push_back(it->second);
  for (sc_col* c : victims) sc_col_free(ctx, c);
  return SC_OK;
}
// Non-owning column over caller-owned device memory (e.g. a torch tensor or an NCCL receive buffer); 16-byte aligned.
int32_t sc_col_wrap(sc_ctx* ctx, void* device_ptr, uint64_t len, sc_col** out) {
  ENTER();
  if (!device_ptr || !out || ((uintptr_t)device_ptr & 15)) return fail(SC_EINVAL, "col_wrap: null or misaligned pointer");
  sc_col* c = new sc_col{(uint32_t*)device_ptr, len};
  c->owned = false;
  track(ctx, c);
  *out = c;
  return SC_OK;
}
uint64_t sc_col_len(const sc_col* col) { return col ? col->len : 0; }
void* sc_col_device_ptr(sc_col* col) { return col ? col->d : nullptr; }

int32_t sc_col_broadcast16(sc_ctx* ctx, const sc_col* src, sc_col** out) {
  ENTER();
  if (!src || !out) return fail(SC_EINVAL, "null argument");
  int32_t r = new_col(ctx, src->len * 16, out);
  if (r) return r;
  { ProfScope ps_(ctx, "broadcast16"); CKL(launch_broadcast16(src->d, (*out)->d, src->len, ctx->st)); }
  return SC_OK;
}

int32_t sc_bit_reverse(sc_ctx* ctx, sc_col* col) {
  ENTER();
  if (!col || !is_pow2(col->len)) return fail(SC_EINVAL, "bit_reverse: length must be a power of two");
  { ProfScope ps_(ctx, "bit_reverse"); CKL(launch_bit_reverse(col->d, ilog2(col->len), ctx->st)); }
  return SC_OK;
}

int32_t sc_batch_inverse_m31(sc_ctx* ctx, const sc_col* src, sc_col* dst) {
  ENTER();
  if (!src || !dst || src->len != dst->len) return fail(SC_EINVAL, "batch_inverse: length mismatch");
  { ProfScope ps_(ctx, "batch_inverse"); CKL(launch_batch_inverse_m31(src->d, dst->d, src->len, ctx->st)); }
  return SC_OK;
}
int32_t sc_batch_inverse_qm31(sc_ctx* ctx, sc_col* const src[4], sc_col* const dst[4]) {
  ENTER();
  const uint32_t* s[4]; uint32_t* d[4];
  for (int k = 0; k < 4; k++) {
    if (!src[k] || !dst[k] || src[k]->len != src[0]->len || dst[k]->len != src[0]->len) return fail(SC_EINVAL, "batch_inverse: bad columns");
    s[k] = src[k]->d; d[k] = dst[k]->d;
  }
  { ProfScope ps_(ctx, "batch_inverse"); CKL(launch_batch_inverse_qm31(s, d, src[0]->len, ctx->st)); }
  return SC_OK;
}

// ------------------------------------------------------------------ twiddles / FFT
int32_t sc_precompute_twiddles(sc_ctx* ctx, uint32_t root_log, sc_twiddles** out) {
  ENTER();
  if (!out || root_log < 2 || root_log > 29) return fail(SC_EINVAL, "precompute_twiddles: root_log must be in [2,29]");
  sc_twiddles* t = new sc_twiddles{root_log, nullptr, nullptr};
  // 2 x 2^root_log words each: the tree, then the same tree doubled for the FFT butterflies (fft.cu mulred)
  CK(cudaMallocAsync((void**)&t->tw, (size_t)8 << root_log, ctx->st));
  CK(cudaMallocAsync((void**)&t->itw, (size_t)8 << root_log, ctx->st));
  { ProfScope ps_(ctx, "twiddles"); CKL(launch_twiddle_tree(t->tw, t->itw, root_log, ctx->st)); }
  *out = t;
  return SC_OK;
}
// Twiddle tree owned and cached by the context (the tree depends only on root_log): computed on first use, returned as a
// borrowed handle afterwards, released by sc_ctx_destroy.  Do not pass the handle to sc_twiddles_free.
int32_t sc_twiddles_cached(sc_ctx* ctx, uint32_t root_log, const sc_twiddles** out) {
  ENTER();
  if (!out) return fail(SC_EINVAL, "null out");
  auto it = ctx->tw_cache.find(root_log);
  if (it == ctx->tw_cache.end()) {
    sc_twiddles* t = nullptr;
    int32_t r = sc_precompute_twiddles(ctx, root_log, &t);
    if (r) return r;
    it = ctx->tw_cache.emplace(root_log, t).first;
  }
  *out = it->second;
  return SC_OK;
}
int32_t sc_twiddles_free(sc_ctx* ctx, sc_twiddles* tw) {
  if (!tw) return SC_OK;
  if (!ctx) return fail(SC_EINVAL, "null context");
  cudaSetDevice(ctx->device);
  cudaFreeAsync(tw->tw, ctx->st);
  cudaFreeAsync(tw->itw, ctx->st);
  delete tw;
  return SC_OK;
}
int32_t sc_twiddles_to_host(sc_ctx* ctx, const sc_twiddles* tw, uint32_t* twiddles, uint32_t* itwiddles) {
  ENTER();
  if (!tw) return fail(SC_EINVAL, "null twiddles");
  if (twiddles) CK(cudaMemcpyAsync(twiddles, tw->tw, (size_t)4 << tw->root_log, cudaMemcpyDeviceToHost, ctx->st));
  if (itwiddles) CK(cudaMemcpyAsync(itwiddles, tw->itw, (size_t)4 << tw->root_log, cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return SC_OK;
}

int32_t sc_interpolate(sc_ctx* ctx, sc_col* const* cols, uint32_t n, const sc_twiddles* tw) {
  ENTER();
  if (!tw || (!cols && n)) return fail(SC_EINVAL, "null argument");
  std::map<uint32_t, std::vector<uint32_t*>> by_log;
  for (uint32_t i = 0; i < n; i++) {
    if (!cols[i] || !is_pow2(cols[i]->len)) return fail(SC_EINVAL, "interpolate: column length must be a power of two");
    uint32_t lg = ilog2(cols[i]->len);
    if (lg < 3) return fail(SC_EINVAL, "interpolate: log size < 3 is not supported on the device");
    if (lg > tw->root_log + 1) return fail(SC_EINVAL, "interpolate: twiddle tree too small for this domain");
    by_log[lg].push_back(cols[i]->d);
  }
  std::vector<uint32_t*> all;
  for (auto& kv : by_log) all.insert(all.end(), kv.second.begin(), kv.second.end());
  if (all.empty()) return SC_OK;
  void* dp;
  int32_t r = stage(ctx, all.data(), all.size() * sizeof(void*), &dp);   // one pointer table for every size class
  if (r) return r;
  size_t off = 0;
  for (auto& kv : by_log) {
    { ProfScope ps_(ctx, "fft_interpolate"); CKL(launch_interpolate((uint32_t* const*)dp + off, (uint32_t)kv.second.size(), kv.first, tw->itw + ((size_t)2 << tw->root_log), ctx->st)); }
    off += kv.second.size();
  }
  return SC_OK;
}

int32_t sc_evaluate(sc_ctx* ctx, sc_col* const* coeffs, uint32_t n, uint32_t log_blowup, const sc_twiddles* tw, sc_col** out) {
  ENTER();
  if (!tw || !out || (!coeffs && n)) return fail(SC_EINVAL, "null argument");
  struct G { std::vector<const uint32_t*> src; std::vector<uint32_t*> dst; };
  std::map<uint32_t, G> by_log;
  for (uint32_t i = 0; i < n; i++) {
    if (!coeffs[i] || !is_pow2(coeffs[i]->len)) return fail(SC_EINVAL, "evaluate: column length must be a power of two");
    uint32_t lg = ilog2(coeffs[i]->len);
    if (lg + log_blowup < 3) return fail(SC_EINVAL, "evaluate: domain log size < 3 is not supported on the device");
    if (lg + log_blowup > tw->root_log + 1) return fail(SC_EINVAL, "evaluate: twiddle tree too small for this domain");
  }
  std::vector<uint32_t*> temps;  // zero-extended coefficient copies for log_blowup > 1 (PolyOps::extend)
  for (uint32_t i = 0; i < n; i++) {
    uint32_t lg = ilog2(coeffs[i]->len);
    int32_t r = new_col(ctx, coeffs[i]->len << log_blowup, &out[i]);
    if (r) return r;
    const uint32_t* src = coeffs[i]->d;
    if (log_blowup > 1) {
      uint32_t* t;
      size_t half = (size_t)coeffs[i]->len << (log_blowup - 1);
      CK(cudaMallocAsync((void**)&t, half * 4, ctx->st));
      CK(cudaMemsetAsync(t, 0, half * 4, ctx->st));
      CK(cudaMemcpyAsync(t, src, coeffs[i]->len * 4, cudaMemcpyDeviceToDevice, ctx->st));
      temps.push_back(t);
      src = t;
      lg += log_blowup - 1;
    }
    by_log[lg].src.push_back(src);
    by_log[lg].dst.push_back(out[i]->d);
  }
  if (log_blowup > 1) log_blowup = 1;
  {
    std::vector<const void*> all;   // every group's sources, then every group's destinations: one staged table
    for (auto& kv : by_log) all.insert(all.end(), kv.second.src.begin(), kv.second.src.end());
    const size_t ndst0 = all.size();
    for (auto& kv : by_log) all.insert(all.end(), kv.second.dst.begin(), kv.second.dst.end());
    void* dall = nullptr;
    if (!all.empty()) { int32_t r = stage(ctx, all.data(), all.size() * sizeof(void*), &dall); if (r) return r; }
    size_t off = 0;
    for (auto& kv : by_log) {
      const uint32_t* const* ds = (const uint32_t* const*)dall + off;
      uint32_t* const* dd = (uint32_t* const*)dall + ndst0 + off;
      { ProfScope ps_(ctx, "fft_evaluate"); CKL(launch_evaluate(ds, dd, (uint32_t)kv.second.src.size(), kv.first,
                          kv.first + log_blowup, tw->tw + ((size_t)2 << tw->root_log), ctx->st)); }
      off += kv.second.src.size();
    }
  }
  for (uint32_t* t : temps) CK(cudaFreeAsync(t, ctx->st));
  return SC_OK;
}

// ---- lane-repeated columns.  A column of 2^m stored values stands for an evaluation of log size m + log_repeat in which
// every value fills 2^log_repeat consecutive rows (the reference writes one table row into all 16 SIMD lanes:
// components/processor/table.rs:86-100).  Its polynomial has a single non-zero coefficient per 2^log_repeat, and the first
// log_repeat FFT layers only scale or replicate, so both transforms run on the 2^m distinct values (fft.cu, LINE).
int32_t sc_interpolate_repeated(sc_ctx* ctx, sc_col* const* cols, uint32_t n, uint32_t log_repeat, const sc_twiddles* tw, sc_col** out) {
  ENTER();
  if (!tw || (!cols && n)) return fail(SC_EINVAL, "null argument");
  struct G { std::vector<const uint32_t*> src; std::vector<uint32_t*> dst; };
  std::map<uint32_t, G> by_log;
  for (uint32_t i = 0; i < n; i++) {
    if (!cols[i] || !is_pow2(cols[i]->len)) return fail(SC_EINVAL, "interpolate_repeated: column length must be a power of two");
    uint32_t lg = ilog2(cols[i]->len);
    if (lg + log_repeat > tw->root_log + 1 || lg > tw->root_log) return fail(SC_EINVAL, "interpolate_repeated: twiddle tree too small for this domain");
  }
  for (uint32_t i = 0; i < n; i++) {
    uint32_t* dst = cols[i]->d;
    if (out) { int32_t r = new_col(ctx, cols[i]->len, &out[i]); if (r) return r; dst = out[i]->d; }
    G& g = by_log[ilog2(cols[i]->len)];
    g.src.push_back(cols[i]->d); g.dst.push_back(dst);
  }
  for (auto& kv : by_log) {
    void *ds, *dd;
    int32_t r = stage(ctx, kv.second.src.data(), kv.second.src.size() * sizeof(void*), &ds); if (r) return r;
    r = stage(ctx, kv.second.dst.data(), kv.second.dst.size() * sizeof(void*), &dd); if (r) return r;
    { ProfScope ps_(ctx, "fft_interpolate"); CKL(launch_interpolate_repeated((const uint32_t* const*)ds, (uint32_t* const*)dd, (uint32_t)kv.second.src.size(), kv.first, tw->itw + ((size_t)2 << tw->root_log), ctx->st)); }
  }
  return SC_OK;
}

// coeffs: compact coefficients (sc_interpolate_repeated).  out: new FULL columns of length (len << log_repeat) << log_blowup.
static int32_t evaluate_repeated_impl(sc_ctx* ctx, sc_col* const* coeffs, uint32_t n, uint32_t log_repeat, uint32_t log_blowup, const sc_twiddles* tw,
                                      const uint64_t* row_off, const uint64_t* row_cnt, sc_col** out) {
  ENTER();
  if (!tw || !out || (!coeffs && n)) return fail(SC_EINVAL, "null argument");
  if (log_blowup > 1) return fail(SC_EINVAL, "evaluate_repeated: log_blowup > 1 is not supported");
  if (log_repeat < 2 || log_repeat > 8) return fail(SC_EINVAL, "evaluate_repeated: log_repeat must be in [2, 8]");
  struct G { std::vector<const uint32_t*> src; std::vector<uint32_t*> tmp, dst; std::vector<const uint32_t*> part; };
  struct Key { uint32_t lg; uint64_t off, cnt; bool operator<(const Key& o) const { return std::tie(lg, off, cnt) < std::tie(o.lg, o.off, o.cnt); } };
  std::map<Key, G> groups;
  const uint64_t rmask = (1ull << log_repeat) - 1;
  for (uint32_t i = 0; i < n; i++) {
    if (!coeffs[i] || !is_pow2(coeffs[i]->len)) return fail(SC_EINVAL, "evaluate_repeated: column length must be a power of two");
    uint32_t lg = ilog2(coeffs[i]->len);
    if (lg + log_blowup + log_repeat > tw->root_log + 1 || lg + log_blowup > tw->root_log) return fail(SC_EINVAL, "evaluate_repeated: twiddle tree too small for this domain");
    if (row_off) {
      uint64_t full = (coeffs[i]->len << log_repeat) << log_blowup;
      if ((row_off[i] & rmask) || (row_cnt[i] & rmask) || row_off[i] + row_cnt[i] > full) return fail(SC_EINVAL, "evaluate_repeated: row range must be aligned to the repetition and inside the domain");
    }
  }
  std::vector<uint32_t*> temps;
  for (uint32_t i = 0; i < n; i++) {
    uint32_t lg = ilog2(coeffs[i]->len);
    uint64_t full = (coeffs[i]->len << log_repeat) << log_blowup;
    uint64_t off = row_off ? row_off[i] : 0, cnt = row_off ? row_cnt[i] : full;
    int32_t r = new_col(ctx, cnt, &out[i]);
    if (r) return r;
    uint32_t* t;
    CK(cudaMallocAsync((void**)&t, (coeffs[i]->len << log_blowup) * 4, ctx->st));
    temps.push_back(t);
    G& g = groups[Key{lg, off, cnt}];
    g.src.push_back(coeffs[i]->d); g.tmp.push_back(t); g.dst.push_back(out[i]->d); g.part.push_back(t + (off >> log_repeat));
  }
  for (auto& kv : groups) {
    void *ds, *dt, *dd, *dpart;
    size_t nc = kv.second.src.size();
    int32_t r = stage(ctx, kv.second.src.data(), nc * sizeof(void*), &ds); if (r) return r;
    r = stage(ctx, kv.second.tmp.data(), nc * sizeof(void*), &dt); if (r) return r;
    r = stage(ctx, kv.second.dst.data(), nc * sizeof(void*), &dd); if (r) return r;
    r = stage(ctx, kv.second.part.data(), nc * sizeof(void*), &dpart); if (r) return r;
    {
      ProfScope ps_(ctx, "fft_evaluate");
      CKL(launch_evaluate_repeated((const uint32_t* const*)ds, (uint32_t* const*)dt, (uint32_t)nc, kv.first.lg, kv.first.lg + log_blowup,
                                   tw->tw + ((size_t)2 << tw->root_log), ctx->st));
    }
    if (kv.first.cnt) { ProfScope ps_(ctx, "broadcast16"); CKL(launch_broadcast_cols((const uint32_t* const*)dpart, (uint32_t* const*)dd, (uint32_t)nc, (size_t)(kv.first.cnt >> log_repeat), log_repeat, ctx->st)); }
  }
  for (uint32_t* t : temps) CK(cudaFreeAsync(t, ctx->st));
  return SC_OK;
}
int32_t sc_evaluate_repeated(sc_ctx* ctx, sc_col* const* coeffs, uint32_t n, uint32_t log_repeat, uint32_t log_blowup, const sc_twiddles* tw, sc_col** out) {
  return evaluate_repeated_impl(ctx, coeffs, n, log_repeat, log_blowup, tw, nullptr, nullptr, out);
}
// Rows [row_off[i], row_off[i] + row_cnt[i]) of the evaluation only (both multiples of 2^log_repeat): what one rank of the
// sharded prover keeps of a main-trace column — every rank transforms the distinct values itself instead of exchanging LDEs.
int32_t sc_evaluate_repeated_range(sc_ctx* ctx, sc_col* const* coeffs, uint32_t n, uint32_t log_repeat, uint32_t log_blowup, const sc_twiddles* tw,
                                   const uint64_t* row_off, const uint64_t* row_cnt, sc_col** out) {
  if (n && (!row_off || !row_cnt)) return fail(SC_EINVAL, "null argument");
  return evaluate_repeated_impl(ctx, coeffs, n, log_repeat, log_blowup, tw, row_off, row_cnt, out);
}

static int32_t eval_at_point_impl(sc_ctx* ctx, sc_col* const* polys, const uint32_t* log_repeats, uint32_t n, const uint32_t* points, uint32_t* out);
int32_t sc_eval_at_point(sc_ctx* ctx, sc_col* const* polys, uint32_t n, const uint32_t* points, uint32_t* out) {
  return eval_at_point_impl(ctx, polys, nullptr, n, points, out);
}
int32_t sc_eval_at_point_repeated(sc_ctx* ctx, sc_col* const* polys, const uint32_t* log_repeats, uint32_t n, const uint32_t* points, uint32_t* out) {
  if (!log_repeats && n) return fail(SC_EINVAL, "null argument");
  return eval_at_point_impl(ctx, polys, log_repeats, n, points, out);
}
// log_repeats[i] = r > 0: polys[i] holds the compact coefficients of sc_interpolate_repeated, i.e. coefficient j stands at
// index j << r of the full vector, whose basis monomial is the product of f_k over the set bits k >= r of that index.
static int32_t eval_at_point_impl(sc_ctx* ctx, sc_col* const* polys, const uint32_t* log_repeats, uint32_t n, const uint32_t* points, uint32_t* out) {
  ENTER();
  if (!n) return SC_OK;
  if (!polys || !points || !out) return fail(SC_EINVAL, "null argument");
  std::vector<EvalTaskHost> tasks(n);
  uint32_t blocks = 0;
  for (uint32_t i = 0; i < n; i++) {
    if (!polys[i] || !is_pow2(polys[i]->len)) return fail(SC_EINVAL, "eval_at_point: length must be a power of two");
    uint32_t lg = ilog2(polys[i]->len);
    const uint32_t rep = log_repeats ? log_repeats[i] : 0;
    if (lg + rep > 28) return fail(SC_EINVAL, "eval_at_point: polynomial too large");
    const uint32_t* p = points + 8 * i;
    QM31 x = q_make(p[0], p[1], p[2], p[3]), y = q_make(p[4], p[5], p[6], p[7]);
    EvalTaskHost& t = tasks[i];
    t.coeffs = polys[i]->d; t.log = lg; t.first_block = blocks;
    for (int k = 0; k < 28; k++) t.f[k] = q_zero();
    if (rep == 0) t.f[0] = y;
    for (uint32_t k = 1; k < lg + rep; k++) { if (k >= rep) t.f[k - rep] = x; x = q_sub(q_mulm(q_sqr(x), 2), q_fromm(1)); }
    blocks += lg > 13 ? (1u << (lg - 13)) : 1u;  // EV_CHUNK_LOG (ops.cu)
  }
  void* dt;
  int32_t r = stage(ctx, tasks.data(), tasks.size() * sizeof(EvalTaskHost), &dt);
  if (r) return r;
  QM31 *partials, *work, *dout;
  CK(cudaMallocAsync((void**)&partials, (size_t)blocks * sizeof(QM31), ctx->st));
  CK(cudaMallocAsync((void**)&work, (size_t)blocks * sizeof(QM31), ctx->st));
  CK(cudaMallocAsync((void**)&dout, (size_t)n * sizeof(QM31), ctx->st));
  { ProfScope ps_(ctx, "eval_at_point"); CKL(launch_eval_at_point_tasks(dt, n, blocks, partials, work, dout, ctx->st)); }
  CK(cudaMemcpyAsync(out, dout, (size_t)n * sizeof(QM31), cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  CK(cudaFreeAsync(partials, ctx->st));
  CK(cudaFreeAsync(work, ctx->st));
  CK(cudaFreeAsync(dout, ctx->st));
  return SC_OK;
}

// ------------------------------------------------------------------ Merkle
static int32_t commit_layer_impl(sc_ctx* ctx, uint32_t log_size, const sc_col* prev, sc_col* const* cols, uint32_t n, uint32_t log_repeat, sc_col** out);
int32_t sc_merkle_commit_layer(sc_ctx* ctx, uint32_t log_size, const sc_col* prev, sc_col* const* cols, uint32_t n, sc_col** out) {
  return commit_layer_impl(ctx, log_size, prev, cols, n, 0, out);
}
// As sc_merkle_commit_layer when every column (and therefore every child pair) repeats each value 2^log_repeat times:
// one node per group is hashed and its digest stored 2^log_repeat times.  The caller guarantees the repetition.
int32_t sc_merkle_commit_layer_repeated(sc_ctx* ctx, uint32_t log_size, const sc_col* prev, sc_col* const* cols, uint32_t n, uint32_t log_repeat, sc_col** out) {
  if (log_repeat > 8) return fail(SC_EINVAL, "commit_on_layer: log_repeat > 8");
  return commit_layer_impl(ctx, log_size, prev, cols, n, log_repeat, out);
}
static int32_t commit_layer_impl(sc_ctx* ctx, uint32_t log_size, const sc_col* prev, sc_col* const* cols, uint32_t n, uint32_t log_repeat, sc_col** out) {
  ENTER();
  if (!out || (!cols && n) || log_size > 30) return fail(SC_EINVAL, "commit_on_layer: bad argument");
  uint64_t rows = 1ull << log_size;
  if (prev && prev->len != rows * 16) return fail(SC_EINVAL, "commit_on_layer: previous layer must have 2^(log_size+1) digests");
  std::vector<const uint32_t*> p(n);
  for (uint32_t i = 0; i < n; i++) {
    if (!cols[i] || cols[i]->len != rows) return fail(SC_EINVAL, "commit_on_layer: column length != 2^log_size");
    p[i] = cols[i]->d;
  }
  void* dp = nullptr;
  if (n) { int32_t r = stage(ctx, p.data(), n * sizeof(void*), &dp); if (r) return r; }
  int32_t r = new_col(ctx, rows * 8, out);
  if (r) return r;
  { ProfScope ps_(ctx, "merkle_commit_layer"); CKL(launch_commit_layer(log_size, prev ? prev->d : nullptr, (const uint32_t* const*)dp, n, (*out)->d, ctx->st, log_repeat)); }
  return SC_OK;
}

int32_t sc_merkle_commit(sc_ctx* ctx, sc_col* const* cols, uint32_t n, sc_col** layers_out, uint32_t* max_log_out, uint32_t root_out[8]) {
  return merkle_commit_impl(ctx, cols, n, 0, layers_out, max_log_out, root_out);
}
// Tree over columns that ALL repeat each value 2^log_repeat times: the nodes of the deepest layer then repeat 2^log_repeat
// times, those of the layer above half as often, and so on; the first log_repeat layers hash one node per group.
int32_t sc_merkle_commit_repeated(sc_ctx* ctx, sc_col* const* cols, uint32_t n, uint32_t log_repeat, sc_col** layers_out, uint32_t* max_log_out, uint32_t root_out[8]) {
  if (log_repeat > 8) return fail(SC_EINVAL, "merkle_commit: log_repeat > 8");
  return merkle_commit_impl(ctx, cols, n, log_repeat, layers_out, max_log_out, root_out);
}
int32_t merkle_commit_impl(sc_ctx* ctx, sc_col* const* cols, uint32_t n, uint32_t log_repeat, sc_col** layers_out, uint32_t* max_log_out, uint32_t root_out[8]) {
  ENTER();
  if (!layers_out || (!cols && n)) return fail(SC_EINVAL, "null argument");
  uint32_t max_log = 0;
  for (uint32_t i = 0; i < n; i++) {
    if (!cols[i] || !is_pow2(cols[i]->len)) return fail(SC_EINVAL, "merkle_commit: column length must be a power of two");
    max_log = std::max(max_log, ilog2(cols[i]->len));
  }
  // layers above `top` one launch each; the remaining small ones (no repetition left there) in a single launch
  int top = (int)std::min<uint32_t>(max_log, MERKLE_TOP_LOG);
  if (log_repeat && (uint32_t)top + log_repeat > max_log) top = (int)max_log - (int)log_repeat;  // may become < 0: nothing fused
  // one pointer table for the whole tree (columns by layer, deepest first, input order within a layer), staged once: a
  // per-layer table would put a small host->device copy in front of every launch
  std::vector<const uint32_t*> cp;
  std::vector<uint32_t> first(max_log + 2, 0);   // columns of layer lg are cp[first[lg+1] .. first[lg]) in this numbering
  for (int lg = (int)max_log; lg >= 0; lg--) {
    first[lg + 1] = (uint32_t)cp.size();
    for (uint32_t i = 0; i < n; i++) if (ilog2(cols[i]->len) == (uint32_t)lg) cp.push_back(cols[i]->d);  // stable
  }
  first[0] = (uint32_t)cp.size();
  const uint32_t* const* dp = nullptr;
  if (!cp.empty()) { void* d; int32_t r = stage(ctx, cp.data(), cp.size() * sizeof(void*), &d); if (r) return r; dp = (const uint32_t* const*)d; }
  // layers top+1 .. sub_from (at most 2^MERKLE_SUB_FROM nodes, none of them in the repeated region): one sub-tree launch
  int sub_from = -1;
  if (top == (int)MERKLE_TOP_LOG && !getenv("SC_MERKLE_NO_SUBTREE")) {
    sub_from = (int)std::min<uint32_t>(max_log > log_repeat ? max_log - log_repeat : 0, MERKLE_SUB_FROM);
    if (log_repeat == 0) sub_from = (int)std::min<uint32_t>(max_log, MERKLE_SUB_FROM);
    if (sub_from - top < 2) sub_from = -1;   // nothing to gain over one or two plain launches
  }
  for (int lg = (int)max_log; lg > top; lg--) {
    if (lg == sub_from) {
      const uint32_t S = (uint32_t)(sub_from - top - 1);
      uint32_t col_off[MERKLE_SUB_MAX + 2];
      uint32_t* outp[MERKLE_SUB_MAX + 1];
      for (uint32_t k = 0; k <= S; k++) {
        int l2 = sub_from - (int)k;
        col_off[k] = first[l2 + 1] - first[sub_from + 1];
        int32_t r = new_col(ctx, 8ull << l2, &layers_out[l2]);
        if (r) return r;
        outp[k] = layers_out[l2]->d;
      }
      col_off[S + 1] = first[sub_from - (int)S] - first[sub_from + 1];
      ProfScope ps_(ctx, "merkle_commit_layer");
      CKL(launch_commit_subtree((uint32_t)sub_from, S, sub_from == (int)max_log ? nullptr : layers_out[sub_from + 1]->d, dp + first[sub_from + 1],
                                col_off, outp, ctx->st));
      lg = top + 1;   // the loop's decrement leaves the walk at `top`
      continue;
    }
    uint32_t depth = max_log - (uint32_t)lg, rep = log_repeat > depth ? log_repeat - depth : 0;
    int32_t r = new_col(ctx, 8ull << lg, &layers_out[lg]);
    if (r) return r;
    ProfScope ps_(ctx, "merkle_commit_layer");
    CKL(launch_commit_layer(lg, lg == (int)max_log ? nullptr : layers_out[lg + 1]->d, dp + first[lg + 1], first[lg] - first[lg + 1],
                            layers_out[lg]->d, ctx->st, rep));
  }
  if (top >= 0) {
    uint32_t col_off[MERKLE_TOP_LOG + 2];
    uint32_t* outp[MERKLE_TOP_LOG + 1];
    for (int k = 0; k <= top; k++) {
      int lg = top - k;
      col_off[k] = first[lg + 1] - first[top + 1];
      int32_t r = new_col(ctx, 8ull << lg, &layers_out[lg]);
      if (r) return r;
      outp[k] = layers_out[lg]->d;
    }
    col_off[top + 1] = first[0] - first[top + 1];
    ProfScope ps_(ctx, "merkle_commit_layer");
    CKL(launch_commit_top((uint32_t)top, top == (int)max_log ? nullptr : layers_out[top + 1]->d, dp + first[top + 1], col_off, outp, ctx->st));
  }
  if (max_log_out) *max_log_out = max_log;
  if (root_out) return sc_col_read(ctx, layers_out[0], 0, 8, root_out);
  return SC_OK;
}

// ------------------------------------------------------------------ FRI
int32_t sc_fold_line(sc_ctx* ctx, sc_col* const src[4], uint32_t log, const uint32_t alpha[4], const sc_twiddles* tw, sc_col* dst_out[4]) {
  ENTER();
  if (!tw || log < 1 || log > tw->root_log) return fail(SC_EINVAL, "fold_line: bad log size / twiddle tree too small");
  const uint32_t* s[4]; uint32_t* d[4];
  for (int k = 0; k < 4; k++) {
    if (!src[k] || src[k]->len != (1ull << log)) return fail(SC_EINVAL, "fold_line: bad source column");
    s[k] = src[k]->d;
  }
  for (int k = 0; k < 4; k++) { int32_t r = new_col(ctx, 1ull << (log - 1), &dst_out[k]); if (r) return r; d[k] = dst_out[k]->d; }
  { ProfScope ps_(ctx, "fold_line"); CKL(launch_fold_line(s, log, q_make(alpha[0], alpha[1], alpha[2], alpha[3]), d, tw->itw + ((size_t)1 << tw->root_log), ctx->st)); }
  return SC_OK;
}
int32_t sc_fold_circle_into_line(sc_ctx* ctx, sc_col* const src[4], uint32_t log, const uint32_t alpha[4], const sc_twiddles* tw, sc_col* const dst[4]) {
  ENTER();
  if (!tw || log < 3 || log > tw->root_log + 1) return fail(SC_EINVAL, "fold_circle_into_line: bad log size / twiddle tree too small");
  const uint32_t* s[4]; uint32_t* d[4];
  for (int k = 0; k < 4; k++) {
    if (!src[k] || !dst[k] || src[k]->len != (1ull << log) || dst[k]->len != (1ull << (log - 1))) return fail(SC_EINVAL, "fold_circle_into_line: bad columns");
    s[k] = src[k]->d; d[k] = dst[k]->d;
  }
  { ProfScope ps_(ctx, "fold_circle_into_line"); CKL(launch_fold_circle_into_line(s, log, q_make(alpha[0], alpha[1], alpha[2], alpha[3]), d, tw->itw + ((size_t)1 << tw->root_log), ctx->st)); }
  return SC_OK;
}

// ------------------------------------------------------------------ quotients
int32_t sc_accumulate_quotients_range(sc_ctx* ctx, uint32_t log, uint64_t row_off, uint64_t n_rows, sc_col* const* cols, uint32_t n,
                                      const uint32_t random_coeff[4], const uint32_t* batch_points, const uint32_t* batch_sizes,
                                      const uint32_t* entry_cols, const uint32_t* entry_vals, uint32_t nb, sc_col* out[4]) {
  ENTER();
  if (log < 2 || log > 30 || (!cols && n) || !out || (row_off & 3) || (n_rows & 3) || row_off + n_rows > (1ull << log))
    return fail(SC_EINVAL, "accumulate_quotients: bad argument");
  std::vector<const uint32_t*> p(n);
  for (uint32_t i = 0; i < n; i++) {
    if (!cols[i] || cols[i]->len != n_rows) return fail(SC_EINVAL, "accumulate_quotients: column length != number of rows");
    p[i] = cols[i]->d;
  }
  QM31 alpha = q_make(random_coeff[0], random_coeff[1], random_coeff[2], random_coeff[3]);
  std::vector<QuotBatch> qb(nb);
  std::vector<QuotEntry> qe;
  size_t e = 0;
  for (uint32_t b = 0; b < nb; b++) {
    const uint32_t* q = batch_points + 8 * b;
    QM31 sx = q_make(q[0], q[1], q[2], q[3]), sy = q_make(q[4], q[5], q[6], q[7]);
    QuotBatch& B = qb[b];
    B.prx = sx.a; B.pix = sx.b; B.pry = sy.a; B.piy = sy.b;
    B.c0 = c_sub(c_mul(B.prx, B.piy), c_mul(B.pry, B.pix));
    B.suma = q_zero(); B.sumb = q_zero(); B.first = (uint32_t)qe.size(); B.count = batch_sizes[b];
    QM31 al = q_fromm(1);
    QM31 c = q_sub(q_conj(sy), sy);
    for (uint32_t j = 0; j < batch_sizes[b]; j++, e++) {
      if (entry_cols[e] >= n) return fail(SC_EINVAL, "accumulate_quotients: column index out of range");
      al = q_mul(al, alpha);
      QM31 v = q_make(entry_vals[4 * e], entry_vals[4 * e + 1], entry_vals[4 * e + 2], entry_vals[4 * e + 3]);
      QM31 a = q_sub(q_conj(v), v);
      QM31 bb = q_sub(q_mul(v, c), q_mul(a, sy));
      QM31 ac = q_mul(al, c);
      B.suma = q_add(B.suma, q_mul(al, a));
      B.sumb = q_add(B.sumb, q_mul(al, bb));
      qe.push_back(QuotEntry{entry_cols[e], {ac.a.a, ac.a.b, ac.b.a, ac.b.b}});
    }
    B.coeff = q_pow(alpha, batch_sizes[b]);
  }
  uint32_t* d[4];
  for (int k = 0; k < 4; k++) { int32_t r = new_col(ctx, n_rows, &out[k]); if (r) return r; d[k] = out[k]->d; }
  if (nb == 0) {
    for (int k = 0; k < 4; k++) CK(cudaMemsetAsync(d[k], 0, n_rows * 4, ctx->st));
    return SC_OK;
  }
  void *dp = nullptr, *db, *de = nullptr;
  int32_t r;
  if (n) { r = stage(ctx, p.data(), n * sizeof(void*), &dp); if (r) return r; }
  r = stage(ctx, qb.data(), qb.size() * sizeof(QuotBatch), &db); if (r) return r;
  if (!qe.empty()) { r = stage(ctx, qe.data(), qe.size() * sizeof(QuotEntry), &de); if (r) return r; }
  uint32_t* scratch = nullptr;
  size_t sw = quotients_scratch_words(log, row_off, n_rows);
  if (sw) CK(cudaMallocAsync((void**)&scratch, sw * 4, ctx->st));
  { ProfScope ps_(ctx, "accumulate_quotients"); CKL(launch_accumulate_quotients(log, row_off, n_rows, (const uint32_t* const*)dp, (const QuotBatch*)db, nb, (const QuotEntry*)de, d, ctx->st, scratch)); }
  if (scratch) CK(cudaFreeAsync(scratch, ctx->st));
  return SC_OK;
}
int32_t sc_accumulate_quotients(sc_ctx* ctx, uint32_t log, sc_col* const* cols, uint32_t n, const uint32_t random_coeff[4],
                                const uint32_t* batch_points, const uint32_t* batch_sizes, const uint32_t* entry_cols,
                                const uint32_t* entry_vals, uint32_t nb, sc_col* out[4]) {
  if (log > 30) return fail(SC_EINVAL, "accumulate_quotients: bad argument");
  return sc_accumulate_quotients_range(ctx, log, 0, 1ull << log, cols, n, random_coeff, batch_points, batch_sizes, entry_cols,
                                       entry_vals, nb, out);
}

// ------------------------------------------------------------------ accumulation / grind / misc
int32_t sc_accumulate(sc_ctx* ctx, sc_col* const dst[4], sc_col* const src[4]) {
  ENTER();
  uint32_t* d[4]; const uint32_t* s[4];
  for (int k = 0; k < 4; k++) {
    if (!dst[k] || !src[k] || dst[k]->len != dst[0]->len || src[k]->len != dst[0]->len) return fail(SC_EINVAL, "accumulate: bad columns");
    d[k] = dst[k]->d; s[k] = src[k]->d;
  }
  { ProfScope ps_(ctx, "accumulate"); CKL(launch_accumulate(d, s, dst[0]->len, ctx->st)); }
  return SC_OK;
}
int32_t sc_secure_powers(const uint32_t felt[4], uint32_t n, uint32_t* out) {
  if (!felt || (!out && n)) return fail(SC_EINVAL, "null argument");
  QM31 f = q_make(felt[0], felt[1], felt[2], felt[3]), acc = q_fromm(1);
  for (uint32_t i = 0; i < n; i++) {
    out[4 * i] = acc.a.a; out[4 * i + 1] = acc.a.b; out[4 * i + 2] = acc.b.a; out[4 * i + 3] = acc.b.b;
    acc = q_mul(acc, f);
  }
  return SC_OK;
}
// Integer-pipe micro-benchmark (microbench.cu): the measured peak the Merkle roofline is reported against.
int32_t sc_microbench_int(sc_ctx* ctx, int32_t kind, uint32_t iters, double out[4]) {
  ENTER();
  if (!out || kind < 0 || kind > 6 || iters == 0) return fail(SC_EINVAL, "microbench_int: bad argument");
  int n_sm = 0;
  CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx->device));
  void* scratch = nullptr;
  CK(cudaMalloc(&scratch, 16 * (size_t)n_sm + 64));
  int e = launch_int_pipe_bench(kind, iters, n_sm, scratch, out, ctx->st);
  cudaFree(scratch);
  CKL(e);
  return SC_OK;
}
int32_t sc_grind(sc_ctx* ctx, const uint32_t digest[8], uint32_t pow_bits, uint64_t* nonce_out) {
  ENTER();
  if (!digest || !nonce_out || pow_bits > 64) return fail(SC_EINVAL, "grind: bad argument");
  unsigned long long* dres;
  CK(cudaMallocAsync((void**)&dres, 8, ctx->st));
  CK(cudaMemsetAsync(dres, 0xff, 8, ctx->st));
  int e = launch_grind(digest, pow_bits, dres, ctx->st);
  unsigned long long r = 0;
  if (e == 0) CK(cudaMemcpyAsync(&r, dres, 8, cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  CK(cudaFreeAsync(dres, ctx->st));
  if (e > 0) { ctx->poisoned = true; return fail(SC_ECUDA, "grind kernel failed"); }
  if (e < 0) return fail(SC_EINVAL, "grind: no nonce found below 2^40");
  *nonce_out = r;
  return SC_OK;
}
int32_t sc_gen_is_first(sc_ctx* ctx, uint32_t log_size, sc_col** out) {
  ENTER();
  if (!out || log_size > 30) return fail(SC_EINVAL, "gen_is_first: bad argument");
  int32_t r = new_col(ctx, 1ull << log_size, out);
  if (r) return r;
  { ProfScope ps_(ctx, "gen_is_first"); CKL(launch_gen_is_first((*out)->d, log_size, ctx->st)); }
  return SC_OK;
}
// The polynomial of gen_is_first(log_size), i.e. sc_gen_is_first followed by sc_interpolate, computed in closed form.
int32_t sc_is_first_coeffs(sc_ctx* ctx, uint32_t log_size, const sc_twiddles* tw, sc_col** out) {
  ENTER();
  if (!out || !tw || log_size < 3) return fail(SC_EINVAL, "is_first_coeffs: bad argument");
  if (log_size > tw->root_log + 1) return fail(SC_EINVAL, "is_first_coeffs: twiddle tree too small for this domain");
  int32_t r = new_col(ctx, 1ull << log_size, out);
  if (r) return r;
  { ProfScope ps_(ctx, "gen_is_first"); CKL(launch_is_first_coeffs((*out)->d, log_size, tw->itw + ((size_t)1 << tw->root_log), ctx->st)); }
  return SC_OK;
}
// Rows [row_off, row_off + n_rows) of the IsFirst column of log_size extended to log_size + log_blowup, written in closed form
// (quotients.cu is_first_lde_kernel): what sc_evaluate gives for sc_is_first_coeffs, without the transform.
int32_t sc_is_first_lde(sc_ctx* ctx, uint32_t log_size, uint32_t log_blowup, const sc_twiddles* tw, uint64_t row_off, uint64_t n_rows, sc_col** out) {
  ENTER();
  const uint32_t dom = log_size + log_blowup;
  if (!out || !tw || log_size < 3 || dom > 30 || (row_off & 3) || (n_rows & 3) || !n_rows || row_off + n_rows > (1ull << dom))
    return fail(SC_EINVAL, "is_first_lde: bad argument");
  if (log_size > tw->root_log + 1) return fail(SC_EINVAL, "is_first_lde: twiddle tree too small for this domain");
  int32_t r = new_col(ctx, n_rows, out);
  if (r) return r;
  uint32_t* scratch = nullptr;
  const size_t sw = quotients_scratch_words(dom, row_off, n_rows);
  if (sw) CK(cudaMallocAsync((void**)&scratch, sw * 4, ctx->st));
  { ProfScope ps_(ctx, "gen_is_first"); CKL(launch_is_first_lde((*out)->d, log_size, dom, row_off, n_rows, tw->itw + ((size_t)1 << tw->root_log), ctx->st, scratch)); }
  if (scratch) CK(cudaFreeAsync(scratch, ctx->st));
  return SC_OK;
}
static inline size_t prefix_scratch_words(uint64_t len) { return ((len + 2 * ((len >> 11) + 2) + 8) + 3) & ~(size_t)3; }
int32_t sc_prefix_sum_bitrev(sc_ctx* ctx, sc_col* col) {
  ENTER();
  if (!col || !is_pow2(col->len) || col->len < 2) return fail(SC_EINVAL, "prefix_sum: length must be a power of two >= 2");
  uint32_t lg = ilog2(col->len);
  uint32_t* scratch;
  if (lg >= 12) {
    CK(cudaMallocAsync((void**)&scratch, prefix_sum_tiled_words(lg) * 4, ctx->st));
    uint32_t* v1[1] = {col->d};
    { ProfScope ps_(ctx, "prefix_sum"); CKL(launch_prefix_sum_bitrev_tiled(v1, 1, lg, scratch, ctx->st)); }
  } else {
    size_t words = prefix_scratch_words(col->len);
    CK(cudaMallocAsync((void**)&scratch, words * 4, ctx->st));
    { ProfScope ps_(ctx, "prefix_sum"); CKL(launch_prefix_sum_bitrev(col->d, lg, scratch, ctx->st)); }
  }
  CK(cudaFreeAsync(scratch, ctx->st));
  return SC_OK;
}

// Batched decommitment gather: out[i*words .. +words) = cols[i][offsets[i] .. +words).  One kernel + one D2H copy instead
// of one tiny copy per queried value (MerkleProver::decommit walks Column::at element by element upstream).
int32_t sc_gather(sc_ctx* ctx, sc_col* const* cols, const uint64_t* offsets, uint32_t n, uint32_t words, uint32_t* out_host) {
  ENTER();
  if (!n) return SC_OK;
  if (!cols || !offsets || !out_host || !words) return fail(SC_EINVAL, "gather: bad argument");
  std::vector<const uint32_t*> p(n);
  for (uint32_t i = 0; i < n; i++) {
    if (!cols[i] || offsets[i] + words > cols[i]->len) return fail(SC_EINVAL, "gather: out of range");
    p[i] = cols[i]->d + offsets[i];
  }
  void* dp;
  int32_t r = stage(ctx, p.data(), n * sizeof(void*), &dp); if (r) return r;
  uint32_t* dout;
  CK(cudaMallocAsync((void**)&dout, (size_t)n * words * 4, ctx->st));
  { ProfScope ps_(ctx, "gather"); CKL(launch_gather((const uint32_t* const*)dp, n, words, dout, ctx->st)); }
  CK(cudaMemcpyAsync(out_host, dout, (size_t)n * words * 4, cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  CK(cudaFreeAsync(dout, ctx->st));
  return SC_OK;
}

// ------------------------------------------------------------------ AIR layer (air_kernels.cu)
int32_t sc_logup_generate(sc_ctx* ctx, int32_t component, sc_col* const* main_cols, uint32_t n_main, uint32_t log_repeat,
                          const uint32_t* elements, sc_col** out, uint32_t claimed_sum[4]) {
  ENTER();
  if (component < 0 || component >= sbf::N_COMPONENTS || !main_cols || !elements || !out) return fail(SC_EINVAL, "logup_generate: bad argument");
  if ((int)n_main != sbf::N_MAIN_COLS[component]) return fail(SC_EINVAL, "logup_generate: wrong number of main columns");
  if (!main_cols[0] || log_repeat > 8) return fail(SC_EINVAL, "logup_generate: bad argument");
  uint64_t len = main_cols[0]->len << log_repeat;
  if (!is_pow2(len) || len < 16) return fail(SC_EINVAL, "logup_generate: column length must be a power of two >= 16");
  std::vector<const uint32_t*> mp(n_main);
  for (uint32_t i = 0; i < n_main; i++) { if (!main_cols[i] || (main_cols[i]->len << log_repeat) != len) return fail(SC_EINVAL, "logup_generate: column length mismatch"); mp[i] = main_cols[i]->d; }
  int nout = 4 * sbf::N_LOGUP_COLS[component];
  std::vector<uint32_t*> op(nout);
  for (int i = 0; i < nout; i++) { int32_t r = new_col(ctx, len, &out[i]); if (r) return r; op[i] = out[i]->d; }
  void *dm, *dout;
  int32_t r = stage(ctx, mp.data(), mp.size() * sizeof(void*), &dm); if (r) return r;
  r = stage(ctx, op.data(), op.size() * sizeof(void*), &dout); if (r) return r;
  AirParams p{};
  p.main = (const uint32_t* const*)dm; p.out = (uint32_t* const*)dout; p.log_size = ilog2(len); p.main_shift = log_repeat;
  memcpy(&p.el, elements, sizeof(p.el));
  { ProfScope ps_(ctx, "logup_generate"); CKL(launch_air(false, component, p, ctx->st)); }
  // LogupTraceGenerator::finalize_last: prefix-sum the last column's coordinates in coset order; claimed_sum = col.at(1)
  {
    uint32_t* scratch;
    uint32_t* v4[4] = {op[nout - 4], op[nout - 3], op[nout - 2], op[nout - 1]};
    if (p.log_size >= 12) {
      CK(cudaMallocAsync((void**)&scratch, 4 * prefix_sum_tiled_words(p.log_size) * 4, ctx->st));
      { ProfScope ps_(ctx, "prefix_sum"); CKL(launch_prefix_sum_bitrev_tiled(v4, 4, p.log_size, scratch, ctx->st)); }
    } else {
      size_t words = prefix_scratch_words(len);
      CK(cudaMallocAsync((void**)&scratch, 4 * words * 4, ctx->st));
      { ProfScope ps_(ctx, "prefix_sum"); CKL(launch_prefix_sum_bitrev4(v4, p.log_size, scratch, words, ctx->st)); }
    }
    CK(cudaFreeAsync(scratch, ctx->st));
  }
  // claimed_sum == NULL: the caller reads element 1 of the last four columns itself (e.g. one sc_gather for all components)
  if (claimed_sum) for (int k = 0; k < 4; k++) { r = sc_col_read(ctx, out[nout - 4 + k], 1, 1, &claimed_sum[k]); if (r) return r; }
  return SC_OK;
}

int32_t sc_eval_constraints(sc_ctx* ctx, int32_t component, uint32_t log_size, sc_col* const* main_lde, uint32_t n_main,
                            sc_col* const* inter_lde, uint32_t n_inter, const sc_col* is_first_lde, const uint32_t* elements,
                            const uint32_t total_sum[4], const uint32_t* coeffs, sc_col* const accum[4]) {
  ENTER();
  if (component < 0 || component >= sbf::N_COMPONENTS || !main_lde || !inter_lde || !is_first_lde || !elements || !coeffs || !accum)
    return fail(SC_EINVAL, "eval_constraints: bad argument");
  if ((int)n_main != sbf::N_MAIN_COLS[component] || (int)n_inter != 4 * sbf::N_LOGUP_COLS[component])
    return fail(SC_EINVAL, "eval_constraints: wrong number of columns");
  uint64_t len = 2ull << log_size;
  std::vector<const uint32_t*> mp(n_main), ip(n_inter);
  for (uint32_t i = 0; i < n_main; i++) { if (!main_lde[i] || main_lde[i]->len != len) return fail(SC_EINVAL, "eval_constraints: main column length"); mp[i] = main_lde[i]->d; }
  for (uint32_t i = 0; i < n_inter; i++) { if (!inter_lde[i] || inter_lde[i]->len != len) return fail(SC_EINVAL, "eval_constraints: interaction column length"); ip[i] = inter_lde[i]->d; }
  if (is_first_lde->len != len) return fail(SC_EINVAL, "eval_constraints: is_first column length");
  for (int k = 0; k < 4; k++) if (!accum[k] || accum[k]->len != len) return fail(SC_EINVAL, "eval_constraints: accumulator length");
  void *dm, *di, *dc;
  int32_t r = stage(ctx, mp.data(), mp.size() * sizeof(void*), &dm); if (r) return r;
  r = stage(ctx, ip.data(), ip.size() * sizeof(void*), &di); if (r) return r;
  r = stage(ctx, coeffs, (size_t)sbf::N_CONSTRAINTS[component] * 16, &dc); if (r) return r;
  AirParams p{};
  p.main = (const uint32_t* const*)dm; p.inter = (const uint32_t* const*)di; p.is_first = is_first_lde->d;
  p.coeff = (const QM31*)dc; p.log_size = log_size;
  memcpy(&p.el, elements, sizeof(p.el));
  p.total_sum = q_make(total_sum[0], total_sum[1], total_sum[2], total_sum[3]);
  vanishing_denom_inv(log_size, p.denom_inv);
  for (int k = 0; k < 4; k++) p.acc[k] = accum[k]->d;
  { ProfScope ps_(ctx, "eval_constraints"); CKL(launch_air(true, component, p, ctx->st)); }
  return SC_OK;
}

}  // extern "C"
