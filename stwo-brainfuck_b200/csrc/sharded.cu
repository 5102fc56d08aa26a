// Multi-GPU entry points of the C ABI: row-range variants of the row-local Backend ops (a rank works on rows
// [row_off, row_off + n) of a bit-reversed domain), the helpers the sharded prover needs (pre-shifted LogUp column,
// selective LogUp outputs, single-column accumulate, device copies) and the NCCL collectives of the exchange steps
// (all-to-all column->row re-sharding, sub-root all-gather, small all-reduces).  BASELINE.json north_star: "split by
// column for interpolation and extension, then re-sharded by row range over NVLink (NCCL all-to-all) for Merkle leaf
// hashing, quotient evaluation and FRI folding".  NCCL is resolved with dlopen so that the library also loads on a box
// without it (single-GPU use); the soname is the one torch ships, so both share one copy.
#include <cstdlib>
#include <dlfcn.h>
#include <nccl.h>

#include "air_params.cuh"
#include "capi_internal.cuh"

namespace sb {

struct Ptr4 { uint32_t* p[4]; };
struct CPtr4 { const uint32_t* p[4]; };

// ---- FRI folds on a row range: outputs [off, off + n) of the folded layer (twiddles are indexed by the global row)
// alpha_p != NULL: the coefficient is read from device memory (sc_dchan: the transcript runs on the device)
__global__ void fold_line_range_kernel(CPtr4 s, Ptr4 d, uint32_t log, QM31 alpha, const uint32_t* __restrict__ itw_end, size_t off, size_t n,
                                       const uint32_t* __restrict__ alpha_p = nullptr) {
  const uint32_t* itw = itw_end - ((size_t)1 << log);
  if (alpha_p) alpha = q_make(alpha_p[0], alpha_p[1], alpha_p[2], alpha_p[3]);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint2 c0 = reinterpret_cast<const uint2*>(s.p[0])[i], c1 = reinterpret_cast<const uint2*>(s.p[1])[i];
    uint2 c2 = reinterpret_cast<const uint2*>(s.p[2])[i], c3 = reinterpret_cast<const uint2*>(s.p[3])[i];
    QM31 a = q_make(c0.x, c1.x, c2.x, c3.x), b = q_make(c0.y, c1.y, c2.y, c3.y);
    QM31 f0 = q_add(a, b), f1 = q_mulm(q_sub(a, b), __ldg(itw + off + i));
    QM31 r = q_add(f0, q_mul(alpha, f1));
    d.p[0][i] = r.a.a; d.p[1][i] = r.a.b; d.p[2][i] = r.b.a; d.p[3][i] = r.b.b;
  }
}
__global__ void fold_circle_range_kernel(CPtr4 s, Ptr4 d, uint32_t log, QM31 alpha, QM31 alpha_sq, const uint32_t* __restrict__ itw_end,
                                         size_t off, size_t n, const uint32_t* __restrict__ alpha_p = nullptr) {
  const uint32_t* l1 = itw_end - ((size_t)1 << (log - 1));
  if (alpha_p) { alpha = q_make(alpha_p[0], alpha_p[1], alpha_p[2], alpha_p[3]); alpha_sq = q_mul(alpha, alpha); }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    size_t gi = off + i;
    size_t pair = (gi >> 2) * 2;
    uint32_t x = __ldg(l1 + pair), y = __ldg(l1 + pair + 1);
    uint32_t sel = (uint32_t)gi & 3u;
    uint32_t t = sel < 2 ? y : x;
    if (sel == 1 || sel == 2) t = P - t;
    uint2 c0 = reinterpret_cast<const uint2*>(s.p[0])[i], c1 = reinterpret_cast<const uint2*>(s.p[1])[i];
    uint2 c2 = reinterpret_cast<const uint2*>(s.p[2])[i], c3 = reinterpret_cast<const uint2*>(s.p[3])[i];
    QM31 a = q_make(c0.x, c1.x, c2.x, c3.x), b = q_make(c0.y, c1.y, c2.y, c3.y);
    QM31 f0 = q_add(a, b), f1 = q_mulm(q_sub(a, b), t);
    QM31 acc = q_make(d.p[0][i], d.p[1][i], d.p[2][i], d.p[3][i]);
    QM31 r = q_add(q_mul(acc, alpha_sq), q_add(f0, q_mul(alpha, f1)));
    d.p[0][i] = r.a.a; d.p[1][i] = r.a.b; d.p[2][i] = r.b.a; d.p[3][i] = r.b.b;
  }
}
// Send buffer of a column->row exchange in ONE launch: block (d, j) of the buffer = rows [d * seg_j, (d + 1) * seg_j) of owned
// column j (the whole column when it is replicated), blocks laid out destination-major.  Replaces world x columns small copies.
struct PackCol { const uint32_t* src; uint32_t seg; uint32_t sharded; uint64_t off; };   // off: word offset inside a destination's block
__global__ void __launch_bounds__(256) pack_exchange_kernel(const PackCol* __restrict__ cols, uint32_t ncols, uint64_t per_dest,
                                                            uint32_t* __restrict__ send) {
  const PackCol c = cols[blockIdx.y];
  const uint32_t d = blockIdx.z;
  const uint32_t* src = c.src + (c.sharded ? (size_t)d * c.seg : 0);
  uint32_t* dst = send + (size_t)d * per_dest + c.off;
  if (((c.seg | c.off | per_dest) & 3u) == 0) {   // everything 16-byte aligned: vector copies
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < c.seg / 4; i += gridDim.x * blockDim.x) d4[i] = s4[i];
  } else {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < c.seg; i += gridDim.x * blockDim.x) dst[i] = src[i];
  }
}
// The same blocks written where they are needed: block (d, j) goes straight into rank d's receive window over NVLink
// (peer-mapped device memory, sc_exchange_begin), at the offset this rank's contribution has in every rank's window.  The
// pack pass, the send buffer and NCCL's send/recv all-to-all (243 GB/s per direction between two B200s, measured) collapse
// into one kernel of peer stores; a one-word all-reduce behind it tells every rank that its window is complete.
__global__ void __launch_bounds__(256) push_exchange_kernel(const PackCol* __restrict__ cols, uint32_t ncols, uint32_t* const* __restrict__ peer,
                                                            uint64_t win_off) {
  const PackCol c = cols[blockIdx.y];
  const uint32_t d = blockIdx.z;
  const uint32_t* src = c.src + (c.sharded ? (size_t)d * c.seg : 0);
  uint32_t* dst = peer[d] + win_off + c.off;
  if (((c.seg | c.off | win_off) & 3u) == 0) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < c.seg / 4; i += gridDim.x * blockDim.x) d4[i] = s4[i];
  } else {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < c.seg; i += gridDim.x * blockDim.x) dst[i] = src[i];
  }
}
// Row ranges -> whole columns on their owners (the composition accumulators): piece j of this rank goes, whole, to ONE rank's
// window at a caller-given offset.  Same windows, same completion barrier as push_exchange_kernel.
struct ScatterCol { const uint32_t* src; uint32_t len; uint32_t dest; uint64_t off; };
__global__ void __launch_bounds__(256) scatter_exchange_kernel(const ScatterCol* __restrict__ cols, uint32_t* const* __restrict__ peer, uint64_t win_off) {
  const ScatterCol c = cols[blockIdx.y];
  uint32_t* dst = peer[c.dest] + win_off + c.off;
  if (((c.len | c.off | win_off) & 3u) == 0) {
    const uint4* s4 = reinterpret_cast<const uint4*>(c.src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < c.len / 4; i += gridDim.x * blockDim.x) d4[i] = s4[i];
  } else {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < c.len; i += gridDim.x * blockDim.x) dst[i] = c.src[i];
  }
}
// out[row] = col[storage index of the coset-order predecessor of row]  (offset_bit_reversed_circle_domain_index(.., -1))
__global__ void shift_prev_kernel(const uint32_t* __restrict__ col, uint32_t* __restrict__ out, uint32_t e) {
  uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= (1u << e)) return;
  uint32_t idx = __brev(row) >> (32 - e), half = 1u << (e - 1);
  uint32_t pidx = idx < half ? ((idx + half - 1) & (half - 1)) : (((idx - half + 1) & (half - 1)) + half);
  out[row] = col[__brev(pidx) >> (32 - e)];
}
__global__ void accumulate_col_kernel(uint32_t* __restrict__ d, const uint32_t* __restrict__ s, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = m_add(d[i], s[i]);
}
static inline unsigned grid_for(size_t n) {
  size_t b = (n + 255) / 256;
  return (unsigned)(b < 1 ? 1 : (b > 148u * 16u ? 148u * 16u : b));
}

// ---- NCCL through dlopen
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  const char* (*GetErrorString)(ncclResult_t);
};
static NcclApi* nccl() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return nullptr;
#define SB_SYM(name) *(void**)(&api.name) = dlsym(h, "nccl" #name); if (!api.name) return nullptr;
    SB_SYM(GetUniqueId) SB_SYM(CommInitRank) SB_SYM(CommDestroy) SB_SYM(GroupStart) SB_SYM(GroupEnd) SB_SYM(Send) SB_SYM(Recv)
    SB_SYM(AllGather) SB_SYM(AllReduce) SB_SYM(GetErrorString)
#undef SB_SYM
    api.h = h;
  }
  return api.h ? &api : nullptr;
}

}  // namespace sb

struct sc_comm {
  ncclComm_t comm;
  int rank, world;
  // receive window of the direct (peer-store) column->row exchange: one cudaMalloc'ed buffer per rank, exported with CUDA IPC
  // and mapped by every other rank; regions are bump-allocated per proof in the same order on every rank (the sizes are the
  // same everywhere), so a region has the same offset in every window.
  uint32_t* win = nullptr;
  size_t win_cap = 0, win_off = 0, win_need = 0, win_want = 0;   // words
  std::vector<uint32_t*> peer;      // peer[d] = rank d's window as mapped here (peer[rank] = win)
  uint32_t** d_peer = nullptr;      // the same table on the device
  uint32_t* d_flag = nullptr;       // one word for the completion all-reduce
  bool mapped = false, push_ok = false;
};
#define CKN(call)                                                                                                   \
  do {                                                                                                              \
    ncclResult_t r_ = (call);                                                                                       \
    if (r_ != ncclSuccess) { ctx->poisoned = true; return fail(SC_ECUDA, std::string(#call) + ": " + nccl()->GetErrorString(r_)); } \
  } while (0)

extern "C" {

// ------------------------------------------------------------------ device copies / views
int32_t sc_col_copy(sc_ctx* ctx, sc_col* dst, uint64_t dst_off, const sc_col* src, uint64_t src_off, uint64_t n) {
  ENTER();
  if (!dst || !src || dst_off + n > dst->len || src_off + n > src->len) return fail(SC_EINVAL, "col_copy: out of range");
  if (n) CK(cudaMemcpyAsync(dst->d + dst_off, src->d + src_off, n * 4, cudaMemcpyDeviceToDevice, ctx->st));
  return SC_OK;
}
int32_t sc_col_view(sc_ctx* ctx, sc_col* col, uint64_t off, uint64_t n, sc_col** out) {
  ENTER();
  // no alignment requirement here: Merkle column reads are scalar; the vectorised kernels (FFT, folds, quotients) are only
  // ever given views at multiples of 16 rows by the sharded prover
  if (!col || !out || off + n > col->len) return fail(SC_EINVAL, "col_view: out of range");
  sc_col* c = new sc_col{col->d + off, n};
  c->owned = false;
  track(ctx, c);
  *out = c;
  return SC_OK;
}

// send[d * per_dest + off_j .. + seg_j) = cols[j][(sharded_j ? d * seg_j : 0) .. + seg_j) for every destination d < world
int32_t sc_pack_exchange(sc_ctx* ctx, sc_col* const* cols, const uint64_t* segs, const uint8_t* sharded, uint32_t n, uint32_t world, sc_col* send) {
  ENTER();
  if (!send || (n && (!cols || !segs || !sharded)) || !world) return fail(SC_EINVAL, "pack_exchange: null argument");
  if (!n) return SC_OK;
  std::vector<PackCol> pc(n);
  uint64_t off = 0;
  uint32_t max_seg = 1;
  for (uint32_t j = 0; j < n; j++) {
    if (!cols[j] || segs[j] > 0xffffffffull || cols[j]->len < (sharded[j] ? segs[j] * world : segs[j])) return fail(SC_EINVAL, "pack_exchange: bad column");
    pc[j] = {cols[j]->d, (uint32_t)segs[j], sharded[j] ? 1u : 0u, off};
    off += segs[j];
    max_seg = std::max<uint32_t>(max_seg, (uint32_t)segs[j]);
  }
  if (send->len < off * world) return fail(SC_EINVAL, "pack_exchange: send buffer too small");
  void* d_pc = nullptr;
  { int32_t r = stage(ctx, pc.data(), pc.size() * sizeof(PackCol), &d_pc); if (r) return r; }
  const uint32_t bx = std::max(1u, std::min(64u, (max_seg / 4 + 255) / 256));
  { ProfScope ps_(ctx, "pack_exchange");
    pack_exchange_kernel<<<dim3(bx, n, world), 256, 0, ctx->st>>>((const PackCol*)d_pc, n, off, send->d);
    g_launch_count++; CK(cudaGetLastError()); }
  return SC_OK;
}

// ------------------------------------------------------------------ row-range FRI folds
int32_t sc_fold_line_range(sc_ctx* ctx, sc_col* const src[4], uint32_t log, uint64_t out_off, uint64_t n_out, const uint32_t alpha[4],
                           const sc_twiddles* tw, sc_col* dst_out[4]) {
  ENTER();
  if (!tw || log < 1 || log > tw->root_log || out_off + n_out > (1ull << (log - 1))) return fail(SC_EINVAL, "fold_line_range: bad argument");
  CPtr4 s; Ptr4 d;
  for (int k = 0; k < 4; k++) { if (!src[k] || src[k]->len != 2 * n_out) return fail(SC_EINVAL, "fold_line_range: bad source"); s.p[k] = src[k]->d; }
  for (int k = 0; k < 4; k++) { int32_t r = new_col(ctx, n_out, &dst_out[k]); if (r) return r; d.p[k] = dst_out[k]->d; }
  { ProfScope ps_(ctx, "fold_line");
    fold_line_range_kernel<<<grid_for(n_out), 256, 0, ctx->st>>>(s, d, log, q_make(alpha[0], alpha[1], alpha[2], alpha[3]),
                                                                tw->itw + ((size_t)1 << tw->root_log), out_off, n_out);
    g_launch_count++; CK(cudaGetLastError()); }
  return SC_OK;
}
int32_t sc_fold_circle_into_line_range(sc_ctx* ctx, sc_col* const src[4], uint32_t log, uint64_t out_off, uint64_t n_out,
                                       const uint32_t alpha[4], const sc_twiddles* tw, sc_col* const dst[4]) {
  ENTER();
  if (!tw || log < 3 || log > tw->root_log + 1 || out_off + n_out > (1ull << (log - 1))) return fail(SC_EINVAL, "fold_circle_range: bad argument");
  CPtr4 s; Ptr4 d;
  for (int k = 0; k < 4; k++) {
    if (!src[k] || !dst[k] || src[k]->len != 2 * n_out || dst[k]->len != n_out) return fail(SC_EINVAL, "fold_circle_range: bad columns");
    s.p[k] = src[k]->d; d.p[k] = dst[k]->d;
  }
  QM31 a = q_make(alpha[0], alpha[1], alpha[2], alpha[3]);
  { ProfScope ps_(ctx, "fold_circle_into_line");
    fold_circle_range_kernel<<<grid_for(n_out), 256, 0, ctx->st>>>(s, d, log, a, q_mul(a, a), tw->itw + ((size_t)1 << tw->root_log), out_off, n_out);
    g_launch_count++; CK(cudaGetLastError()); }
  return SC_OK;
}

// The same folds with the coefficient taken from a device-resident transcript (sc_dchan): coefficient #k of `dc`.
int32_t sc_fold_line_range_dc(sc_ctx* ctx, sc_col* const src[4], uint32_t log, uint64_t out_off, uint64_t n_out, const sc_dchan* dc, uint32_t k,
                              const sc_twiddles* tw, sc_col* dst_out[4]) {
  ENTER();
  const uint32_t* ap = sc_dchan_coeff_ptr(dc, k);
  if (!tw || !ap || log < 1 || log > tw->root_log || out_off + n_out > (1ull << (log - 1))) return fail(SC_EINVAL, "fold_line_range_dc: bad argument");
  CPtr4 s; Ptr4 d;
  for (int j = 0; j < 4; j++) { if (!src[j] || src[j]->len != 2 * n_out) return fail(SC_EINVAL, "fold_line_range_dc: bad source"); s.p[j] = src[j]->d; }
  for (int j = 0; j < 4; j++) { int32_t r = new_col(ctx, n_out, &dst_out[j]); if (r) return r; d.p[j] = dst_out[j]->d; }
  { ProfScope ps_(ctx, "fold_line");
    fold_line_range_kernel<<<grid_for(n_out), 256, 0, ctx->st>>>(s, d, log, q_make(0, 0, 0, 0), tw->itw + ((size_t)1 << tw->root_log), out_off, n_out, ap);
    g_launch_count++; CK(cudaGetLastError()); }
  return SC_OK;
}
int32_t sc_fold_circle_into_line_range_dc(sc_ctx* ctx, sc_col* const src[4], uint32_t log, uint64_t out_off, uint64_t n_out, const sc_dchan* dc,
                                          uint32_t k, const sc_twiddles* tw, sc_col* const dst[4]) {
  ENTER();
  const uint32_t* ap = sc_dchan_coeff_ptr(dc, k);
  if (!tw || !ap || log < 3 || log > tw->root_log + 1 || out_off + n_out > (1ull << (log - 1))) return fail(SC_EINVAL, "fold_circle_range_dc: bad argument");
  CPtr4 s; Ptr4 d;
  for (int j = 0; j < 4; j++) {
    if (!src[j] || !dst[j] || src[j]->len != 2 * n_out || dst[j]->len != n_out) return fail(SC_EINVAL, "fold_circle_range_dc: bad columns");
    s.p[j] = src[j]->d; d.p[j] = dst[j]->d;
  }
  { ProfScope ps_(ctx, "fold_circle_into_line");
    fold_circle_range_kernel<<<grid_for(n_out), 256, 0, ctx->st>>>(s, d, log, q_make(0, 0, 0, 0), q_make(0, 0, 0, 0), tw->itw + ((size_t)1 << tw->root_log),
                                                                  out_off, n_out, ap);
    g_launch_count++; CK(cudaGetLastError()); }
  return SC_OK;
}

// ------------------------------------------------------------------ AIR helpers
// The LDE of a LogUp cumulative-sum column read at coset offset -1, as a column of its own (so that it can be re-sharded by
// rows like any other column).  col: LDE on CanonicCoset(trace_log + 1).
int32_t sc_shift_prev(sc_ctx* ctx, const sc_col* col, uint32_t trace_log, sc_col** out) {
  ENTER();
  if (!col || !out || col->len != (2ull << trace_log)) return fail(SC_EINVAL, "shift_prev: bad column");
  int32_t r = new_col(ctx, col->len, out);
  if (r) return r;
  uint32_t n = (uint32_t)col->len;
  shift_prev_kernel<<<(n + 255) / 256, 256, 0, ctx->st>>>(col->d, (*out)->d, trace_log + 1);
  g_launch_count++;
  CK(cudaGetLastError());
  return SC_OK;
}
int32_t sc_accumulate_col(sc_ctx* ctx, sc_col* dst, const sc_col* src) {
  ENTER();
  if (!dst || !src || dst->len != src->len) return fail(SC_EINVAL, "accumulate_col: length mismatch");
  { ProfScope ps_(ctx, "accumulate");
    accumulate_col_kernel<<<grid_for(dst->len), 256, 0, ctx->st>>>(dst->d, src->d, dst->len);
    g_launch_count++; CK(cudaGetLastError()); }
  return SC_OK;
}
// LogUp generation with selective outputs: want[i] != 0 -> out[i] receives a new column, else out[i] = NULL.  No prefix sum.
int32_t sc_logup_generate_sel(sc_ctx* ctx, int32_t component, sc_col* const* main_cols, uint32_t n_main, uint32_t log_repeat,
                              const uint32_t* elements, const uint8_t* want, sc_col** out) {
  ENTER();
  if (component < 0 || component >= sbf::N_COMPONENTS || !main_cols || !elements || !out || !want || log_repeat > 8)
    return fail(SC_EINVAL, "logup_generate_sel: bad argument");
  if ((int)n_main != sbf::N_MAIN_COLS[component]) return fail(SC_EINVAL, "logup_generate_sel: wrong number of main columns");
  uint64_t len = main_cols[0]->len << log_repeat;
  std::vector<const uint32_t*> mp(n_main);
  for (uint32_t i = 0; i < n_main; i++) { if (!main_cols[i] || (main_cols[i]->len << log_repeat) != len) return fail(SC_EINVAL, "logup_generate_sel: column length"); mp[i] = main_cols[i]->d; }
  int nout = 4 * sbf::N_LOGUP_COLS[component];
  std::vector<uint32_t*> op(nout, nullptr);
  for (int i = 0; i < nout; i++) {
    out[i] = nullptr;
    if (want[i]) { int32_t r = new_col(ctx, len, &out[i]); if (r) return r; op[i] = out[i]->d; }
  }
  void *dm, *dout;
  int32_t r = stage(ctx, mp.data(), mp.size() * sizeof(void*), &dm); if (r) return r;
  r = stage(ctx, op.data(), op.size() * sizeof(void*), &dout); if (r) return r;
  AirParams p{};
  p.main = (const uint32_t* const*)dm; p.out = (uint32_t* const*)dout; p.log_size = ilog2(len); p.main_shift = log_repeat;
  memcpy(&p.el, elements, sizeof(p.el));
  { ProfScope ps_(ctx, "logup_generate"); CKL(launch_air(false, component, p, ctx->st)); }
  return SC_OK;
}
// Constraint quotients on rows [row_off, row_off + n_rows) of the LDE.  prev: the 4 coordinates of the last LogUp column
// shifted by sc_shift_prev (same rows).  Column handles cover exactly the row range.
int32_t sc_eval_constraints_range(sc_ctx* ctx, int32_t component, uint32_t log_size, uint64_t row_off, uint64_t n_rows,
                                  sc_col* const* main_lde, uint32_t n_main, sc_col* const* inter_lde, uint32_t n_inter,
                                  sc_col* const prev[4], const sc_col* is_first_lde, const uint32_t* elements,
                                  const uint32_t total_sum[4], const uint32_t* coeffs, sc_col* const accum[4]) {
  ENTER();
  if (component < 0 || component >= sbf::N_COMPONENTS || !main_lde || !inter_lde || !prev || !is_first_lde || !elements || !coeffs || !accum)
    return fail(SC_EINVAL, "eval_constraints_range: bad argument");
  if ((int)n_main != sbf::N_MAIN_COLS[component] || (int)n_inter != 4 * sbf::N_LOGUP_COLS[component] || row_off + n_rows > (2ull << log_size))
    return fail(SC_EINVAL, "eval_constraints_range: wrong shape");
  std::vector<const uint32_t*> mp(n_main), ip(n_inter);
  for (uint32_t i = 0; i < n_main; i++) { if (!main_lde[i] || main_lde[i]->len != n_rows) return fail(SC_EINVAL, "eval_constraints_range: main column length"); mp[i] = main_lde[i]->d; }
  for (uint32_t i = 0; i < n_inter; i++) { if (!inter_lde[i] || inter_lde[i]->len != n_rows) return fail(SC_EINVAL, "eval_constraints_range: interaction column length"); ip[i] = inter_lde[i]->d; }
  if (is_first_lde->len != n_rows) return fail(SC_EINVAL, "eval_constraints_range: is_first length");
  void *dm, *di, *dc;
  int32_t r = stage(ctx, mp.data(), mp.size() * sizeof(void*), &dm); if (r) return r;
  r = stage(ctx, ip.data(), ip.size() * sizeof(void*), &di); if (r) return r;
  r = stage(ctx, coeffs, (size_t)sbf::N_CONSTRAINTS[component] * 16, &dc); if (r) return r;
  AirParams p{};
  p.main = (const uint32_t* const*)dm; p.inter = (const uint32_t* const*)di; p.is_first = is_first_lde->d;
  p.coeff = (const QM31*)dc; p.log_size = log_size; p.row_off = (uint32_t)row_off; p.n_rows = (uint32_t)n_rows;
  memcpy(&p.el, elements, sizeof(p.el));
  p.total_sum = q_make(total_sum[0], total_sum[1], total_sum[2], total_sum[3]);
  vanishing_denom_inv(log_size, p.denom_inv);
  for (int k = 0; k < 4; k++) {
    if (!accum[k] || accum[k]->len != n_rows || !prev[k] || prev[k]->len != n_rows) return fail(SC_EINVAL, "eval_constraints_range: accumulator / prev length");
    p.acc[k] = accum[k]->d; p.prev[k] = prev[k]->d;
  }
  { ProfScope ps_(ctx, "eval_constraints"); CKL(launch_air(true, component, p, ctx->st)); }
  return SC_OK;
}

// ------------------------------------------------------------------ NCCL collectives
int32_t sc_comm_unique_id(uint8_t out[128]) {
  sc_ctx* ctx = nullptr; (void)ctx;
  if (!nccl()) return fail(SC_EINVAL, "NCCL (libnccl.so.2) is not available");
  ncclUniqueId id;
  if (nccl()->GetUniqueId(&id) != ncclSuccess) return fail(SC_ECUDA, "ncclGetUniqueId failed");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  memcpy(out, &id, 128);
  return SC_OK;
}
int32_t sc_comm_init(sc_ctx* ctx, int32_t rank, int32_t world, const uint8_t id[128], sc_comm** out) {
  ENTER();
  if (!out || !id || world < 1 || rank < 0 || rank >= world) return fail(SC_EINVAL, "comm_init: bad argument");
  if (!nccl()) return fail(SC_EINVAL, "NCCL (libnccl.so.2) is not available");
  ncclUniqueId uid;
  memcpy(&uid, id, 128);
  sc_comm* c = new sc_comm{nullptr, rank, world};
  CKN(nccl()->CommInitRank(&c->comm, world, uid, rank));
  *out = c;
  return SC_OK;
}
int32_t sc_comm_destroy(sc_ctx* ctx, sc_comm* c) {
  if (!c) return SC_OK;
  if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->st); }
  for (int d = 0; d < (int)c->peer.size(); d++) if (d != c->rank && c->peer[d]) cudaIpcCloseMemHandle(c->peer[d]);
  if (nccl()) nccl()->CommDestroy(c->comm);   // after every rank's unmapping was issued; the windows are freed below
  if (c->win) cudaFree(c->win);
  if (c->d_peer) cudaFree(c->d_peer);
  if (c->d_flag) cudaFree(c->d_flag);
  delete c;
  return SC_OK;
}
// all-to-all with per-peer counts (words): send block for peer d starts at sum(send_counts[0..d)), same for recv.
int32_t sc_all_to_all(sc_ctx* ctx, sc_comm* c, const sc_col* send, const uint64_t* send_counts, sc_col* recv, const uint64_t* recv_counts) {
  ENTER();
  if (!c || !send || !recv || !send_counts || !recv_counts) return fail(SC_EINVAL, "all_to_all: null argument");
  uint64_t so = 0, ro = 0;
  for (int d = 0; d < c->world; d++) { so += send_counts[d]; ro += recv_counts[d]; }
  if (so > send->len || ro > recv->len) return fail(SC_EINVAL, "all_to_all: counts exceed the buffers");
  ProfScope ps_(ctx, "nccl_all_to_all");
  so = ro = 0;
  CKN(nccl()->GroupStart());
  for (int d = 0; d < c->world; d++) {
    if (d == c->rank) {
      if (send_counts[d] != recv_counts[d]) return fail(SC_EINVAL, "all_to_all: self counts differ");
      if (send_counts[d]) CK(cudaMemcpyAsync(recv->d + ro, send->d + so, send_counts[d] * 4, cudaMemcpyDeviceToDevice, ctx->st));
    } else {
      if (send_counts[d]) CKN(nccl()->Send(send->d + so, send_counts[d], ncclUint32, d, c->comm, ctx->st));
      if (recv_counts[d]) CKN(nccl()->Recv(recv->d + ro, recv_counts[d], ncclUint32, d, c->comm, ctx->st));
    }
    so += send_counts[d]; ro += recv_counts[d];
  }
  CKN(nccl()->GroupEnd());
  return SC_OK;
}
int32_t sc_all_gather(sc_ctx* ctx, sc_comm* c, const sc_col* send, sc_col* recv, uint64_t n) {
  ENTER();
  if (!c || !send || !recv || send->len < n || recv->len < n * c->world) return fail(SC_EINVAL, "all_gather: bad buffers");
  ProfScope ps_(ctx, "nccl_all_gather");
  CKN(nccl()->AllGather(send->d, recv->d, n, ncclUint32, c->comm, ctx->st));
  return SC_OK;
}
// In-place sum of a small host table over the ranks (each slot has exactly one non-zero contributor, so a plain u32 sum is a
// gather); synchronous.
int32_t sc_allreduce_host_u32(sc_ctx* ctx, sc_comm* c, uint32_t* buf, uint64_t n) {
  ENTER();
  if (!c || (!buf && n)) return fail(SC_EINVAL, "allreduce: null argument");
  if (!n) return SC_OK;
  uint32_t* d;
  CK(cudaMallocAsync((void**)&d, n * 4, ctx->st));
  CK(cudaMemcpyAsync(d, buf, n * 4, cudaMemcpyHostToDevice, ctx->st));
  { ProfScope ps_(ctx, "nccl_all_reduce"); CKN(nccl()->AllReduce(d, d, n, ncclUint32, ncclSum, c->comm, ctx->st)); }
  CK(cudaMemcpyAsync(buf, d, n * 4, cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  CK(cudaFreeAsync(d, ctx->st));
  return SC_OK;
}
// ---- direct exchange (peer stores over NVLink), see push_exchange_kernel
// Once per proof, collective: start bump allocation in the window; when a rank wants a larger window than it has (the previous
// proof recorded what it would have needed) or nothing is mapped yet, every rank unmaps, reallocates, exports and re-maps.
// A steady-state call is one one-word all-reduce (which is also the guarantee that no rank still reads the previous proof's
// regions when the first peer store of this proof arrives: a rank enters it only after its previous proof has drained).
int32_t sc_exchange_begin(sc_ctx* ctx, sc_comm* c) {
  ENTER();
  if (!c) return fail(SC_EINVAL, "exchange_begin: null communicator");
  static const bool disabled = getenv("SC_NO_PUSH_EXCHANGE") != nullptr;
  c->win_off = 0;
  c->win_want = std::max(c->win_want, c->win_need);
  c->win_need = 0;
  if (disabled || c->world < 2) { c->push_ok = false; return SC_OK; }
  uint32_t flag = (!c->mapped || c->win_want > c->win_cap) ? 1u : 0u;
  { int32_t r = sc_allreduce_host_u32(ctx, c, &flag, 1); if (r) return r; }
  if (!flag) return SC_OK;
  // ---- (re)map: nobody may free a window that a peer still has mapped
  for (int d = 0; d < (int)c->peer.size(); d++) if (d != c->rank && c->peer[d]) cudaIpcCloseMemHandle(c->peer[d]);
  c->peer.assign((size_t)c->world, nullptr);
  c->mapped = false; c->push_ok = false;
  { uint32_t z = 0; int32_t r = sc_allreduce_host_u32(ctx, c, &z, 1); if (r) return r; }   // every rank has unmapped
  if (c->win_want > c->win_cap || !c->win) {
    if (c->win) CK(cudaFree(c->win));
    c->win = nullptr; c->win_cap = 0;
    const size_t want = std::max<size_t>(c->win_want + c->win_want / 16, (size_t)1 << 20);
    if (cudaMalloc((void**)&c->win, want * 4) == cudaSuccess) c->win_cap = want;
    else { cudaGetLastError(); c->win = nullptr; }
  }
  if (!c->d_peer) CK(cudaMalloc((void**)&c->d_peer, sizeof(uint32_t*) * (size_t)c->world));
  if (!c->d_flag) { CK(cudaMalloc((void**)&c->d_flag, 4)); CK(cudaMemsetAsync(c->d_flag, 0, 4, ctx->st)); }
  // handle (64 bytes) + capacity, all-gathered through device memory
  constexpr size_t REC = 20;   // words per rank: 16 (handle) + 2 (capacity) + 2 (padding)
  std::vector<uint32_t> mine(REC, 0), all(REC * (size_t)c->world, 0);
  uint32_t fails = 0;
  if (c->win) {
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, c->win) == cudaSuccess) { static_assert(sizeof(h) == 64, "IPC handle size"); memcpy(mine.data(), &h, 64); }
    else { cudaGetLastError(); fails = 1; }
    const uint64_t cap = c->win_cap;
    memcpy(mine.data() + 16, &cap, 8);
  } else fails = 1;
  {
    uint32_t* d;
    CK(cudaMallocAsync((void**)&d, (REC + REC * (size_t)c->world) * 4, ctx->st));
    CK(cudaMemcpyAsync(d, mine.data(), REC * 4, cudaMemcpyHostToDevice, ctx->st));
    CKN(nccl()->AllGather(d, d + REC, REC, ncclUint32, c->comm, ctx->st));
    CK(cudaMemcpyAsync(all.data(), d + REC, REC * (size_t)c->world * 4, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    CK(cudaFreeAsync(d, ctx->st));
  }
  for (int d = 0; d < c->world && !fails; d++) {
    uint64_t cap;
    memcpy(&cap, all.data() + REC * (size_t)d + 16, 8);
    if (cap != c->win_cap) { fails = 1; break; }   // every rank sizes its window by the same rule; anything else: no direct exchange
    if (d == c->rank) { c->peer[(size_t)d] = c->win; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, all.data() + REC * (size_t)d, 64);
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); fails = 1; break; }
    c->peer[(size_t)d] = (uint32_t*)p;
  }
  c->mapped = true;
  { int32_t r = sc_allreduce_host_u32(ctx, c, &fails, 1); if (r) return r; }   // all or nothing
  c->push_ok = fails == 0;
  if (c->push_ok) CK(cudaMemcpyAsync(c->d_peer, c->peer.data(), sizeof(uint32_t*) * (size_t)c->world, cudaMemcpyHostToDevice, ctx->st));
  return SC_OK;
}
// One column->row exchange by peer stores.  cols / segs / sharded: the columns this rank owns, as in sc_pack_exchange;
// recv_counts[s]: words rank s contributes to every rank's receive buffer (its block starts at sum(recv_counts[0..s))).
// *recv_out: the receive buffer (a region of this rank's window; freeing the handle releases nothing), or NULL when the
// direct path is not available (no window, window too small — it grows at the next sc_exchange_begin): the caller then uses
// sc_pack_exchange + sc_all_to_all.  Every rank takes the same branch (the sizes are the same on every rank).
int32_t sc_exchange_push(sc_ctx* ctx, sc_comm* c, sc_col* const* cols, const uint64_t* segs, const uint8_t* sharded, uint32_t n,
                         const uint64_t* recv_counts, sc_col** recv_out) {
  ENTER();
  if (!c || !recv_counts || !recv_out || (n && (!cols || !segs || !sharded))) return fail(SC_EINVAL, "exchange_push: null argument");
  *recv_out = nullptr;
  uint64_t rtot = 0, roff_me = 0;
  for (int s = 0; s < c->world; s++) { if (s < c->rank) roff_me += recv_counts[s]; rtot += recv_counts[s]; }
  const uint64_t region = (rtot + 63) & ~63ull;
  c->win_need += region;
  if (!c->push_ok || c->win_off + region > c->win_cap) return SC_OK;
  std::vector<PackCol> pc(n);
  uint64_t off = 0;
  uint32_t max_seg = 1;
  for (uint32_t j = 0; j < n; j++) {
    if (!cols[j] || segs[j] > 0xffffffffull || cols[j]->len < (sharded[j] ? segs[j] * c->world : segs[j])) return fail(SC_EINVAL, "exchange_push: bad column");
    pc[j] = {cols[j]->d, (uint32_t)segs[j], sharded[j] ? 1u : 0u, roff_me + off};
    off += segs[j];
    max_seg = std::max<uint32_t>(max_seg, (uint32_t)segs[j]);
  }
  if (off != recv_counts[c->rank]) return fail(SC_EINVAL, "exchange_push: this rank's columns do not add up to its receive count");
  if (n) {
    void* d_pc = nullptr;
    { int32_t r = stage(ctx, pc.data(), pc.size() * sizeof(PackCol), &d_pc); if (r) return r; }
    const uint32_t bx = std::max(1u, std::min(64u, (max_seg / 4 + 255) / 256));
    ProfScope ps_(ctx, "push_exchange");
    push_exchange_kernel<<<dim3(bx, n, (unsigned)c->world), 256, 0, ctx->st>>>((const PackCol*)d_pc, n, c->d_peer, c->win_off);
    g_launch_count++; CK(cudaGetLastError());
  }
  { ProfScope ps_(ctx, "nccl_all_reduce"); CKN(nccl()->AllReduce(c->d_flag, c->d_flag, 1, ncclUint32, ncclSum, c->comm, ctx->st)); }
  sc_col* r = new sc_col{c->win + c->win_off, rtot};
  r->owned = false;
  track(ctx, r);
  *recv_out = r;
  c->win_off += region;
  return SC_OK;
}
// Pieces of this rank to arbitrary places in arbitrary ranks' windows: piece j (cols[j], whole) lands at word dst_offs[j] of
// rank dest_ranks[j]'s region.  Every rank calls this with the same region_words (the windows advance together); the caller's
// layout must make the pieces of all ranks disjoint.  *recv_out: this rank's region, or NULL when the direct path is not
// available (the caller then falls back to its all-to-all).  Used for the composition accumulators: row ranges -> whole
// coordinate columns on their owner ranks.
int32_t sc_exchange_scatter(sc_ctx* ctx, sc_comm* c, sc_col* const* cols, const uint32_t* dest_ranks, const uint64_t* dst_offs, uint32_t n,
                            uint64_t region_words, sc_col** recv_out) {
  ENTER();
  if (!c || !recv_out || (n && (!cols || !dest_ranks || !dst_offs))) return fail(SC_EINVAL, "exchange_scatter: null argument");
  *recv_out = nullptr;
  const uint64_t region = (region_words + 63) & ~63ull;
  c->win_need += region;
  if (!c->push_ok || c->win_off + region > c->win_cap) return SC_OK;
  std::vector<ScatterCol> sc(n);
  uint32_t max_len = 1;
  for (uint32_t j = 0; j < n; j++) {
    if (!cols[j] || cols[j]->len > 0xffffffffull || (int)dest_ranks[j] >= c->world || dst_offs[j] + cols[j]->len > region_words)
      return fail(SC_EINVAL, "exchange_scatter: bad piece");
    sc[j] = {cols[j]->d, (uint32_t)cols[j]->len, dest_ranks[j], dst_offs[j]};
    max_len = std::max<uint32_t>(max_len, (uint32_t)cols[j]->len);
  }
  if (n) {
    void* d_sc = nullptr;
    { int32_t r = stage(ctx, sc.data(), sc.size() * sizeof(ScatterCol), &d_sc); if (r) return r; }
    const uint32_t bx = std::max(1u, std::min(128u, (max_len / 4 + 255) / 256));
    ProfScope ps_(ctx, "push_exchange");
    scatter_exchange_kernel<<<dim3(bx, n), 256, 0, ctx->st>>>((const ScatterCol*)d_sc, c->d_peer, c->win_off);
    g_launch_count++; CK(cudaGetLastError());
  }
  { ProfScope ps_(ctx, "nccl_all_reduce"); CKN(nccl()->AllReduce(c->d_flag, c->d_flag, 1, ncclUint32, ncclSum, c->comm, ctx->st)); }
  sc_col* r = new sc_col{c->win + c->win_off, region_words};
  r->owned = false;
  track(ctx, r);
  *recv_out = r;
  c->win_off += region;
  return SC_OK;
}
int32_t sc_comm_rank(const sc_comm* c) { return c ? c->rank : 0; }
int32_t sc_comm_world(const sc_comm* c) { return c ? c->world : 1; }

}  // extern "C"
