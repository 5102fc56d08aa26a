// Device-side table building: Vec<Registers> -> the main-trace columns of the 13 components.
//
// Replaces the host builders behind the reference's 13 `trace_evaluation` calls (crates/brainfuck_prover/src/brainfuck_air/
// mod.rs:511-547): memory/table.rs:249-318 + :85-117 (sort by (mp, clk), fill the clk gaps, pad, pair with next),
// instruction/table.rs:250-281 + :85-110 (program rows ++ trace rows, stable sort by (ip, clk), pad), program/table.rs:36-46,
// processor/table.rs:117-142,195-207, processor/instructions/table.rs:293-328, jump/table.rs:264-297 and
// end_of_execution/table.rs:71-77.  Output is the lane-compact form the rest of the path consumes: ONE word per table row
// (the reference then writes that word into all 16 SIMD lanes, `data[vec_row] = value.into()`, e.g. processor/table.rs:86-100;
// csrc/fft.cu and merkle.cu work on the distinct values).
//
// The 7-word register rows (crates/brainfuck_vm/src/registers.rs:5-21) are uploaded once (28 B per VM step instead of the
// ~300 B per step of finished columns), unpacked into seven arrays, and every table is produced by data-parallel kernels:
//   * Processor / Program / EndOfExecution: one thread per row;
//   * the eight opcode tables: ordered stream compaction (per-block counts -> scan -> ranks from warp ballots), then one
//     thread per table row gathering the step and its successor;
//   * Memory and Instruction: a stable LSD radix sort (8-bit digits; the trace is in clk order, so a stable sort on the key
//     alone is the reference's sort on (key, clk)), then for Memory a prefix sum of the per-entry row offsets and a binary
//     search per output row that places the real entries and synthesises the gap / padding dummies.
// Table sizes are inputs: the host derives them while the VM runs (csrc/host/vm.hpp TraceStats); every kernel that can
// see a disagreement raises a flag in `status` instead of writing out of bounds.
#include "kernels.cuh"

namespace sb {

// ---------------------------------------------------------------------------------------------------------------- unpack
// regs: n x 7 words (clk ip ci ni mp mv mvi) -> seven arrays of n words.  fill_mvi: compute mvi = mv^-1 (0 for 0) here
// instead of trusting the input (machine.rs:224-228).  status |= 1 when clk is not strictly increasing.
__global__ void __launch_bounds__(256) tb_unpack_kernel(const uint32_t* __restrict__ regs, uint32_t n, uint32_t fill_mvi, TraceSoA t,
                                                        uint32_t* __restrict__ status) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t* r = regs + (size_t)i * 7;
  const uint32_t clk = r[0], mv = r[5];
  t.clk[i] = clk; t.ip[i] = r[1]; t.ci[i] = r[2]; t.ni[i] = r[3]; t.mp[i] = r[4]; t.mv[i] = mv;
  t.mvi[i] = fill_mvi ? (mv ? m_inv(mv) : 0u) : r[6];
  if (i && regs[(size_t)(i - 1) * 7] >= clk) atomicOr(status, 1u);
}

// ---------------------------------------------------------------------------------------------------------------- simple tables
// processor/table.rs: row i < m is step i; row i >= m is dummy(last.clk + (i - m + 1), last.ip); next_clk pairs with row i+1.
__global__ void __launch_bounds__(256) tb_processor_kernel(TraceSoA t, uint32_t m, uint32_t n, ColPtrs c) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t last_clk = t.clk[m - 1], last_ip = t.ip[m - 1];
  auto clk_of = [&](uint32_t j) { return j < m ? t.clk[j] : (uint32_t)(((uint64_t)last_clk + (j - m + 1)) % P); };
  if (i < m) {
    c.p[0][i] = t.clk[i]; c.p[1][i] = t.ip[i]; c.p[2][i] = t.ci[i]; c.p[3][i] = t.ni[i]; c.p[4][i] = t.mp[i]; c.p[5][i] = t.mv[i];
    c.p[6][i] = t.mvi[i]; c.p[7][i] = 0;
  } else {
    c.p[0][i] = clk_of(i); c.p[1][i] = last_ip; c.p[2][i] = 0; c.p[3][i] = 0; c.p[4][i] = 0; c.p[5][i] = 0; c.p[6][i] = 0; c.p[7][i] = 1;
  }
  c.p[8][i] = clk_of(i + 1);
}
// program/table.rs: (ip, ci, ni, d) for every program word, padded with (last ip, 0, 0, 1).
__global__ void __launch_bounds__(256) tb_program_kernel(const uint32_t* __restrict__ code, uint32_t np, uint32_t n, ColPtrs c) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i < np) { c.p[0][i] = i; c.p[1][i] = code[i]; c.p[2][i] = i + 1 == np ? 0u : code[i + 1]; c.p[3][i] = 0; }
  else { c.p[0][i] = np - 1; c.p[1][i] = 0; c.p[2][i] = 0; c.p[3][i] = 1; }
}
// end_of_execution/table.rs: the one row with ci == 0 (its index comes from the host's TraceStats; checked here).
__global__ void tb_eoe_kernel(TraceSoA t, uint32_t idx, ColPtrs c, uint32_t* status) {
  if (threadIdx.x || blockIdx.x) return;
  if (t.ci[idx] != 0) atomicOr(status, 2u);
  c.p[0][0] = t.clk[idx]; c.p[1][0] = t.ip[idx]; c.p[2][0] = t.ci[idx]; c.p[3][0] = t.ni[idx]; c.p[4][0] = t.mp[idx]; c.p[5][0] = t.mv[idx];
  c.p[6][0] = t.mvi[idx];
}

// ---------------------------------------------------------------------------------------------------------------- opcode tables
__device__ __forceinline__ int tb_op_slot(uint32_t ci) {  // component id - 4: ] [ , < - . + >
  switch (ci) { case ']': return 0; case '[': return 1; case ',': return 2; case '<': return 3; case '-': return 4; case '.': return 5;
                case '+': return 6; case '>': return 7; default: return -1; }
}
constexpr uint32_t TB_BLOCK = 1024;
// Pass 1: per-block count of every opcode among steps [0, m-1) (a step needs a successor to be paired with).
// cnt[s * nblk + b].  Pass 3 (WRITE): the same walk, ranks from ballots, steps[s][base + rank] = i.
template <bool WRITE>
__global__ void __launch_bounds__(TB_BLOCK) tb_opcode_kernel(const uint32_t* __restrict__ ci, uint32_t m1, uint32_t nblk,
                                                             uint32_t* __restrict__ cnt, OpSteps steps, OpCounts cap) {
  __shared__ uint32_t wc[TB_BLOCK / 32][8];
  const uint32_t i = blockIdx.x * TB_BLOCK + threadIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int slot = i < m1 ? tb_op_slot(ci[i]) : -1;
  uint32_t my_rank = 0;
#pragma unroll
  for (int s = 0; s < 8; s++) {
    const uint32_t mask = __ballot_sync(0xffffffffu, slot == s);
    if (lane == 0) wc[w][s] = __popc(mask);
    if (slot == s) my_rank = __popc(mask & ((1u << lane) - 1));
  }
  __syncthreads();
  if (!WRITE) {
    if (threadIdx.x < 8) {
      uint32_t tot = 0;
      for (uint32_t k = 0; k < TB_BLOCK / 32; k++) tot += wc[k][threadIdx.x];
      cnt[threadIdx.x * nblk + blockIdx.x] = tot;
    }
  } else if (slot >= 0) {
    uint32_t base = cnt[slot * nblk + blockIdx.x];  // exclusive scan of the block counts
    for (uint32_t k = 0; k < w; k++) base += wc[k][slot];
    if (base + my_rank < cap.n[slot]) steps.p[slot][base + my_rank] = i;  // a count that disagrees with the host's is flagged by pass 2
  }
}
// Pass 2: exclusive scan of each opcode's block counts (one CTA per opcode); totals checked against the host's counts.
__global__ void __launch_bounds__(1024) tb_scan_counts_kernel(uint32_t* __restrict__ cnt, uint32_t nblk, OpCounts expect, uint32_t* status) {
  __shared__ uint32_t sh[1024];
  uint32_t* c = cnt + blockIdx.x * nblk;
  uint32_t carry = 0;
  for (uint32_t base = 0; base < nblk; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < nblk ? c[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (uint32_t d = 1; d < 1024; d <<= 1) {
      uint32_t x = threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
      __syncthreads();
      sh[threadIdx.x] += x;
      __syncthreads();
    }
    if (i < nblk) c[i] = carry + sh[threadIdx.x] - v;
    carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0 && carry != expect.n[blockIdx.x]) atomicOr(status, 4u);
}
// One thread per row of one opcode table (blockIdx.y = opcode slot).  instructions/table.rs:293-328, jump/table.rs:264-297:
// row < cnt: (step, next step); beyond: dummy(last_clk + j, last_ip) for entry j = 2(row - cnt) (+1 for the second half).
__global__ void __launch_bounds__(256) tb_op_fill_kernel(TraceSoA t, OpSteps steps, OpCounts cnt, OpTables tabs) {
  const uint32_t s = blockIdx.y, row = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = tabs.rows[s], k = cnt.n[s];
  if (row >= n) return;
  uint32_t* const* c = tabs.cols[s];
  uint32_t last_clk = 0, last_ip = 0;
  if (k) { const uint32_t l = steps.p[s][k - 1] + 1; last_clk = t.clk[l]; last_ip = t.ip[l]; }
  uint32_t a[8], b[4];  // a: clk ip ci ni mp mv mvi d;  b: clk ip mp mv
  if (row < k) {
    const uint32_t i = steps.p[s][row];
    a[0] = t.clk[i]; a[1] = t.ip[i]; a[2] = t.ci[i]; a[3] = t.ni[i]; a[4] = t.mp[i]; a[5] = t.mv[i]; a[6] = t.mvi[i]; a[7] = 0;
    b[0] = t.clk[i + 1]; b[1] = t.ip[i + 1]; b[2] = t.mp[i + 1]; b[3] = t.mv[i + 1];
  } else {
    const uint64_t j = 2ull * (row - k);
    a[0] = (uint32_t)((last_clk + j) % P); a[1] = last_ip; a[2] = a[3] = a[4] = a[5] = a[6] = 0; a[7] = 1;
    b[0] = (uint32_t)((last_clk + j + 1) % P); b[1] = last_ip; b[2] = b[3] = 0;
  }
#pragma unroll
  for (int q = 0; q < 7; q++) c[q][row] = a[q];
  if (s < 2) {  // JumpColumn: ... next_clk next_ip next_mp next_mv d is_mv_zero
    c[7][row] = b[0]; c[8][row] = b[1]; c[9][row] = b[2]; c[10][row] = b[3]; c[11][row] = a[7];
    c[12][row] = m_sub(1, m_mul(a[5], a[6]));
  } else {      // ProcessorInstructionColumn: ... d next_ip next_mp next_mv
    c[7][row] = a[7]; c[8][row] = b[1]; c[9][row] = b[2]; c[10][row] = b[3];
  }
}

// ---------------------------------------------------------------------------------------------------------------- stable radix sort
// One 8-bit digit per pass over (key, payload) pairs.  A CTA owns a tile of RS_TILE consecutive elements; warp w owns the
// RS_TILE/8 consecutive elements [w * 256, (w + 1) * 256) of it and walks them 32 at a time, so "earlier in the input" is
// "earlier chunk, lower lane" everywhere and the scatter below is stable.
constexpr uint32_t RS_THREADS = 256, RS_ITEMS = 8, RS_TILE = RS_THREADS * RS_ITEMS;
// iota: the payload of the first pass is the element's own index (not read from memory)
__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const uint32_t* __restrict__ keys, uint32_t n, uint32_t shift, uint32_t nblk,
                                                             uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t base = blockIdx.x * RS_TILE;
  for (uint32_t j = 0; j < RS_ITEMS; j++) {
    const uint32_t i = base + j * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];  // digit-major: one exclusive scan gives every (digit, tile) base
}
// Exclusive scan of `len` words in place by ONE CTA (len <= a few hundred thousand: 256 digits x tiles).
__global__ void __launch_bounds__(1024) scan_single_cta_kernel(uint32_t* __restrict__ v, uint32_t len) {
  __shared__ uint32_t sh[1024];
  const uint32_t per = (len + 1023) / 1024, lo = threadIdx.x * per, hi = min(len, lo + per);
  uint32_t sum = 0;
  for (uint32_t i = lo; i < hi; i++) sum += v[i];
  sh[threadIdx.x] = sum;
  __syncthreads();
  for (uint32_t d = 1; d < 1024; d <<= 1) {
    uint32_t x = threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
    __syncthreads();
    sh[threadIdx.x] += x;
    __syncthreads();
  }
  uint32_t run = sh[threadIdx.x] - sum;
  for (uint32_t i = lo; i < hi; i++) { const uint32_t x = v[i]; v[i] = run; run += x; }
}
template <bool IOTA>
__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint32_t n,
                                                                uint32_t shift, uint32_t nblk, const uint32_t* __restrict__ hist,
                                                                uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
  __shared__ uint32_t wcnt[RS_THREADS / 32][256];  // per warp: elements of each digit seen so far in the warp's segment
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (uint32_t d = threadIdx.x; d < 256 * (RS_THREADS / 32); d += RS_THREADS) (&wcnt[0][0])[d] = 0;
  __syncthreads();
  const uint32_t seg = blockIdx.x * RS_TILE + w * (RS_TILE / (RS_THREADS / 32));
  uint32_t key[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
  for (uint32_t j = 0; j < RS_ITEMS; j++) {
    const uint32_t i = seg + j * 32 + lane;
    const bool live = i < n;
    key[j] = live ? keys[i] : 0xffffffffu;
    const uint32_t dg = (key[j] >> shift) & 255u;
    const uint32_t peers = __match_any_sync(0xffffffffu, live ? dg : 256u + lane);  // dead lanes match nobody
    const uint32_t leader = __ffs(peers) - 1;
    uint32_t before = 0;
    if (live && lane == leader) { before = wcnt[w][dg]; wcnt[w][dg] = before + __popc(peers); }
    before = __shfl_sync(0xffffffffu, before, leader);
    rank[j] = before + __popc(peers & ((1u << lane) - 1));
    __syncwarp();
  }
  __syncthreads();
#pragma unroll
  for (uint32_t j = 0; j < RS_ITEMS; j++) {
    const uint32_t i = seg + j * 32 + lane;
    if (i >= n) continue;
    const uint32_t dg = (key[j] >> shift) & 255u;
    uint32_t pos = hist[dg * nblk + blockIdx.x] + rank[j];
    for (uint32_t k = 0; k < w; k++) pos += wcnt[k][dg];
    keys_out[pos] = key[j];
    vals_out[pos] = IOTA ? i : vals[i];
  }
}

// ---------------------------------------------------------------------------------------------------------------- memory table
// After the sort: ord[k] = step of the k-th entry in (mp, clk) order.  delta[k] = rows this entry adds = clk gap to the
// previous entry of the same cell, or 1 for the first entry of a cell (memory/table.rs:259-283).
__global__ void __launch_bounds__(256) tb_mem_delta_kernel(TraceSoA t, const uint32_t* __restrict__ ord, uint32_t m, uint32_t* __restrict__ delta) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= m) return;
  const uint32_t i = ord[k];
  uint32_t d = 1;
  if (k) { const uint32_t p = ord[k - 1]; if (t.mp[p] == t.mp[i]) d = t.clk[i] - t.clk[p]; }
  delta[k] = d;
}
// Inclusive scan, three phases (tiles of 1024 x 4 words).
constexpr uint32_t SC_TILE = 4096;
__global__ void __launch_bounds__(1024) scan_tile_sums_kernel(const uint32_t* __restrict__ v, uint32_t n, uint32_t* __restrict__ sums) {
  __shared__ uint32_t sh[32];
  const uint32_t base = blockIdx.x * SC_TILE + threadIdx.x * 4;
  uint32_t s = 0;
  for (uint32_t q = 0; q < 4; q++) if (base + q < n) s += v[base + q];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = sh[threadIdx.x];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) sums[blockIdx.x] = s;
  }
}
__global__ void __launch_bounds__(1024) scan_tiles_kernel(uint32_t* __restrict__ v, uint32_t n, const uint32_t* __restrict__ tile_base) {
  __shared__ uint32_t sh[1024];
  const uint32_t base = blockIdx.x * SC_TILE + threadIdx.x * 4;
  uint32_t x[4], s = 0;
  for (uint32_t q = 0; q < 4; q++) { x[q] = base + q < n ? v[base + q] : 0; s += x[q]; }
  sh[threadIdx.x] = s;
  __syncthreads();
  for (uint32_t d = 1; d < 1024; d <<= 1) {
    uint32_t y = threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
    __syncthreads();
    sh[threadIdx.x] += y;
    __syncthreads();
  }
  uint32_t run = tile_base[blockIdx.x] + sh[threadIdx.x] - s;
  for (uint32_t q = 0; q < 4; q++) { run += x[q]; if (base + q < n) v[base + q] = run; }
}
// One thread per output row w of the Memory table.  end[k] = inclusive scan of delta = 1 + row index of real entry k.
// Row w < rows: the first k with end[k] - 1 >= w is the entry at or after w; equality -> the real entry, otherwise a gap dummy
// (clk counts back from that entry, mp / mv of the previous one, d = 1).  Rows beyond: padding from the last row.
struct MemRow { uint32_t clk, mp, mv, d; };
__device__ __forceinline__ MemRow tb_mem_row(const TraceSoA& t, const uint32_t* ord, const uint32_t* end, uint32_t m, uint32_t rows, uint32_t w) {
  if (w >= rows) {
    const uint32_t l = ord[m - 1];
    return {(uint32_t)(((uint64_t)t.clk[l] + (w - rows) + 1) % P), t.mp[l], t.mv[l], 1u};
  }
  uint32_t lo = 0, hi = m - 1;  // smallest k with end[k] > w
  while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (end[mid] > w) hi = mid; else lo = mid + 1; }
  const uint32_t i = ord[lo], g = end[lo] - 1 - w;
  if (g == 0 || lo == 0) return {t.clk[i], t.mp[i], t.mv[i], 0u};  // lo == 0 with a gap only if `rows` is inconsistent (flagged)
  const uint32_t p = ord[lo - 1];
  return {t.clk[i] - g, t.mp[p], t.mv[p], 1u};
}
__global__ void __launch_bounds__(256) tb_mem_fill_kernel(TraceSoA t, const uint32_t* __restrict__ ord, const uint32_t* __restrict__ end, uint32_t m,
                                                          uint32_t rows, uint32_t n, ColPtrs c, uint32_t* status) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n) return;
  if (w == 0 && end[m - 1] != rows) atomicOr(status, 8u);
  const MemRow a = tb_mem_row(t, ord, end, m, rows, w), b = tb_mem_row(t, ord, end, m, rows, w + 1);
  c.p[0][w] = a.clk; c.p[1][w] = a.mp; c.p[2][w] = a.mv; c.p[3][w] = a.d;
  c.p[4][w] = b.clk; c.p[5][w] = b.mp; c.p[6][w] = b.mv; c.p[7][w] = b.d;
}

// ---------------------------------------------------------------------------------------------------------------- instruction table
// keys of the concatenation program rows ++ trace rows: element i < np has ip = i, element np + s has the ip of step s
__global__ void __launch_bounds__(256) tb_ins_keys_kernel(const uint32_t* __restrict__ ip, uint32_t np, uint32_t total, uint32_t* __restrict__ keys) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) keys[i] = i < np ? i : ip[i - np];
}
struct InsRow { uint32_t ip, ci, ni, d; };
__device__ __forceinline__ InsRow tb_ins_row(const TraceSoA& t, const uint32_t* code, const uint32_t* ord, uint32_t np, uint32_t total, uint32_t k) {
  if (k >= total) {  // dummy(last.ip)
    const uint32_t i = ord[total - 1];
    return {i < np ? i : t.ip[i - np], 0u, 0u, 1u};
  }
  const uint32_t i = ord[k];
  if (i < np) return {i, code[i], i + 1 == np ? 0u : code[i + 1], 0u};
  return {t.ip[i - np], t.ci[i - np], t.ni[i - np], 0u};
}
__global__ void __launch_bounds__(256) tb_ins_fill_kernel(TraceSoA t, const uint32_t* __restrict__ code, const uint32_t* __restrict__ ord, uint32_t np,
                                                          uint32_t total, uint32_t n, ColPtrs c) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const InsRow a = tb_ins_row(t, code, ord, np, total, k), b = tb_ins_row(t, code, ord, np, total, k + 1);
  c.p[0][k] = a.ip; c.p[1][k] = a.ci; c.p[2][k] = a.ni; c.p[3][k] = a.d;
  c.p[4][k] = b.ip; c.p[5][k] = b.ci; c.p[6][k] = b.ni; c.p[7][k] = b.d;
}

// ================================================================================================================ launchers
#define TB_LAUNCHED() do { g_launch_count++; cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return (int)e_; } while (0)
static inline uint32_t cdiv(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

int launch_tb_unpack(const uint32_t* d_regs, uint32_t n, bool fill_mvi, const TraceSoA& t, uint32_t* d_status, cudaStream_t st) {
  tb_unpack_kernel<<<cdiv(n, 256), 256, 0, st>>>(d_regs, n, fill_mvi ? 1u : 0u, t, d_status); TB_LAUNCHED();
  return 0;
}
int launch_tb_processor(const TraceSoA& t, uint32_t m, uint32_t n, const ColPtrs& c, cudaStream_t st) {
  tb_processor_kernel<<<cdiv(n, 256), 256, 0, st>>>(t, m, n, c); TB_LAUNCHED();
  return 0;
}
int launch_tb_program(const uint32_t* d_code, uint32_t np, uint32_t n, const ColPtrs& c, cudaStream_t st) {
  tb_program_kernel<<<cdiv(n, 256), 256, 0, st>>>(d_code, np, n, c); TB_LAUNCHED();
  return 0;
}
int launch_tb_eoe(const TraceSoA& t, uint32_t idx, const ColPtrs& c, uint32_t* d_status, cudaStream_t st) {
  tb_eoe_kernel<<<1, 32, 0, st>>>(t, idx, c, d_status); TB_LAUNCHED();
  return 0;
}
size_t tb_opcode_scratch_words(uint32_t m) { return 8 * (size_t)cdiv(m ? m : 1, TB_BLOCK); }
// steps.p[s] must hold cnt.n[s] words; tabs.rows[s] rows and 13 / 11 column pointers per table; d_cnt: tb_opcode_scratch_words(m)
int launch_tb_opcodes(const TraceSoA& t, uint32_t m, const OpCounts& cnt, const OpSteps& steps, const OpTables& tabs, uint32_t* d_cnt,
                      uint32_t* d_status, cudaStream_t st) {
  const uint32_t m1 = m ? m - 1 : 0, nblk = cdiv(m ? m : 1, TB_BLOCK);
  tb_opcode_kernel<false><<<nblk, TB_BLOCK, 0, st>>>(t.ci, m1, nblk, d_cnt, steps, cnt); TB_LAUNCHED();
  tb_scan_counts_kernel<<<8, 1024, 0, st>>>(d_cnt, nblk, cnt, d_status); TB_LAUNCHED();
  tb_opcode_kernel<true><<<nblk, TB_BLOCK, 0, st>>>(t.ci, m1, nblk, d_cnt, steps, cnt); TB_LAUNCHED();
  uint32_t max_rows = 1;
  for (int s = 0; s < 8; s++) max_rows = tabs.rows[s] > max_rows ? tabs.rows[s] : max_rows;
  tb_op_fill_kernel<<<dim3(cdiv(max_rows, 256), 8), 256, 0, st>>>(t, steps, cnt, tabs); TB_LAUNCHED();
  return 0;
}
size_t rs_scratch_words(uint32_t n) { return 256 * (size_t)cdiv(n ? n : 1, RS_TILE); }
// Stable sort of n (key, index) pairs by the low `key_bits` bits of the key.  keys_in is left untouched; kbuf / vbuf: two
// buffers of n words each; on return *ord_out points at the vbuf that holds the permutation.  d_hist: rs_scratch_words(n).
int launch_radix_sort_index(const uint32_t* keys_in, uint32_t* const kbuf[2], uint32_t* const vbuf[2], uint32_t n, uint32_t key_bits,
                            uint32_t* d_hist, uint32_t** ord_out, cudaStream_t st) {
  const uint32_t nblk = cdiv(n ? n : 1, RS_TILE);
  const uint32_t passes = cdiv(key_bits ? key_bits : 1, 8);
  for (uint32_t p = 0; p < passes; p++) {
    const uint32_t* kin = p == 0 ? keys_in : kbuf[(p - 1) & 1];
    rs_hist_kernel<<<nblk, RS_THREADS, 0, st>>>(kin, n, 8 * p, nblk, d_hist); TB_LAUNCHED();
    scan_single_cta_kernel<<<1, 1024, 0, st>>>(d_hist, 256 * nblk); TB_LAUNCHED();
    if (p == 0) rs_scatter_kernel<true><<<nblk, RS_THREADS, 0, st>>>(kin, nullptr, n, 0, nblk, d_hist, kbuf[0], vbuf[0]);
    else rs_scatter_kernel<false><<<nblk, RS_THREADS, 0, st>>>(kin, vbuf[(p - 1) & 1], n, 8 * p, nblk, d_hist, kbuf[p & 1], vbuf[p & 1]);
    TB_LAUNCHED();
  }
  *ord_out = vbuf[(passes - 1) & 1];
  return 0;
}
size_t scan_scratch_words(uint32_t n) { return cdiv(n ? n : 1, SC_TILE) + 1; }
int launch_inclusive_scan(uint32_t* v, uint32_t n, uint32_t* d_sums, cudaStream_t st) {
  const uint32_t tiles = cdiv(n ? n : 1, SC_TILE);
  scan_tile_sums_kernel<<<tiles, 1024, 0, st>>>(v, n, d_sums); TB_LAUNCHED();
  scan_single_cta_kernel<<<1, 1024, 0, st>>>(d_sums, tiles); TB_LAUNCHED();
  scan_tiles_kernel<<<tiles, 1024, 0, st>>>(v, n, d_sums); TB_LAUNCHED();
  return 0;
}
// Memory table from the (mp, clk)-sorted permutation.  d_delta: m words (becomes the inclusive scan), d_sums: scan_scratch_words(m).
int launch_tb_memory(const TraceSoA& t, const uint32_t* d_ord, uint32_t m, uint32_t rows, uint32_t n, const ColPtrs& c, uint32_t* d_delta,
                     uint32_t* d_sums, uint32_t* d_status, cudaStream_t st) {
  tb_mem_delta_kernel<<<cdiv(m, 256), 256, 0, st>>>(t, d_ord, m, d_delta); TB_LAUNCHED();
  int e = launch_inclusive_scan(d_delta, m, d_sums, st);
  if (e) return e;
  tb_mem_fill_kernel<<<cdiv(n, 256), 256, 0, st>>>(t, d_ord, d_delta, m, rows, n, c, d_status); TB_LAUNCHED();
  return 0;
}
int launch_tb_ins_keys(const uint32_t* d_ip, uint32_t np, uint32_t total, uint32_t* d_keys, cudaStream_t st) {
  tb_ins_keys_kernel<<<cdiv(total, 256), 256, 0, st>>>(d_ip, np, total, d_keys); TB_LAUNCHED();
  return 0;
}
int launch_tb_instruction(const TraceSoA& t, const uint32_t* d_code, const uint32_t* d_ord, uint32_t np, uint32_t total, uint32_t n,
                          const ColPtrs& c, cudaStream_t st) {
  tb_ins_fill_kernel<<<cdiv(n, 256), 256, 0, st>>>(t, d_code, d_ord, np, total, n, c); TB_LAUNCHED();
  return 0;
}

}  // namespace sb
