// Host-side table builders: Vec<Registers> -> the 13 component tables (one value per table row; the 16-lane
// broadcast of the reference's `data[row] = x.into()` is applied on the device by sc_col_broadcast16).
// Follows crates/brainfuck_prover/src/components/{memory,instruction,program,processor}/table.rs and
// processor/instructions/{table.rs,jump/table.rs,end_of_execution/table.rs} — cited per function.
#pragma once
#include <algorithm>
#include <string>
#include <thread>
#include <utility>
#include "air_ids.hpp"
#include "vm.hpp"

namespace sbf {

// Table columns can live in a backend-provided host arena (pinned memory in the CUDA backend, so that the upload is a
// true asynchronous DMA instead of a staged pageable copy).  The arena is installed by the prover around build_tables
// and the tables' lifetime; without one the columns are ordinary heap vectors.
struct HostArena {
  virtual ~HostArena() {}
  virtual void* alloc(size_t bytes) = 0;  // thread-safe, 64-byte aligned, released wholesale by the owner
};
inline HostArena*& current_arena() { static HostArena* a = nullptr; return a; }
template <class T>
struct ArenaAlloc {
  typedef T value_type;
  ArenaAlloc() = default;
  template <class U> ArenaAlloc(const ArenaAlloc<U>&) {}
  T* allocate(size_t n) {
    if (HostArena* a = current_arena()) return static_cast<T*>(a->alloc(n * sizeof(T)));
    return static_cast<T*>(::operator new(n * sizeof(T)));
  }
  void deallocate(T* p, size_t) { if (!current_arena()) ::operator delete(p); }
  // `ColVec(n)` leaves its words uninitialised (every builder writes every row of every column); `ColVec(n, v)` fills
  template <class U> void construct(U*) noexcept {}
  template <class U, class A0, class... A> void construct(U* p, A0&& a0, A&&... a) { ::new ((void*)p) U(std::forward<A0>(a0), std::forward<A>(a)...); }
  template <class U> bool operator==(const ArenaAlloc<U>&) const { return true; }
  template <class U> bool operator!=(const ArenaAlloc<U>&) const { return false; }
};
typedef std::vector<uint32_t, ArenaAlloc<uint32_t>> ColVec;
// builder scratch: large heap vectors are mmap'ed and page-faulted afresh on every proof, the arena's pages are not
template <class T> using Scratch = std::vector<T, ArenaAlloc<T>>;
typedef std::vector<ColVec> ColVecs;

struct Table {
  uint32_t log_size = 0;                    // column log size = log2(rows) + LOG_N_LANES
  ColVecs cols;  // cols[c][row], rows = 2^(log_size - 4)
  size_t rows() const { return cols.empty() ? 0 : cols[0].size(); }
};

inline ColVecs make_cols(size_t ncols, size_t rows) {  // uninitialised, no prototype copies
  ColVecs c;
  c.reserve(ncols);
  for (size_t k = 0; k < ncols; k++) c.emplace_back(rows);
  return c;
}
// fn(lo, hi) over [0, n) on `threads` host threads (the caller's included); small ranges stay on the caller
template <class Fn>
inline void parallel_ranges(size_t n, unsigned threads, Fn fn) {
  if (threads <= 1 || n < ((size_t)1 << 14)) { fn((size_t)0, n); return; }
  std::vector<std::thread> th;
  size_t per = (n + threads - 1) / threads;
  for (unsigned t = 1; t < threads; t++) {
    size_t lo = std::min(n, t * per), hi = std::min(n, lo + per);
    if (lo < hi) th.emplace_back([=] { fn(lo, hi); });
  }
  fn((size_t)0, std::min(n, per));
  for (auto& x : th) x.join();
}
constexpr unsigned TABLE_THREADS = 4;  // extra host threads inside the three big builders

// max of key(i) over [0, n) and a check that ok(i) holds everywhere, on TABLE_THREADS threads
template <class KeyFn, class OkFn>
inline uint32_t scan_max(size_t n, KeyFn key, OkFn ok, const char* what) {
  uint32_t mx[TABLE_THREADS] = {0, 0, 0, 0};
  bool bad[TABLE_THREADS] = {false, false, false, false};
  const size_t per = (n + TABLE_THREADS - 1) / TABLE_THREADS;
  parallel_ranges(n, TABLE_THREADS, [&](size_t lo, size_t hi) {
    const size_t t = per ? std::min<size_t>(lo / per, TABLE_THREADS - 1) : 0;
    uint32_t m = 0;
    bool b = false;
    for (size_t i = lo; i < hi; i++) { m = std::max(m, key(i)); b |= !ok(i); }
    mx[t] = std::max(mx[t], m); bad[t] |= b;
  });
  for (unsigned t = 0; t < TABLE_THREADS; t++) if (bad[t]) throw std::runtime_error(what);
  return *std::max_element(mx, mx + TABLE_THREADS);
}

inline size_t next_pow2(size_t n) { size_t p = 1; while (p < n) p <<= 1; return p; }
inline uint32_t ilog2_exact(size_t n) { uint32_t l = 0; while (((size_t)1 << l) < n) l++; return l; }

inline Table finish(ColVecs cols) {
  Table t;
  size_t rows = cols[0].size();
  if (rows == 0 || (rows & (rows - 1))) throw std::runtime_error("table length must be a non-zero power of two");
  t.log_size = ilog2_exact(rows) + LOG_N_LANES;
  t.cols = std::move(cols);
  return t;
}

// memory/table.rs:249-318 (sort by (mp,clk), fill clk gaps with dummies, pad) and :85-117 (pair with next, extra dummy)
// Stable order of the trace by `key` then clk.  The VM emits rows in clk order, so a counting sort on the key is the
// reference's `sort_by_key(|x| (x.key, x.clk))`; falls back to std::stable_sort for huge key ranges.
template <class KeyFn>
inline Scratch<uint32_t> order_by_key(size_t n, uint32_t max_key, KeyFn key) {
  Scratch<uint32_t> idx(n);
  if (max_key >= (1u << 22)) {
    for (size_t i = 0; i < n; i++) idx[i] = (uint32_t)i;
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return key(a) < key(b); });
    return idx;
  }
  // counting sort; large inputs on TABLE_THREADS threads: per-thread histograms of contiguous chunks, offsets taken in
  // (key, thread) order so that equal keys keep their input order
  const size_t K = (size_t)max_key + 1;
  const unsigned T = (n >= ((size_t)1 << 16) && K * TABLE_THREADS <= n) ? TABLE_THREADS : 1;
  const size_t per = (n + T - 1) / T;
  std::vector<uint32_t> cnt(K * T, 0);
  auto run = [&](auto fn) {
    std::vector<std::thread> th;
    for (unsigned t = 1; t < T; t++) th.emplace_back([=] { fn(t, std::min(n, t * per), std::min(n, (t + 1) * per)); });
    fn(0u, (size_t)0, std::min(n, per));
    for (auto& x : th) x.join();
  };
  run([&](unsigned t, size_t lo, size_t hi) { uint32_t* c = cnt.data() + (size_t)t * K; for (size_t i = lo; i < hi; i++) c[key(i)]++; });
  uint32_t acc = 0;
  for (size_t k = 0; k < K; k++)
    for (unsigned t = 0; t < T; t++) { uint32_t c = cnt[(size_t)t * K + k]; cnt[(size_t)t * K + k] = acc; acc += c; }
  run([&](unsigned t, size_t lo, size_t hi) { uint32_t* c = cnt.data() + (size_t)t * K; for (size_t i = lo; i < hi; i++) idx[c[key(i)]++] = (uint32_t)i; });
  return idx;
}

// log_max_rows < 32: fail before the columns are allocated when the table cannot fit (sierpinski.bf needs log size 29:
// 2^25 rows x 8 columns = 1 GiB of host memory that would be thrown away by the size check of the driver)
inline Table memory_table(const std::vector<Registers>& regs, uint32_t log_max_rows = 32) {
  if (regs.empty()) throw std::runtime_error("empty trace");
  const size_t m = regs.size();
  const uint32_t max_mp = scan_max(m, [&](size_t i) { return regs[i].mp; }, [&](size_t i) { return !i || regs[i].clk > regs[i - 1].clk; },
                                   "trace is not in clk order");
  // the (mp, clk)-sorted entries as three dense arrays: the passes below then read memory in order
  struct S { uint32_t clk, mp, mv; };
  Scratch<S> e(m);
  {
    Scratch<uint32_t> ord = order_by_key(m, max_mp, [&](size_t i) { return regs[i].mp; });
    parallel_ranges(m, TABLE_THREADS, [&](size_t lo, size_t hi) {
      for (size_t k = lo; k < hi; k++) { const Registers& r = regs[ord[k]]; e[k] = {r.clk, r.mp, r.mv}; }
    });
  }
  // row offset of every real entry after gap filling (per mp run: last.clk - first.clk + 1 rows)
  Scratch<size_t> off(m);
  size_t rows = 0;
  for (size_t k = 0; k < m; k++) {
    if (k && e[k - 1].mp == e[k].mp) rows += e[k].clk - e[k - 1].clk; else rows += 1;
    off[k] = rows - 1;
  }
  size_t n = next_pow2(rows);
  if (log_max_rows < 32 && log_max_rows >= LOG_N_LANES && n > ((size_t)1 << (log_max_rows - LOG_N_LANES)))
    throw std::runtime_error("component too large: memory (" + std::to_string(rows) + " rows after filling the clk gaps)");
  ColVecs c = make_cols(8, n);
  uint32_t *clk = c[0].data(), *mp = c[1].data(), *mv = c[2].data(), *d = c[3].data();
  parallel_ranges(m, TABLE_THREADS, [&](size_t lo, size_t hi) {
    for (size_t k = lo; k < hi; k++) {
      size_t w = off[k];
      if (k && e[k - 1].mp == e[k].mp) {
        const S& p = e[k - 1];
        size_t g = w - (e[k].clk - p.clk - 1);
        for (uint32_t x = p.clk + 1; x < e[k].clk; x++, g++) { clk[g] = x; mp[g] = p.mp; mv[g] = p.mv; d[g] = 1; }
      }
      clk[w] = e[k].clk; mp[w] = e[k].mp; mv[w] = e[k].mv; d[w] = 0;
    }
  });
  const uint32_t last_clk = clk[rows - 1], last_mp = mp[rows - 1], last_mv = mv[rows - 1];
  parallel_ranges(n - rows, TABLE_THREADS, [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++) {
      size_t w = rows + i;
      clk[w] = (uint32_t)(((uint64_t)last_clk + i + 1) % P); mp[w] = last_mp; mv[w] = last_mv; d[w] = 1;
    }
  });
  parallel_ranges(n - 1, TABLE_THREADS, [&](size_t lo, size_t hi) {  // next_* = the following entry
    for (int k = 0; k < 4; k++) std::copy(c[k].begin() + 1 + lo, c[k].begin() + 1 + hi, c[4 + k].begin() + lo);
  });
  c[4][n - 1] = sb::m_add(clk[n - 1], 1); c[5][n - 1] = mp[n - 1]; c[6][n - 1] = mv[n - 1]; c[7][n - 1] = 1;  // one more dummy
  return finish(std::move(c));
}

// instruction/table.rs:250-281 (program rows ++ trace rows, stable sort by (ip,clk), pad with dummy(last.ip)) and :85-110
inline Table instruction_table(const std::vector<Registers>& regs, const std::vector<uint32_t>& code) {
  // program rows (clk 0) come first in the concatenation and the trace is in clk order, so a stable sort on ip alone is
  // the reference's stable sort on (ip, clk)
  const size_t np = code.size(), total = np + regs.size();
  if (total == 0) throw std::runtime_error("empty trace");
  auto ip_of = [&](size_t i) { return i < np ? (uint32_t)i : regs[i - np].ip; };
  const uint32_t max_ip = scan_max(total, ip_of, [&](size_t i) { return i <= np || regs[i - np].clk > regs[i - np - 1].clk; },
                                   "trace is not in clk order");
  Scratch<uint32_t> ord = order_by_key(total, max_ip, ip_of);
  size_t n = next_pow2(total);
  ColVecs c = make_cols(8, n);
  parallel_ranges(total, TABLE_THREADS, [&](size_t lo, size_t hi) {
    for (size_t k = lo; k < hi; k++) {
      size_t i = ord[k];
      if (i < np) { c[0][k] = (uint32_t)i; c[1][k] = code[i]; c[2][k] = i + 1 == np ? 0 : code[i + 1]; }
      else { const Registers& r = regs[i - np]; c[0][k] = r.ip; c[1][k] = r.ci; c[2][k] = r.ni; }
      c[3][k] = 0;
    }
  });
  uint32_t last_ip = c[0][total - 1];
  for (size_t k = total; k < n; k++) { c[0][k] = last_ip; c[1][k] = 0; c[2][k] = 0; c[3][k] = 1; }
  parallel_ranges(n - 1, TABLE_THREADS, [&](size_t lo, size_t hi) {
    for (int k = 0; k < 4; k++) std::copy(c[k].begin() + 1 + lo, c[k].begin() + 1 + hi, c[4 + k].begin() + lo);
  });
  c[4][n - 1] = c[0][n - 1]; c[5][n - 1] = 0; c[6][n - 1] = 0; c[7][n - 1] = 1;
  return finish(std::move(c));
}

// program/table.rs:36-46,91-121
inline Table program_table(const std::vector<uint32_t>& code) {
  size_t n0 = code.size();
  if (n0 == 0) throw std::runtime_error("empty program");
  size_t n = next_pow2(n0);
  ColVecs c(4, ColVec(n, 0));
  for (size_t i = 0; i < n; i++) {
    if (i < n0) { c[0][i] = (uint32_t)i; c[1][i] = code[i]; c[2][i] = i + 1 == n0 ? 0 : code[i + 1]; c[3][i] = 0; }
    else { c[0][i] = (uint32_t)(n0 - 1); c[3][i] = 1; }
  }
  return finish(std::move(c));
}

// processor/table.rs:117-142 (pair with next + extra dummy), :195-207 (pad with dummy(last.clk+i, last.ip))
inline Table processor_table(const std::vector<Registers>& regs) {
  const size_t m = regs.size();
  if (m == 0) throw std::runtime_error("empty trace");
  const size_t n = next_pow2(m);
  const Registers last = regs.back();
  ColVecs c = make_cols(9, n);
  // row i < m: the step itself; row i >= m: dummy(last.clk + (i - m + 1), last.ip); next_clk pairs with row i + 1 (or one more dummy)
  auto clk_of = [&](size_t i) { return i < m ? regs[i].clk : (uint32_t)(((uint64_t)last.clk + (i - m + 1)) % P); };
  parallel_ranges(n, TABLE_THREADS, [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++) {
      if (i < m) {
        const Registers& r = regs[i];
        c[0][i] = r.clk; c[1][i] = r.ip; c[2][i] = r.ci; c[3][i] = r.ni; c[4][i] = r.mp; c[5][i] = r.mv; c[6][i] = r.mvi; c[7][i] = 0;
      } else {
        c[0][i] = clk_of(i); c[1][i] = last.ip; c[2][i] = 0; c[3][i] = 0; c[4][i] = 0; c[5][i] = 0; c[6][i] = 0; c[7][i] = 1;
      }
      c[8][i] = clk_of(i + 1);
    }
  });
  return finish(std::move(c));
}

// processor/instructions/table.rs:293-328 and jump/table.rs:264-297: for every step with ci == op the pair (step, next step),
// padded in ENTRIES with dummy(last_clk + i, last_ip), i from 0, then chunked in twos; an empty table is one dummy row.
struct PairEntry { uint32_t clk, ip, ci, ni, mp, mv, mvi, d; };
// Step indices by opcode, found in ONE pass over the trace and shared by the eight instruction / jump tables (each of them
// scanning the 28-byte registers twice costs more than building the table once the trace is large).
inline int op_slot(uint32_t ci) {
  switch (ci) { case ']': return 0; case '[': return 1; case ',': return 2; case '<': return 3; case '-': return 4; case '.': return 5;
                case '+': return 6; case '>': return 7; default: return -1; }
}
struct TraceIndex {
  Scratch<uint32_t> steps[8];
  explicit TraceIndex(const std::vector<Registers>& regs) {
    size_t cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const size_t m = regs.empty() ? 0 : regs.size() - 1;   // a step needs a successor to be paired with
    for (size_t i = 0; i < m; i++) { int sl = op_slot(regs[i].ci); if (sl >= 0) cnt[sl]++; }
    for (int sl = 0; sl < 8; sl++) steps[sl].reserve(cnt[sl]);
    for (size_t i = 0; i < m; i++) { int sl = op_slot(regs[i].ci); if (sl >= 0) steps[sl].push_back((uint32_t)i); }
  }
};
struct PairRows {
  const std::vector<Registers>& regs;
  Scratch<uint32_t> own;             // i with regs[i].ci == op (and a following step), when no shared index was given
  const Scratch<uint32_t>& steps;
  uint32_t last_clk = 0, last_ip = 0;  // of the last real entry
  size_t n = 1;                       // table rows
  PairRows(const std::vector<Registers>& r, uint32_t op, const TraceIndex* ix) : regs(r), steps(ix ? ix->steps[op_slot(op)] : own) {
    if (!ix) {
      size_t cnt = 0;
      for (size_t i = 0; i + 1 < regs.size(); i++) cnt += regs[i].ci == op;
      own.reserve(cnt);
      for (size_t i = 0; i + 1 < regs.size(); i++) if (regs[i].ci == op) own.push_back((uint32_t)i);
    }
    const size_t cnt = steps.size();
    if (cnt) { const Registers& l = regs[steps.back() + 1]; last_clk = l.clk; last_ip = l.ip; n = next_pow2(2 * cnt) / 2; }
  }
  PairEntry dummy(size_t j) const { return {(uint32_t)(((uint64_t)last_clk + j) % P), last_ip, 0, 0, 0, 0, 0, 1}; }
  static PairEntry real(const Registers& r) { return {r.clk, r.ip, r.ci, r.ni, r.mp, r.mv, r.mvi, 0}; }
  PairEntry first(size_t row) const { return row < steps.size() ? real(regs[steps[row]]) : dummy(2 * (row - steps.size())); }
  PairEntry second(size_t row) const { return row < steps.size() ? real(regs[steps[row] + 1]) : dummy(2 * (row - steps.size()) + 1); }
};
// columns: clk ip ci ni mp mv mvi d next_ip next_mp next_mv  (ProcessorInstructionColumn)
inline Table instruction_op_table(const std::vector<Registers>& regs, uint32_t op, const TraceIndex* ix = nullptr) {
  PairRows t(regs, op, ix);
  ColVecs c = make_cols(11, t.n);
  for (size_t i = 0; i < t.n; i++) {
    const PairEntry a = t.first(i), b = t.second(i);
    c[0][i] = a.clk; c[1][i] = a.ip; c[2][i] = a.ci; c[3][i] = a.ni; c[4][i] = a.mp; c[5][i] = a.mv; c[6][i] = a.mvi;
    c[7][i] = a.d; c[8][i] = b.ip; c[9][i] = b.mp; c[10][i] = b.mv;
  }
  return finish(std::move(c));
}
// columns: clk ip ci ni mp mv mvi next_clk next_ip next_mp next_mv d is_mv_zero  (JumpColumn)
inline Table jump_table(const std::vector<Registers>& regs, uint32_t op, const TraceIndex* ix = nullptr) {
  PairRows t(regs, op, ix);
  ColVecs c = make_cols(13, t.n);
  for (size_t i = 0; i < t.n; i++) {
    const PairEntry a = t.first(i), b = t.second(i);
    c[0][i] = a.clk; c[1][i] = a.ip; c[2][i] = a.ci; c[3][i] = a.ni; c[4][i] = a.mp; c[5][i] = a.mv; c[6][i] = a.mvi;
    c[7][i] = b.clk; c[8][i] = b.ip; c[9][i] = b.mp; c[10][i] = b.mv; c[11][i] = a.d;
    c[12][i] = sb::m_sub(1, sb::m_mul(a.mv, a.mvi));
  }
  return finish(std::move(c));
}
// end_of_execution/table.rs:71-77,100-111: exactly one row with ci == 0
inline Table eoe_table(const std::vector<Registers>& regs) {
  std::vector<const Registers*> rows;
  for (auto& r : regs) if (r.ci == 0) rows.push_back(&r);
  if (rows.size() != 1) throw std::runtime_error("InvalidEndOfExecution");
  const Registers& r = *rows[0];
  ColVecs c = {{r.clk}, {r.ip}, {r.ci}, {r.ni}, {r.mp}, {r.mv}, {r.mvi}};
  return finish(std::move(c));
}

inline Table build_table(int k, const std::vector<Registers>& regs, const std::vector<uint32_t>& code, const TraceIndex* ix = nullptr,
                         uint32_t log_max_rows = 32) {
  switch (k) {
    case MEMORY: return memory_table(regs, log_max_rows);
    case INSTRUCTION: return instruction_table(regs, code);
    case PROGRAM: return program_table(code);
    case PROCESSOR: return processor_table(regs);
    case JNZ: return jump_table(regs, ']', ix);
    case JZ: return jump_table(regs, '[', ix);
    case EOE: return eoe_table(regs);
    default: return instruction_op_table(regs, opcode_of(k), ix);
  }
}

// The 13 tables are independent: one host thread each (the reference builds them one after the other, mod.rs:511-547).
inline std::vector<Table> build_tables(const std::vector<Registers>& regs, const std::vector<uint32_t>& code, uint32_t log_max_rows = 32) {
  std::vector<Table> t(N_COMPONENTS);
  std::vector<std::string> err(N_COMPONENTS);
  std::vector<std::thread> th;
  auto spawn = [&](int k, const TraceIndex* ix) {
    th.emplace_back([&, k, ix] {
      try { t[k] = build_table(k, regs, code, ix, log_max_rows); } catch (const std::exception& e) { err[k] = e.what(); }
    });
  };
  // the four tables that do not need the opcode index start first; this thread indexes the trace meanwhile
  for (int k : {(int)MEMORY, (int)INSTRUCTION, (int)PROCESSOR, (int)PROGRAM}) spawn(k, nullptr);
  TraceIndex ix(regs);
  for (int k = 0; k < N_COMPONENTS; k++)
    if (k != MEMORY && k != INSTRUCTION && k != PROCESSOR && k != PROGRAM) spawn(k, &ix);
  for (auto& x : th) x.join();
  for (auto& e : err) if (!e.empty()) throw std::runtime_error(e);
  return t;
}

}  // namespace sbf
