// Host-side table builders: Vec<Registers> -> the 13 component tables (one value per table row; the 16-lane
// broadcast of the reference's `data[row] = x.into()` is applied on the device by sc_col_broadcast16).
// Follows crates/brainfuck_prover/src/components/{memory,instruction,program,processor}/table.rs and
// processor/instructions/{table.rs,jump/table.rs,end_of_execution/table.rs} — cited per function.
#pragma once
#include <algorithm>
#include <string>
#include <thread>
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
  template <class U> bool operator==(const ArenaAlloc<U>&) const { return true; }
  template <class U> bool operator!=(const ArenaAlloc<U>&) const { return false; }
};
typedef std::vector<uint32_t, ArenaAlloc<uint32_t>> ColVec;
typedef std::vector<ColVec> ColVecs;

struct Table {
  uint32_t log_size = 0;                    // column log size = log2(rows) + LOG_N_LANES
  ColVecs cols;  // cols[c][row], rows = 2^(log_size - 4)
  size_t rows() const { return cols.empty() ? 0 : cols[0].size(); }
};

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
inline std::vector<uint32_t> order_by_key(size_t n, uint32_t max_key, KeyFn key) {
  std::vector<uint32_t> idx(n);
  if (max_key < (1u << 22)) {
    std::vector<uint32_t> cnt((size_t)max_key + 2, 0);
    for (size_t i = 0; i < n; i++) cnt[key(i) + 1]++;
    for (size_t k = 1; k < cnt.size(); k++) cnt[k] += cnt[k - 1];
    for (size_t i = 0; i < n; i++) idx[cnt[key(i)]++] = (uint32_t)i;
  } else {
    for (size_t i = 0; i < n; i++) idx[i] = (uint32_t)i;
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return key(a) < key(b); });
  }
  return idx;
}

inline Table memory_table(const std::vector<Registers>& regs) {
  if (regs.empty()) throw std::runtime_error("empty trace");
  for (size_t i = 1; i < regs.size(); i++)
    if (regs[i].clk <= regs[i - 1].clk) throw std::runtime_error("trace is not in clk order");
  uint32_t max_mp = 0;
  for (auto& r : regs) max_mp = std::max(max_mp, r.mp);
  std::vector<uint32_t> ord = order_by_key(regs.size(), max_mp, [&](size_t i) { return regs[i].mp; });
  // rows after gap filling: per mp run, last.clk - first.clk + 1
  size_t rows = 0;
  for (size_t k = 0; k < ord.size(); k++) {
    const Registers& e = regs[ord[k]];
    if (k && regs[ord[k - 1]].mp == e.mp) rows += e.clk - regs[ord[k - 1]].clk; else rows += 1;
  }
  size_t n = next_pow2(rows);
  ColVecs c(8, ColVec(n));
  uint32_t *clk = c[0].data(), *mp = c[1].data(), *mv = c[2].data(), *d = c[3].data();
  size_t w = 0;
  for (size_t k = 0; k < ord.size(); k++) {
    const Registers& e = regs[ord[k]];
    if (k) {
      const Registers& p = regs[ord[k - 1]];
      if (p.mp == e.mp)
        for (uint32_t x = p.clk + 1; x < e.clk; x++) { clk[w] = x; mp[w] = p.mp; mv[w] = p.mv; d[w] = 1; w++; }
    }
    clk[w] = e.clk; mp[w] = e.mp; mv[w] = e.mv; d[w] = 0; w++;
  }
  uint32_t last_clk = clk[w - 1], last_mp = mp[w - 1], last_mv = mv[w - 1];
  for (uint32_t i = 1; w < n; i++, w++) { clk[w] = sb::m_add(last_clk, i); mp[w] = last_mp; mv[w] = last_mv; d[w] = 1; }
  for (int k = 0; k < 4; k++) {  // next_* = the following entry; the last row pairs with one more dummy
    std::copy(c[k].begin() + 1, c[k].end(), c[4 + k].begin());
  }
  c[4][n - 1] = sb::m_add(clk[n - 1], 1); c[5][n - 1] = mp[n - 1]; c[6][n - 1] = mv[n - 1]; c[7][n - 1] = 1;
  return finish(std::move(c));
}

// instruction/table.rs:250-281 (program rows ++ trace rows, stable sort by (ip,clk), pad with dummy(last.ip)) and :85-110
inline Table instruction_table(const std::vector<Registers>& regs, const std::vector<uint32_t>& code) {
  // program rows (clk 0) come first in the concatenation and the trace is in clk order, so a stable sort on ip alone is
  // the reference's stable sort on (ip, clk)
  const size_t np = code.size(), total = np + regs.size();
  if (total == 0) throw std::runtime_error("empty trace");
  for (size_t i = 1; i < regs.size(); i++)
    if (regs[i].clk <= regs[i - 1].clk) throw std::runtime_error("trace is not in clk order");
  auto ip_of = [&](size_t i) { return i < np ? (uint32_t)i : regs[i - np].ip; };
  uint32_t max_ip = 0;
  for (size_t i = 0; i < total; i++) max_ip = std::max(max_ip, ip_of(i));
  std::vector<uint32_t> ord = order_by_key(total, max_ip, ip_of);
  size_t n = next_pow2(total);
  ColVecs c(8, ColVec(n));
  for (size_t k = 0; k < total; k++) {
    size_t i = ord[k];
    if (i < np) { c[0][k] = (uint32_t)i; c[1][k] = code[i]; c[2][k] = i + 1 == np ? 0 : code[i + 1]; }
    else { const Registers& r = regs[i - np]; c[0][k] = r.ip; c[1][k] = r.ci; c[2][k] = r.ni; }
    c[3][k] = 0;
  }
  uint32_t last_ip = c[0][total - 1];
  for (size_t k = total; k < n; k++) { c[0][k] = last_ip; c[1][k] = 0; c[2][k] = 0; c[3][k] = 1; }
  for (int k = 0; k < 4; k++) std::copy(c[k].begin() + 1, c[k].end(), c[4 + k].begin());
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
  struct E { uint32_t clk, ip, ci, ni, mp, mv, mvi, d; };
  std::vector<E> t;
  for (auto& r : regs) t.push_back({r.clk, r.ip, r.ci, r.ni, r.mp, r.mv, r.mvi, 0});
  if (t.empty()) throw std::runtime_error("empty trace");
  E last = t.back();
  size_t pad = next_pow2(t.size()) - t.size();
  for (uint32_t i = 1; i <= pad; i++) t.push_back({sb::m_add(last.clk, i), last.ip, 0, 0, 0, 0, 0, 1});
  last = t.back();
  t.push_back({sb::m_add(last.clk, 1), last.ip, 0, 0, 0, 0, 0, 1});
  size_t n = t.size() - 1;
  ColVecs c(9, ColVec(n));
  for (size_t i = 0; i < n; i++) {
    c[0][i] = t[i].clk; c[1][i] = t[i].ip; c[2][i] = t[i].ci; c[3][i] = t[i].ni; c[4][i] = t[i].mp;
    c[5][i] = t[i].mv; c[6][i] = t[i].mvi; c[7][i] = t[i].d; c[8][i] = t[i + 1].clk;
  }
  return finish(std::move(c));
}

// processor/instructions/table.rs:293-328 and jump/table.rs:264-297: for every step with ci == op the pair (step, next step),
// padded in ENTRIES with dummy(last_clk + i, last_ip), i from 0, then chunked in twos; an empty table is one dummy row.
struct PairEntry { uint32_t clk, ip, ci, ni, mp, mv, mvi, d; };
inline std::vector<PairEntry> pair_entries(const std::vector<Registers>& regs, uint32_t op) {
  std::vector<PairEntry> t;
  for (size_t i = 0; i + 1 < regs.size(); i++)
    if (regs[i].ci == op)
      for (int k = 0; k < 2; k++) {
        const Registers& r = regs[i + k];
        t.push_back({r.clk, r.ip, r.ci, r.ni, r.mp, r.mv, r.mvi, 0});
      }
  uint32_t last_clk = t.empty() ? 0 : t.back().clk, last_ip = t.empty() ? 0 : t.back().ip;
  size_t len = t.size();
  size_t pad = (len == 0 ? 1 : next_pow2(len)) - len;
  for (uint32_t i = 0; i < pad; i++) t.push_back({sb::m_add(last_clk, i), last_ip, 0, 0, 0, 0, 0, 1});
  if (t.size() == 1) t.push_back({sb::m_add(t[0].clk, 1), t[0].ip, 0, 0, 0, 0, 0, 1});
  return t;
}
// columns: clk ip ci ni mp mv mvi d next_ip next_mp next_mv  (ProcessorInstructionColumn)
inline Table instruction_op_table(const std::vector<Registers>& regs, uint32_t op) {
  auto t = pair_entries(regs, op);
  size_t n = t.size() / 2;
  ColVecs c(11, ColVec(n));
  for (size_t i = 0; i < n; i++) {
    const PairEntry &a = t[2 * i], &b = t[2 * i + 1];
    c[0][i] = a.clk; c[1][i] = a.ip; c[2][i] = a.ci; c[3][i] = a.ni; c[4][i] = a.mp; c[5][i] = a.mv; c[6][i] = a.mvi;
    c[7][i] = a.d; c[8][i] = b.ip; c[9][i] = b.mp; c[10][i] = b.mv;
  }
  return finish(std::move(c));
}
// columns: clk ip ci ni mp mv mvi next_clk next_ip next_mp next_mv d is_mv_zero  (JumpColumn)
inline Table jump_table(const std::vector<Registers>& regs, uint32_t op) {
  auto t = pair_entries(regs, op);
  size_t n = t.size() / 2;
  ColVecs c(13, ColVec(n));
  for (size_t i = 0; i < n; i++) {
    const PairEntry &a = t[2 * i], &b = t[2 * i + 1];
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

inline Table build_table(int k, const std::vector<Registers>& regs, const std::vector<uint32_t>& code) {
  switch (k) {
    case MEMORY: return memory_table(regs);
    case INSTRUCTION: return instruction_table(regs, code);
    case PROGRAM: return program_table(code);
    case PROCESSOR: return processor_table(regs);
    case JNZ: return jump_table(regs, ']');
    case JZ: return jump_table(regs, '[');
    case EOE: return eoe_table(regs);
    default: return instruction_op_table(regs, opcode_of(k));
  }
}

// The 13 tables are independent: one host thread each (the reference builds them one after the other, mod.rs:511-547).
inline std::vector<Table> build_tables(const std::vector<Registers>& regs, const std::vector<uint32_t>& code) {
  std::vector<Table> t(N_COMPONENTS);
  std::vector<std::string> err(N_COMPONENTS);
  std::vector<std::thread> th;
  for (int k = 0; k < N_COMPONENTS; k++)
    th.emplace_back([&, k] {
      try { t[k] = build_table(k, regs, code); } catch (const std::exception& e) { err[k] = e.what(); }
    });
  for (auto& x : th) x.join();
  for (auto& e : err) if (!e.empty()) throw std::runtime_error(e);
  return t;
}

}  // namespace sbf
