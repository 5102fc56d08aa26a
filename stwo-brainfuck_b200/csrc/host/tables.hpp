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

struct Table {
  uint32_t log_size = 0;                    // column log size = log2(rows) + LOG_N_LANES
  std::vector<std::vector<uint32_t>> cols;  // cols[c][row], rows = 2^(log_size - 4)
  size_t rows() const { return cols.empty() ? 0 : cols[0].size(); }
};

inline size_t next_pow2(size_t n) { size_t p = 1; while (p < n) p <<= 1; return p; }
inline uint32_t ilog2_exact(size_t n) { uint32_t l = 0; while (((size_t)1 << l) < n) l++; return l; }

inline Table finish(std::vector<std::vector<uint32_t>> cols) {
  Table t;
  size_t rows = cols[0].size();
  if (rows == 0 || (rows & (rows - 1))) throw std::runtime_error("table length must be a non-zero power of two");
  t.log_size = ilog2_exact(rows) + LOG_N_LANES;
  t.cols = std::move(cols);
  return t;
}

// memory/table.rs:249-318 (sort by (mp,clk), fill clk gaps with dummies, pad) and :85-117 (pair with next, extra dummy)
inline Table memory_table(const std::vector<Registers>& regs) {
  struct E { uint32_t clk, mp, mv, d; };
  std::vector<E> src;
  for (auto& r : regs) src.push_back({r.clk, r.mp, r.mv, 0});
  std::stable_sort(src.begin(), src.end(), [](const E& a, const E& b) { return a.mp != b.mp ? a.mp < b.mp : a.clk < b.clk; });
  std::vector<E> t;
  if (!src.empty()) {
    const E* prev = &src[0];
    for (auto& e : src) {
      uint32_t next_clk = sb::m_add(prev->clk, 1);
      if (e.mp == prev->mp && e.clk > next_clk)
        for (uint32_t clk = next_clk; clk < e.clk; clk++) t.push_back({clk, prev->mp, prev->mv, 1});
      t.push_back(e);
      prev = &e;
    }
  }
  if (t.empty()) throw std::runtime_error("empty trace");
  E last = t.back();
  size_t pad = next_pow2(t.size()) - t.size();
  for (uint32_t i = 1; i <= pad; i++) t.push_back({sb::m_add(last.clk, i), last.mp, last.mv, 1});
  last = t.back();
  t.push_back({sb::m_add(last.clk, 1), last.mp, last.mv, 1});
  size_t n = t.size() - 1;
  std::vector<std::vector<uint32_t>> c(8, std::vector<uint32_t>(n));
  for (size_t i = 0; i < n; i++) {
    c[0][i] = t[i].clk; c[1][i] = t[i].mp; c[2][i] = t[i].mv; c[3][i] = t[i].d;
    c[4][i] = t[i + 1].clk; c[5][i] = t[i + 1].mp; c[6][i] = t[i + 1].mv; c[7][i] = t[i + 1].d;
  }
  return finish(std::move(c));
}

// instruction/table.rs:250-281 (program rows ++ trace rows, stable sort by (ip,clk), pad with dummy(last.ip)) and :85-110
inline Table instruction_table(const std::vector<Registers>& regs, const std::vector<uint32_t>& code) {
  struct E { uint32_t ip, ci, ni, d, clk; };
  std::vector<E> t;
  for (size_t i = 0; i < code.size(); i++) t.push_back({(uint32_t)i, code[i], i + 1 == code.size() ? 0 : code[i + 1], 0, 0});
  for (auto& r : regs) t.push_back({r.ip, r.ci, r.ni, 0, r.clk});
  std::stable_sort(t.begin(), t.end(), [](const E& a, const E& b) { return a.ip != b.ip ? a.ip < b.ip : a.clk < b.clk; });
  if (t.empty()) throw std::runtime_error("empty trace");
  uint32_t last_ip = t.back().ip;
  size_t pad = next_pow2(t.size()) - t.size();
  for (size_t i = 0; i < pad; i++) t.push_back({last_ip, 0, 0, 1, 0});
  t.push_back({t.back().ip, 0, 0, 1, 0});
  size_t n = t.size() - 1;
  std::vector<std::vector<uint32_t>> c(8, std::vector<uint32_t>(n));
  for (size_t i = 0; i < n; i++) {
    c[0][i] = t[i].ip; c[1][i] = t[i].ci; c[2][i] = t[i].ni; c[3][i] = t[i].d;
    c[4][i] = t[i + 1].ip; c[5][i] = t[i + 1].ci; c[6][i] = t[i + 1].ni; c[7][i] = t[i + 1].d;
  }
  return finish(std::move(c));
}

// program/table.rs:36-46,91-121
inline Table program_table(const std::vector<uint32_t>& code) {
  size_t n0 = code.size();
  if (n0 == 0) throw std::runtime_error("empty program");
  size_t n = next_pow2(n0);
  std::vector<std::vector<uint32_t>> c(4, std::vector<uint32_t>(n, 0));
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
  std::vector<std::vector<uint32_t>> c(9, std::vector<uint32_t>(n));
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
  std::vector<std::vector<uint32_t>> c(11, std::vector<uint32_t>(n));
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
  std::vector<std::vector<uint32_t>> c(13, std::vector<uint32_t>(n));
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
  std::vector<std::vector<uint32_t>> c = {{r.clk}, {r.ip}, {r.ci}, {r.ni}, {r.mp}, {r.mv}, {r.mvi}};
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
