// Brainfuck VM — input generator for the prover (north_star: "VM execution and trace filling stay on the host").
// Follows crates/brainfuck_vm/src/compiler.rs:13-37 (compile), machine.rs:141-238 (execute) and registers.rs:5-21.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <algorithm>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "../m31.cuh"

namespace sbf {
using sb::P;

struct Registers {
  uint32_t clk = 0, ip = 0, ci = 0, ni = 0, mp = 0, mv = 0, mvi = 0;
};

// What the table builders need to know about a trace before they allocate anything: the size of every table follows from
// these counts (memory/table.rs:259-283: one row per clk of every cell's life-span; instructions/table.rs:293-328: one entry
// pair per step of the opcode).  The VM keeps them while it runs; trace_stats() recomputes them for a trace that came from
// elsewhere.  Device-side table building (csrc/tables.cu) takes its sizes from here and cross-checks them.
struct TraceStats {
  uint64_t steps = 0;          // trace rows (the last one has ci = 0)
  uint64_t memory_rows = 0;    // sum over touched cells of (last clk - first clk + 1)
  uint32_t op_count[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // steps with a successor, by opcode: ] [ , < - . + >
  uint32_t zero_ci = 0;        // rows with ci == 0 (EndOfExecution needs exactly one)
  uint64_t zero_ci_index = 0;  // index of the first of them
  uint32_t max_mp = 0, max_ip = 0;
};
inline int op_slot_of(uint32_t ci) {
  switch (ci) { case ']': return 0; case '[': return 1; case ',': return 2; case '<': return 3; case '-': return 4; case '.': return 5;
                case '+': return 6; case '>': return 7; default: return -1; }
}
inline TraceStats trace_stats(const Registers* regs, size_t n) {
  TraceStats st;
  st.steps = n;
  std::vector<std::pair<uint32_t, uint32_t>> seen;  // mp -> last clk + 1, dense up to the largest mp
  for (size_t i = 0; i < n; i++) {
    const Registers& r = regs[i];
    if (i + 1 < n) { int sl = op_slot_of(r.ci); if (sl >= 0) st.op_count[sl]++; }
    if (r.ci == 0 && st.zero_ci++ == 0) st.zero_ci_index = i;
    st.max_mp = std::max(st.max_mp, r.mp); st.max_ip = std::max(st.max_ip, r.ip);
    if (r.mp >= seen.size()) seen.resize((size_t)r.mp + 1, {0u, 0u});
    auto& c = seen[r.mp];
    st.memory_rows += c.first ? (uint64_t)(r.clk - c.second) : 1;   // rows are in clk order
    c.first = 1; c.second = r.clk;
  }
  return st;
}

// `[` is followed by the index of the matching `]`'s argument slot, `]` by the index after the `[`'s slot.
inline std::vector<uint32_t> compile(const std::string& code) {
  std::vector<uint32_t> ins;
  std::vector<size_t> stack;
  for (unsigned char ch : code) {
    if (ch == ' ' || ch == '\n' || ch == '\t' || ch == '\r' || ch == '\f' || ch == '\v') continue;
    ins.push_back(ch);
    if (ch == '[') {
      ins.push_back(0);
      stack.push_back(ins.size() - 1);
    } else if (ch == ']') {
      if (stack.empty()) throw std::runtime_error("unbalanced ]");
      size_t start = stack.back();
      stack.pop_back();
      ins[start] = (uint32_t)ins.size();
      ins.push_back((uint32_t)(start + 1));
    }
  }
  return ins;
}

struct Machine {
  std::vector<uint32_t> program;
  std::vector<uint32_t> ram;
  std::vector<uint8_t> input, output;
  size_t in_pos = 0;
  std::vector<Registers> trace;
  Registers r;
  // Optional external trace buffer (the CUDA path hands in pinned memory so that the upload is a plain DMA): when set, rows
  // go to sink[0 .. sink_len) and `trace` stays empty.  `skip_inverses`: leave mvi = 0 for the device to fill (tables.cu).
  Registers* sink = nullptr;
  size_t sink_cap = 0, sink_len = 0;
  bool skip_inverses = false;
  TraceStats stats;
  const Registers* rows() const { return sink ? sink : trace.data(); }
  size_t n_rows() const { return sink ? sink_len : trace.size(); }

  Machine(std::vector<uint32_t> code, std::vector<uint8_t> in, size_t ram_size = 30000)
      : program(std::move(code)), ram(ram_size, 0), input(std::move(in)) {}

  ~Machine() { recycle(trace); }

  void execute() {
    // Growing the trace from empty costs more than the run itself (reallocation + a page fault per 4 KB), so the buffer of
    // the previous proof is taken over when there is one; 2^20 + 1 rows is the most the AIR can take at LOG_MAX_ROWS 24
    // (the processor table holds one row per step) and untouched pages cost nothing.
    if (!sink) {
      recycle(trace, /*take=*/true);
      trace.clear();
      trace.reserve(((size_t)1 << 20) + 1);
    }
    sink_len = 0;
    stats = TraceStats();
    if (sink) run<true>(); else run<false>();
    if (!skip_inverses) fill_inverses();
  }

  // The interpreter loop.  Every register lives in a local for the whole run (the loop is the critical path of a proof on
  // eight GPUs: the device waits for the trace once the program-independent phase is done); the statistics the table
  // builders need (TraceStats) are kept on the way instead of in a second pass over the trace.
  template <bool SINK>
  void run() {
    const size_t n = program.size();
    const uint32_t* prog = program.data();
    uint32_t* cells = ram.data();
    const uint32_t ram_size = (uint32_t)ram.size();
    if (r.mp >= ram_size) throw std::runtime_error("memory pointer out of range");
    std::vector<uint32_t> last_seen(ram.size(), 0);  // clk + 1 of the last access of every cell (0: untouched)
    uint32_t* seen = last_seen.data();
    uint64_t mem_rows = 0;
    uint32_t opc[256] = {0};
    uint32_t clk = r.clk, ip = r.ip, mp = r.mp, mv = r.mv, max_mp = r.mp, max_ip = 0;
    const uint32_t mvi0 = r.mvi;
    Registers* out = sink;
    size_t len = 0;
    const size_t cap = sink_cap;
    // the pointer is range-checked when it moves; every access in between is to a checked cell
    while (ip < n) {
      const uint32_t ci = prog[ip], ni = (ip == n - 1) ? 0 : prog[ip + 1];
      if (SINK) {
        if (len == cap) throw std::runtime_error("component too large: processor (the trace does not fit the buffer)");
        out[len++] = Registers{clk, ip, ci, ni, mp, mv, mvi0};
      } else {
        trace.push_back(Registers{clk, ip, ci, ni, mp, mv, mvi0});
      }
      const uint32_t l = seen[mp];
      mem_rows += l ? (uint64_t)(clk + 1 - l) : 1;
      seen[mp] = clk + 1;
      max_ip = ip > max_ip ? ip : max_ip;
      opc[ci & 255u]++;
      bool early = false;
      switch (ci) {
        case '>': mp = sb::m_add(mp, 1); if (mp >= ram_size) throw std::runtime_error("memory pointer out of range"); max_mp = mp > max_mp ? mp : max_mp; break;
        case '<': mp = sb::m_sub(mp, 1); if (mp >= ram_size) throw std::runtime_error("memory pointer out of range"); max_mp = mp > max_mp ? mp : max_mp; break;
        case '+': cells[mp] = sb::m_add(cells[mp], 1); break;
        case '-': cells[mp] = sb::m_sub(cells[mp], 1); break;
        case ',':
          if (in_pos >= input.size()) throw std::runtime_error("input exhausted");
          cells[mp] = input[in_pos++];
          break;
        case '.': output.push_back((uint8_t)cells[mp]); break;
        case '[': {
          if (ip + 1 >= n) throw std::out_of_range("jump without an argument");
          const uint32_t arg = prog[ip + 1];
          if (cells[mp] == 0) { ip = arg; early = true; } else ip += 1;
          break;
        }
        case ']': {
          if (ip + 1 >= n) throw std::out_of_range("jump without an argument");
          const uint32_t arg = prog[ip + 1];
          if (cells[mp] != 0) { ip = arg - 1; early = true; } else ip += 1;
          break;
        }
        default: throw std::runtime_error("invalid instruction");
      }
      if (!early) mv = cells[mp];  // mvi is filled in afterwards (fill_inverses): it does not influence execution
      clk += 1;
      ip += 1;
    }
    // the final row: ci = ni = 0
    {
      const Registers last{clk, ip, 0, 0, mp, mv, mvi0};
      if (SINK) {
        if (len == cap) throw std::runtime_error("component too large: processor (the trace does not fit the buffer)");
        out[len++] = last;
      } else trace.push_back(last);
      const uint32_t l = seen[mp];
      mem_rows += l ? (uint64_t)(clk + 1 - l) : 1;
      max_ip = ip > max_ip ? ip : max_ip;
    }
    if (SINK) sink_len = len;
    r.clk = clk; r.ip = ip; r.ci = 0; r.ni = 0; r.mp = mp; r.mv = mv;
    stats.steps = n_rows();
    stats.memory_rows = mem_rows;
    static const char ops[8] = {']', '[', ',', '<', '-', '.', '+', '>'};
    for (int k = 0; k < 8; k++) stats.op_count[k] = opc[(unsigned char)ops[k]];   // every counted step has a successor (the final row)
    stats.zero_ci = 1 + opc[0];          // program words are never 0 (compile() keeps instruction bytes only; jump targets are >= 1)
    stats.zero_ci_index = n_rows() - 1;
    if (opc[0]) {                        // cannot happen with compile()'s output; stay exact if it ever does
      const Registers* p = rows();
      for (size_t i = 0; i < n_rows(); i++) if (p[i].ci == 0) { stats.zero_ci_index = i; break; }
    }
    stats.max_mp = max_mp; stats.max_ip = max_ip;
  }

  // One spare trace buffer per process: a finished machine leaves its (already faulted-in) buffer for the next one.
  static void recycle(std::vector<Registers>& v, bool take = false) {
    static std::mutex mu;
    static std::vector<Registers> spare;
    std::lock_guard<std::mutex> lk(mu);
    if (take) { if (spare.capacity() > v.capacity()) v.swap(spare); }
    else if (v.capacity() > spare.capacity()) { v.clear(); spare.swap(v); }
  }

  // mvi = mv^-1 (0 for 0) for every row (machine.rs:224-228 computes it per step).  A jump keeps the previous mv/mvi pair,
  // which the pass below reproduces because it inverts whatever mv the row recorded.  Montgomery's trick over chunks of the
  // trace (3 multiplications per row and one inversion per chunk), chunks on a few host threads.
  void fill_inverses() {
    const size_t n = n_rows(), chunk = (size_t)1 << 14;
    Registers* const trace = sink ? sink : this->trace.data();
    auto work = [&](size_t lo, size_t hi) {
      std::vector<uint32_t> pre(chunk);
      for (size_t c0 = lo; c0 < hi; c0 += chunk) {
        const size_t c1 = c0 + chunk < hi ? c0 + chunk : hi;
        uint32_t run = 1;
        for (size_t i = c0; i < c1; i++) { pre[i - c0] = run; if (trace[i].mv) run = sb::m_mul(run, trace[i].mv); }
        uint32_t inv = sb::m_inv(run);
        for (size_t i = c1; i-- > c0;) {
          const uint32_t v = trace[i].mv;
          if (v) { trace[i].mvi = sb::m_mul(inv, pre[i - c0]); inv = sb::m_mul(inv, v); } else trace[i].mvi = 0;
        }
      }
    };
    const unsigned T = n >= ((size_t)1 << 17) ? 4 : 1;
    if (T == 1) { work(0, n); return; }
    std::vector<std::thread> th;
    const size_t per = ((n + T - 1) / T + chunk - 1) / chunk * chunk;
    for (unsigned t = 1; t < T; t++) { size_t lo = std::min(n, t * per), hi = std::min(n, lo + per); if (lo < hi) th.emplace_back(work, lo, hi); }
    work(0, std::min(n, per));
    for (auto& x : th) x.join();
  }

 private:
  uint32_t& at(uint32_t mp) {
    if (mp >= ram.size()) throw std::runtime_error("memory pointer out of range");
    return ram[mp];
  }
};

}  // namespace sbf
