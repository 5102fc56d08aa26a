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

  Machine(std::vector<uint32_t> code, std::vector<uint8_t> in, size_t ram_size = 30000)
      : program(std::move(code)), ram(ram_size, 0), input(std::move(in)) {}

  ~Machine() { recycle(trace); }

  void execute() {
    const size_t n = program.size();
    // Growing the trace from empty costs more than the run itself (reallocation + a page fault per 4 KB), so the buffer of
    // the previous proof is taken over when there is one; 2^20 + 1 rows is the most the AIR can take at LOG_MAX_ROWS 24
    // (the processor table holds one row per step) and untouched pages cost nothing.
    recycle(trace, /*take=*/true);
    trace.clear();
    trace.reserve(((size_t)1 << 20) + 1);
    const uint32_t* prog = program.data();
    uint32_t* cells = ram.data();
    const uint32_t ram_size = (uint32_t)ram.size();
    if (r.mp >= ram_size) throw std::runtime_error("memory pointer out of range");
    // the pointer is range-checked when it moves; every access in between is to a checked cell
    while (r.ip < n) {
      r.ci = prog[r.ip];
      r.ni = (r.ip == n - 1) ? 0 : prog[r.ip + 1];
      trace.push_back(r);
      bool early = false;
      switch (r.ci) {
        case '>': r.mp = sb::m_add(r.mp, 1); if (r.mp >= ram_size) throw std::runtime_error("memory pointer out of range"); break;
        case '<': r.mp = sb::m_sub(r.mp, 1); if (r.mp >= ram_size) throw std::runtime_error("memory pointer out of range"); break;
        case '+': cells[r.mp] = sb::m_add(cells[r.mp], 1); break;
        case '-': cells[r.mp] = sb::m_sub(cells[r.mp], 1); break;
        case ',':
          if (in_pos >= input.size()) throw std::runtime_error("input exhausted");
          cells[r.mp] = input[in_pos++];
          break;
        case '.': output.push_back((uint8_t)cells[r.mp]); break;
        case '[': {
          uint32_t arg = program.at(r.ip + 1);
          if (cells[r.mp] == 0) { r.ip = arg; early = true; } else r.ip += 1;
          break;
        }
        case ']': {
          uint32_t arg = program.at(r.ip + 1);
          if (cells[r.mp] != 0) { r.ip = arg - 1; early = true; } else r.ip += 1;
          break;
        }
        default: throw std::runtime_error("invalid instruction");
      }
      if (!early) r.mv = cells[r.mp];  // mvi is filled in afterwards (fill_inverses): it does not influence execution
      r.clk += 1;
      r.ip += 1;
    }
    r.ci = 0;
    r.ni = 0;
    trace.push_back(r);
    fill_inverses();
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
    const size_t n = trace.size(), chunk = (size_t)1 << 14;
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
