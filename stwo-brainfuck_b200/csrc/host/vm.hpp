// Brainfuck VM — input generator for the prover (north_star: "VM execution and trace filling stay on the host").
// Follows crates/brainfuck_vm/src/compiler.rs:13-37 (compile), machine.rs:141-238 (execute) and registers.rs:5-21.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
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

  void execute() {
    const size_t n = program.size();
    // growing from empty costs more than the run itself (reallocation + page faults); 2^20 + 1 rows is the most the AIR can
    // take at LOG_MAX_ROWS 24 (the processor table holds one row per step), untouched pages cost nothing
    trace.reserve(((size_t)1 << 20) + 1);
    while (r.ip < n) {
      r.ci = program[r.ip];
      r.ni = (r.ip == n - 1) ? 0 : program[r.ip + 1];
      trace.push_back(r);
      bool early = false;
      switch (r.ci) {
        case '>': r.mp = sb::m_add(r.mp, 1); break;
        case '<': r.mp = sb::m_sub(r.mp, 1); break;
        case '+': at(r.mp) = sb::m_add(at(r.mp), 1); break;
        case '-': at(r.mp) = sb::m_sub(at(r.mp), 1); break;
        case ',':
          if (in_pos >= input.size()) throw std::runtime_error("input exhausted");
          at(r.mp) = input[in_pos++];
          break;
        case '.': output.push_back((uint8_t)at(r.mp)); break;
        case '[': {
          uint32_t arg = program.at(r.ip + 1);
          if (at(r.mp) == 0) { r.ip = arg; early = true; } else r.ip += 1;
          break;
        }
        case ']': {
          uint32_t arg = program.at(r.ip + 1);
          if (at(r.mp) != 0) { r.ip = arg - 1; early = true; } else r.ip += 1;
          break;
        }
        default: throw std::runtime_error("invalid instruction");
      }
      if (!early) {
        r.mv = at(r.mp);
        r.mvi = inverse(r.mv);
      }
      r.clk += 1;
      r.ip += 1;
    }
    r.ci = 0;
    r.ni = 0;
    trace.push_back(r);
  }

 private:
  // mv.inverse() is needed at every step (machine.rs:224-228); cell values are small in practice, so remember them
  std::vector<uint32_t> inv_cache = std::vector<uint32_t>(1 << 16, 0);
  uint32_t inverse(uint32_t v) {
    if (v == 0) return 0;
    if (v < inv_cache.size()) {
      if (!inv_cache[v]) inv_cache[v] = sb::m_inv(v);
      return inv_cache[v];
    }
    return sb::m_inv(v);
  }
  uint32_t& at(uint32_t mp) {
    if (mp >= ram.size()) throw std::runtime_error("memory pointer out of range");
    return ram[mp];
  }
};

}  // namespace sbf
