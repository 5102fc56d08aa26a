// Component ids, column/constraint counts and opcodes of the 13 AIR components (no STL: usable from device code).
// Order = BrainfuckClaim / provers() order, crates/brainfuck_prover/src/brainfuck_air/mod.rs:79-93,399-415.
#pragma once
#include <cstdint>
#include "../m31.cuh"

namespace sbf {

enum ComponentId { MEMORY = 0, INSTRUCTION, PROGRAM, PROCESSOR, JNZ, JZ, INPUT, LEFT, MINUS, OUTPUT, PLUS, RIGHT, EOE, N_COMPONENTS };
static const char* const COMPONENT_NAMES[N_COMPONENTS] = {"memory", "instruction", "program", "processor", "jump_if_not_zero",
    "jump_if_zero", "input_instruction", "left_instruction", "minus_instruction", "output_instruction", "plus_instruction",
    "right_instruction", "end_of_execution"};
// (main columns, LogUp columns) = TraceColumn::count() of each component
static const int N_MAIN_COLS[N_COMPONENTS] = {8, 8, 4, 9, 13, 13, 11, 11, 11, 11, 11, 11, 7};
static const int N_LOGUP_COLS[N_COMPONENTS] = {1, 1, 1, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1};
static const int N_CONSTRAINTS[N_COMPONENTS] = {12, 11, 5, 10, 9, 9, 7, 7, 8, 8, 8, 7, 2};
SB_HD uint32_t opcode_of(int comp) {  // ASCII codes, crates/brainfuck_vm/src/instruction.rs:65-76
  switch (comp) {
    case JNZ: return ']'; case JZ: return '['; case INPUT: return ','; case LEFT: return '<';
    case MINUS: return '-'; case OUTPUT: return '.'; case PLUS: return '+'; case RIGHT: return '>';
    default: return 0;
  }
}
constexpr uint32_t LOG_N_LANES = 4;


}  // namespace sbf
