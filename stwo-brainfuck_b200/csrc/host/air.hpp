// The Brainfuck AIR: the thirteen `FrameworkEval::evaluate` bodies, written once over an evaluator concept `E`
// (Stwo's `EvalAtRow`) and instantiated for (i) the CUDA domain evaluator (air_kernels.cu), (ii) the host point
// evaluator (OODS check in prover and verifier), (iii) the host assert evaluator (tests).
// Constraint order = random-coefficient order, exactly as in the reference:
//   memory/component.rs:62-137, instruction/component.rs:65-142, program/component.rs:60-104,
//   processor/component.rs:79-153, processor/instructions/{plus,minus,left,right,input,output}_component.rs:62-121,
//   jump/{jump_if_not_zero,jump_if_zero}_component.rs:61-130, end_of_execution/component.rs:61-90.
// LogUp constraints (constraint_framework/logup.rs @ 31e8dbc) come after a component's own constraints.
//
// Evaluator concept:
//   using F  — mask value type (M31 on a domain row, QM31 at the OODS point); using EF — secure field value.
//              The AIR only uses + - * on F/EF and EF*F, so the CPU oracle can plug in its own arithmetic types.
//   EF ef(F);  EF ef_neg_one();       conversions
//   F next();                         next_trace_mask(): main-trace columns in *Column::index() order, offset 0
//   F is_first();                     get_preprocessed_column(IsFirst(log_size))
//   F cst(uint32_t);                  constant
//   void add(F) / void add(EF);       add_constraint
//   void relation(int rel, EF num, const F* vals, int n);   add_to_relation(RelationEntry::new(elements, num, vals))
//   void finalize_logup();
#pragma once
#include "../m31.cuh"
#include "air_ids.hpp"

namespace sbf {
using namespace sb;

struct Fm { uint32_t v; };
struct Fq { QM31 v; };
SB_HD Fm operator+(Fm a, Fm b) { return {m_add(a.v, b.v)}; }
SB_HD Fm operator-(Fm a, Fm b) { return {m_sub(a.v, b.v)}; }
SB_HD Fm operator*(Fm a, Fm b) { return {m_mul(a.v, b.v)}; }
SB_HD Fq operator+(Fq a, Fq b) { return {q_add(a.v, b.v)}; }
SB_HD Fq operator-(Fq a, Fq b) { return {q_sub(a.v, b.v)}; }
SB_HD Fq operator*(Fq a, Fq b) { return {q_mul(a.v, b.v)}; }
SB_HD Fq operator*(Fq a, Fm b) { return {q_mulm(a.v, b.v)}; }
SB_HD Fq to_ef(Fm a) { return {q_fromm(a.v)}; }
SB_HD Fq to_ef(Fq a) { return a; }

enum RelationId { REL_MEMORY = 0, REL_INSTRUCTION = 1, REL_PROCESSOR = 2 };
struct LookupElements { QM31 z; QM31 alpha_pow[7]; };   // LookupElements<N>::{z, alpha_powers}
struct InteractionElements { LookupElements rel[3]; };  // memory (3), instruction (3), processor (7)

// Relation::combine: sum_i alpha^i * v_i - z   (memory/table.rs:426-465, instruction/table.rs:391-430, processor/table.rs:393-432)
template <class F>
SB_HD Fq combine(const LookupElements& le, const F* vals, int n) {
  Fq acc{q_zero()};
  for (int i = 0; i < n; i++) acc = acc + Fq{le.alpha_pow[i]} * vals[i];
  return acc - Fq{le.z};
}

template <class E>
SB_HD void eval_memory(E& e) {
  typedef typename E::F F;
  F one = e.cst(1);
  F is_first = e.is_first();
  F clk = e.next(), mp = e.next(), mv = e.next(), d = e.next();
  F next_clk = e.next(), next_mp = e.next(), next_mv = e.next(), next_d = e.next();
  e.add(is_first * clk);
  e.add(is_first * mp);
  e.add(is_first * mv);
  e.add(is_first * d);
  e.add(d * (d - one));
  e.add(next_d * (next_d - one));
  e.add((next_mp - mp) * (next_mp - mp - one));
  e.add((next_mp - mp - one) * (next_clk - clk - one));
  e.add((next_mp - mp) * next_mv);
  e.add(d * (next_mp - mp));
  e.add(d * (next_mv - mv));
  F vals[3] = {clk, mp, mv};
  e.relation(REL_MEMORY, e.ef(d - one), vals, 3);
  e.finalize_logup();
}

template <class E>
SB_HD void eval_instruction(E& e) {
  typedef typename E::F F;
  F one = e.cst(1);
  F is_first = e.is_first();
  F ip = e.next(), ci = e.next(), ni = e.next(), d = e.next();
  F next_ip = e.next(), next_ci = e.next(), next_ni = e.next(), next_d = e.next();
  e.add(is_first * ip);
  e.add(d * (d - one));
  e.add(next_d * (next_d - one));
  e.add(d * ci);
  e.add(d * ni);
  e.add(next_d * next_ci);
  e.add(next_d * next_ni);
  e.add((next_ip - ip) * (next_ip - ip - one));
  e.add((next_ip - ip - one) * (next_ci - ci));
  e.add((next_ip - ip - one) * (next_ni - ni));
  F vals[3] = {ip, ci, ni};
  e.relation(REL_INSTRUCTION, e.ef(d - one), vals, 3);
  e.finalize_logup();
}

template <class E>
SB_HD void eval_program(E& e) {
  typedef typename E::F F;
  F one = e.cst(1);
  F is_first = e.is_first();
  F ip = e.next(), ci = e.next(), ni = e.next(), d = e.next();
  e.add(is_first * ip);
  e.add(d * (d - one));
  e.add(d * ci);
  e.add(d * ni);
  F vals[3] = {ip, ci, ni};
  e.relation(REL_INSTRUCTION, e.ef(one - d), vals, 3);
  e.finalize_logup();
}

template <class E>
SB_HD void eval_processor(E& e) {
  typedef typename E::F F;
  F one = e.cst(1);
  F is_first = e.is_first();
  F clk = e.next(), ip = e.next(), ci = e.next(), ni = e.next(), mp = e.next(), mv = e.next(), mvi = e.next();
  F d = e.next(), next_clk = e.next();
  e.add(is_first * clk);
  e.add(is_first * ip);
  e.add(is_first * mp);
  e.add(is_first * mv);
  e.add(mv * (mv * mvi - one));
  e.add(mvi * (mv * mvi - one));
  e.add(next_clk - clk - one);
  typename E::EF num = e.ef(one) - e.ef(d);
  F v7[7] = {clk, ip, ci, ni, mp, mv, mvi};
  e.relation(REL_PROCESSOR, num, v7, 7);
  F vi[3] = {ip, ci, ni};
  e.relation(REL_INSTRUCTION, num, vi, 3);
  F vm[3] = {clk, mp, mv};
  e.relation(REL_MEMORY, num, vm, 3);
  e.finalize_logup();
}

// `+ - < > , .`: columns clk ip ci ni mp mv mvi d next_ip next_mp next_mv
template <class E>
SB_HD void eval_instruction_op(E& e, int comp) {
  typedef typename E::F F;
  F one = e.cst(1);
  F clk = e.next(), ip = e.next(), ci = e.next(), ni = e.next(), mp = e.next(), mv = e.next(), mvi = e.next();
  F d = e.next(), next_ip = e.next(), next_mp = e.next(), next_mv = e.next();
  e.add(ci * (ci - e.cst(opcode_of(comp))));
  e.add(d * (d - one));
  e.add(d * mv);
  e.add(d * ci);
  e.add((one - d) * (next_ip - ip - one));
  switch (comp) {
    case PLUS: e.add(next_mp - mp); e.add((one - d) * (next_mv - mv - one)); break;
    case MINUS: e.add(next_mp - mp); e.add((one - d) * (next_mv - mv + one)); break;
    case LEFT: e.add((one - d) * (next_mp - mp + one)); break;
    case RIGHT: e.add((one - d) * (next_mp - mp - one)); break;
    case INPUT: e.add(next_mp - mp); break;
    case OUTPUT: e.add(next_mp - mp); e.add(next_mv - mv); break;
  }
  F v7[7] = {clk, ip, ci, ni, mp, mv, mvi};
  e.relation(REL_PROCESSOR, e.ef(d - one), v7, 7);
  e.finalize_logup();
}

// `[ ]`: columns clk ip ci ni mp mv mvi next_clk next_ip next_mp next_mv d is_mv_zero
template <class E>
SB_HD void eval_jump(E& e, int comp) {
  typedef typename E::F F;
  F one = e.cst(1), two = e.cst(2);
  F clk = e.next(), ip = e.next(), ci = e.next(), ni = e.next(), mp = e.next(), mv = e.next(), mvi = e.next();
  F next_clk = e.next(), next_ip = e.next(), next_mp = e.next(), next_mv = e.next(), d = e.next(), is_mv_zero = e.next();
  e.add(ci * (ci - e.cst(opcode_of(comp))));
  e.add(next_clk - clk - one);
  e.add(d * (d - one));
  e.add(d * mv);
  e.add(d * ci);
  if (comp == JNZ) e.add((d - one) * (is_mv_zero * (next_ip - ip - two) + mv * (next_ip - ni)));
  else e.add((d - one) * (mv * (next_ip - ip - two) + is_mv_zero * (next_ip - (ni + one))));
  e.add(next_mp - mp);
  e.add(next_mv - mv);
  F v7[7] = {clk, ip, ci, ni, mp, mv, mvi};
  e.relation(REL_PROCESSOR, e.ef(d - one), v7, 7);
  e.finalize_logup();
}

template <class E>
SB_HD void eval_eoe(E& e) {
  typedef typename E::F F;
  F clk = e.next(), ip = e.next(), ci = e.next(), ni = e.next(), mp = e.next(), mv = e.next(), mvi = e.next();
  e.add(ci);
  F v7[7] = {clk, ip, ci, ni, mp, mv, mvi};
  e.relation(REL_PROCESSOR, e.ef_neg_one(), v7, 7);
  e.finalize_logup();
}

template <class E>
SB_HD void eval_component(int comp, E& e) {
  switch (comp) {
    case MEMORY: eval_memory(e); break;
    case INSTRUCTION: eval_instruction(e); break;
    case PROGRAM: eval_program(e); break;
    case PROCESSOR: eval_processor(e); break;
    case JNZ: case JZ: eval_jump(e, comp); break;
    case EOE: eval_eoe(e); break;
    default: eval_instruction_op(e, comp); break;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// LogUp bookkeeping shared by every evaluator (LogupAtRow): one batch per fraction (finalize_logup()).
// The evaluator supplies:  EF ext_mask_cur(int batch)            — cumsum column of `batch` at offset 0
//                          void ext_mask_last(EF& prev, EF& cur) — last cumsum column at offsets [-1, 0]
//                          EF total_sum();  F is_first();  EF ef_zero();  void add(EF)
template <class E>
struct LogupState {
  typedef typename E::EF EF;
  EF num[3], den[3];
  int n = 0;
  SB_HD void push(EF nu, EF de) {
    if (n == 0) { num[0] = nu; den[0] = de; } else if (n == 1) { num[1] = nu; den[1] = de; } else { num[2] = nu; den[2] = de; }
    n++;
  }
  SB_HD void finalize(E& e) {
    // fixed trip count and select chains instead of run-time indices: `n` is a compile-time fact after inlining (1 or 3 pushes
    // per component), and the arrays then stay in registers on the device (the Processor kernel kept them in 696 bytes of
    // local memory and ran 5x slower per row than the others)
    EF prev_col = e.ef_zero();
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int b = 0; b < 2; b++) {
      if (b + 1 < n) {
        EF cur = e.ext_mask_cur(b);
        EF diff = cur - prev_col;
        prev_col = cur;
        e.add(diff * den[b] - num[b]);
      }
    }
    EF prev_row, cur;
    e.ext_mask_last(prev_row, cur);
    EF diff = cur - prev_row - prev_col;
    EF fixed = diff + e.total_sum() * e.is_first();
    const EF dl = n == 1 ? den[0] : (n == 2 ? den[1] : den[2]);
    const EF nl = n == 1 ? num[0] : (n == 2 ? num[1] : num[2]);
    e.add(fixed * dl - nl);
  }
};

// Relation::combine for the product's own field types.
SB_HD Fq combine_q(const LookupElements& le, const Fm* vals, int n) { return combine<Fm>(le, vals, n); }
SB_HD Fq combine_q(const LookupElements& le, const Fq* vals, int n) { return combine<Fq>(le, vals, n); }

}  // namespace sbf
