// AIR layer on the device: LogUp interaction-trace generation and constraint-quotient evaluation on the blown-up domain
// for the 13 Brainfuck components.
//
// Replaces (a) the seven `interaction_trace_evaluation` functions of the reference (e.g. crates/brainfuck_prover/src/
// components/processor/table.rs:456-533, memory/table.rs:485-518, instruction/table.rs:456-491, program/table.rs:233-267,
// processor/instructions/table.rs:466-507, jump/table.rs:436-477, end_of_execution/table.rs:220-257), which drive Stwo's
// SimdBackend-only LogupTraceGenerator (write_frac / finalize_col / finalize_last), and (b) `impl ComponentProver<SimdBackend>
// for FrameworkComponent<E>`::evaluate_constraint_quotients_on_domain (constraint_framework/{component,simd_domain}.rs),
// entered from prover::prove at brainfuck_air/mod.rs:732.  Both instantiate the shared AIR definition in host/air.hpp.
//
// One thread per row; every column word is read once, coalesced.  Constraint evaluation is HBM-bound:
// R_eval * (4*(C_main + C_int + 1) + 32) bytes per component (read the masks, read-modify-write the accumulator).
#include "air_params.cuh"

namespace sb {
using namespace sbf;

// ---- LogUp generation: fraction per relation entry, column-cumulative (LogupColGenerator::finalize_col)
struct LogupGenEval {
  typedef Fm F;
  typedef Fq EF;
  const AirParams& p;
  uint32_t row;
  int col = 0, batch = 0;
  Fq cum{q_zero()};
  __device__ LogupGenEval(const AirParams& pp, uint32_t r) : p(pp), row(r) {}
  __device__ F next() { return {__ldg(p.main[col++] + (row >> p.main_shift))}; }
  __device__ F is_first() { return {0u}; }
  __device__ F cst(uint32_t c) { return {c}; }
  __device__ EF ef(F x) { return {q_fromm(x.v)}; }
  __device__ EF ef_neg_one() { return {q_fromm(P - 1)}; }
  __device__ void add(F) {}
  __device__ void add(EF) {}
  __device__ void relation(int rel, EF num, const F* vals, int n) {
    Fq den = combine_q(p.el.rel[rel], vals, n);
    cum = cum + num * Fq{q_inv(den.v)};
    // a null output pointer means "not wanted here" (multi-GPU: a rank only materialises the coordinate columns it owns)
    uint32_t* o0 = p.out[4 * batch + 0]; uint32_t* o1 = p.out[4 * batch + 1];
    uint32_t* o2 = p.out[4 * batch + 2]; uint32_t* o3 = p.out[4 * batch + 3];
    if (o0) o0[row] = cum.v.a.a;
    if (o1) o1[row] = cum.v.a.b;
    if (o2) o2[row] = cum.v.b.a;
    if (o3) o3[row] = cum.v.b.b;
    batch++;
  }
  __device__ void finalize_logup() {}
};

template <int COMP>
__global__ void __launch_bounds__(256) logup_gen_kernel(AirParams p) {
  uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= (1u << p.log_size)) return;
  LogupGenEval e(p, row);
  eval_component(COMP, e);
}

// ---- constraint evaluation on the LDE (SimdDomainEvaluator + ColumnAccumulator::accumulate)
struct DomainEval {
  typedef Fm F;
  typedef Fq EF;
  const AirParams& p;
  uint32_t row, prev_row;
  int col = 0, k = 0, pending = 0;
  // sum_k coeff_k * constraint_k in four 64-bit lanes: one IMAD.WIDE per coordinate and base-field constraint, a partial
  // fold after every third product (high word back in with weight 2^32 == 2: one IMAD.WIDE, result < 3 * 2^32, and
  // 3 * 2^62 + that stays below 2^64), one full reduction at the end
  unsigned long long acc[4] = {0, 0, 0, 0};
  LogupState<DomainEval> lg;
  __device__ DomainEval(const AirParams& pp, uint32_t r, uint32_t pr) : p(pp), row(r), prev_row(pr) {}
  __device__ F next() { return {__ldg(p.main[col++] + row)}; }
  __device__ F is_first() { return {__ldg(p.is_first + row)}; }
  __device__ F cst(uint32_t c) { return {c}; }
  __device__ EF ef(F x) { return {q_fromm(x.v)}; }
  __device__ EF ef_zero() { return {q_zero()}; }
  __device__ EF ef_neg_one() { return {q_fromm(P - 1)}; }
  __device__ EF total_sum() { return {p.total_sum}; }
  __device__ void add(F c) {
    const QM31 q = p.coeff[k++];
    acc[0] += (unsigned long long)q.a.a * c.v; acc[1] += (unsigned long long)q.a.b * c.v;
    acc[2] += (unsigned long long)q.b.a * c.v; acc[3] += (unsigned long long)q.b.b * c.v;
    if (++pending == 3) {
      pending = 0;
#pragma unroll
      for (int j = 0; j < 4; j++) acc[j] = (unsigned long long)(uint32_t)(acc[j] >> 32) * 2u + (uint32_t)acc[j];
    }
  }
  __device__ void add(EF c) {
    const QM31 t = q_mul(p.coeff[k++], c.v);
    acc[0] += t.a.a; acc[1] += t.a.b; acc[2] += t.b.a; acc[3] += t.b.b;
  }
  __device__ Fq row_result() const { return {q_make(m_red_wide(acc[0]), m_red_wide(acc[1]), m_red_wide(acc[2]), m_red_wide(acc[3]))}; }
  __device__ void relation(int rel, EF num, const F* vals, int n) { lg.push(num, combine_q(p.el.rel[rel], vals, n)); }
  __device__ EF ext_at(int b, uint32_t r) {
    return {q_make(__ldg(p.inter[4 * b] + r), __ldg(p.inter[4 * b + 1] + r), __ldg(p.inter[4 * b + 2] + r), __ldg(p.inter[4 * b + 3] + r))};
  }
  __device__ EF ext_mask_cur(int b) { return ext_at(b, row); }
  __device__ void ext_mask_last(EF& prev, EF& cur) {
    if (p.prev[0]) prev = {q_make(__ldg(p.prev[0] + row), __ldg(p.prev[1] + row), __ldg(p.prev[2] + row), __ldg(p.prev[3] + row))};
    else prev = ext_at(lg.n - 1, prev_row);
    cur = ext_at(lg.n - 1, row);
  }
  __device__ void finalize_logup() { lg.finalize(*this); }
};

template <int COMP>
__global__ void __launch_bounds__(256) constraint_kernel(AirParams p) {
  const uint32_t e = p.log_size + 1;
  uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;       // local row: index into the (possibly partial) columns
  if (row >= (p.n_rows ? p.n_rows : (1u << e))) return;
  const uint32_t grow = row + p.row_off;                       // row of the whole domain
  // offset_bit_reversed_circle_domain_index(grow, log_size, log_size + 1, -1); only used when the columns are whole
  uint32_t idx = __brev(grow) >> (32 - e), half = 1u << (e - 1);
  uint32_t pidx = idx < half ? ((idx + half - 1) & (half - 1)) : (((idx - half + 1) & (half - 1)) + half);
  uint32_t prev_row = __brev(pidx) >> (32 - e);
  DomainEval ev(p, row, prev_row);
  eval_component(COMP, ev);
  Fq res = ev.row_result() * Fm{p.denom_inv[grow >> p.log_size]};
  p.acc[0][row] = m_add(p.acc[0][row], res.v.a.a);
  p.acc[1][row] = m_add(p.acc[1][row], res.v.a.b);
  p.acc[2][row] = m_add(p.acc[2][row], res.v.b.a);
  p.acc[3][row] = m_add(p.acc[3][row], res.v.b.b);
}

template <int COMP>
static void launch_both(bool constraints, const AirParams& p, cudaStream_t st) {
  uint32_t n = (constraints && p.n_rows) ? p.n_rows : (1u << (p.log_size + (constraints ? 1 : 0)));
  uint32_t threads = n < 256 ? n : 256;
  if (constraints) constraint_kernel<COMP><<<(n + threads - 1) / threads, threads, 0, st>>>(p);
  else logup_gen_kernel<COMP><<<(n + threads - 1) / threads, threads, 0, st>>>(p);
  g_launch_count++;
}

int launch_air(bool constraints, int comp, const AirParams& p, cudaStream_t st) {
  switch (comp) {
    case MEMORY: launch_both<MEMORY>(constraints, p, st); break;
    case INSTRUCTION: launch_both<INSTRUCTION>(constraints, p, st); break;
    case PROGRAM: launch_both<PROGRAM>(constraints, p, st); break;
    case PROCESSOR: launch_both<PROCESSOR>(constraints, p, st); break;
    case JNZ: launch_both<JNZ>(constraints, p, st); break;
    case JZ: launch_both<JZ>(constraints, p, st); break;
    case INPUT: launch_both<INPUT>(constraints, p, st); break;
    case LEFT: launch_both<LEFT>(constraints, p, st); break;
    case MINUS: launch_both<MINUS>(constraints, p, st); break;
    case OUTPUT: launch_both<OUTPUT>(constraints, p, st); break;
    case PLUS: launch_both<PLUS>(constraints, p, st); break;
    case RIGHT: launch_both<RIGHT>(constraints, p, st); break;
    case EOE: launch_both<EOE>(constraints, p, st); break;
    default: return -1;
  }
  return (int)cudaGetLastError();
}

// host helper: 1 / coset_vanishing(CanonicCoset(log_size).coset, CanonicCoset(log_size+1).circle_domain().at(i)), i = 0, 1.
// For a canonic coset the translation cancels, so the vanishing polynomial is pi^(log_size-1)(x)  (SURVEY.md A.7).
void vanishing_denom_inv(uint32_t log_size, uint32_t out[2]) {
  Pt g{GEN_X, GEN_Y};
  auto at_index = [&](uint32_t idx) {
    Pt r{1u, 0u}, b = g;
    idx &= 0x7fffffffu;
    while (idx) { if (idx & 1u) r = p_add(r, b); b = p_dbl(b); idx >>= 1; }
    return r;
  };
  uint32_t e = log_size + 1;                       // half_odds(e-1): initial 2^(30-e), step 2^(32-e)
  for (uint32_t i = 0; i < 2; i++) {
    uint32_t idx = (1u << (30 - e)) + i * (1u << (32 - e));
    uint32_t x = at_index(idx).x;
    for (uint32_t k = 1; k < log_size; k++) x = m_sub(m_mul(2, m_sqr(x)), 1);
    out[i] = m_inv(x);
  }
}

}  // namespace sb
