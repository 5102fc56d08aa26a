//! The two constraint-framework pieces Stwo writes concretely against `SimdBackend`, so a new backend has to bring
//! its own (SURVEY.md §8b "off-trait pieces"):
//!
//!  * `impl ComponentProver<SimdBackend> for FrameworkComponent<E>` (upstream `constraint_framework/component.rs` +
//!    `simd_domain.rs`) → `CudaComponent<E>`, whose `evaluate_constraint_quotients_on_domain` is one
//!    `sc_eval_constraints` call.  The reference's `evaluate<E: EvalAtRow>` bodies (components/*/component.rs) stay as
//!    they are and keep serving `evaluate_constraint_quotients_at_point` (verifier and OODS side, host arithmetic);
//!    the device kernel restates the same constraint list per component and is checked against them by
//!    `assert_constraints`-style tests and by proof equality.
//!  * `LogupTraceGenerator` / `LogupColGenerator` (upstream `constraint_framework/logup.rs`), driven by the
//!    reference's seven `interaction_trace_evaluation` functions → `CudaLogup::generate`, one `sc_logup_generate` call.

use std::ptr;

use stwo_prover::constraint_framework::logup::LookupElements;
use stwo_prover::constraint_framework::{FrameworkComponent, FrameworkEval, PREPROCESSED_TRACE_IDX};
use stwo_prover::core::air::accumulation::{DomainEvaluationAccumulator, PointEvaluationAccumulator};
use stwo_prover::core::air::{Component, ComponentProver, Trace};
use stwo_prover::core::backend::Column;
use stwo_prover::core::circle::CirclePoint;
use stwo_prover::core::fields::m31::BaseField;
use stwo_prover::core::fields::qm31::SecureField;
use stwo_prover::core::pcs::TreeVec;
use stwo_prover::core::poly::circle::{CanonicCoset, CircleEvaluation};
use stwo_prover::core::poly::BitReversedOrder;
use stwo_prover::core::ColumnVec;

use crate::column::CudaBaseColumn;
use crate::fri::coord_handles;
use crate::{ck, ctx, ffi, words, CudaBackend};

/// Component numbering of the ABI = field order of `BrainfuckClaim` (brainfuck_air/mod.rs:79-93).
#[derive(Copy, Clone, Debug, PartialEq, Eq)]
#[repr(i32)]
pub enum ComponentId {
    Memory = 0,
    Instruction = 1,
    Program = 2,
    Processor = 3,
    JumpIfNotZero = 4,
    JumpIfZero = 5,
    Input = 6,
    Left = 7,
    Minus = 8,
    Output = 9,
    Plus = 10,
    Right = 11,
    EndOfExecution = 12,
}

impl ComponentId {
    /// `TraceColumn::count()` of the component's column enum: (main columns, LogUp columns).
    pub const fn columns(self) -> (usize, usize) {
        match self {
            ComponentId::Memory | ComponentId::Instruction => (8, 1),
            ComponentId::Program => (4, 1),
            ComponentId::Processor => (9, 3),
            ComponentId::JumpIfNotZero | ComponentId::JumpIfZero => (13, 1),
            ComponentId::EndOfExecution => (7, 1),
            _ => (11, 1),
        }
    }
}

/// The three relations' lookup elements in the ABI's layout: 3 × { z[4], alpha_powers[7][4] } words for the memory
/// (N = 3), instruction (N = 3) and processor (N = 7) relations (brainfuck_air/mod.rs:149-165); unused powers are zero.
#[derive(Clone, Debug)]
pub struct RelationElements(pub [u32; 96]);

impl RelationElements {
    /// The reference wraps each `LookupElements<N>` in a tuple struct with a private field (`MemoryElements`,
    /// `InstructionElements`, `ProcessorElements`: components/memory/table.rs:425-426 and siblings); the trait-level patch
    /// gives those wrappers a `pub fn lookup_elements(&self) -> &LookupElements<N>` accessor to call this with.
    pub fn new(memory: &LookupElements<3>, instruction: &LookupElements<3>, processor: &LookupElements<7>) -> Self {
        fn put<const N: usize>(dst: &mut [u32], e: &LookupElements<N>) {
            dst[..4].copy_from_slice(&words::qm31(e.z));
            for (i, a) in e.alpha_powers.iter().enumerate() {
                dst[4 + 4 * i..8 + 4 * i].copy_from_slice(&words::qm31(*a));
            }
        }
        let mut w = [0u32; 96];
        put(&mut w[0..32], memory);
        put(&mut w[32..64], instruction);
        put(&mut w[64..96], processor);
        RelationElements(w)
    }
}

/// Stands in for `LogupTraceGenerator` as the reference drives it.
pub struct CudaLogup;

impl CudaLogup {
    /// What `interaction_trace_evaluation(main_trace_eval, lookup_elements)` computes (e.g. components/processor/
    /// table.rs:456-533, components/memory/table.rs:485-518): for every relation entry the fraction
    /// `±(1 − d) / (Σ vᵢ·αⁱ − z)`, batched per column, inverted, accumulated, and the last column prefix-summed in
    /// coset order.  Returns the 4·k base columns of the k LogUp columns and the claimed sum.
    ///
    /// `log_repeat = 4` takes the lane-compact form of the main trace (one value per 16 rows, `lanes::upload`),
    /// `0` the full columns.
    pub fn generate(
        component: ComponentId,
        main_trace: &[&CudaBaseColumn],
        log_repeat: u32,
        elements: &RelationElements,
    ) -> (ColumnVec<CircleEvaluation<CudaBackend, BaseField, BitReversedOrder>>, SecureField) {
        let (n_main, n_logup) = component.columns();
        assert_eq!(main_trace.len(), n_main, "EmptyTrace / wrong column count");
        let h: Vec<_> = main_trace.iter().map(|c| c.handle()).collect();
        let mut out = vec![ptr::null_mut(); 4 * n_logup];
        let mut sum = [0u32; 4];
        ck(unsafe {
            ffi::sc_logup_generate(ctx(), component as i32, h.as_ptr(), n_main as u32, log_repeat, elements.0.as_ptr(),
                                   out.as_mut_ptr(), sum.as_mut_ptr())
        });
        let log_size = (main_trace[0].len() << log_repeat).ilog2();
        let domain = CanonicCoset::new(log_size).circle_domain();
        let evals = out.into_iter().map(|o| CircleEvaluation::new(domain, CudaBaseColumn::from_handle(o))).collect();
        (evals, words::to_qm31(&sum))
    }
}

/// `FrameworkComponent<E>` with a device-side `ComponentProver`.  Orphan rules forbid implementing
/// `ComponentProver<CudaBackend>` for the upstream type from this crate, hence the newtype; `Component` delegates.
pub struct CudaComponent<E: FrameworkEval> {
    pub inner: FrameworkComponent<E>,
    pub id: ComponentId,
    pub elements: RelationElements,
    pub total_sum: SecureField,
}

impl<E: FrameworkEval> CudaComponent<E> {
    pub fn new(inner: FrameworkComponent<E>, id: ComponentId, elements: RelationElements, total_sum: SecureField) -> Self {
        Self { inner, id, elements, total_sum }
    }
}

impl<E: FrameworkEval> Component for CudaComponent<E> {
    fn n_constraints(&self) -> usize {
        self.inner.n_constraints()
    }
    fn max_constraint_log_degree_bound(&self) -> u32 {
        self.inner.max_constraint_log_degree_bound()
    }
    fn trace_log_degree_bounds(&self) -> TreeVec<ColumnVec<u32>> {
        self.inner.trace_log_degree_bounds()
    }
    fn mask_points(&self, point: CirclePoint<SecureField>) -> TreeVec<ColumnVec<Vec<CirclePoint<SecureField>>>> {
        self.inner.mask_points(point)
    }
    fn preproccessed_column_indices(&self) -> ColumnVec<usize> {
        self.inner.preproccessed_column_indices()
    }
    fn evaluate_constraint_quotients_at_point(
        &self,
        point: CirclePoint<SecureField>,
        mask: &TreeVec<ColumnVec<Vec<SecureField>>>,
        evaluation_accumulator: &mut PointEvaluationAccumulator,
    ) {
        self.inner.evaluate_constraint_quotients_at_point(point, mask, evaluation_accumulator)
    }
}

impl<E: FrameworkEval + Sync> ComponentProver<CudaBackend> for CudaComponent<E> {
    /// accum += Σₖ coeffₖ · Cₖ(row) / vanishing(row) over `CanonicCoset(log_size + 1)`.
    ///
    /// The reference commits with blowup 1 and every component's constraint degree bound is `log_size + 1`
    /// (`max_constraint_log_degree_bound`, e.g. components/memory/component.rs:54-56), so the committed LDEs already
    /// live on the evaluation domain and are used as they are — the branch upstream calls "no need to extend".
    fn evaluate_constraint_quotients_on_domain(
        &self,
        trace: &Trace<'_, CudaBackend>,
        evaluation_accumulator: &mut DomainEvaluationAccumulator<CudaBackend>,
    ) {
        let log_size = self.inner.log_size();
        let eval_log = self.inner.max_constraint_log_degree_bound();
        assert_eq!(eval_log, log_size + 1, "the device evaluator is written for the reference's degree-2 bound");

        let evals = trace.evals.sub_tree(self.inner.trace_locations());
        let (n_main, n_logup) = self.id.columns();
        let main: Vec<_> = evals[1].iter().map(|e| e.values.handle()).collect();
        let inter: Vec<_> = evals[2].iter().map(|e| e.values.handle()).collect();
        assert_eq!((main.len(), inter.len()), (n_main, 4 * n_logup));
        assert!(evals[1].iter().chain(evals[2].iter()).all(|e| e.values.len() == 1 << eval_log));
        // the component's single preprocessed column: IsFirst(log_size)
        let pre = self.inner.preproccessed_column_indices();
        assert_eq!(pre.len(), 1);
        let is_first = trace.evals[PREPROCESSED_TRACE_IDX][pre[0]].values.handle();

        let [mut accum] = evaluation_accumulator.columns([(eval_log, self.inner.n_constraints())]);
        // upstream reverses the powers so that the first constraint carries the highest one
        accum.random_coeff_powers.reverse();
        let coeffs: Vec<u32> = accum.random_coeff_powers.iter().flat_map(|c| words::qm31(*c)).collect();
        let (acc, total) = (coord_handles(accum.col), words::qm31(self.total_sum));
        ck(unsafe {
            ffi::sc_eval_constraints(ctx(), self.id as i32, log_size, main.as_ptr(), n_main as u32, inter.as_ptr(),
                                     inter.len() as u32, is_first, self.elements.0.as_ptr(), total.as_ptr(),
                                     coeffs.as_ptr(), acc.as_ptr())
        });
    }
}
