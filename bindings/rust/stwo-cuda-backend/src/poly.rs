//! `PolyOps` (upstream `core/poly/circle/ops.rs`; parity target `core/backend/simd/circle.rs` + `simd/fft/*`).
//!
//! Reference call sites: `precompute_twiddles` brainfuck_air/mod.rs:480-484; `interpolate_columns` through
//! `tree_builder.extend_evals` :497,550-562,690-702; `evaluate_polynomials` through `tree_builder.commit` :500,583,723;
//! `eval_at_point`, `evaluate`, `interpolate` inside `prover::prove` :732.

use std::ptr;
use std::rc::Rc;

use stwo_prover::core::backend::{Col, Column, ColumnOps};
use stwo_prover::core::circle::{CirclePoint, Coset};
use stwo_prover::core::fields::m31::BaseField;
use stwo_prover::core::fields::qm31::SecureField;
use stwo_prover::core::poly::circle::{CanonicCoset, CircleDomain, CircleEvaluation, CirclePoly, PolyOps};
use stwo_prover::core::poly::twiddles::TwiddleTree;
use stwo_prover::core::poly::BitReversedOrder;
use stwo_prover::core::ColumnVec;

use crate::column::CudaBaseColumn;
use crate::{ck, ctx, ffi, words, CudaBackend};

#[derive(Debug)]
struct RawTwiddles(*mut ffi::ScTwiddles);
impl Drop for RawTwiddles {
    fn drop(&mut self) {
        unsafe { ffi::sc_twiddles_free(ctx(), self.0) };
    }
}

/// `PolyOps::Twiddles`.  The library keeps the forward tree and its inverses behind one handle, so the `twiddles` and
/// `itwiddles` fields of `TwiddleTree<CudaBackend>` share it.
#[derive(Clone, Debug)]
pub struct CudaTwiddles(Rc<RawTwiddles>);
impl CudaTwiddles {
    pub fn handle(&self) -> *const ffi::ScTwiddles {
        self.0 .0
    }
}

fn is_canonic(domain: CircleDomain) -> bool {
    domain == CanonicCoset::new(domain.log_size()).circle_domain()
}

impl PolyOps for CudaBackend {
    type Twiddles = CudaTwiddles;

    /// Host permutation as in upstream `cpu/circle.rs`, then one upload.  Not on the reference's path.
    fn new_canonical_ordered(
        coset: CanonicCoset,
        values: Col<Self, BaseField>,
    ) -> CircleEvaluation<Self, BaseField, BitReversedOrder> {
        let domain = coset.circle_domain();
        let v = values.to_cpu();
        assert_eq!(v.len(), domain.size());
        let half = 1usize << (coset.log_size() - 1);
        let mut out = Vec::with_capacity(v.len());
        out.extend((0..half).map(|i| v[i << 1]));
        out.extend((0..half).map(|i| v[domain.size() - 1 - (i << 1)]));
        let mut col: CudaBaseColumn = out.into_iter().collect();
        <Self as ColumnOps<BaseField>>::bit_reverse_column(&mut col);
        CircleEvaluation::new(domain, col)
    }

    fn interpolate(
        eval: CircleEvaluation<Self, BaseField, BitReversedOrder>,
        twiddles: &TwiddleTree<Self>,
    ) -> CirclePoly<Self> {
        assert!(is_canonic(eval.domain), "CudaBackend transforms canonic domains only (all the reference uses)");
        let h = [eval.values.handle()];
        ck(unsafe { ffi::sc_interpolate(ctx(), h.as_ptr(), 1, twiddles.itwiddles.handle()) });
        CirclePoly::new(eval.values)
    }

    /// One batched call for all columns of a tree (mixed sizes): the library groups them by size and runs one launch
    /// set per size instead of one per column.
    fn interpolate_columns(
        columns: impl IntoIterator<Item = CircleEvaluation<Self, BaseField, BitReversedOrder>>,
        twiddles: &TwiddleTree<Self>,
    ) -> Vec<CirclePoly<Self>> {
        let evals: Vec<_> = columns.into_iter().collect();
        assert!(evals.iter().all(|e| is_canonic(e.domain)));
        let h: Vec<_> = evals.iter().map(|e| e.values.handle()).collect();
        ck(unsafe { ffi::sc_interpolate(ctx(), h.as_ptr(), h.len() as u32, twiddles.itwiddles.handle()) });
        evals.into_iter().map(|e| CirclePoly::new(e.values)).collect()
    }

    fn eval_at_point(poly: &CirclePoly<Self>, point: CirclePoint<SecureField>) -> SecureField {
        let (h, p) = ([poly.coeffs.handle()], words::point(point));
        let mut out = [0u32; 4];
        ck(unsafe { ffi::sc_eval_at_point(ctx(), h.as_ptr(), 1, p.as_ptr(), out.as_mut_ptr()) });
        words::to_qm31(&out)
    }

    fn extend(poly: &CirclePoly<Self>, log_size: u32) -> CirclePoly<Self> {
        assert!(log_size >= poly.log_size());
        let out = CudaBaseColumn::zeros(1 << log_size);
        ck(unsafe { ffi::sc_col_copy(ctx(), out.handle(), 0, poly.coeffs.handle(), 0, poly.coeffs.len() as u64) });
        CirclePoly::new(out)
    }

    fn evaluate(
        poly: &CirclePoly<Self>,
        domain: CircleDomain,
        twiddles: &TwiddleTree<Self>,
    ) -> CircleEvaluation<Self, BaseField, BitReversedOrder> {
        assert!(is_canonic(domain) && domain.log_size() >= poly.log_size());
        let (h, mut out) = ([poly.coeffs.handle()], [ptr::null_mut()]);
        let log_blowup = domain.log_size() - poly.log_size();
        ck(unsafe { ffi::sc_evaluate(ctx(), h.as_ptr(), 1, log_blowup, twiddles.twiddles.handle(), out.as_mut_ptr()) });
        CircleEvaluation::new(domain, CudaBaseColumn::from_handle(out[0]))
    }

    fn evaluate_polynomials(
        polynomials: &ColumnVec<CirclePoly<Self>>,
        log_blowup_factor: u32,
        twiddles: &TwiddleTree<Self>,
    ) -> Vec<CircleEvaluation<Self, BaseField, BitReversedOrder>> {
        let h: Vec<_> = polynomials.iter().map(|p| p.coeffs.handle()).collect();
        let mut out = vec![ptr::null_mut(); h.len()];
        ck(unsafe {
            ffi::sc_evaluate(ctx(), h.as_ptr(), h.len() as u32, log_blowup_factor, twiddles.twiddles.handle(), out.as_mut_ptr())
        });
        polynomials
            .iter()
            .zip(out)
            .map(|(p, o)| {
                let domain = CanonicCoset::new(p.log_size() + log_blowup_factor).circle_domain();
                CircleEvaluation::new(domain, CudaBaseColumn::from_handle(o))
            })
            .collect()
    }

    /// The reference passes `CanonicCoset::new(LOG_MAX_ROWS + log_blowup + 2).circle_domain().half_coset`
    /// (brainfuck_air/mod.rs:480-484).  The library's tree is rooted at the canonic half coset of the same size.
    fn precompute_twiddles(coset: Coset) -> TwiddleTree<Self> {
        assert_eq!(coset, CanonicCoset::new(coset.log_size() + 1).circle_domain().half_coset);
        let mut h = ptr::null_mut();
        ck(unsafe { ffi::sc_precompute_twiddles(ctx(), coset.log_size(), &mut h) });
        let tw = CudaTwiddles(Rc::new(RawTwiddles(h)));
        TwiddleTree { root_coset: coset, twiddles: tw.clone(), itwiddles: tw }
    }
}
