//! `FriOps` (upstream `core/fri.rs`; parity target `core/backend/simd/fri.rs`).  Reached from `FriProver::commit`
//! inside `prover::prove` (brainfuck_air/mod.rs:732): one circle fold at the composition LDE size, then line folds
//! down to the last layer.

use std::ptr;

use num_traits::Zero;
use stwo_prover::core::backend::Column;
use stwo_prover::core::fields::m31::BaseField;
use stwo_prover::core::fields::qm31::SecureField;
use stwo_prover::core::fields::secure_column::SecureColumnByCoords;
use stwo_prover::core::fri::FriOps;
use stwo_prover::core::poly::circle::SecureEvaluation;
use stwo_prover::core::poly::line::LineEvaluation;
use stwo_prover::core::poly::twiddles::TwiddleTree;
use stwo_prover::core::poly::BitReversedOrder;

use crate::column::CudaBaseColumn;
use crate::{ck, ctx, ffi, words, CudaBackend};

pub(crate) fn coord_handles(c: &SecureColumnByCoords<CudaBackend>) -> [*mut ffi::ScCol; 4] {
    [c.columns[0].handle(), c.columns[1].handle(), c.columns[2].handle(), c.columns[3].handle()]
}

impl FriOps for CudaBackend {
    fn fold_line(eval: &LineEvaluation<Self>, alpha: SecureField, twiddles: &TwiddleTree<Self>) -> LineEvaluation<Self> {
        let log = eval.domain().log_size();
        assert!(log >= 1, "evaluation too small");
        let (src, a) = (coord_handles(&eval.values), words::qm31(alpha));
        let mut out = [ptr::null_mut(); 4];
        ck(unsafe { ffi::sc_fold_line(ctx(), src.as_ptr(), log, a.as_ptr(), twiddles.itwiddles.handle(), out.as_mut_ptr()) });
        let columns = out.map(CudaBaseColumn::from_handle);
        LineEvaluation::new(eval.domain().double(), SecureColumnByCoords { columns })
    }

    /// `dst <- dst·alpha² + fold(src)`
    fn fold_circle_into_line(
        dst: &mut LineEvaluation<Self>,
        src: &SecureEvaluation<Self, BitReversedOrder>,
        alpha: SecureField,
        twiddles: &TwiddleTree<Self>,
    ) {
        let log = src.domain.log_size();
        assert_eq!(src.len() >> 1, dst.len());
        let (s, d, a) = (coord_handles(&src.values), coord_handles(&dst.values), words::qm31(alpha));
        ck(unsafe {
            ffi::sc_fold_circle_into_line(ctx(), s.as_ptr(), log, a.as_ptr(), twiddles.itwiddles.handle(), d.as_ptr())
        });
    }

    /// Not on the reference's path (its quotient columns live on canonic domains, where λ = 0 and FRI never asks for
    /// the decomposition).  Restated from the definition on the host so the trait is complete: λ = (Σ first half −
    /// Σ second half) / n in bit-reversed order; g = f − λ on the first half, f + λ on the second.
    fn decompose(
        eval: &SecureEvaluation<Self, BitReversedOrder>,
    ) -> (SecureEvaluation<Self, BitReversedOrder>, SecureField) {
        let v = eval.values.to_vec();
        let (n, half) = (v.len(), v.len() / 2);
        let a: SecureField = v[..half].iter().fold(SecureField::zero(), |s, x| s + *x);
        let b: SecureField = v[half..].iter().fold(SecureField::zero(), |s, x| s + *x);
        let lambda = (a - b) / BaseField::from_u32_unchecked(n as u32);
        // upstream implements FromIterator<SecureField> for the CPU column only: split the coordinates by hand
        let mut coords: [Vec<BaseField>; 4] = Default::default();
        for (i, x) in v.iter().enumerate() {
            let g = if i < half { *x - lambda } else { *x + lambda };
            for (k, c) in g.to_m31_array().into_iter().enumerate() {
                coords[k].push(c);
            }
        }
        let columns = coords.map(|c| c.into_iter().collect::<CudaBaseColumn>());
        (SecureEvaluation::new(eval.domain, SecureColumnByCoords { columns }), lambda)
    }
}
