//! `QuotientOps::accumulate_quotients` (upstream `core/pcs/quotients.rs`; parity target `core/backend/simd/quotients.rs`).
//! Called by `compute_fri_quotients` inside `prover::prove` (brainfuck_air/mod.rs:732), once per distinct LDE size.

use std::ptr;

use stwo_prover::core::fields::m31::BaseField;
use stwo_prover::core::fields::qm31::SecureField;
use stwo_prover::core::fields::secure_column::SecureColumnByCoords;
use stwo_prover::core::pcs::quotients::{ColumnSampleBatch, QuotientOps};
use stwo_prover::core::poly::circle::{CanonicCoset, CircleDomain, CircleEvaluation, SecureEvaluation};
use stwo_prover::core::poly::BitReversedOrder;

use crate::column::CudaBaseColumn;
use crate::{ck, ctx, ffi, words, CudaBackend};

impl QuotientOps for CudaBackend {
    fn accumulate_quotients(
        domain: CircleDomain,
        columns: &[&CircleEvaluation<Self, BaseField, BitReversedOrder>],
        random_coeff: SecureField,
        sample_batches: &[ColumnSampleBatch],
        _log_blowup_factor: u32,
    ) -> SecureEvaluation<Self, BitReversedOrder> {
        // The SIMD backend evaluates on a sub-domain and extends by `log_blowup_factor`; the result is the same
        // column as direct evaluation on `domain`, which is what the kernel does (one pass over the LDE columns).
        assert_eq!(domain, CanonicCoset::new(domain.log_size()).circle_domain());
        let cols: Vec<_> = columns.iter().map(|c| c.values.handle()).collect();
        let (mut points, mut sizes, mut entry_cols, mut entry_vals) = (vec![], vec![], vec![], vec![]);
        for b in sample_batches {
            points.extend_from_slice(&words::point(b.point));
            sizes.push(b.columns_and_values.len() as u32);
            for (c, v) in &b.columns_and_values {
                entry_cols.push(*c as u32);
                entry_vals.extend_from_slice(&words::qm31(*v));
            }
        }
        let rc = words::qm31(random_coeff);
        let mut out = [ptr::null_mut(); 4];
        ck(unsafe {
            ffi::sc_accumulate_quotients(ctx(), domain.log_size(), cols.as_ptr(), cols.len() as u32, rc.as_ptr(),
                                         points.as_ptr(), sizes.as_ptr(), entry_cols.as_ptr(), entry_vals.as_ptr(),
                                         sample_batches.len() as u32, out.as_mut_ptr())
        });
        SecureEvaluation::new(domain, SecureColumnByCoords { columns: out.map(CudaBaseColumn::from_handle) })
    }
}
