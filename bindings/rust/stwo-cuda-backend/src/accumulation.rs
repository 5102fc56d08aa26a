//! `AccumulationOps` (upstream `core/air/accumulation.rs`; parity target `core/backend/simd/accumulation.rs`).
//! Used by `DomainEvaluationAccumulator` inside `prover::prove` (brainfuck_air/mod.rs:732).

use stwo_prover::core::air::accumulation::AccumulationOps;
use stwo_prover::core::fields::qm31::SecureField;
use stwo_prover::core::fields::secure_column::SecureColumnByCoords;

use crate::fri::coord_handles;
use crate::{ck, ctx, ffi, words, CudaBackend};

impl AccumulationOps for CudaBackend {
    fn accumulate(column: &mut SecureColumnByCoords<Self>, other: &SecureColumnByCoords<Self>) {
        let (d, s) = (coord_handles(column), coord_handles(other));
        ck(unsafe { ffi::sc_accumulate(ctx(), d.as_ptr(), s.as_ptr()) });
    }

    /// A few hundred field multiplications at most (one power per constraint: 103 for the reference's 13 components);
    /// computed by the library on the host, no launch.
    fn generate_secure_powers(felt: SecureField, n_powers: usize) -> Vec<SecureField> {
        let (f, mut out) = (words::qm31(felt), vec![0u32; 4 * n_powers]);
        ck(unsafe { ffi::sc_secure_powers(f.as_ptr(), n_powers as u32, out.as_mut_ptr()) });
        out.chunks_exact(4).map(words::to_qm31).collect()
    }
}
