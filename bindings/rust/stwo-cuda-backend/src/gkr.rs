//! `GkrOps` / `MleOps`: super-traits of `Backend` that the reference never calls (no GKR lookup in its AIR; SURVEY.md
//! §2.2 K16).  Present so that `CudaBackend: Backend` holds; every method panics with a message that says so.

use stwo_prover::core::fields::m31::BaseField;
use stwo_prover::core::fields::qm31::SecureField;
use stwo_prover::core::lookups::gkr_prover::{GkrMultivariatePolyOracle, GkrOps, Layer};
use stwo_prover::core::lookups::mle::{Mle, MleOps};
use stwo_prover::core::lookups::utils::UnivariatePoly;

use crate::CudaBackend;

const WHY: &str = "CudaBackend: GKR/MLE operations are outside stwo-brainfuck's proving path";

impl MleOps<BaseField> for CudaBackend {
    fn fix_first_variable(_mle: Mle<Self, BaseField>, _assignment: SecureField) -> Mle<Self, SecureField> {
        unimplemented!("{WHY}")
    }
}
impl MleOps<SecureField> for CudaBackend {
    fn fix_first_variable(_mle: Mle<Self, SecureField>, _assignment: SecureField) -> Mle<Self, SecureField> {
        unimplemented!("{WHY}")
    }
}
impl GkrOps for CudaBackend {
    fn gen_eq_evals(_y: &[SecureField], _v: SecureField) -> Mle<Self, SecureField> {
        unimplemented!("{WHY}")
    }
    fn next_layer(_layer: &Layer<Self>) -> Layer<Self> {
        unimplemented!("{WHY}")
    }
    fn sum_as_poly_in_first_variable(
        _h: &GkrMultivariatePolyOracle<'_, Self>,
        _claim: SecureField,
    ) -> UnivariatePoly<SecureField> {
        unimplemented!("{WHY}")
    }
}
