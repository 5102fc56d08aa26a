//! Smoke tests of `CudaBackend` against `CpuBackend` on a machine with a B200, a Rust toolchain and the Stwo checkout
//! (UNCOMPILED in this repository's image).  They mirror what tests/test_backend_gpu.py checks through ctypes: every trait
//! method equals the CPU backend's result on the same input, bit for bit.
use stwo_cuda_backend::CudaBackend;
use stwo_prover::core::backend::{Column, ColumnOps, CpuBackend};
use stwo_prover::core::fields::m31::BaseField;
use stwo_prover::core::fields::FieldOps;
use stwo_prover::core::poly::circle::{CanonicCoset, CircleEvaluation, PolyOps};
use stwo_prover::core::poly::BitReversedOrder;
use stwo_prover::core::vcs::blake2_merkle::Blake2sMerkleHasher;
use stwo_prover::core::vcs::ops::MerkleOps;

fn values(log: u32) -> Vec<BaseField> {
    // deterministic, full-range: x -> 3^x style walk in M31
    let mut v = BaseField::from_u32_unchecked(1);
    (0..1u32 << log).map(|_| { v = v * BaseField::from_u32_unchecked(1_234_567) + BaseField::from_u32_unchecked(89); v }).collect()
}

#[test]
fn bit_reverse_and_batch_inverse_match_cpu() {
    let host = values(12);
    let mut cpu = host.clone();
    <CpuBackend as ColumnOps<BaseField>>::bit_reverse_column(&mut cpu);
    let mut dev: <CudaBackend as ColumnOps<BaseField>>::Column = host.iter().copied().collect();
    <CudaBackend as ColumnOps<BaseField>>::bit_reverse_column(&mut dev);
    assert_eq!(dev.to_cpu(), cpu);

    let mut cpu_inv = vec![BaseField::from_u32_unchecked(0); host.len()];
    <CpuBackend as FieldOps<BaseField>>::batch_inverse(&host, &mut cpu_inv);
    let src: <CudaBackend as ColumnOps<BaseField>>::Column = host.iter().copied().collect();
    let mut dst = <CudaBackend as ColumnOps<BaseField>>::Column::zeros(host.len());
    <CudaBackend as FieldOps<BaseField>>::batch_inverse(&src, &mut dst);
    assert_eq!(dst.to_cpu(), cpu_inv);
}

#[test]
fn interpolate_evaluate_commit_match_cpu() {
    let log = 10;
    let domain = CanonicCoset::new(log).circle_domain();
    let host = values(log);
    let cpu_tw = CpuBackend::precompute_twiddles(CanonicCoset::new(log + 2).circle_domain().half_coset);
    let dev_tw = CudaBackend::precompute_twiddles(CanonicCoset::new(log + 2).circle_domain().half_coset);

    let cpu_poly = CircleEvaluation::<CpuBackend, BaseField, BitReversedOrder>::new(domain, host.clone()).interpolate_with_twiddles(&cpu_tw);
    let dev_poly = CircleEvaluation::<CudaBackend, BaseField, BitReversedOrder>::new(domain, host.iter().copied().collect())
        .interpolate_with_twiddles(&dev_tw);
    assert_eq!(dev_poly.coeffs.to_cpu(), cpu_poly.coeffs);

    let big = CanonicCoset::new(log + 1).circle_domain();
    let cpu_lde = cpu_poly.evaluate_with_twiddles(big, &cpu_tw);
    let dev_lde = dev_poly.evaluate_with_twiddles(big, &dev_tw);
    assert_eq!(dev_lde.values.to_cpu(), cpu_lde.values);

    let cpu_layer = <CpuBackend as MerkleOps<Blake2sMerkleHasher>>::commit_on_layer(log + 1, None, &[&cpu_lde.values]);
    let dev_layer = <CudaBackend as MerkleOps<Blake2sMerkleHasher>>::commit_on_layer(log + 1, None, &[&dev_lde.values]);
    assert_eq!(dev_layer.to_cpu(), cpu_layer);
}

#[test]
fn whole_proof_is_accepted_by_the_library_verifier() {
    let code = std::fs::read_to_string(concat!(env!("CARGO_MANIFEST_DIR"), "/../../../tests/golden/programs/hello_kakarot.bf")).unwrap();
    let proof = stwo_cuda_backend::whole_proof::prove(&code, b"", 20, 0).expect("prove");
    proof.verify().expect("verify");
    assert_eq!(proof.output(), b"Hello Kakarot World!\n");
    // the reference's own check, once this crate is a dependency of brainfuck_prover (bindings/rust/reference-cuda-feature.patch):
    //   let bf: BrainfuckProof<Blake2sMerkleHasher> = proof.parse().unwrap();  verify_brainfuck(bf).unwrap();
}
