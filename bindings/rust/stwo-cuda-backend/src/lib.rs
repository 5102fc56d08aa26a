//! `CudaBackend` — Stwo's `Backend` trait family over `libstwo_cuda.so` (B200, sm_100a).
//!
//! The reference selects its backend by type: `SimdBackend` at
//! `crates/brainfuck_prover/src/brainfuck_air/mod.rs:56,399,480-497,732` and `components/mod.rs:42`.  This crate is the
//! type a maintainer puts there instead.  Every trait method is one call through the C ABI of `include/stwo_cuda.h`
//! (raw declarations: `ffi.rs`, generated from the headers); no arithmetic happens on the Rust side.
//!
//! Status: written against stwo-prover 0.1.1 @ 31e8dbc from its published trait definitions, UNCOMPILED — the build
//! image of this repository has no Rust toolchain.  The same entry points, with the same argument meaning, are
//! exercised from Python (`stwo-brainfuck_b200/__init__.py`) by the GPU parity tests.
//!
//! Modules follow the trait family:
//!  * `column`        — `Column<T>`, `ColumnOps<T>`, `FieldOps<F>`
//!  * `poly`          — `PolyOps` (twiddles, interpolate, evaluate, eval_at_point, extend)
//!  * `merkle`        — `MerkleOps<Blake2sMerkleHasher>`
//!  * `fri`           — `FriOps`
//!  * `quotients`     — `QuotientOps`
//!  * `accumulation`  — `AccumulationOps`
//!  * `grind`         — `GrindOps<Blake2sChannel>`
//!  * `gkr`           — `GkrOps` / `MleOps`: trait bounds only, never called by the reference
//!  * `framework`     — the two pieces Stwo writes concretely against `SimdBackend`: `ComponentProver` for
//!                      `FrameworkComponent<E>` and `LogupTraceGenerator`
//!  * `lanes`         — lane-repeated trace columns (1/16 of the upload and of the transform work)
//!  * `whole_proof`   — `sbf_prove` / `sbf_verify`: the entire `prove_brainfuck` on the device behind one call
//!  * `sharded`       — (feature `sharded`) the multi-GPU entry points

pub mod ffi;

pub mod accumulation;
pub mod column;
pub mod framework;
pub mod fri;
pub mod gkr;
pub mod grind;
pub mod lanes;
pub mod merkle;
pub mod poly;
pub mod quotients;
#[cfg(feature = "sharded")]
pub mod sharded;
pub mod whole_proof;

use std::cell::Cell;
use std::ffi::CStr;
use std::ptr;

use serde::{Deserialize, Serialize};
use stwo_prover::core::backend::{Backend, BackendForChannel};
use stwo_prover::core::vcs::blake2_merkle::Blake2sMerkleChannel;

pub use column::{CudaBaseColumn, CudaHashColumn, CudaSecureColumn};
pub use framework::{CudaComponent, CudaLogup, ComponentId};
pub use poly::CudaTwiddles;

/// The unit struct that takes `SimdBackend`'s place as the type parameter `B`.
#[derive(Copy, Clone, Debug, Default, Serialize, Deserialize)]
pub struct CudaBackend;

impl Backend for CudaBackend {}
impl BackendForChannel<Blake2sMerkleChannel> for CudaBackend {}

thread_local! {
    /// One context (device + stream + scratch + twiddle cache) per host thread.  `prove_brainfuck` is strictly
    /// sequential (brainfuck_air/mod.rs:471-735), so one thread drives one context; the ABI serialises calls on a
    /// context on its stream and is thread-safe across contexts.
    static CTX: Cell<*mut ffi::ScCtx> = const { Cell::new(ptr::null_mut()) };
}

/// Selects the CUDA device for the calling thread's context (default: device 0, created on first use).
/// `STWO_CUDA_DEVICE` overrides the default, which is how a torchrun-style launcher pins one process per GPU.
pub fn init(device: i32) {
    CTX.with(|c| {
        assert!(c.get().is_null(), "stwo-cuda-backend: context already created on this thread");
        let mut h = ptr::null_mut();
        ck(unsafe { ffi::sc_ctx_create(device, ptr::null_mut(), &mut h) });
        c.set(h);
    })
}

/// Destroys the calling thread's context.  Every column / twiddle handle created on it must have been dropped.
pub fn shutdown() {
    CTX.with(|c| {
        let h = c.replace(ptr::null_mut());
        if !h.is_null() {
            ck(unsafe { ffi::sc_ctx_destroy(h) });
        }
    })
}

pub(crate) fn ctx() -> *mut ffi::ScCtx {
    CTX.with(|c| {
        if c.get().is_null() {
            let device = std::env::var("STWO_CUDA_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
            let mut h = ptr::null_mut();
            // There is no CPU path: without a CUDA device this panics with the library's message.
            ck(unsafe { ffi::sc_ctx_create(device, ptr::null_mut(), &mut h) });
            c.set(h);
        }
        c.get()
    })
}

/// Status → panic, mirroring the `assert!`s Stwo's own backends use for the same preconditions (non-power-of-two
/// length, twiddle tree too small, mismatched column sizes).  Nothing unwinds across the ABI itself.
#[track_caller]
pub(crate) fn ck(status: i32) {
    if status != ffi::SC_OK {
        let msg = unsafe { CStr::from_ptr(ffi::sc_last_error()) }.to_string_lossy().into_owned();
        panic!("libstwo_cuda: status {status}: {msg}");
    }
}

/// Blocks until everything queued on the calling thread's context has finished.
pub fn synchronize() {
    ck(unsafe { ffi::sc_ctx_sync(ctx()) });
}

/// Frees every column created after `mark` that is still alive — what a caller does after catching a panic out of
/// `prove`, where `Drop` of half-built structures may not have run (`sc_ctx_mark` / `sc_ctx_release_since`).
pub struct LeakGuard(u64);
impl LeakGuard {
    pub fn new() -> Self {
        LeakGuard(unsafe { ffi::sc_ctx_mark(ctx()) })
    }
    pub fn release(self) {
        ck(unsafe { ffi::sc_ctx_release_since(ctx(), self.0) });
    }
}
impl Default for LeakGuard {
    fn default() -> Self {
        Self::new()
    }
}

/// `SecureField` ↔ the ABI's four words `{a, b, c, d}` of `(a + bi) + (c + di)u`.
pub(crate) mod words {
    use stwo_prover::core::circle::CirclePoint;
    use stwo_prover::core::fields::m31::BaseField;
    use stwo_prover::core::fields::qm31::SecureField;

    pub fn qm31(v: SecureField) -> [u32; 4] {
        let [a, b, c, d] = v.to_m31_array();
        [a.0, b.0, c.0, d.0]
    }
    pub fn to_qm31(w: &[u32]) -> SecureField {
        SecureField::from_m31_array([BaseField::from_u32_unchecked(w[0]), BaseField::from_u32_unchecked(w[1]),
                                     BaseField::from_u32_unchecked(w[2]), BaseField::from_u32_unchecked(w[3])])
    }
    /// `{x[4], y[4]}` as `sc_eval_at_point` and `sc_accumulate_quotients` take a point.
    pub fn point(p: CirclePoint<SecureField>) -> [u32; 8] {
        let (x, y) = (qm31(p.x), qm31(p.y));
        [x[0], x[1], x[2], x[3], y[0], y[1], y[2], y[3]]
    }
}
