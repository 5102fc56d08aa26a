//! Lane-repeated trace columns.
//!
//! The reference's seven `table.rs` files write every table row into all 16 SIMD lanes of a main-trace column
//! (`data[vec_row] = value.into()`, e.g. components/processor/table.rs:86-100), so a column of log size m + 4 holds 2^m
//! distinct values, each filling 16 consecutive rows of the bit-reversed evaluation.  Such a column's polynomial has
//! one non-zero coefficient in 16 and the first four FFT layers only scale and replicate; the library therefore
//! transforms, samples and hashes the 2^m values ("compact" columns) and returns LDEs that are bit-identical to
//! what `PolyOps` on the expanded column gives.  Only 1/16 of the main trace crosses PCIe.
//!
//! Two ways to use this from the reference's `trace_evaluation`:
//!  * stay on the trait path: `upload` the table-height vector, `broadcast16` it and hand the full column to
//!    `CircleEvaluation::new` — every `PolyOps` / `MerkleOps` call then sees an ordinary column;
//!  * keep the columns compact and call the `*_repeated` functions from a commitment scheme that knows about them
//!    (what the in-library prover behind `whole_proof::prove` does).

use std::ptr;

use stwo_prover::core::backend::Column;
use stwo_prover::core::circle::CirclePoint;
use stwo_prover::core::fields::m31::BaseField;
use stwo_prover::core::fields::qm31::SecureField;
use stwo_prover::core::poly::circle::{CanonicCoset, CircleEvaluation};
use stwo_prover::core::poly::twiddles::TwiddleTree;
use stwo_prover::core::poly::BitReversedOrder;
use stwo_prover::core::vcs::blake2_hash::Blake2sHash;

use crate::column::{CudaBaseColumn, CudaHashColumn};
use crate::{ck, ctx, ffi, words, CudaBackend};

/// log2 of the SIMD lane count the reference broadcasts over (`LOG_N_LANES`).
pub const LOG_REPEAT: u32 = 4;

/// Page-locked staging memory owned by the context: filling table columns here instead of in a `Vec` turns the upload
/// into a true asynchronous DMA beside kernels that are already queued.
pub struct PinnedBuf {
    ptr: *mut u32,
    len: usize,
}
impl PinnedBuf {
    pub fn new(len: usize) -> Self {
        let mut p = ptr::null_mut();
        ck(unsafe { ffi::sc_host_arena_alloc(ctx(), 4 * len as u64, &mut p) });
        PinnedBuf { ptr: p.cast(), len }
    }
    pub fn as_mut_slice(&mut self) -> &mut [u32] {
        unsafe { std::slice::from_raw_parts_mut(self.ptr, self.len) }
    }
    /// Queues the copy on the context's copy stream; the buffer must stay untouched until the next synchronising
    /// call (`crate::synchronize`, any read-back) — the arena is reset per proof, not per column.
    pub fn upload_async(&self) -> CudaBaseColumn {
        let mut h = ptr::null_mut();
        ck(unsafe { ffi::sc_col_from_host_async(ctx(), self.ptr, self.len as u64, &mut h) });
        CudaBaseColumn::from_handle(h)
    }
}
/// Releases every `PinnedBuf` of the context at once (between proofs).
pub fn reset_pinned_arena() {
    ck(unsafe { ffi::sc_host_arena_reset(ctx()) });
}

/// One value per table row (the column a `table.rs` builder would have broadcast).
pub fn upload(values: &[BaseField]) -> CudaBaseColumn {
    values.iter().copied().collect()
}

/// 16× expansion on the device: the `PackedBaseField` broadcast of `trace_evaluation`.
pub fn broadcast16(compact: &CudaBaseColumn) -> CudaBaseColumn {
    let mut h = ptr::null_mut();
    ck(unsafe { ffi::sc_col_broadcast16(ctx(), compact.handle(), &mut h) });
    CudaBaseColumn::from_handle(h)
}

fn handles(cols: &[&CudaBaseColumn]) -> Vec<*mut ffi::ScCol> {
    cols.iter().map(|c| c.handle()).collect()
}

/// `interpolate_columns` on compact columns: returns the compact coefficients (coefficient j = coefficient j·16 of the
/// full polynomial), inputs untouched.
pub fn interpolate_repeated(cols: &[&CudaBaseColumn], tw: &TwiddleTree<CudaBackend>) -> Vec<CudaBaseColumn> {
    let h = handles(cols);
    let mut out = vec![ptr::null_mut(); h.len()];
    ck(unsafe {
        ffi::sc_interpolate_repeated(ctx(), h.as_ptr(), h.len() as u32, LOG_REPEAT, tw.itwiddles.handle(), out.as_mut_ptr())
    });
    out.into_iter().map(CudaBaseColumn::from_handle).collect()
}

/// `evaluate_polynomials` from compact coefficients: ordinary full-length LDE columns.
pub fn evaluate_repeated(
    coeffs: &[&CudaBaseColumn],
    log_blowup: u32,
    tw: &TwiddleTree<CudaBackend>,
) -> Vec<CircleEvaluation<CudaBackend, BaseField, BitReversedOrder>> {
    let h = handles(coeffs);
    let mut out = vec![ptr::null_mut(); h.len()];
    ck(unsafe {
        ffi::sc_evaluate_repeated(ctx(), h.as_ptr(), h.len() as u32, LOG_REPEAT, log_blowup, tw.twiddles.handle(), out.as_mut_ptr())
    });
    coeffs
        .iter()
        .zip(out)
        .map(|(c, o)| {
            let log = (c.len() << LOG_REPEAT).ilog2() + log_blowup;
            CircleEvaluation::new(CanonicCoset::new(log).circle_domain(), CudaBaseColumn::from_handle(o))
        })
        .collect()
}

/// `eval_at_point` for a batch of (polynomial, point) pairs in one launch and one read-back; `log_repeats[i]` is 4 for
/// a compact polynomial and 0 for an ordinary one, so a whole tree's samples go out together.
pub fn eval_at_points(
    polys: &[&CudaBaseColumn],
    log_repeats: &[u32],
    points: &[CirclePoint<SecureField>],
) -> Vec<SecureField> {
    assert!(polys.len() == log_repeats.len() && polys.len() == points.len());
    let h = handles(polys);
    let p: Vec<u32> = points.iter().flat_map(|p| words::point(*p)).collect();
    let mut out = vec![0u32; 4 * h.len()];
    ck(unsafe {
        ffi::sc_eval_at_point_repeated(ctx(), h.as_ptr(), log_repeats.as_ptr(), h.len() as u32, p.as_ptr(), out.as_mut_ptr())
    });
    out.chunks_exact(4).map(words::to_qm31).collect()
}

/// `MerkleProver::commit` over full-length columns the caller promises are 16-repeated: the deepest four layers hash
/// each distinct node once.  Same layers and root as `merkle::commit_tree`.
pub fn commit_tree_repeated(columns: &[&CudaBaseColumn]) -> (Vec<CudaHashColumn>, Blake2sHash) {
    let h = handles(columns);
    let max_log = columns.iter().map(|c| c.len().ilog2()).max().expect("at least one column");
    let mut layers = vec![ptr::null_mut(); max_log as usize + 1];
    let (mut got_log, mut root) = (0u32, [0u32; 8]);
    ck(unsafe {
        ffi::sc_merkle_commit_repeated(ctx(), h.as_ptr(), h.len() as u32, LOG_REPEAT, layers.as_mut_ptr(), &mut got_log,
                                       root.as_mut_ptr())
    });
    let mut bytes = [0u8; 32];
    for (k, w) in root.iter().enumerate() {
        bytes[4 * k..4 * k + 4].copy_from_slice(&w.to_le_bytes());
    }
    (layers.into_iter().map(CudaHashColumn::from_handle).collect(), Blake2sHash(bytes))
}
