//! `MerkleOps<Blake2sMerkleHasher>` (upstream `core/vcs/ops.rs`; parity target `core/backend/simd/blake2s.rs`).
//!
//! `MerkleProver::commit` calls `commit_on_layer` once per log size from the largest down; the reference reaches it
//! from the three `tree_builder.commit` calls (brainfuck_air/mod.rs:500,583,723), the composition commit and every FRI
//! layer inside `prover::prove` (:732).  Node = Blake2s(prev_left ‖ prev_right ‖ column words), 32-byte digest.

use std::ptr;

use stwo_prover::core::backend::{Col, Column};
use stwo_prover::core::fields::m31::BaseField;
use stwo_prover::core::vcs::blake2_hash::Blake2sHash;
use stwo_prover::core::vcs::blake2_merkle::Blake2sMerkleHasher;
use stwo_prover::core::vcs::ops::MerkleOps;

use crate::column::{CudaBaseColumn, CudaHashColumn};
use crate::{ck, ctx, ffi, CudaBackend};

impl MerkleOps<Blake2sMerkleHasher> for CudaBackend {
    fn commit_on_layer(
        log_size: u32,
        prev_layer: Option<&Col<Self, Blake2sHash>>,
        columns: &[&Col<Self, BaseField>],
    ) -> Col<Self, Blake2sHash> {
        let h: Vec<_> = columns.iter().map(|c| c.handle()).collect();
        let prev = prev_layer.map_or(ptr::null(), |p| p.handle() as *const ffi::ScCol);
        let mut out = ptr::null_mut();
        ck(unsafe { ffi::sc_merkle_commit_layer(ctx(), log_size, prev, h.as_ptr(), h.len() as u32, &mut out) });
        CudaHashColumn::from_handle(out)
    }
}

/// Whole-tree commit in one call (`MerkleProver::commit` upstream walks the sizes on the host): fuses the top layers
/// into one launch and returns every layer, `layers[k]` = layer of log size `k`.  A caller that owns its commitment
/// scheme (a Stwo fork) uses this instead of the per-layer trait method; the trait method above is bit-identical.
pub fn commit_tree(columns: &[&CudaBaseColumn]) -> (Vec<CudaHashColumn>, Blake2sHash) {
    let h: Vec<_> = columns.iter().map(|c| c.handle()).collect();
    let max_log = columns.iter().map(|c| c.len().ilog2()).max().expect("at least one column");
    let mut layers = vec![ptr::null_mut(); max_log as usize + 1];
    let (mut got_log, mut root) = (0u32, [0u32; 8]);
    ck(unsafe { ffi::sc_merkle_commit(ctx(), h.as_ptr(), h.len() as u32, layers.as_mut_ptr(), &mut got_log, root.as_mut_ptr()) });
    debug_assert_eq!(got_log, max_log);
    let mut bytes = [0u8; 32];
    for (k, w) in root.iter().enumerate() {
        bytes[4 * k..4 * k + 4].copy_from_slice(&w.to_le_bytes());
    }
    (layers.into_iter().map(CudaHashColumn::from_handle).collect(), Blake2sHash(bytes))
}

/// Batched `Column::at` for decommitment.  Upstream's `MerkleProver::decommit` and `FriProver::decommit` read single
/// elements with `at(i)`; on a device backend that is thousands of 4- and 32-byte copies.  One gather kernel and one
/// read-back instead: `out[i] = cols[i][offsets[i] .. offsets[i] + words]`.
pub fn gather(cols: &[*mut ffi::ScCol], offsets: &[u64], words: u32) -> Vec<u32> {
    assert_eq!(cols.len(), offsets.len());
    let mut out = vec![0u32; cols.len() * words as usize];
    ck(unsafe { ffi::sc_gather(ctx(), cols.as_ptr(), offsets.as_ptr(), cols.len() as u32, words, out.as_mut_ptr()) });
    out
}
