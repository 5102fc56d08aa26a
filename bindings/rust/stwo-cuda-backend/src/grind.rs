//! `GrindOps<Blake2sChannel>` (upstream `core/proof_of_work.rs`; parity target `core/backend/simd/grind.rs`):
//! the SMALLEST nonce whose Blake2s(digest ‖ nonce) has `pow_bits` trailing zero bits — the proof carries the nonce, so
//! "any valid nonce" would not be bit-identical.  The reference's PcsConfig uses pow_bits = 5 (brainfuck_air/mod.rs:473).

use stwo_prover::core::channel::Blake2sChannel;
use stwo_prover::core::proof_of_work::GrindOps;

use crate::{ck, ctx, ffi, CudaBackend};

impl GrindOps<Blake2sChannel> for CudaBackend {
    fn grind(channel: &Blake2sChannel, pow_bits: u32) -> u64 {
        let bytes = channel.digest().0;
        let digest: [u32; 8] = std::array::from_fn(|k| u32::from_le_bytes(bytes[4 * k..4 * k + 4].try_into().unwrap()));
        let mut nonce = 0u64;
        ck(unsafe { ffi::sc_grind(ctx(), digest.as_ptr(), pow_bits, &mut nonce) });
        nonce
    }
}
