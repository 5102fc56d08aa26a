//! The whole of `prove_brainfuck` behind one call.
//!
//! `brainfuck_prover prove` (crates/brainfuck_prover/src/bin/brainfuck_prover.rs:79-142) compiles the program, runs the
//! VM, calls `prove_brainfuck(&machine)` (brainfuck_air/mod.rs:471-735) and serialises the proof with serde_json.
//! `prove` below does the same inside the library: VM and table building on host threads, everything from the filled
//! columns to the proof object on the device, with the transcript (channel mixing order, claims, PCS config
//! pow_bits 5 / blowup 1 / 3 queries / last-layer bound 0) of the reference.  The JSON it returns has serde's shape for
//! `BrainfuckProof<Blake2sMerkleHasher>`, so the reference's own `verify_brainfuck` (mod.rs:738-797) can check it:
//!
//! ```ignore
//! let proof = stwo_cuda_backend::whole_proof::prove(&code, &stdin, 24, 0)?;
//! let bf: BrainfuckProof<Blake2sMerkleHasher> = proof.parse()?;     // serde_json::from_str
//! verify_brainfuck(bf)?;
//! ```

use std::ffi::{CStr, CString};
use std::ptr;

use serde::de::DeserializeOwned;

use crate::{ctx, ffi};

/// `ProvingError` / `VerificationError` as text (the library reports them through `sbf_last_error`).
#[derive(Debug)]
pub struct ProofError(pub i32, pub String);

fn last_error(code: i32) -> ProofError {
    ProofError(code, unsafe { CStr::from_ptr(ffi::sbf_last_error()) }.to_string_lossy().into_owned())
}

/// Handle to a `BrainfuckProof { claim, interaction_claim, proof }` held by the library.
pub struct CudaProof(*mut ffi::SbfProof);

impl Drop for CudaProof {
    fn drop(&mut self) {
        unsafe { ffi::sbf_proof_free(self.0) };
    }
}

/// `log_max_rows` = the reference's `LOG_MAX_ROWS` (24; brainfuck_air/mod.rs:427-428).  `flags`: `ffi::SBF_NO_OVERLAP`,
/// `ffi::SBF_NO_TWIDDLE_CACHE` (recompute the twiddle tree in every proof, as the reference does), `ffi::SBF_SHARDED_DRIVER`,
/// `ffi::SBF_CACHE_PREPROCESSED` (keep the program-independent preprocessed tree on the context between proofs).
pub fn prove(code: &str, input: &[u8], log_max_rows: u32, flags: u32) -> Result<CudaProof, ProofError> {
    let code = CString::new(code).map_err(|_| ProofError(ffi::SC_EINVAL, "program text contains a NUL byte".into()))?;
    let mut out = ptr::null_mut();
    let rc = unsafe { ffi::sbf_prove(ctx(), code.as_ptr(), input.as_ptr(), input.len(), log_max_rows, flags, &mut out) };
    if rc != ffi::SC_OK {
        return Err(last_error(rc));
    }
    Ok(CudaProof(out))
}

/// Drops the tree kept by `SBF_CACHE_PREPROCESSED` (destroying the context does too).
pub fn clear_preprocessed_cache() -> Result<(), ProofError> {
    match unsafe { ffi::sbf_preprocessed_cache_clear(ctx()) } {
        ffi::SC_OK => Ok(()),
        rc => Err(last_error(rc)),
    }
}

impl CudaProof {
    pub(crate) fn from_raw(h: *mut ffi::SbfProof) -> Self {
        CudaProof(h)
    }

    /// The library's own restatement of `verify_brainfuck` (host arithmetic only).
    pub fn verify(&self) -> Result<(), ProofError> {
        match unsafe { ffi::sbf_verify(self.0) } {
            ffi::SC_OK => Ok(()),
            rc => Err(last_error(rc)),
        }
    }

    fn take_string(p: *mut std::ffi::c_char) -> String {
        let s = unsafe { CStr::from_ptr(p) }.to_string_lossy().into_owned();
        unsafe { ffi::sbf_string_free(p) };
        s
    }

    /// serde-shaped JSON, what `brainfuck_prover prove --output` writes (bin/brainfuck_prover.rs:127-131).
    pub fn to_json(&self) -> String {
        Self::take_string(unsafe { ffi::sbf_proof_json(self.0) })
    }

    /// `serde_json::from_str` into the reference's `BrainfuckProof<Blake2sMerkleHasher>` (or any type of that shape).
    pub fn parse<T: DeserializeOwned>(&self) -> Result<T, serde_json::Error> {
        serde_json::from_str(&self.to_json())
    }

    /// Steps, component log sizes and per-stage milliseconds.
    pub fn report(&self) -> String {
        Self::take_string(unsafe { ffi::sbf_proof_report(self.0) })
    }

    /// What the program wrote to stdout while the VM ran.
    pub fn output(&self) -> Vec<u8> {
        let n = unsafe { ffi::sbf_proof_output(self.0, ptr::null_mut(), 0) };
        let mut buf = vec![0u8; n];
        unsafe { ffi::sbf_proof_output(self.0, buf.as_mut_ptr(), buf.len()) };
        buf
    }
}
