//! One proof over several GPUs (include/stwo_cuda_sharded.h): one process per GPU, one context per process, NCCL over
//! NVLink for the exchange steps.  The reference is single-process; this is the part of BASELINE.json's north_star that
//! has no counterpart there.  Every rank calls `prove_sharded` with the same program and ends with the same proof bytes.

use std::ffi::CString;
use std::ptr;

use crate::whole_proof::{CudaProof, ProofError};
use crate::{ck, ctx, ffi};

/// NCCL unique id: rank 0 creates it and hands it to the other ranks out of band (a file, a socket, MPI …).
pub const UNIQUE_ID_BYTES: usize = 128;

pub fn unique_id() -> [u8; UNIQUE_ID_BYTES] {
    let mut id = [0u8; UNIQUE_ID_BYTES];
    ck(unsafe { ffi::sc_comm_unique_id(id.as_mut_ptr()) });
    id
}

pub struct Comm(*mut ffi::ScComm);

impl Comm {
    pub fn new(rank: i32, world: i32, id: &[u8; UNIQUE_ID_BYTES]) -> Self {
        let mut h = ptr::null_mut();
        ck(unsafe { ffi::sc_comm_init(ctx(), rank, world, id.as_ptr(), &mut h) });
        Comm(h)
    }
    pub fn rank(&self) -> i32 {
        unsafe { ffi::sc_comm_rank(self.0) }
    }
    pub fn world(&self) -> i32 {
        unsafe { ffi::sc_comm_world(self.0) }
    }
}
impl Drop for Comm {
    fn drop(&mut self) {
        unsafe { ffi::sc_comm_destroy(ctx(), self.0) };
    }
}

pub fn prove_sharded(comm: &Comm, code: &str, input: &[u8], log_max_rows: u32, flags: u32) -> Result<CudaProof, ProofError> {
    let code = CString::new(code).map_err(|_| ProofError(ffi::SC_EINVAL, "program text contains a NUL byte".into()))?;
    let mut out = ptr::null_mut();
    let rc = unsafe {
        ffi::sbf_prove_sharded(ctx(), comm.0, code.as_ptr(), input.as_ptr(), input.len(), log_max_rows, flags, &mut out)
    };
    if rc != ffi::SC_OK {
        let msg = unsafe { std::ffi::CStr::from_ptr(ffi::sbf_last_error()) }.to_string_lossy().into_owned();
        return Err(ProofError(rc, msg));
    }
    Ok(CudaProof::from_raw(out))
}
