//! `bulletproofs_gadgets` on a B200: drop-in `Prover` / `Verifier` for `src/prove.rs` and `src/verify.rs` of the
//! reference, backed by `libbpg.so` through the C ABI in `include/bpg.h`.
pub mod batch;
pub mod ffi;
pub mod prover;
pub mod verifier;

use bulletproofs::r1cs::R1CSError;

/// One GPU context (stream + work buffers; generator tables shared between the contexts of a GPU).
pub struct Context(pub(crate) *mut ffi::bpg_ctx);
unsafe impl Send for Context {}

impl Context {
    pub fn new(device: i32) -> Result<Context, String> {
        let mut h = std::ptr::null_mut();
        let rc = unsafe { ffi::bpg_ctx_create(device, &mut h) };
        if rc != ffi::BPG_OK { return Err(ffi::last_error()); } // no sm_100a device: there is no CPU fallback
        Ok(Context(h))
    }
    /// A further context on the same GPU: own stream, shared tables; one per worker thread.
    pub fn shared(&self) -> Result<Context, String> {
        let mut h = std::ptr::null_mut();
        let rc = unsafe { ffi::bpg_ctx_create_shared(self.0, &mut h) };
        if rc != ffi::BPG_OK { return Err(ffi::last_error()); }
        Ok(Context(h))
    }
    /// `BulletproofGens::new(capacity, 1)` once per GPU instead of once per proof (`src/prove.rs:78`).
    pub fn ensure_generators(&self, capacity: u64) -> Result<(), String> {
        if unsafe { ffi::bpg_gens_ensure(self.0, capacity) } != ffi::BPG_OK { return Err(ffi::last_error()); }
        Ok(())
    }
}
impl Drop for Context {
    fn drop(&mut self) { unsafe { ffi::bpg_ctx_destroy(self.0) } }
}

pub(crate) fn map_err(rc: i32) -> R1CSError {
    match rc {
        ffi::BPG_E_FORMAT => R1CSError::FormatError,
        ffi::BPG_E_GENS_LEN => R1CSError::InvalidGeneratorsLength,
        ffi::BPG_E_MISSING_ASSIGNMENT => R1CSError::MissingAssignment,
        ffi::BPG_E_VERIFY => R1CSError::VerificationError,
        _ => R1CSError::GadgetError { description: ffi::last_error() },
    }
}
