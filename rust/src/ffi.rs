//! `extern "C"` declarations of include/bpg.h and include/bulletproofs_gadgets.h (libbpg.so).
#![allow(non_camel_case_types)]
use std::ffi::CStr;
use std::os::raw::{c_char, c_int};

#[repr(C)] pub struct bpg_ctx { _p: [u8; 0] }
#[repr(C)] pub struct bpg_transcript { _p: [u8; 0] }
#[repr(C)] pub struct bpg_prover { _p: [u8; 0] }
#[repr(C)] pub struct bpg_verifier { _p: [u8; 0] }
#[repr(C)] pub struct bpg_circuit { _p: [u8; 0] }

pub const BPG_OK: c_int = 0;
pub const BPG_E_FORMAT: c_int = -1;             // R1CSError::FormatError
pub const BPG_E_VERIFY: c_int = -2;             // R1CSError::VerificationError
pub const BPG_E_GENS_LEN: c_int = -3;           // R1CSError::InvalidGeneratorsLength
pub const BPG_E_MISSING_ASSIGNMENT: c_int = -4; // R1CSError::MissingAssignment
pub const BPG_E_CUDA: c_int = -5;
pub const BPG_E_ARG: c_int = -6;
pub const BPG_E_GADGET: c_int = -7;
pub const BPG_JOB_VERIFY: u32 = 1;

/// include/bpg.h: bpg_proof_artifacts
#[repr(C)]
pub struct bpg_proof_artifacts {
    pub commitments: *mut c_char,
    pub proof: *mut u8,
    pub proof_len: usize,
    pub num_constraints: u64,
}
/// include/bulletproofs_gadgets.h = /root/reference/interfaces/ios/src/lib.rs:11-19
#[repr(C)]
pub struct ProofArtifacts {
    pub commitments: *const c_char,
    pub proof: *const u8,
    pub proof_len: usize,
    pub proof_cap: usize,
}
/// include/bpg.h: bpg_prove_job
#[repr(C)]
pub struct bpg_prove_job {
    pub label: *const u8, pub label_len: usize,
    pub v32m: *const u8, pub vbl32m: *const u8, pub m: u64,
    pub circuit: *const bpg_circuit,
    pub a_l32n: *const u8, pub a_r32n: *const u8, pub n: u64,
    pub row_start: *const u32, pub term_var: *const u32, pub term_coef32: *const u8, pub q: u64,
    pub rng_seed32: *const u8, pub verify_seed32: *const u8, pub flags: u32,
    pub v_out32m: *mut u8, pub proof_out: *mut u8, pub proof_cap: usize, pub proof_len: usize,
    pub status: c_int,
}
/// include/bpg.h: bpg_verify_job
#[repr(C)]
pub struct bpg_verify_job {
    pub label: *const u8, pub label_len: usize,
    pub v32m: *const u8, pub m: u64,
    pub circuit: *const bpg_circuit,
    pub n: u64, pub row_start: *const u32, pub term_var: *const u32, pub term_coef32: *const u8, pub q: u64,
    pub proof: *const u8, pub proof_len: usize, pub rng_seed32: *const u8,
    pub status: c_int,
}

#[link(name = "bpg")]
extern "C" {
    pub fn bpg_last_error() -> *const c_char;
    pub fn bpg_ctx_create(device: c_int, out: *mut *mut bpg_ctx) -> c_int;
    pub fn bpg_ctx_create_shared(parent: *mut bpg_ctx, out: *mut *mut bpg_ctx) -> c_int;
    pub fn bpg_ctx_destroy(ctx: *mut bpg_ctx);
    pub fn bpg_gens_ensure(ctx: *mut bpg_ctx, capacity: u64) -> c_int;
    pub fn bpg_host_alloc(bytes: usize) -> *mut u8;
    pub fn bpg_host_free(p: *mut u8);

    pub fn bpg_transcript_new(label: *const u8, len: usize) -> *mut bpg_transcript;
    pub fn bpg_transcript_free(t: *mut bpg_transcript);
    pub fn bpg_transcript_append_message(t: *mut bpg_transcript, label: *const u8, ll: usize, msg: *const u8, ml: usize);
    pub fn bpg_transcript_challenge_bytes(t: *mut bpg_transcript, label: *const u8, ll: usize, out: *mut u8, n: usize);

    pub fn bpg_prover_new(ctx: *mut bpg_ctx, t: *mut bpg_transcript, out: *mut *mut bpg_prover) -> c_int;
    pub fn bpg_prover_free(p: *mut bpg_prover);
    pub fn bpg_prover_commit(p: *mut bpg_prover, v: *const u8, v_blinding: *const u8, v_out: *mut u8, var_out: *mut u32) -> c_int;
    pub fn bpg_prover_commit_batch(p: *mut bpg_prover, v: *const u8, vb: *const u8, k: u64, v_out: *mut u8, first_var: *mut u32) -> c_int;
    pub fn bpg_prover_allocate_multiplier(p: *mut bpg_prover, l: *const u8, r: *const u8, vars_out: *mut u32) -> c_int;
    pub fn bpg_prover_multiply(p: *mut bpg_prover, lvars: *const u32, lcoef: *const u8, ln: usize,
                               rvars: *const u32, rcoef: *const u8, rn: usize, vars_out: *mut u32) -> c_int;
    pub fn bpg_prover_constrain(p: *mut bpg_prover, vars: *const u32, coef: *const u8, n: usize) -> c_int;
    pub fn bpg_prover_load_cs(p: *mut bpg_prover, a_l: *const u8, a_r: *const u8, n: u64,
                              row_start: *const u32, term_var: *const u32, term_coef: *const u8, q: u64) -> c_int;
    pub fn bpg_prover_num_constraints(p: *const bpg_prover) -> u64;
    pub fn bpg_prover_num_multipliers(p: *const bpg_prover) -> u64;
    pub fn bpg_prover_prove(p: *mut bpg_prover, rng_seed32: *const u8, proof_out: *mut u8, cap: usize, len: *mut usize) -> c_int;

    pub fn bpg_verifier_new(ctx: *mut bpg_ctx, t: *mut bpg_transcript, out: *mut *mut bpg_verifier) -> c_int;
    pub fn bpg_verifier_free(v: *mut bpg_verifier);
    pub fn bpg_verifier_commit(v: *mut bpg_verifier, com: *const u8, var_out: *mut u32) -> c_int;
    pub fn bpg_verifier_commit_batch(v: *mut bpg_verifier, coms: *const u8, k: u64, first_var: *mut u32) -> c_int;
    pub fn bpg_verifier_allocate_multiplier(v: *mut bpg_verifier, vars_out: *mut u32) -> c_int;
    pub fn bpg_verifier_multiply(v: *mut bpg_verifier, lvars: *const u32, lcoef: *const u8, ln: usize,
                                 rvars: *const u32, rcoef: *const u8, rn: usize, vars_out: *mut u32) -> c_int;
    pub fn bpg_verifier_constrain(v: *mut bpg_verifier, vars: *const u32, coef: *const u8, n: usize) -> c_int;
    pub fn bpg_verifier_load_cs(v: *mut bpg_verifier, n: u64, row_start: *const u32, term_var: *const u32,
                                term_coef: *const u8, q: u64) -> c_int;
    pub fn bpg_verifier_num_vars(v: *const bpg_verifier) -> u64;
    pub fn bpg_verifier_verify(v: *mut bpg_verifier, proof: *const u8, len: usize, rng_seed32: *const u8) -> c_int;

    pub fn bpg_r1cs_prove_batch(ctxs: *const *mut bpg_ctx, n_ctx: usize, jobs: *mut bpg_prove_job, n_jobs: usize) -> c_int;
    pub fn bpg_r1cs_verify_batch(ctxs: *const *mut bpg_ctx, n_ctx: usize, jobs: *mut bpg_verify_job, n_jobs: usize) -> c_int;

    pub fn bpg_prove(ctx: *mut bpg_ctx, name: *const c_char, instance: *const c_char, witness: *const c_char,
                     gadgets: *const c_char, blinding_seed32: *const u8, rng_seed32: *const u8,
                     out: *mut *mut bpg_proof_artifacts) -> c_int;
    pub fn bpg_verify(ctx: *mut bpg_ctx, name: *const c_char, instance: *const c_char, gadgets: *const c_char,
                      commitments: *const c_char, proof: *const u8, proof_len: usize, rng_seed32: *const u8,
                      accepted: *mut c_int) -> c_int;
    pub fn bpg_free_proof(a: *mut bpg_proof_artifacts);

    // include/bulletproofs_gadgets.h: the symbols interfaces/ios/src/lib.rs:21,45,55 export, served by libbpg.so
    pub fn c_prove(name: *const c_char, instance: *const c_char, witness: *const c_char, gadgets: *const c_char) -> *mut ProofArtifacts;
    pub fn c_verify(name: *const c_char, instance: *const c_char, gadgets: *const c_char, commitments: *const c_char,
                    proof: *const u8, proof_len: usize) -> bool;
    pub fn free_proof(artifacts_pointer: *mut ProofArtifacts);
}

pub fn last_error() -> String {
    unsafe { CStr::from_ptr(bpg_last_error()).to_string_lossy().into_owned() }
}
