//! Throughput: N independent flat statements through `bpg_r1cs_prove_batch` / `bpg_r1cs_verify_batch`.  The library owns
//! the worker threads (one per context); the caller only lays the jobs out.  This is what bench.py's headline legs call.
use crate::ffi::*;
use crate::Context;

/// One flat statement: what `assign_buffer` (src/prove.rs:84-99) would replay, plus the committed values.
pub struct Statement<'a> {
    pub label: &'a [u8],
    pub v: &'a [u8],        // m x 32
    pub v_blinding: &'a [u8],
    pub a_l: &'a [u8],      // n x 32
    pub a_r: &'a [u8],
    pub row_start: &'a [u32],
    pub term_var: &'a [u32],
    pub term_coef: &'a [u8],
}
pub struct Proved { pub status: i32, pub commitments: Vec<u8>, pub proof: Vec<u8> }

/// Proves (and, with `verify`, immediately verifies) every statement; `ctxs` = contexts of one or several GPUs.
pub fn prove_batch(ctxs: &[Context], sts: &[Statement], verify: bool) -> Vec<Proved> {
    let handles: Vec<*mut bpg_ctx> = ctxs.iter().map(|c| c.0).collect();
    let mut outs: Vec<(Vec<u8>, Vec<u8>)> = sts.iter().map(|s| (vec![0u8; s.v.len().max(32)], vec![0u8; 1 + 14 * 32 + 66 * 32])).collect();
    let mut jobs: Vec<bpg_prove_job> = sts.iter().zip(outs.iter_mut()).map(|(s, o)| bpg_prove_job {
        label: s.label.as_ptr(), label_len: s.label.len(),
        v32m: s.v.as_ptr(), vbl32m: s.v_blinding.as_ptr(), m: (s.v.len() / 32) as u64,
        circuit: std::ptr::null(),
        a_l32n: s.a_l.as_ptr(), a_r32n: s.a_r.as_ptr(), n: (s.a_l.len() / 32) as u64,
        row_start: s.row_start.as_ptr(), term_var: s.term_var.as_ptr(), term_coef32: s.term_coef.as_ptr(),
        q: (s.row_start.len() - 1) as u64,
        rng_seed32: std::ptr::null(), verify_seed32: std::ptr::null(),      // OS entropy, like thread_rng()
        flags: if verify { BPG_JOB_VERIFY } else { 0 },
        v_out32m: o.0.as_mut_ptr(), proof_out: o.1.as_mut_ptr(), proof_cap: o.1.len(), proof_len: 0, status: 0,
    }).collect();
    let rc = unsafe { bpg_r1cs_prove_batch(handles.as_ptr(), handles.len(), jobs.as_mut_ptr(), jobs.len()) };
    assert!(rc >= 0, "{}", last_error());
    jobs.iter().zip(outs.into_iter()).zip(sts.iter()).map(|((j, (mut coms, mut proof)), s)| {
        coms.truncate(s.v.len());
        proof.truncate(j.proof_len);
        Proved { status: j.status, commitments: coms, proof }
    }).collect()
}
