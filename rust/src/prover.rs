//! `GpuProver`: what `src/prove.rs:45-47,78-81` constructs instead of `bulletproofs::r1cs::Prover`.
//! Variables cross the boundary as u32 tags `kind << 29 | index` (0 Committed, 1 MultiplierLeft, 2 MultiplierRight,
//! 3 MultiplierOutput, 4 One); scalars as `Scalar::as_bytes()`, points as `CompressedRistretto::as_bytes()`.
use crate::ffi::*;
use crate::{map_err, Context};
use bulletproofs::r1cs::{ConstraintSystem, LinearCombination, R1CSError, Variable};
use curve25519_dalek::ristretto::CompressedRistretto;
use curve25519_dalek::scalar::Scalar;

pub(crate) fn tag(v: Variable) -> u32 {
    match v {
        Variable::Committed(i) => i as u32,
        Variable::MultiplierLeft(i) => (1 << 29) | i as u32,
        Variable::MultiplierRight(i) => (2 << 29) | i as u32,
        Variable::MultiplierOutput(i) => (3 << 29) | i as u32,
        Variable::One() => 4 << 29,
    }
}
pub(crate) fn untag(t: u32) -> Variable {
    let i = (t & ((1 << 29) - 1)) as usize;
    match t >> 29 {
        0 => Variable::Committed(i),
        1 => Variable::MultiplierLeft(i),
        2 => Variable::MultiplierRight(i),
        3 => Variable::MultiplierOutput(i),
        _ => Variable::One(),
    }
}
/// term list as parallel arrays (variable tags, 32-byte coefficients)
pub(crate) fn pack(lc: &LinearCombination) -> (Vec<u32>, Vec<u8>) {
    let mut vars = Vec::new();
    let mut coef = Vec::new();
    for (v, c) in lc.get_terms() {           // FairAds fork accessor, as used at src/cs_buffer.rs
        vars.push(tag(*v));
        coef.extend_from_slice(c.as_bytes());
    }
    (vars, coef)
}

pub struct GpuProver<'a> { p: *mut bpg_prover, t: *mut bpg_transcript, _ctx: &'a Context }

impl<'a> GpuProver<'a> {
    /// `Transcript::new(name)` + `Prover::new(&pc_gens, &mut transcript)` (src/prove.rs:45-47)
    pub fn new(ctx: &'a Context, name: &[u8]) -> GpuProver<'a> {
        unsafe {
            let t = bpg_transcript_new(name.as_ptr(), name.len());
            let mut p = std::ptr::null_mut();
            assert_eq!(bpg_prover_new(ctx.0, t, &mut p), BPG_OK, "{}", last_error());
            GpuProver { p, t, _ctx: ctx }
        }
    }
    /// `Prover::commit` (src/gadget.rs:32, src/commitments.rs:28,40)
    pub fn commit(&mut self, v: Scalar, v_blinding: Scalar) -> (CompressedRistretto, Variable) {
        let mut out = [0u8; 32];
        let mut var = 0u32;
        unsafe {
            assert_eq!(bpg_prover_commit(self.p, v.as_bytes().as_ptr(), v_blinding.as_bytes().as_ptr(), out.as_mut_ptr(), &mut var),
                       BPG_OK, "{}", last_error());
        }
        (CompressedRistretto(out), Variable::Committed(var as usize))
    }
    /// All commitments of a statement in one launch (the .wtns replay + every gadget's derived witnesses).
    pub fn commit_batch(&mut self, v: &[Scalar], v_blinding: &[Scalar]) -> Vec<(CompressedRistretto, Variable)> {
        let pack32 = |xs: &[Scalar]| xs.iter().flat_map(|s| s.as_bytes().to_vec()).collect::<Vec<u8>>();
        let (vb, bb) = (pack32(v), pack32(v_blinding));
        let mut out = vec![0u8; 32 * v.len()];
        let mut first = 0u32;
        unsafe {
            assert_eq!(bpg_prover_commit_batch(self.p, vb.as_ptr(), bb.as_ptr(), v.len() as u64, out.as_mut_ptr(), &mut first),
                       BPG_OK, "{}", last_error());
        }
        out.chunks(32).enumerate().map(|(k, c)| {
            let mut a = [0u8; 32];
            a.copy_from_slice(c);
            (CompressedRistretto(a), Variable::Committed(first as usize + k))
        }).collect()
    }
    /// Bulk replacement of `assign_buffer` (src/prove.rs:84-99): flat multiplier assignments + CSR constraints.
    pub fn load_cs(&mut self, a_l: &[u8], a_r: &[u8], row_start: &[u32], term_var: &[u32], term_coef: &[u8]) -> Result<(), R1CSError> {
        let rc = unsafe {
            bpg_prover_load_cs(self.p, a_l.as_ptr(), a_r.as_ptr(), (a_l.len() / 32) as u64, row_start.as_ptr(), term_var.as_ptr(),
                               term_coef.as_ptr(), (row_start.len() - 1) as u64)
        };
        if rc != BPG_OK { Err(map_err(rc)) } else { Ok(()) }
    }
    pub fn num_constraints(&self) -> usize { unsafe { bpg_prover_num_constraints(self.p) as usize } }
    pub fn get_num_multiplications(&self) -> usize { unsafe { bpg_prover_num_multipliers(self.p) as usize } }
    /// `prover.prove(&bp_gens)?.to_bytes()` (src/prove.rs:78-81); the generators live in the context
    pub fn prove(self) -> Result<Vec<u8>, R1CSError> {
        let mut buf = vec![0u8; 1 + 14 * 32 + 66 * 32];
        let mut len = 0usize;
        let rc = unsafe { bpg_prover_prove(self.p, std::ptr::null(), buf.as_mut_ptr(), buf.len(), &mut len) };
        if rc != BPG_OK { return Err(map_err(rc)); }
        buf.truncate(len);
        Ok(buf)
    }
}
impl<'a> Drop for GpuProver<'a> {
    fn drop(&mut self) { unsafe { bpg_prover_free(self.p); bpg_transcript_free(self.t); } }
}

impl<'a> ConstraintSystem for GpuProver<'a> {
    fn transcript(&mut self) -> &mut merlin::Transcript { unimplemented!("the transcript lives in the library (bpg_transcript_*)") }
    fn multiply(&mut self, l: LinearCombination, r: LinearCombination) -> (Variable, Variable, Variable) {
        let (lv, lc) = pack(&l);
        let (rv, rc) = pack(&r);
        let mut out = [0u32; 3];
        unsafe {
            assert_eq!(bpg_prover_multiply(self.p, lv.as_ptr(), lc.as_ptr(), lv.len(), rv.as_ptr(), rc.as_ptr(), rv.len(), out.as_mut_ptr()),
                       BPG_OK, "{}", last_error());
        }
        (untag(out[0]), untag(out[1]), untag(out[2]))
    }
    fn allocate(&mut self, _: Option<Scalar>) -> Result<Variable, R1CSError> {
        Err(R1CSError::GadgetError { description: "allocate() is not used by the reference's gadgets".into() })
    }
    fn allocate_multiplier(&mut self, a: Option<(Scalar, Scalar)>) -> Result<(Variable, Variable, Variable), R1CSError> {
        let (l, r) = a.ok_or(R1CSError::MissingAssignment)?;
        let mut out = [0u32; 3];
        let rc = unsafe { bpg_prover_allocate_multiplier(self.p, l.as_bytes().as_ptr(), r.as_bytes().as_ptr(), out.as_mut_ptr()) };
        if rc != BPG_OK { return Err(map_err(rc)); }
        Ok((untag(out[0]), untag(out[1]), untag(out[2])))
    }
    fn constrain(&mut self, lc: LinearCombination) {
        let (v, c) = pack(&lc);
        unsafe { assert_eq!(bpg_prover_constrain(self.p, v.as_ptr(), c.as_ptr(), v.len()), BPG_OK, "{}", last_error()); }
    }
}
