//! `GpuVerifier`: what `src/verify.rs:44-46,70-72` constructs instead of `bulletproofs::r1cs::Verifier`.
use crate::ffi::*;
use crate::prover::{pack, untag};
use crate::{map_err, Context};
use bulletproofs::r1cs::{ConstraintSystem, LinearCombination, R1CSError, Variable};
use curve25519_dalek::ristretto::CompressedRistretto;
use curve25519_dalek::scalar::Scalar;

pub struct GpuVerifier<'a> { v: *mut bpg_verifier, t: *mut bpg_transcript, _ctx: &'a Context }

impl<'a> GpuVerifier<'a> {
    /// `Transcript::new(name)` + `Verifier::new(&mut transcript)` (src/verify.rs:44-46)
    pub fn new(ctx: &'a Context, name: &[u8]) -> GpuVerifier<'a> {
        unsafe {
            let t = bpg_transcript_new(name.as_ptr(), name.len());
            let mut v = std::ptr::null_mut();
            assert_eq!(bpg_verifier_new(ctx.0, t, &mut v), BPG_OK, "{}", last_error());
            GpuVerifier { v, t, _ctx: ctx }
        }
    }
    /// `Verifier::commit` (src/lalrpop/assignment_parser.rs:138)
    pub fn commit(&mut self, com: CompressedRistretto) -> Variable {
        let mut var = 0u32;
        unsafe { assert_eq!(bpg_verifier_commit(self.v, com.as_bytes().as_ptr(), &mut var), BPG_OK); }
        untag(var)
    }
    /// the whole .coms replay (assignment_parser.rs:133-141) in one call; variables are Committed(first .. first + k)
    pub fn commit_batch(&mut self, coms: &[CompressedRistretto]) -> usize {
        let packed: Vec<u8> = coms.iter().flat_map(|c| c.as_bytes().to_vec()).collect();
        let mut first = 0u32;
        unsafe { assert_eq!(bpg_verifier_commit_batch(self.v, packed.as_ptr(), coms.len() as u64, &mut first), BPG_OK); }
        first as usize
    }
    /// Bulk replacement of `assign_buffer` (src/verify.rs:75-90)
    pub fn load_cs(&mut self, n: u64, row_start: &[u32], term_var: &[u32], term_coef: &[u8]) -> Result<(), R1CSError> {
        let rc = unsafe { bpg_verifier_load_cs(self.v, n, row_start.as_ptr(), term_var.as_ptr(), term_coef.as_ptr(), (row_start.len() - 1) as u64) };
        if rc != BPG_OK { Err(map_err(rc)) } else { Ok(()) }
    }
    pub fn get_num_vars(&self) -> usize { unsafe { bpg_verifier_num_vars(self.v) as usize } }
    /// `R1CSProof::from_bytes(&proof)` + `verifier.verify(&proof, &pc_gens, &bp_gens)` (src/verify.rs:53,70-72).
    /// Ok(()) = accepted; Err(VerificationError) = rejected; Err(FormatError) where the reference's unwrap() panics.
    pub fn verify(self, proof: &[u8]) -> Result<(), R1CSError> {
        let rc = unsafe { bpg_verifier_verify(self.v, proof.as_ptr(), proof.len(), std::ptr::null()) };
        if rc != BPG_OK { Err(map_err(rc)) } else { Ok(()) }
    }
}
impl<'a> Drop for GpuVerifier<'a> {
    fn drop(&mut self) { unsafe { bpg_verifier_free(self.v); bpg_transcript_free(self.t); } }
}

impl<'a> ConstraintSystem for GpuVerifier<'a> {
    fn transcript(&mut self) -> &mut merlin::Transcript { unimplemented!("the transcript lives in the library (bpg_transcript_*)") }
    fn multiply(&mut self, l: LinearCombination, r: LinearCombination) -> (Variable, Variable, Variable) {
        let (lv, lc) = pack(&l);
        let (rv, rc) = pack(&r);
        let mut out = [0u32; 3];
        unsafe {
            assert_eq!(bpg_verifier_multiply(self.v, lv.as_ptr(), lc.as_ptr(), lv.len(), rv.as_ptr(), rc.as_ptr(), rv.len(), out.as_mut_ptr()),
                       BPG_OK, "{}", last_error());
        }
        (untag(out[0]), untag(out[1]), untag(out[2]))
    }
    fn allocate(&mut self, _: Option<Scalar>) -> Result<Variable, R1CSError> {
        Err(R1CSError::GadgetError { description: "allocate() is not used by the reference's gadgets".into() })
    }
    fn allocate_multiplier(&mut self, _: Option<(Scalar, Scalar)>) -> Result<(Variable, Variable, Variable), R1CSError> {
        let mut out = [0u32; 3];
        unsafe { assert_eq!(bpg_verifier_allocate_multiplier(self.v, out.as_mut_ptr()), BPG_OK); }
        Ok((untag(out[0]), untag(out[1]), untag(out[2])))
    }
    fn constrain(&mut self, lc: LinearCombination) {
        let (v, c) = pack(&lc);
        unsafe { assert_eq!(bpg_verifier_constrain(self.v, v.as_ptr(), c.as_ptr(), v.len()), BPG_OK, "{}", last_error()); }
    }
}
