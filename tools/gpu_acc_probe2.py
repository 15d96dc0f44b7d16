"""Dev probe (GPU): per-MSM accumulate times of one resident config-2 step (BPG_ACC_TRACE)."""
import os, sys
os.environ["BPG_ACC_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bulletproof_gadgets_b200 as bpg
from bulletproof_gadgets_b200 import workloads as W
ctx = bpg.Context(0)
st = W.bounds_check_statement(1024)
ctx.gens_ensure(st.n)
circ = bpg.Circuit(ctx, st.n, st.m, st.row_start, st.term_var, st.term_coef, st.q).set_witness(st.aL, st.aR)
def step():
    p = bpg.Prover(ctx, bpg.Transcript(st.label)); coms = p.commit_batch_packed(st.v_bytes, st.vbl_bytes)
    p.attach(circ); proof = p.prove(b"\x07" * 32)
    vf = bpg.Verifier(ctx, bpg.Transcript(st.label)); vf.commit_batch(coms); vf.attach(circ)
    assert vf.verify(proof, b"\x09" * 32)
step()
ctx.set("time_accum", 1)
step()
