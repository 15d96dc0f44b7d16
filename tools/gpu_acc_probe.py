"""Dev probe (GPU): k_accumulate time per step of config 2 in an instrumented (synchronous) resident step."""
import os, sys
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
    print("   after prove: sum_accum %.3f ms entries %d" % (ctx.get("sum_accum_ns") / 1e6, ctx.get("sum_entries")))
    vf = bpg.Verifier(ctx, bpg.Transcript(st.label)); vf.commit_batch(coms); vf.attach(circ)
    assert vf.verify(proof, b"\x09" * 32)
for it in range(3):
    ctx.set("time_accum", 1)
    step()
    print("step %d: sum_accum %.3f ms entries %d" % (it, ctx.get("sum_accum_ns") / 1e6, ctx.get("sum_entries")))
    ctx.set("time_accum", 0)
    step()
for tl in (16, 24, 28, 32, 40, 48, 64):
    ctx.set("task_len", tl)
    step()
    ctx.set("time_accum", 1); step()
    print("task_len %d: sum_accum %.3f ms" % (tl, ctx.get("sum_accum_ns") / 1e6)); ctx.set("time_accum", 0)
