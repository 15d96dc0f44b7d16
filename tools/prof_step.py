"""Profiling driver: one warm-up and ONE measured resident prove+verify of config 2 (used under ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bulletproof_gadgets_b200 as bpg
from bulletproof_gadgets_b200 import workloads as W
count = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = bpg.Context(0)
st = W.bounds_check_statement(count)
ctx.gens_ensure(st.n)
circ = bpg.Circuit(ctx, st.n, st.m, st.row_start, st.term_var, st.term_coef, st.q).set_witness(st.aL, st.aR)
for it in range(reps):
    T = bpg.Transcript(st.label); p = bpg.Prover(ctx, T)
    coms = [c for c, _ in p.commit_batch(st.v, st.vbl)]
    p.attach(circ); proof = p.prove(b"\x07" * 32)
    T = bpg.Transcript(st.label); vf = bpg.Verifier(ctx, T); vf.commit_batch(coms); vf.attach(circ)
    assert vf.verify(proof, b"\x09" * 32)
print("ok launches", ctx.get("launches"))
