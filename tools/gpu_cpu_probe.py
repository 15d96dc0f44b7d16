"""Dev probe (GPU): where the host CPU of one resident step goes (thread CPU time), single thread, no overlap."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("BPG_SYNC", "block")
import bulletproof_gadgets_b200 as bpg
from bulletproof_gadgets_b200 import workloads as W
ctx = bpg.Context(0)
st = W.bounds_check_statement(1024)
ctx.gens_ensure(st.n)
circ = bpg.Circuit(ctx, st.n, st.m, st.row_start, st.term_var, st.term_coef, st.q).set_witness(st.aL, st.aR)
def step(i):
    seed = (i + 1).to_bytes(32, "little")
    p = bpg.Prover(ctx, bpg.Transcript(st.label)); coms = p.commit_batch_packed(st.v_bytes, st.vbl_bytes)
    p.attach(circ); proof = p.prove(seed)
    vf = bpg.Verifier(ctx, bpg.Transcript(st.label)); vf.commit_batch(coms); vf.attach(circ)
    assert vf.verify(proof, seed)
for i in range(3): step(i)
keys = ("sync", "commit", "prove", "verify", "rng")
c0 = {k: ctx.get("cpu_%s_ns" % k) for k in keys}
t0, p0, w0 = time.thread_time(), time.process_time(), time.perf_counter()
N = 10
for i in range(N): step(10 + i)
t1, p1, w1 = time.thread_time(), time.process_time(), time.perf_counter()
c1 = {k: ctx.get("cpu_%s_ns" % k) for k in keys}
print("per step: wall %.1f ms | python thread cpu %.1f ms | process cpu %.1f ms" % (1e3*(w1-w0)/N, 1e3*(t1-t0)/N, 1e3*(p1-p0)/N))
print("  C-ABI thread cpu per step (ms):", {k: round((c1[k]-c0[k])/N/1e6, 2) for k in keys}, "(sync and rng are inside prove/verify)")
