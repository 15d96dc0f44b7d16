"""Dev check (GPU): BASELINE config 2 (BOUND 64-bit x COUNT in one proof): timing + parity vs the C oracle."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bulletproof_gadgets_b200 as bpg  # noqa: E402
from bulletproof_gadgets_b200 import workloads as W  # noqa: E402
from oracle import coracle  # noqa: E402

count = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
check_count = int(sys.argv[2]) if len(sys.argv) > 2 else 64
ctx = bpg.Context(0)
for cnt in sorted({check_count, count}):
    st = W.bounds_check_statement(cnt)
    t0 = time.time()
    ctx.gens_ensure(st.n)
    print("count %d: n=%d m=%d q=%d nnz=%d gens %.3fs" % (cnt, st.n, st.m, st.q, st.nnz, time.time() - t0))
    for it in range(3):
        t0 = time.time()
        proof, coms = W.prove_statement(bpg, ctx, st)
        t1 = time.time()
        ok = W.verify_statement(bpg, ctx, st, proof, coms)
        t2 = time.time()
        print("  host-buffer path: prove %.1f ms  verify %.1f ms  ok=%s len=%d" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, ok, len(proof)))
    circ = bpg.Circuit(ctx, st.n, st.m, st.row_start, st.term_var, st.term_coef, st.q).set_witness(st.aL, st.aR)
    for it in range(3):
        T = bpg.Transcript(st.label)
        p = bpg.Prover(ctx, T)
        t0 = time.time()
        coms2 = [c for c, _ in p.commit_batch(st.v, st.vbl)]
        t1 = time.time()
        p.attach(circ)
        proof2 = p.prove(b"\x07" * 32)
        t2 = time.time()
        T = bpg.Transcript(st.label)
        vf = bpg.Verifier(ctx, T)
        for c in coms2:
            vf.commit(c)
        vf.attach(circ)
        t3 = time.time()
        ok2 = vf.verify(proof2, b"\x09" * 32)
        t4 = time.time()
        print("  resident path: commit %.1f ms  prove %.1f ms  verify %.1f ms ok=%s same=%s" %
              ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t4 - t3) * 1e3, ok2, proof2 == proof))
    assert ok and ok2 and proof2 == proof
    if cnt == check_count:
        t0 = time.time()
        p_c, coms_c = coracle.prove_flat(st, b"\x07" * 32)
        t1 = time.time()
        v_c = coracle.verify_flat(st, coms_c, p_c, b"\x09" * 32)
        t2 = time.time()
        print("  C oracle: prove %.2f s verify %.2f s accepted=%s | proof identical=%s coms identical=%s" %
              (t1 - t0, t2 - t1, v_c, p_c == proof, coms_c == coms))
        assert p_c == proof and coms_c == coms and v_c is True
print("OK")
