"""Dev probe (GPU): where the host CPU of a statement goes when many are in flight.
Wraps every call the bench's resident step makes into the library with per-thread CPU and wall accumulators
(time.thread_time is CPU of the calling thread only), runs `inflight` host threads for `count` statements and prints
CPU and wall per statement by call, plus the part of the worker threads' CPU that is outside all wrapped calls
(python glue, GIL hand-offs) and the process CPU that is outside the worker threads (driver / runtime threads).

    python tools/gpu_host_cpu_breakdown.py [inflight=48] [count=192]
"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bulletproof_gadgets_b200 as bpg
from bulletproof_gadgets_b200 import workloads as W

inflight = int(sys.argv[1]) if len(sys.argv) > 1 else 48
count = int(sys.argv[2]) if len(sys.argv) > 2 else 192
ctx0 = bpg.Context(0)
ctxs = [ctx0] + [ctx0.shared() for _ in range(inflight - 1)]
st = W.bounds_check_statement(1024).pin(bpg)
ctx0.gens_ensure(st.n)
circ = bpg.Circuit(ctx0, st.n, st.m, st.row_start, st.term_var, st.term_coef, st.q).set_witness(st.aL, st.aR)

acc, acc_lock = {}, threading.Lock()


def timed(name, fn, *a):
    c0, w0 = time.thread_time(), time.perf_counter()
    r = fn(*a)
    c1, w1 = time.thread_time(), time.perf_counter()
    with acc_lock:
        e = acc.setdefault(name, [0.0, 0.0, 0])
        e[0] += c1 - c0
        e[1] += w1 - w0
        e[2] += 1
    return r


def step(ctx, i):
    seed = (i + 1).to_bytes(32, "little")
    t = timed("Transcript()", bpg.Transcript, st.label)
    p = timed("Prover()", bpg.Prover, ctx, t)
    coms = timed("commit_batch_packed", p.commit_batch_packed, st.v_bytes, st.vbl_bytes)
    timed("attach", p.attach, circ)
    proof = timed("prove", p.prove, seed)
    timed("del prover", p.__del__)
    t2 = timed("Transcript()", bpg.Transcript, st.label)
    vf = timed("Verifier()", bpg.Verifier, ctx, t2)
    timed("verifier.commit_batch", vf.commit_batch, coms)
    timed("attach", vf.attach, circ)
    ok = timed("verify", vf.verify, proof, seed)
    timed("del verifier", vf.__del__)
    assert ok


def run(n):
    lock, nxt, cpu = threading.Lock(), [0], [0.0]

    def work(c):
        while True:
            with lock:
                i = nxt[0]
                nxt[0] += 1
            if i >= n:
                break
            step(c, i)
        with lock:
            cpu[0] += time.thread_time()

    ts = [threading.Thread(target=work, args=(c,)) for c in ctxs]
    p0, w0 = time.process_time(), time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return time.process_time() - p0, time.perf_counter() - w0, cpu[0]


run(inflight)          # warm-up
acc.clear()
proc, wall, workers = run(count)
print("%d in flight, %d statements: %.1f prove+verify/s; process CPU %.1f ms / statement, worker threads %.1f ms, other threads %.1f ms"
      % (inflight, count, count / wall, 1e3 * proc / count, 1e3 * workers / count, 1e3 * (proc - workers) / count))
inside = 0.0
for name, (c, w, k) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    inside += c
    print("  %-24s cpu %7.2f ms  wall %8.2f ms  per statement (%d calls)" % (name, 1e3 * c / count, 1e3 * w / count, k))
print("  %-24s cpu %7.2f ms  per statement" % ("outside wrapped calls", 1e3 * (workers - inside) / count))
keys = ("sync", "commit", "prove", "verify", "rng")
print("  C-ABI thread-CPU counters (ms / statement, whole run incl. warm-up):",
      {k: round(sum(c.get("cpu_%s_ns" % k) for c in ctxs) / 1e6 / (count + inflight), 2) for k in keys})
