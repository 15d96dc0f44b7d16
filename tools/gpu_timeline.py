"""Concurrent-kernel timeline of the throughput regime (nsys is not installed: CUPTI through torch.profiler).

Runs `n` resident config-2 statements through bpg_r1cs_prove_batch with `inflight` contexts under torch.profiler (CUDA
activities only), then reduces the kernel records to: the time at least one kernel was running (union), the per-kernel
count / mean duration under concurrency (to compare with the serialised durations of the ncu launch lists), the time-
weighted number of kernels resident at once, and how much of the wall time had a k_accumulate running.

usage: python tools/gpu_timeline.py [n=64] [inflight=48] [count=1024]  -> one JSON line
"""
import collections, json, os, re, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bulletproof_gadgets_b200 as bpg
from bulletproof_gadgets_b200 import workloads as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
inflight = int(sys.argv[2]) if len(sys.argv) > 2 else 48
count = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
ctx0 = bpg.Context(0)
ctxs = [ctx0] + [ctx0.shared() for _ in range(inflight - 1)]
st = W.bounds_check_statement(count, seed=20261018, label=b"bench-bound-0").pin(bpg)
ctx0.gens_ensure(st.n)
circuit = bpg.Circuit(ctx0, st.n, st.m, st.row_start, st.term_var, st.term_coef, st.q).set_witness(st.aL, st.aR)


VERIFY = not os.environ.get("TIMELINE_NOVERIFY")
C4 = os.environ.get("TIMELINE_MODE") == "c4"   # BASELINE config 4 (small LESS_THAN / SET_MEMBER statements, text in)
if C4:
    texts = W.batch_texts(max(n, 4 * inflight))
    jobs = [("batch-%d" % i, t[1], t[2], t[0]) for i, t in enumerate(texts)]


def run(first, k):
    if C4:
        sd = [(i + 1).to_bytes(32, "little") for i in range(k)]
        out = bpg.prove_text_batch(ctxs, jobs[:k], sd, sd, verify=True)
        assert all(o[0] == 0 and o[3] for o in out)
        return
    sd = [(first + i + 1).to_bytes(32, "little") for i in range(k)]
    out = bpg.prove_batch(ctxs, [st] * k, sd, circuits=circuit, verify=VERIFY, verify_seeds=sd if VERIFY else None)
    assert os.environ.get("BPG_X_SKIP") or all(o[0] == 0 for o in out)


PHASES = ["setup", "rng", "phase1", "poly", "t_commit", "ipp_early", "ipp_fold", "ipp_late", "final", "v_launch", "v_scalars", "v_wait"]
run(0, 4 * inflight if C4 else max(inflight, 32))
torch.cuda.synchronize()
ph0 = [sum(c.get("phase_ns_%d" % i) for c in ctxs) for i in range(len(PHASES))]
CPU_KEYS = ("sync", "commit", "prove", "verify", "rng", "load")
cpu0 = {k: sum(c.get("cpu_%s_ns" % k) for c in ctxs) for k in CPU_KEYS}
pcpu0 = time.process_time()
t0 = time.perf_counter()
acts = [ProfilerActivity.CUDA] + ([ProfilerActivity.CPU] if os.environ.get("TIMELINE_DUMP") else [])
with profile(activities=acts) as prof:
    run(1000, n)
    torch.cuda.synchronize()
wall = time.perf_counter() - t0
pcpu = time.process_time() - pcpu0
cpu = {k: round((sum(c.get("cpu_%s_ns" % k) for c in ctxs) - cpu0[k]) / 1e6 / n, 3) for k in CPU_KEYS}
ph = {nm: round((sum(c.get("phase_ns_%d" % i) for c in ctxs) - ph0[i]) / 1e6 / n, 2) for i, nm in enumerate(PHASES)}
kev = [e for e in prof.profiler.kineto_results.events() if e.device_type() == torch.autograd.DeviceType.CUDA]
raw = [(e.start_ns() / 1e3, (e.start_ns() + e.duration_ns()) / 1e3, re.sub(r"\(.*", "", e.name()).replace("void ", ""),
        e.device_resource_id(), e.correlation_id()) for e in kev]
raw.sort()
if os.environ.get("TIMELINE_DUMP"):  # raw records for offline analysis: start_us end_us stream correlation kernel|memcpy|api
    import gzip
    t_base = raw[0][0]
    with gzip.open(os.environ["TIMELINE_DUMP"], "wt") as f:
        for s_, e_, nm, sid, cid in raw:
            f.write("G %.3f %.3f %d %d %s\n" % (s_ - t_base, e_ - t_base, sid, cid, nm.replace(" ", "_")))
        for e in prof.profiler.kineto_results.events():
            if e.device_type() == torch.autograd.DeviceType.CPU and e.name().startswith("cuda"):
                f.write("H %.3f %.3f %d %d %s\n" % (e.start_ns() / 1e3 - t_base, (e.start_ns() + e.duration_ns()) / 1e3 - t_base,
                                                    e.start_thread_id(), e.correlation_id(), e.name()))
raw = [(s, e, nm, sid) for s, e, nm, sid, _ in raw if not nm.lower().startswith("mem")]
ev = [(s, e, nm) for s, e, nm, _ in raw]
lo, hi = ev[0][0], max(e for _, e, _ in ev)
span = hi - lo
# sweep line: union, concurrency histogram, accumulate coverage
pts = []
for s, e, nm in ev:
    acc = nm.startswith("k_accumulate")
    pts.append((s, 1, acc)); pts.append((e, -1, acc))
pts.sort(key=lambda p: (p[0], p[1]))
live = live_acc = 0
last = lo
union = acc_cov = 0.0
hist = collections.Counter()
for t, d, acc in pts:
    dt = t - last
    if dt > 0:
        if live: union += dt
        if live_acc: acc_cov += dt
        hist[min(live, 8)] += dt
        hist["acc%d" % min(live_acc, 4)] += dt
    last = t
    live += d
    if acc: live_acc += d
agg = collections.OrderedDict()
for s, e, nm in ev:
    a = agg.setdefault(nm, [0, 0.0]); a[0] += 1; a[1] += e - s
top = sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]
print(json.dumps({
    "statements": n, "inflight": inflight, "wall_ms": wall * 1e3, "kernel_span_ms": span / 1e3,
    "per_statement_ms": span / 1e3 / n, "union_busy_frac": union / span, "accumulate_running_frac": acc_cov / span,
    "host_cpu_ms_per_statement": {"process": round(pcpu * 1e3 / n, 3), "inside_library_calls": cpu, "cores": os.cpu_count()},
    "wall_ms_per_statement_by_phase": ph, "phase_sum_ms": round(sum(ph.values()), 1),
    "kernels": len(ev), "sum_kernel_ms_per_statement": sum(v[1] for v in agg.values()) / 1e3 / n,
    "concurrency_time_frac": {str(k): round(v / span, 4) for k, v in sorted(hist.items(), key=lambda kv: str(kv[0]))},
    "top": [{"kernel": k, "per_statement": round(c / n, 2), "mean_us": round(t / c, 1), "ms_per_statement": round(t / 1e3 / n, 3)} for k, (c, t) in top],
}))
