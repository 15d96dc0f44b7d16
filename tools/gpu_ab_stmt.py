#!/usr/bin/env python3
"""Dev probe (GPU): throughput of the text-statement path (bpg_prove_batch with verify) for config 2, a few hundred statements,
printing statements/s and the host CPU seconds per statement.  BPG_F3=0/1 switches device-side witness generation."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bulletproof_gadgets_b200 as bpg
from bulletproof_gadgets_b200 import workloads as W
n = int(sys.argv[1]) if len(sys.argv) > 1 else 192
inflight = int(sys.argv[2]) if len(sys.argv) > 2 else 48
ctx0 = bpg.Context(0)
ctxs = [ctx0] + [ctx0.shared() for _ in range(inflight - 1)]
gad, inst, wtns = W.bounds_check_text(1024)
ctx0.gens_ensure(1 << 17)
job = ("ab", inst, wtns, gad)
sd = [(i + 1).to_bytes(32, "little") for i in range(n)]
bpg.prove_text_batch(ctxs, [job] * inflight, sd[:inflight], sd[:inflight], verify=True)
c0, t0 = time.process_time(), time.perf_counter()
out = bpg.prove_text_batch(ctxs, [job] * n, sd, sd, verify=True)
dt, cpu = time.perf_counter() - t0, time.process_time() - c0
assert all(o[0] == 0 and o[3] for o in out)
print("BPG_F3=%s: %.1f statements/s, %.1f ms CPU per statement" % (os.environ.get("BPG_F3", "1"), n / dt, 1e3 * cpu / n), flush=True)
