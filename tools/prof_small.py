"""Profiling driver for small statements (BASELINE config 4): `reps` x (LESS_THAN n=379, SET_MEMBER n=32) through the text
entry points on ONE context, proved and verified.  Used under ncu (launch list, full capture of the small-statement kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bulletproof_gadgets_b200 as bpg
from bulletproof_gadgets_b200 import workloads as W
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ctx = bpg.Context(0)
texts = W.batch_texts(2 * reps)
jobs = [("batch-%d" % i, t[1], t[2], t[0]) for i, t in enumerate(texts)]
sd = [(i + 1).to_bytes(32, "little") for i in range(len(jobs))]
out = bpg.prove_text_batch([ctx], jobs, sd, sd, verify=True)
assert all(o[0] == 0 and o[3] for o in out)
print("ok launches", ctx.get("launches"))
