"""Profiling driver: a few raw fixed-base MSMs (uniform scalars, 2^lg points) -- used under ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bulletproof_gadgets_b200 as bpg
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 18
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = bpg.Context(0)
n = 1 << lg
ctx.gens_ensure(n // 2)
a = np.random.default_rng(lg).integers(0, 256, size=(n, 32), dtype=np.uint8)
a[:, 31] &= 0x0F
d = torch.from_numpy(a).to("cuda:0")
for _ in range(reps):
    r = ctx.msm_gens_dev(d.data_ptr(), n // 2, d.data_ptr() + 16 * n, n // 2)
print("ok", r.hex()[:16], "launches", ctx.get("launches"))
