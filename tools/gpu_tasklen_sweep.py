"""Dev probe (GPU): whole-MSM time at 2^18 points (uniform scalars) for several chunk lengths (task_len)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bulletproof_gadgets_b200 as bpg
ctx = bpg.Context(0)
n = 1 << 17
ctx.gens_ensure(n)
a = np.random.default_rng(1).integers(0, 256, size=(2 * n, 32), dtype=np.uint8); a[:, 31] &= 0x0F
d = torch.from_numpy(a).cuda()
ref = None
for tl in (16, 24, 32, 40, 48, 64, 96):
    ctx.set("task_len", tl)
    for _ in range(3): out = ctx.msm_gens_dev(d.data_ptr(), n, d.data_ptr() + 32 * n, n)
    if ref is None: ref = out
    assert out == ref
    t0 = time.perf_counter()
    for _ in range(20): ctx.msm_gens_dev(d.data_ptr(), n, d.data_ptr() + 32 * n, n)
    dt = (time.perf_counter() - t0) / 20
    ctx.set("time_accum", 1); ctx.msm_gens_dev(d.data_ptr(), n, d.data_ptr() + 32 * n, n); acc = ctx.get("accum_us"); ctx.set("time_accum", 0)
    print("task_len %3d: whole MSM %.1f us, accumulate %d us" % (tl, dt * 1e6, acc))
