#!/usr/bin/env python3
"""Dev probe (GPU): k_accumulate time and whole-MSM wall time against the chunk geometry (fixed task_len, or the number of
chunks aimed at by the device-derived geometry) at 2^18 / 2^20 uniform scalars."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bulletproof_gadgets_b200 as bpg  # noqa: E402

ctx = bpg.Context(0)
lgs = [int(x) for x in sys.argv[1:]] or [18, 20]
ctx.gens_ensure(1 << (max(lgs) - 1))
dev = torch.device("cuda", 0)
for lg in lgs:
    n = 1 << lg
    h = n // 2
    a = np.random.default_rng(lg).integers(0, 256, size=(n, 32), dtype=np.uint8)
    a[:, 31] &= 0x0F
    d = torch.from_numpy(a).to(dev)
    call = lambda: ctx.msm_gens_dev(d.data_ptr(), h, d.data_ptr() + 32 * h, h)
    waves = 148 * 4 * 128
    for key, val in [("acc_variant", 0), ("acc_variant", 1), ("acc_variant", 2)] + [("task_len", t) for t in (8, 16, 32)] + \
                    [("target_chunks", int(waves * f)) for f in (1.0, 1.5, 2.0, 4.0)]:
        ctx.set("task_len", 0)
        ctx.set("target_chunks", 0)
        ctx.set("acc_variant", 0)
        ctx.set(key, val)
        for _ in range(3):
            call()
        t0 = time.perf_counter()
        for _ in range(20):
            call()
        wall = (time.perf_counter() - t0) / 20
        ctx.set("time_accum", 1)
        for _ in range(5):
            call()
        k = ctx.get("timed_msms")
        acc = ctx.get("stage_ns_3") / k / 1e3
        red = ctx.get("stage_ns_4") / k / 1e3
        cl = ctx.get("last_chunk_len")
        ctx.set("time_accum", 0)
        print(json.dumps({"lg": lg, key: val, "chunk_len": cl, "whole_msm_us": round(wall * 1e6, 1), "accumulate_us": round(acc, 1),
                          "reduce_us": round(red, 1)}), flush=True)
ctx.close()
