#!/usr/bin/env python3
"""Dev probe (GPU): per-stage CUDA-event times of the fixed-base MSM (time_accum mode) and whole-MSM wall time, for
uniform and 0/1 scalars at several sizes.  One JSON line per measurement.
    python tools/gpu_msm_stages.py [lg ...]     (default 14 16 18 20 22)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bulletproof_gadgets_b200 as bpg  # noqa: E402

STAGES = ["digits0", "scan", "digits1", "accumulate", "bucket_reduce", "final", "total"]
IMAD_WIDE_PEAK = float(os.environ.get("BPG_IMAD_PEAK", "8.157e12"))
lgs = [int(x) for x in sys.argv[1:]] or [14, 16, 18, 20, 22]
ctx = bpg.Context(0)
ctx.gens_ensure(1 << (max(lgs) - 1))
dev = torch.device("cuda", 0)
for lg in lgs:
    n = 1 << lg
    h = n // 2
    for kind in ("uniform", "bits"):
        rng = np.random.default_rng(lg)
        if kind == "uniform":
            a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
            a[:, 31] &= 0x0F
        else:
            a = np.zeros((n, 32), dtype=np.uint8)
            a[:, 0] = rng.integers(0, 2, size=n, dtype=np.uint8)
        d = torch.from_numpy(a).to(dev)
        call = lambda: ctx.msm_gens_dev(d.data_ptr(), h, d.data_ptr() + 32 * h, h)
        for _ in range(3):
            call()
        reps = 20 if lg <= 20 else 8
        runs = []
        for _ in range(3):
            t0 = time.perf_counter()
            for _ in range(reps):
                call()
            runs.append((time.perf_counter() - t0) / reps)
        wall = sorted(runs)[1]
        ctx.set("time_accum", 1)
        for _ in range(5):
            call()
        k = ctx.get("timed_msms")
        st = {s: ctx.get("stage_ns_%d" % i) / k / 1e3 for i, s in enumerate(STAGES)}
        entries, cl = ctx.get("accum_entries"), ctx.get("last_chunk_len")
        ctx.set("time_accum", 0)
        print(json.dumps({"lg": lg, "scalars": kind, "whole_msm_us": round(wall * 1e6, 1), "mpoints_per_s": round(n / wall / 1e6, 1),
                          "frac_of_imad_peak_whole_msm": round(16 * n * 504 / wall / IMAD_WIDE_PEAK, 4) if kind == "uniform" else None,
                          "entries": entries, "chunk_len": cl, "stage_us": {s: round(v, 1) for s, v in st.items()},
                          "accumulate_frac_of_imad_peak": round(entries * 504 / (st["accumulate"] * 1e-6) / IMAD_WIDE_PEAK, 4)}), flush=True)
ctx.close()
