#!/usr/bin/env python3
"""BASELINE config 5, non-fixed-base variant: bpg_msm over random ristretto points, 2^10 .. 2^20 (host buffers in, 32 bytes
out: decompression, upload and the window combine are inside the time).  Points are 4096 distinct random points repeated
(decoding cost is per input either way).  One JSON line per size; BPG_VARBASE_MIN=999999999 forces the thread-per-point path."""
import hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bulletproof_gadgets_b200 as bpg
from bulletproof_gadgets_b200 import _capi
import ctypes
ctx = bpg.Context(0)
ctx.gens_ensure(4096)
base = b"".join(ctx.gens_compressed("G", 0, 4096))
for lg in [int(x) for x in sys.argv[1:]] or [10, 12, 14, 16, 18, 20]:
    n = 1 << lg
    pts = (base * (n // 4096 + 1))[: 32 * n]
    a = np.random.default_rng(lg).integers(0, 256, size=(n, 32), dtype=np.uint8)
    a[:, 31] &= 0x0F
    # page-locked host buffers (bpg_host_alloc): the upload is DMA at PCIe speed instead of a staged pageable copy
    sc_pin, pt_pin = bpg.pinned_copy(a.tobytes(), np.uint8), bpg.pinned_copy(pts, np.uint8)
    sc_p, pt_p = ctypes.cast(sc_pin.ctypes.data, ctypes.c_char_p), ctypes.cast(pt_pin.ctypes.data, ctypes.c_char_p)
    out = ctypes.create_string_buffer(32)
    call = lambda: _capi.check(bpg.lib().bpg_msm(ctx._h, sc_p, pt_p, n, out))
    for _ in range(2):
        call()
    reps = 5 if lg >= 18 else 20
    t0 = time.perf_counter()
    for _ in range(reps):
        call()
    dt = (time.perf_counter() - t0) / reps
    print(json.dumps({"lg": lg, "variable_base": True, "ms": round(dt * 1e3, 3), "mpoints_per_s": round(n / dt / 1e6, 2),
                      "path": "pippenger" if n >= int(os.environ.get("BPG_VARBASE_MIN", "512")) else "thread per point",
                      "out": out.raw.hex()[:16]}), flush=True)
ctx.close()
