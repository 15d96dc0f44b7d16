"""Dev check (GPU): generators, MSM vs the python oracle, timing sweep.  Not part of the test-suite."""
import ctypes, hashlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.pyref import ed, r1cs
from oracle.pyref.merlin import L
lib = ctypes.CDLL(os.path.join(os.path.dirname(__file__), "..", "bulletproof_gadgets_b200", "libbpg.so"))
lib.bpg_last_error.restype = ctypes.c_char_p
lib.bpg_ctx_get.restype = ctypes.c_int64
lib.bpg_ctx_get.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
lib.bpg_ctx_set.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64]
lib.bpg_gens_ensure.argtypes = [ctypes.c_void_p, ctypes.c_uint64]
lib.bpg_gens_compressed.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_char_p]
lib.bpg_msm_gens.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_uint64, ctypes.c_char_p, ctypes.c_uint64, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]
def ck(rc):
    if rc: raise SystemExit("rc=%d %s" % (rc, lib.bpg_last_error()))
ctx = ctypes.c_void_p(); ck(lib.bpg_ctx_create(0, ctypes.byref(ctx)))
t0 = time.time(); ck(lib.bpg_gens_ensure(ctx, 64)); print("gens(64) %.3fs" % (time.time() - t0))
pg = r1cs.PedersenGens(); bg = r1cs.BulletproofGens(64)
buf = ctypes.create_string_buffer(32 * 64)
ck(lib.bpg_gens_compressed(ctx, 0, 0, 64, buf)); assert all(buf.raw[32*i:32*i+32] == bg.G[i].compress() for i in range(64)), "G mismatch"
ck(lib.bpg_gens_compressed(ctx, 1, 0, 64, buf)); assert all(buf.raw[32*i:32*i+32] == bg.H[i].compress() for i in range(64)), "H mismatch"
ck(lib.bpg_gens_compressed(ctx, 2, 0, 1, buf)); assert buf.raw[:32] == pg.B.compress(), "B"
ck(lib.bpg_gens_compressed(ctx, 3, 0, 1, buf)); assert buf.raw[:32] == pg.B_blinding.compress(), "Bb"
print("generators match oracle")
def sc(seed, i): return int.from_bytes(hashlib.shake_256(b"%s-%d" % (seed, i)).digest(64), "little") % L
def msm(sG, sH, sB, sBb):
    out = ctypes.create_string_buffer(32)
    bG = b"".join(s.to_bytes(32, "little") for s in sG); bH = b"".join(s.to_bytes(32, "little") for s in sH)
    ck(lib.bpg_msm_gens(ctx, bG or None, len(sG), bH or None, len(sH), None if sB is None else sB.to_bytes(32, "little"), None if sBb is None else sBb.to_bytes(32, "little"), out))
    return out.raw
def omsm(sG, sH, sB, sBb):
    acc = ed.msm(sG, bg.G[:len(sG)]) + ed.msm(sH, bg.H[:len(sH)])
    if sB is not None: acc = acc + pg.B * sB
    if sBb is not None: acc = acc + pg.B_blinding * sBb
    return acc.compress()
cases = [([], [], None, None), ([1], [], None, None), ([], [], 1, None), ([], [], None, 5), ([0, 0], [0], 0, 0),
         ([L - 1], [2], 3, L - 4), ([1] * 64, [1] * 64, None, None), ([2**255 - 1], [2**252], None, None),
         ([sc(b"a", i) for i in range(64)], [sc(b"b", i) for i in range(33)], sc(b"c", 0), sc(b"d", 0)),
         ([i & 1 for i in range(64)], [1 - (i & 1) for i in range(64)], None, sc(b"e", 1)),
         ([32768, 32769, 65535, 65536, 65537, 2**16 * 32768, 2**240], [], None, None)]
for k, c in enumerate(cases):
    got, want = msm(*c), omsm(*c)
    print("case", k, "OK" if got == want else "MISMATCH %s %s" % (got.hex(), want.hex()))
    assert got == want
print("launches", lib.bpg_ctx_get(ctx, b"launches"))
# timing sweep with random scalars (linearity check: msm(s) + msm(t) == msm(s+t) via oracle add of results)
import numpy as np
lib.bpg_ctx_set(ctx, b"time_accum", 1)
for lg in [10, 14, 16, 17, 18]:
    n = 1 << (lg - 1)
    t0 = time.time(); ck(lib.bpg_gens_ensure(ctx, n)); tg = time.time() - t0
    rng = np.random.default_rng(lg)
    def rand_scalars(n):
        a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); a[:, 31] &= 0x0f; return a
    s1G, s1H, s2G, s2H = (rand_scalars(n) for _ in range(4))
    def addmod(a, b):
        out = np.zeros_like(a)
        for i in range(a.shape[0]):
            out[i] = np.frombuffer(((int.from_bytes(a[i].tobytes(), "little") + int.from_bytes(b[i].tobytes(), "little")) % L).to_bytes(32, "little"), dtype=np.uint8)
        return out
    def run(g, h):
        out = ctypes.create_string_buffer(32)
        ck(lib.bpg_msm_gens(ctx, g.tobytes(), n, h.tobytes(), n, None, None, out)); return out.raw
    r1 = run(s1G, s1H); r2 = run(s2G, s2H)
    t0 = time.time(); r3 = run(addmod(s1G, s2G), addmod(s1H, s2H)); 
    best = 1e9
    for _ in range(5):
        t0 = time.time(); run(s1G, s1H); best = min(best, time.time() - t0)
    lin = (ed.decompress(r1) + ed.decompress(r2)).compress() == r3
    acc_us = lib.bpg_ctx_get(ctx, b"accum_us"); ent = lib.bpg_ctx_get(ctx, b"accum_entries")
    print("N=2^%d gens %.2fs  msm e2e %.3f ms  accumulate %d us  entries %d  => %.2f Gadd/s  linear=%s" % (lg, tg, best * 1e3, acc_us, ent, ent / max(acc_us, 1) / 1e3, lin))
    assert lin
print("ALL OK")
