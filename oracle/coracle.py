"""ORACLE (test infrastructure): ctypes access to the C restatement (oracle/c/bp_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this."""
import ctypes
from ctypes import c_char_p, c_int, c_long, c_size_t, c_void_p

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(_build.build())
        L.bpo_prove.restype = c_long
        L.bpo_prove.argtypes = [c_char_p, c_size_t, c_char_p, c_char_p, c_size_t, c_char_p, c_char_p, c_char_p, c_size_t,
                                c_void_p, c_void_p, c_char_p, c_size_t, c_char_p, c_int, c_char_p, c_char_p, c_size_t]
        L.bpo_verify.restype = c_int
        L.bpo_verify.argtypes = [c_char_p, c_size_t, c_char_p, c_size_t, c_size_t, c_void_p, c_void_p, c_char_p, c_size_t,
                                 c_char_p, c_size_t, c_char_p, c_int]
        L.bpo_msm.argtypes = [c_char_p, c_char_p, c_size_t, c_int, c_char_p]
        L.bpo_msm_gens.argtypes = [c_char_p, c_size_t, c_char_p, c_size_t, c_char_p, c_char_p, c_int, c_char_p]
        L.bpo_gens.argtypes = [c_int, c_size_t, c_size_t, c_char_p]
        L.bpo_merlin_kat.argtypes = [c_char_p, c_size_t, c_char_p, c_char_p, c_size_t, c_char_p, c_char_p, c_size_t]
        _lib = L
    return _lib


def _sb(x):
    return x if isinstance(x, (bytes, bytearray)) else int(x).to_bytes(32, "little")


class FlatCS:
    """Constraint system flattened to the arrays the C oracle (and the C ABI tests) consume."""

    def __init__(self, constraints):
        import numpy as np
        rs, tv, tc = [0], [], bytearray()
        for terms in constraints:
            for (kind, idx), coeff in terms:
                tv.append((kind << 29) | idx)
                tc += _sb(coeff)
            rs.append(len(tv))
        self.row_start = np.asarray(rs, dtype=np.uint32)
        self.term_var = np.asarray(tv if tv else [0], dtype=np.uint32)
        self.term_coef = bytes(tc) if tc else bytes(32)
        self.q = len(constraints)


def _flat_cs(st):
    cs = FlatCS([])
    cs.row_start, cs.term_var, cs.term_coef, cs.q = st.row_start, st.term_var, st.term_coef, st.q
    return cs


def prove_flat(st, seed, cache_gens=True):
    """Proves a bulletproof_gadgets_b200.workloads.FlatStatement (a_O = a_L * a_R mod l on the host)."""
    import numpy as np
    m, n = st.m, st.n
    L = 2**252 + 27742317777372353535851937790883648493
    aO = bytearray(32 * n)
    for i in range(n):
        l = int.from_bytes(st.aL[32 * i: 32 * i + 32], "little")
        r = int.from_bytes(st.aR[32 * i: 32 * i + 32], "little")
        if l and r:
            aO[32 * i: 32 * i + 32] = (l * r % L).to_bytes(32, "little")
    cs = _flat_cs(st)
    V = ctypes.create_string_buffer(32 * max(m, 1))
    cap = 1 + 14 * 32 + 66 * 32
    out = ctypes.create_string_buffer(cap)
    j = lambda xs: b"".join(_sb(x) for x in xs) or None
    ln = lib().bpo_prove(st.label, len(st.label), j(st.v), j(st.vbl), m, st.aL or None, st.aR or None, bytes(aO) or None,
                         n, cs.row_start.ctypes.data, cs.term_var.ctypes.data, cs.term_coef, cs.q, seed,
                         int(cache_gens), V, out, cap)
    if ln < 0:
        raise RuntimeError("bpo_prove failed")
    return out.raw[:ln], [V.raw[32 * i: 32 * i + 32] for i in range(m)]


def verify_flat(st, V, proof, seed, cache_gens=True):
    return verify(st.label, V, st.n, _flat_cs(st), proof, seed, cache_gens)


def prove(label, v, vbl, aL, aR, aO, constraints, seed, cache_gens=True):
    cs = constraints if isinstance(constraints, FlatCS) else FlatCS(constraints)
    m, n = len(v), len(aL)
    V = ctypes.create_string_buffer(32 * max(m, 1))
    cap = 1 + 14 * 32 + 66 * 32
    out = ctypes.create_string_buffer(cap)
    j = lambda xs: b"".join(_sb(x) for x in xs) or None
    ln = lib().bpo_prove(label, len(label), j(v), j(vbl), m, j(aL), j(aR), j(aO), n, cs.row_start.ctypes.data,
                         cs.term_var.ctypes.data, cs.term_coef, cs.q, seed, int(cache_gens), V, out, cap)
    if ln < 0:
        raise RuntimeError("bpo_prove failed")
    return out.raw[:ln], [V.raw[32 * i: 32 * i + 32] for i in range(m)]


def verify(label, V, n, constraints, proof, seed, cache_gens=True):
    cs = constraints if isinstance(constraints, FlatCS) else FlatCS(constraints)
    rc = lib().bpo_verify(label, len(label), b"".join(V) or None, len(V), n, cs.row_start.ctypes.data,
                          cs.term_var.ctypes.data, cs.term_coef, cs.q, proof, len(proof), seed, int(cache_gens))
    return {1: True, 0: False, -1: "format"}[rc]


def msm(scalars, points, constant_time=False):
    out = ctypes.create_string_buffer(32)
    rc = lib().bpo_msm(b"".join(_sb(s) for s in scalars), b"".join(points), len(points), int(constant_time), out)
    return None if rc else out.raw


def msm_gens(sG=b"", sH=b"", sB=None, sBb=None, constant_time=False):
    out = ctypes.create_string_buffer(32)
    lib().bpo_msm_gens(sG or None, len(sG) // 32, sH or None, len(sH) // 32, sB, sBb, int(constant_time), out)
    return out.raw


def gens(which, start, count):
    out = ctypes.create_string_buffer(32 * max(count, 1))
    lib().bpo_gens({"G": 0, "H": 1, "B": 2, "B_blinding": 3}[which], start, count, out)
    return [out.raw[32 * i: 32 * i + 32] for i in range(count)]
