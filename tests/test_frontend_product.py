"""The product's C++ statement front end (libbpg.so: bpg_frontend_flatten_*; pure host code) against the oracle's
python front end on the reference's 13 fixtures and on small synthetic statements: identical flat circuits
(commit order and names, blindings, multiplier assignments, CSR constraints) on both the prover and the verifier side."""
import ctypes

import numpy as np
import pytest

import bulletproof_gadgets_b200 as bpg
from bulletproof_gadgets_b200 import _capi, build
from oracle.pyref import frontend as F
from tests import frontend_glue as G


@pytest.fixture(scope="module")
def lib():
    build.build_lib()
    return bpg.lib()


def _c_flat(lib, side, name, inst, other, gad, seed=G.SEED_BLIND[:32].ljust(32, b"\0")):
    out = ctypes.POINTER(_capi.FlatStatementC)()
    if side == "prover":
        rc = lib.bpg_frontend_flatten_prover(name.encode(), inst.encode(), other.encode(), gad.encode(), seed, ctypes.byref(out))
    else:
        rc = lib.bpg_frontend_flatten_verifier(name.encode(), inst.encode(), other.encode(), gad.encode(), ctypes.byref(out))
    if rc:
        return rc, (lib.bpg_last_error() or b"").decode()
    f = out.contents
    n, m, q, nnz = f.n, f.m, f.q, f.nnz
    get = lambda p, k: bytes(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8 * k)).contents) if k and p else b""
    d = {"n": n, "m": m, "q": q, "nnz": nnz,
         "v": get(f.v32m, 32 * m) if side == "prover" else b"", "vbl": get(f.vbl32m, 32 * m) if side == "prover" else b"",
         "V": get(f.V32m, 32 * m) if side == "verifier" else b"",
         "aL": get(f.aL32n, 32 * n) if side == "prover" else b"", "aR": get(f.aR32n, 32 * n) if side == "prover" else b"",
         "row_start": np.frombuffer(get(f.row_start, 4 * (q + 1)), dtype=np.uint32).copy(),
         "term_var": np.frombuffer(get(f.term_var, 4 * nnz), dtype=np.uint32).copy(),
         "term_coef": get(f.term_coef32, 32 * nnz), "names": f.com_names.decode().split("\n")[:-1]}
    lib.bpg_flat_statement_free(out)
    return 0, d


def _c_blinding(seed32):
    """the C++ front end derives blindings as SHAKE256(seed32 || LE64(k)) -> wide reduction"""
    return G.blinding(seed32)


def _canon(coef_bytes):
    L = F.L
    return b"".join((int.from_bytes(coef_bytes[i: i + 32], "little") % L).to_bytes(32, "little") for i in range(0, len(coef_bytes), 32))


def _assert_same_prover(d, st):
    assert (d["n"], d["m"], d["q"], d["nnz"]) == (st.n, st.m, st.q, st.nnz)
    assert d["names"] == st.com_names
    assert d["v"] == st.v_bytes and d["vbl"] == st.vbl_bytes
    assert d["aL"] == st.aL and d["aR"] == st.aR
    assert (d["row_start"] == st.row_start).all() and (d["term_var"][: st.nnz] == st.term_var[: st.nnz]).all()
    assert _canon(d["term_coef"]) == _canon(st.term_coef[: 32 * st.nnz])


@pytest.mark.parametrize("stem", G.STEMS)
def test_fixture_flat_circuits_identical(lib, stem):
    inst, wtns, gad = G.load(stem)
    seed = b"\x42" * 32
    st = F.compile_prover(stem, inst, wtns, gad, _c_blinding(seed))
    rc, d = _c_flat(lib, "prover", stem, inst, wtns, gad, seed)
    assert rc == 0, d
    _assert_same_prover(d, st)
    # verifier side: any 32-byte strings will do as commitments for the flattening
    text = st.coms_text([bytes([i % 251]) * 32 for i in range(st.m)])
    vs = F.compile_verifier(stem, inst, text, gad)
    rc, dv = _c_flat(lib, "verifier", stem, inst, text, gad)
    assert rc == 0, dv
    assert (dv["n"], dv["m"], dv["q"], dv["nnz"]) == (vs.n, vs.m, vs.q, vs.nnz)
    assert dv["V"] == b"".join(vs.V) and dv["names"] == vs.com_names
    assert (dv["row_start"] == vs.row_start).all() and (dv["term_var"][: vs.nnz] == vs.term_var[: vs.nnz]).all()
    assert _canon(dv["term_coef"]) == _canon(vs.term_coef[: 32 * vs.nnz])
    # and the verifier's circuit is the prover's
    assert (dv["row_start"] == d["row_start"]).all() and (dv["term_var"] == d["term_var"]).all()


SMALL = [
    ("EQUALS W0 I0", "I0 = 0x05", "W0 = 0x06"),
    ("EQUALS W0 W1", "", "W0 = 0x%s\nW1 = 0x07" % ("11" * 40)),
    ("LESS_THAN W0 W1", "", "W0 = 0x09\nW1 = 0x08"),
    ("SET_MEMBER W0 I0 W1 I1", "I0 = 0x01\nI1 = 0x02", "W0 = 0x03\nW1 = 0x03"),
    ("SET_MEMBER I0 W0 W1", "I0 = 0x%s" % ("ab" * 40), "W0 = 0x%s\nW1 = 0x05" % ("ab" * 40)),     # hashing path, instance member
    ("UNEQUAL W0 I0", "I0 = 0x2a", "W0 = 0x%s" % ("ff" * 32)),
    ("UNEQUAL I0 W0", "I0 = 0x%s" % ("2a" * 33), "W0 = 0x2b"),
    ("HASH W1 W0", "", "W0 = 0x%s\nW1 = 0x01" % ("cd" * 32)),                                     # full last block: extra padding block
    ("MERKLE W2 ((W0 I0) W1)", "I0 = 0x1234", "W0 = 0x01\nW1 = 0x02\nW2 = 0x03"),
    ("OR\n[\n{\nEQUALS W0 I0\nBOUND W0 I0 I1\n}\n{\nEQUALS W0 I1\n}\n]\nLESS_THAN W0 W1", "I0 = 0x01\nI1 = 0x02", "W0 = 0x02\nW1 = 0x09"),
    ("", "", ""),
]


@pytest.mark.parametrize("case", range(len(SMALL)))
def test_small_statements_identical(lib, case):
    gad, inst, wtns = SMALL[case]
    seed = bytes([case + 1]) * 32
    st = F.compile_prover("small", inst, wtns, gad, _c_blinding(seed))
    rc, d = _c_flat(lib, "prover", "small", inst, wtns, gad, seed)
    assert rc == 0, d
    _assert_same_prover(d, st)


PANICS = [("FROBNICATE W0", "", "W0 = 0x01"), ("EQUALS W0 I9", "", "W0 = 0x01"), ("EQUALS W0 W0", "", "W0 = 0x1"),
          ("BOUND W0 I0 I1", "I0 = 0x00\nI1 = 0xff", "W0 = 0x%s" % ("01" * 33)), ("OR\n[\n{", "", "W0 = 0x01") ,
          ("MERKLE I0 (W0 I1", "I0 = 0x01\nI1 = 0x02", "W0 = 0x01"), ("LESS_THAN W0 I0", "I0 = 0x01", "W0 = 0x01")]


@pytest.mark.parametrize("case", range(len(PANICS)))
def test_front_end_errors_match_oracle_panics(lib, case):
    gad, inst, wtns = PANICS[case]
    try:
        F.compile_prover("p", inst, wtns, gad, G.blinding())
        oracle_ok = True
    except F.FrontendPanic:
        oracle_ok = False
    rc, msg = _c_flat(lib, "prover", "p", inst, wtns, gad)
    assert (rc == 0) == oracle_ok, (rc, msg)
    if not oracle_ok:
        assert rc == _capi.E_GADGET and msg.startswith("front end:")


# ---------------------------------------------------------------------------------------------------------------------
# random statements: every gadget kind, nested OR blocks, variables of 1..40 bytes (multi-limb variables take the hashing
# and panic paths).  The C++ front end must agree with the oracle's on the flat circuit, or both must reject the input.
def _random_statement(seed):
    rng = np.random.default_rng(seed)
    n_i, n_w = int(rng.integers(2, 6)), int(rng.integers(2, 6))

    def val():
        k = int(rng.choice([1, 2, 8, 15, 31, 32, 33, 40], p=[.2, .15, .2, .1, .1, .1, .1, .05]))
        b = rng.integers(0, 256, size=k, dtype=np.uint8).tobytes()
        return b if b.strip(b"\0") else b"\x01" + b[1:]

    pool = [val() for _ in range(3)]           # repeated values make EQUALS / SET_MEMBER hit
    pick = lambda: pool[int(rng.integers(0, len(pool)))] if rng.random() < 0.4 else val()
    inst = "".join("I%d = 0x%s\n" % (i, pick().hex()) for i in range(n_i))
    wtns = "".join("W%d = 0x%s\n" % (i, pick().hex()) for i in range(n_w))
    I = lambda: "I%d" % int(rng.integers(0, n_i))
    Wv = lambda: "W%d" % int(rng.integers(0, n_w))
    any_v = lambda: I() if rng.random() < 0.5 else Wv()

    def tree(depth):
        kids = []
        for _ in range(2):
            kids.append(tree(depth - 1) if depth > 0 and rng.random() < 0.4 else any_v())
        return "(%s %s)" % (kids[0], kids[1])

    def line(in_or):
        # range proofs inside OR clauses multiply into 10^5+ product constraints (or_conjunction.rs:20-37): top level only
        k = int(rng.choice([1, 2, 3, 5, 6])) if in_or else int(rng.integers(0, 7))
        if k == 0:
            return "BOUND %s %s %s" % (Wv(), I(), I())
        if k == 1:
            return "HASH %s %s" % (any_v(), Wv())
        if k == 2:
            a, b = (Wv(), any_v()) if rng.random() < 0.7 else (I(), Wv())
            return "EQUALS %s %s" % (a, b)
        if k == 3:
            a, b = (Wv(), any_v()) if rng.random() < 0.7 else (I(), Wv())
            return "UNEQUAL %s %s" % (a, b)
        if k == 4:
            return "LESS_THAN %s %s" % (Wv(), Wv())
        if k == 5:
            return "SET_MEMBER %s %s" % (any_v(), " ".join(any_v() for _ in range(int(rng.integers(1, 4)))))
        return "MERKLE %s %s" % (any_v(), tree(1))

    def block(depth):
        out = []
        for _ in range(int(rng.integers(1, 3))):
            if depth < 2 and rng.random() < 0.3:
                out += ["OR", "["]
                for _ in range(int(rng.integers(1, 4))):
                    out += ["{"] + block(depth + 1) + ["}"]
                out += ["]"]
            else:
                out.append(line(depth > 0))
        return out

    return "\n".join(block(0)) + "\n", inst, wtns


@pytest.mark.parametrize("seed", range(150))
def test_random_statements_agree_with_oracle(lib, seed):
    gad, inst, wtns = _random_statement(1000 + seed)
    bseed = bytes([seed % 255 + 1]) * 32
    try:
        st = F.compile_prover("rnd", inst, wtns, gad, _c_blinding(bseed))
    except F.FrontendPanic:
        st = None
    rc, d = _c_flat(lib, "prover", "rnd", inst, wtns, gad, bseed)
    assert (rc == 0) == (st is not None), (gad, inst, wtns, rc, d if rc else "")
    if st is None:
        assert rc == _capi.E_GADGET
        return
    _assert_same_prover(d, st)
    text = st.coms_text([bytes([(7 * i + seed) % 251]) * 32 for i in range(st.m)])
    vs = F.compile_verifier("rnd", inst, text, gad)
    rc, dv = _c_flat(lib, "verifier", "rnd", inst, text, gad)
    assert rc == 0, dv
    assert (dv["n"], dv["m"], dv["q"], dv["nnz"]) == (vs.n, vs.m, vs.q, vs.nnz)
    assert dv["V"] == b"".join(vs.V) and dv["names"] == vs.com_names
    assert (dv["row_start"] == vs.row_start).all() and (dv["term_var"][: vs.nnz] == vs.term_var[: vs.nnz]).all()
    assert _canon(dv["term_coef"]) == _canon(vs.term_coef[: 32 * vs.nnz])


def test_front_end_is_thread_safe(lib):
    """The front end keeps per-thread scratch between statements; concurrent callers must get what a lone caller gets."""
    import threading
    cases = [_random_statement(1000 + s) for s in (0, 1, 2, 3, 5, 6, 7, 8)]
    seeds = [bytes([k + 1]) * 32 for k in range(len(cases))]

    def flat(k):
        gad, inst, wtns = cases[k]
        rc, d = _c_flat(lib, "prover", "rnd", inst, wtns, gad, seeds[k])
        return rc, (d if rc else (d["n"], d["q"], d["aL"], d["aR"], d["term_coef"], d["term_var"].tobytes(), d["v"], d["vbl"]))

    alone = [flat(k) for k in range(len(cases))]
    got, errs = {}, []

    def work(t):
        try:
            for rep in range(6):
                k = (t + rep) % len(cases)
                got[(t, rep)] = (k, flat(k))
        except BaseException as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs
    for (t, rep), (k, res) in got.items():
        if alone[k][0]:
            assert res[0] == alone[k][0]
        else:
            assert res == alone[k], (t, rep, k)
