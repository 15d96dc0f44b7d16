"""CPU-side checks of the product's host code (no GPU, no compute through the device):

  * libbpg.so loads and exports every symbol include/bpg.h declares; the ctypes prototype table covers them
  * without a CUDA device the library fails loudly (BPG_E_CUDA) -- there is no CPU fallback
  * the host Merlin transcript behind bpg_transcript_* equals the oracle's (merlin 2.0.1 KAT + random op mixes)
  * the portable (host) branch of the device math headers fe25519/ge25519 agrees with big-int arithmetic
"""
import ctypes
import hashlib
import os
import random
import re
import subprocess

import pytest

import bulletproof_gadgets_b200 as bpg
from bulletproof_gadgets_b200 import _capi, build
from oracle.pyref import ed
from oracle.pyref.merlin import Transcript as OTranscript

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 2**255 - 19


@pytest.fixture(scope="module")
def lib():
    build.build_lib()
    return bpg.lib()


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "bpg.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(bpg_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_are_exported(lib):
    syms = _declared_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), "libbpg.so does not export %s" % s
    missing = [s for s in syms if s not in _capi.PROTOTYPES]
    assert not missing, "ctypes prototypes missing for %s" % missing
    extra = [s for s in _capi.PROTOTYPES if s not in syms]
    assert not extra, "prototypes without a declaration in include/bpg.h: %s" % extra
    # the reference's own C ABI (include/bulletproofs_gadgets.h): same names as interfaces/ios/src/lib.rs:21,45,55
    ref_hdr = open(os.path.join(ROOT, "include", "bulletproofs_gadgets.h")).read()
    ref_syms = sorted(set(re.findall(r"\b(c_prove|c_verify|free_proof)\s*\(", ref_hdr)))
    assert ref_syms == ["c_prove", "c_verify", "free_proof"] == sorted(_capi.REF_PROTOTYPES)
    for s in ref_syms:
        assert hasattr(lib, s), "libbpg.so does not export %s" % s
    # struct ProofArtifacts: four pointer-sized fields in the reference's order (#[repr(C)], usize lengths)
    assert [f[0] for f in _capi.RefProofArtifacts._fields_] == ["commitments", "proof", "proof_len", "proof_cap"]
    assert ctypes.sizeof(_capi.RefProofArtifacts) == 4 * ctypes.sizeof(ctypes.c_void_p)


def test_reference_abi_fails_loudly_without_a_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert not lib.c_prove(b"x", b"", b"W0 = 0x01", b"BOUND W0 I0 I1")     # NULL, no CPU fallback
    assert b"no CPU fallback" in lib.bpg_last_error()
    assert lib.c_verify(b"x", b"", b"", b"", b"\0", 1) is False
    lib.free_proof(None)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    rc = lib.bpg_ctx_create(0, ctypes.byref(h))
    assert rc == _capi.E_CUDA and not h
    assert b"no CPU fallback" in lib.bpg_last_error()
    with pytest.raises(bpg.BpgError):
        bpg.Context(0)


def test_product_does_not_import_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "bulletproof_gadgets_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "libbp_oracle" not in src and "coracle" not in src, f


def test_host_transcript_known_answer(lib):
    T = bpg.Transcript(b"test protocol")
    T.append_message(b"some label", b"some data")
    assert T.challenge_bytes(b"challenge", 32).hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"


def test_host_transcript_matches_oracle_on_random_ops(lib):
    rnd = random.Random(2001)
    for trial in range(20):
        label = bytes(rnd.randrange(256) for _ in range(rnd.randrange(0, 40)))
        a, b = bpg.Transcript(label), OTranscript(label)
        for _ in range(rnd.randrange(1, 30)):
            lab = bytes(rnd.randrange(1, 256) for _ in range(rnd.randrange(1, 12)))
            if rnd.random() < 0.6:
                msg = bytes(rnd.randrange(256) for _ in range(rnd.choice([0, 1, 8, 32, 165, 166, 167, 400])))
                a.append_message(lab, msg)
                b.append_message(lab, msg)
            else:
                n = rnd.choice([1, 32, 64, 166, 200, 333])
                assert a.challenge_bytes(lab, n) == b.challenge_bytes(lab, n)
        c = a.clone()  # a clone continues the same stream independently
        want = b.challenge_bytes(b"x", 64)
        assert c.challenge_bytes(b"x", 64) == want and a.challenge_bytes(b"x", 64) == want


# ---------------------------------------------------------------------------- device math headers on the host
@pytest.fixture(scope="module")
def hm():
    out = os.path.join(ROOT, "build", "libhost_math.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    src = os.path.join(ROOT, "tests", "host_math", "host_math.cpp")
    inc = os.path.join(ROOT, "bulletproof_gadgets_b200", "csrc")
    deps = [src] + [os.path.join(inc, f) for f in ("fe25519.cuh", "ge25519.cuh", "sc25519.cuh", "consts.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-I", inc, src, "-o", out], check=True)
    return ctypes.CDLL(out)


def _fe(x):
    return (ctypes.c_uint32 * 8)(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)])


def _int(arr):
    return sum(int(arr[i]) << (32 * i) for i in range(8))


def test_field_arithmetic_portable_branch(hm):
    rnd = random.Random(25519)
    edge = [0, 1, 2, 19, P - 1, P - 2, P, P + 1, (1 << 255) - 1, (1 << 256) - 1, (1 << 256) - 38, (1 << 255) + 5]
    vals = edge + [rnd.randrange(1 << 256) for _ in range(200)]
    out, canon = (ctypes.c_uint32 * 8)(), (ctypes.c_uint32 * 8)()
    for i, a in enumerate(vals):
        b = vals[(i * 7 + 3) % len(vals)]
        for fn, op in ((hm.hm_fe_mul, lambda x, y: x * y), (hm.hm_fe_add, lambda x, y: x + y),
                       (hm.hm_fe_sub, lambda x, y: x - y)):
            fn(_fe(a), _fe(b), out)
            hm.hm_fe_canon(out, canon)
            assert _int(canon) == op(a, b) % P, (fn, hex(a), hex(b))
        hm.hm_fe_canon(_fe(a), canon)
        assert _int(canon) == a % P
        if a % P:
            hm.hm_fe_invert(_fe(a), out)
            assert _int(out) * a % P == 1


def test_group_and_codec_portable_branch(hm):
    from tests.test_oracle_anchors import RFC9496_BAD, RFC9496_MULTIPLES
    o1, o2, o3 = (ctypes.create_string_buffer(32) for _ in range(3))
    for k, enc in enumerate(RFC9496_MULTIPLES):
        assert hm.hm_decompress_ops(bytes.fromhex(enc), o1, o2, o3) == 1
        assert o1.raw.hex() == enc
        assert o2.raw == (ed.BASEPOINT * (2 * k)).compress() == o3.raw
    for enc in RFC9496_BAD:
        assert hm.hm_decompress_ops(bytes.fromhex(enc), o1, o2, o3) == 0
    for i in range(32):
        u = hashlib.shake_256(b"hm-%d" % i).digest(64)
        hm.hm_from_uniform(u, o1)
        assert o1.raw == ed.from_uniform_bytes(u).compress()
    a, b = bytes.fromhex(RFC9496_MULTIPLES[5]), bytes.fromhex(RFC9496_MULTIPLES[9])
    assert hm.hm_add(a, b, o1) == 1 and o1.raw.hex() == RFC9496_MULTIPLES[14]


@pytest.fixture(scope="module", params=["host64", "portable"])
def hm_sc(request):
    """tests/host_math built with the host's 4 x 64-bit scalar routines, and with -DBPG_SC_PORTABLE (the limb code the
    device compiles)."""
    out = os.path.join(ROOT, "build", "libhost_math_sc_%s.so" % request.param)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    src = os.path.join(ROOT, "tests", "host_math", "host_math.cpp")
    inc = os.path.join(ROOT, "bulletproof_gadgets_b200", "csrc")
    deps = [src] + [os.path.join(inc, f) for f in ("fe25519.cuh", "ge25519.cuh", "sc25519.cuh", "consts.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        flags = ["-DBPG_SC_PORTABLE=1"] if request.param == "portable" else []
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-I", inc] + flags + [src, "-o", out], check=True)
    return ctypes.CDLL(out)


def test_scalar_field_host_routines(hm_sc):
    """mod-l arithmetic of sc25519.cuh as the host runs it (statement front end, challenge arithmetic) against python
    integers: canonical operands, Scalar::from_bits style operands up to 2^255, and the edges around l and 2^256."""
    L = 2**252 + 27742317777372353535851937790883648493
    R = 1 << 256
    rnd = random.Random(252)
    edge = [0, 1, 2, L - 2, L - 1, L, L + 1, 2 * L - 1, 2 * L, (1 << 252) - 1, 1 << 252, (1 << 255) - 1, (1 << 255), R - 1]
    vals = edge + [rnd.randrange(L) for _ in range(150)] + [rnd.randrange(1 << 255) for _ in range(100)] + [rnd.randrange(R) for _ in range(50)]
    out = (ctypes.c_uint32 * 8)()
    rinv = pow(R, -1, L)
    for i, a in enumerate(vals):
        b = vals[(i * 11 + 5) % len(vals)]
        # raw add / sub with carry / borrow: any 256-bit operands
        c = hm_sc.hm_sc_add_raw(_fe(a), _fe(b), out)
        assert (_int(out), c) == ((a + b) % R, (a + b) >> 256)
        c = hm_sc.hm_sc_sub_raw(_fe(a), _fe(b), out)
        assert (_int(out), c) == ((a - b) % R, 1 if a < b else 0)
        # modular add / sub: canonical operands
        ac, bc = a % L, b % L
        hm_sc.hm_sc_add(_fe(ac), _fe(bc), out)
        assert _int(out) == (ac + bc) % L
        hm_sc.hm_sc_sub(_fe(ac), _fe(bc), out)
        assert _int(out) == (ac - bc) % L
        # products: one operand below 2^255 and one canonical keep a*b < R*l (the contract of sc_montmul)
        a255 = a % (1 << 255)
        hm_sc.hm_sc_montmul(_fe(a255), _fe(bc), out)
        assert _int(out) == a255 * bc * rinv % L
        hm_sc.hm_sc_mul(_fe(a255), _fe(bc), out)
        assert _int(out) == a255 * bc % L
        hm_sc.hm_sc_reduce(_fe(a), out)
        assert _int(out) == a % L


def test_bench_statement_is_the_reference_gadget():
    """The numpy-built statement bench.py times (workloads.bounds_check_statement) equals, array for array, the library
    front end's flattening of the same workload written in the reference's text formats (`BOUND W<i> I0 I1` lines,
    /root/reference/src/bounds_check/bounds_check_gadget.rs:13-48) -- blindings aside, which the text path derives from
    its seed.  Host only."""
    import numpy as np
    import bulletproof_gadgets_b200 as bpg
    from bulletproof_gadgets_b200 import workloads as W
    for count, nbytes in ((1, 8), (5, 8), (40, 8), (3, 2)):
        st = W.bounds_check_statement(count, max_bytes=nbytes)
        gad, inst, wtns = W.bounds_check_text(count, max_bytes=nbytes)
        fs = bpg.flatten_prover("bench", inst, wtns, gad, b"\x01" * 32)
        assert (st.n, st.m, st.q, st.nnz) == (fs.n, fs.m, fs.q, fs.nnz)
        assert bytes(st.aL) == bytes(fs.aL) and bytes(st.aR) == bytes(fs.aR)
        assert np.array_equal(st.row_start, fs.row_start) and np.array_equal(st.term_var, fs.term_var)
        assert bytes(st.term_coef) == bytes(fs.term_coef)
        assert list(st.v) == list(fs.v)


def test_msm_reduction_index_model():
    """Integer model of msm.cu stages 4-6 (chunks, collapsed CTAs, slot lists, bit-marginal butterfly)."""
    from tests.model import msm_reduce_model
    assert msm_reduce_model.main(cases=60, seed=7) == 60
