"""GPU parity on the reference's 13 fixture statements (/root/reference/example.*,
/root/reference/tests/resources/*; copies under tests/golden/fixtures): the oracle front end flattens each
statement, the CUDA library proves and verifies it through the C ABI bulk loaders, and the bytes must equal the
committed oracle goldens (tests/golden/fixtures.json: SHA-256 of the proof and of the .coms text)."""
import hashlib
import json
import os

import pytest

from oracle import coracle
from oracle.pyref import frontend as F
from tests import frontend_glue as G

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "fixtures.json")))


@pytest.fixture(scope="module")
def engine(ctx):
    import bulletproof_gadgets_b200 as bpg
    from bulletproof_gadgets_b200 import workloads as W
    return bpg, W, ctx


@pytest.mark.parametrize("stem", G.STEMS)
def test_fixture_bytes_identical_to_golden(engine, stem):
    bpg, W, ctx = engine
    inst, wtns, gad = G.load(stem)
    st = F.compile_prover(stem, inst, wtns, gad, G.blinding())
    proof, coms = W.prove_statement(bpg, ctx, st, G.SEED_PROVE)
    text = st.coms_text(coms)
    g = GOLD[stem]
    assert len(proof) == g["proof_len"]
    assert hashlib.sha256(text.encode()).hexdigest() == g["coms_sha256"], "commitments differ from the oracle's"
    assert proof[:65].hex() == g["proof_head"]
    assert hashlib.sha256(proof).hexdigest() == g["proof_sha256"], "proof bytes differ from the oracle's"
    vs = F.compile_verifier(stem, inst, text, gad)
    assert W.verify_statement(bpg, ctx, vs, proof, vs.V, G.SEED_VERIFY) is True
    bad = bytearray(proof)
    bad[-40] ^= 4
    assert W.verify_statement(bpg, ctx, vs, bytes(bad), vs.V, G.SEED_VERIFY) is False
    other = F.compile_verifier("another-name", inst, text, gad)          # transcript label = file stem
    assert W.verify_statement(bpg, ctx, other, proof, other.V, G.SEED_VERIFY) is False


@pytest.mark.parametrize("stem", ["equality", "or3", "inequality", "bounds_check", "less_than"])
def test_fixture_cross_verification(engine, stem):
    """GPU-made proofs verify with the oracle verifier; oracle-made proofs verify with the GPU verifier."""
    bpg, W, ctx = engine
    inst, wtns, gad = G.load(stem)
    st = F.compile_prover(stem, inst, wtns, gad, G.blinding(b"cross"))
    p_gpu, coms_gpu = W.prove_statement(bpg, ctx, st, b"\x21" * 32)
    p_cpu, coms_cpu = coracle.prove_flat(st, b"\x21" * 32)
    assert coms_gpu == coms_cpu and p_gpu == p_cpu
    vs = F.compile_verifier(stem, inst, st.coms_text(coms_gpu), gad)
    assert coracle.verify_flat(vs, vs.V, p_gpu, b"\x31" * 32) is True
    assert W.verify_statement(bpg, ctx, vs, p_cpu, vs.V, b"\x31" * 32) is True


NEG = [("EQUALS W0 I0", "I0 = 0x05", "W0 = 0x06"), ("LESS_THAN W0 W1", "", "W0 = 0x09\nW1 = 0x08"),
       ("SET_MEMBER W0 I0 I1", "I0 = 0x01\nI1 = 0x02", "W0 = 0x03"), ("BOUND W0 I0 I1", "I0 = 0x10\nI1 = 0x20", "W0 = 0x21"),
       ("UNEQUAL W0 I0", "I0 = 0x2a", "W0 = 0x2a")]


@pytest.mark.parametrize("case", range(len(NEG)))
def test_false_statements_rejected_by_gpu_and_oracle(engine, case):
    bpg, W, ctx = engine
    gad, inst, wtns = NEG[case]
    st = F.compile_prover("neg", inst, wtns, gad, G.blinding())
    proof, coms = W.prove_statement(bpg, ctx, st, G.SEED_PROVE)
    assert (proof, coms) == coracle.prove_flat(st, G.SEED_PROVE)        # same bytes even for an unsatisfied circuit
    vs = F.compile_verifier("neg", inst, st.coms_text(coms), gad)
    assert W.verify_statement(bpg, ctx, vs, proof, vs.V, G.SEED_VERIFY) is False
    assert coracle.verify_flat(vs, vs.V, proof, G.SEED_VERIFY) is False
