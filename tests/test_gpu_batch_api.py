"""GPU: the reference's own C ABI (c_prove / c_verify / free_proof, include/bulletproofs_gadgets.h) and the batch entry
points (bpg_r1cs_prove_batch / bpg_r1cs_verify_batch / bpg_prove_batch / bpg_verify_batch) against the oracle."""
import hashlib
import json
import os

import pytest

from oracle import coracle
from tests import frontend_glue as G

pytestmark = pytest.mark.gpu
FIX = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "fixtures.json")))


@pytest.fixture(scope="module")
def eng(ctx):
    import bulletproof_gadgets_b200 as bpg
    from bulletproof_gadgets_b200 import workloads as W
    return bpg, W, ctx


@pytest.mark.parametrize("stem", ["example", "bounds_check", "merkle_tree", "or3"])
def test_reference_c_abi_round_trip(eng, stem):
    """fixture -> c_prove -> c_verify -> free_proof (OS randomness, so only accept / reject and the oracle verifier are
    checked); tampered proof, wrong instance and wrong commitments must give false, not abort."""
    bpg, W, ctx = eng
    inst, wtns, gad = G.load(stem)
    out = bpg.c_prove(stem, inst, wtns, gad)
    assert out is not None
    proof, coms = out
    assert len(proof) == FIX[stem]["proof_len"]
    assert bpg.c_verify(stem, inst, gad, coms, proof) is True
    vs = bpg.flatten_verifier(stem, inst, coms, gad)
    assert coracle.verify_flat(vs, vs.V, proof, b"\x09" * 32) is True      # GPU proof, oracle verifier
    bad = bytearray(proof)
    bad[40] ^= 4
    assert bpg.c_verify(stem, inst, gad, coms, bytes(bad)) is False
    assert bpg.c_verify(stem + "x", inst, gad, coms, proof) is False        # other transcript label
    assert bpg.c_verify(stem, inst, gad, coms, proof[:-7]) is False         # malformed: false, not a panic
    assert bpg.c_prove(stem, inst, "W0 = zz", gad) is None                  # unparsable witness: NULL + message
    assert bpg.lib().bpg_last_error()


def test_reference_c_abi_is_thread_safe(eng):
    import threading
    bpg, W, ctx = eng
    texts = W.batch_texts(24)
    res = [None] * len(texts)

    def work(k):
        gad, inst, wtns = texts[k]
        out = bpg.c_prove("t%d" % k, inst, wtns, gad)
        res[k] = out is not None and bpg.c_verify("t%d" % k, inst, gad, out[1], out[0])

    ts = [threading.Thread(target=work, args=(k,)) for k in range(len(texts))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert all(res)


def test_flat_batch_matches_oracle(eng):
    """16 independent statements of four shapes through bpg_r1cs_prove_batch on 5 contexts: every proof byte-identical to
    the oracle's; bpg_r1cs_verify_batch accepts them and rejects tampered ones; BPG_JOB_VERIFY fuses both."""
    bpg, W, ctx = eng
    ctxs = [ctx] + [ctx.shared() for _ in range(4)]
    sts = [W.bounds_check_statement(count=1 + (k % 4), max_bytes=1 + (k % 3), seed=50 + k, label=b"batch-%d" % k) for k in range(16)]
    seeds = [bytes([k + 1]) * 32 for k in range(16)]
    out = bpg.prove_batch(ctxs, sts, seeds)
    for k, (status, proof, coms) in enumerate(out):
        assert status == 0
        p_c, coms_c = coracle.prove_flat(sts[k], seeds[k])
        assert proof == p_c and coms == b"".join(coms_c), k
    assert bpg.verify_batch(ctxs, sts, [o[2] for o in out], [o[1] for o in out], seeds) == [0] * 16
    proofs = [bytearray(o[1]) for o in out]
    proofs[3][70] ^= 1
    proofs[9] = proofs[9][:-1]
    st = bpg.verify_batch(ctxs, sts, [o[2] for o in out], [bytes(p) for p in proofs])
    assert st[3] == -2 and st[9] == -1 and all(s == 0 for k, s in enumerate(st) if k not in (3, 9))
    fused = bpg.prove_batch(ctxs, sts, seeds, verify=True, verify_seeds=seeds)
    assert [f[0] for f in fused] == [0] * 16 and [f[1] for f in fused] == [o[1] for o in out]
    # resident circuit shared by all jobs (same statement, different seeds)
    s0 = sts[5]
    circ = bpg.Circuit(ctx, s0.n, s0.m, s0.row_start, s0.term_var, s0.term_coef, s0.q).set_witness(s0.aL, s0.aR)
    res = bpg.prove_batch(ctxs, [s0] * 6, seeds[:6], circuits=circ, verify=True)
    assert [r[0] for r in res] == [0] * 6 and res[5][1] != res[4][1]
    assert res[0][1] == coracle.prove_flat(s0, seeds[0])[0]
    del circ
    for c in ctxs[1:]:
        c.close()


def test_text_batch_matches_golden(eng):
    """The 13 reference fixtures as ONE bpg_prove_batch call: proof hashes equal tests/golden/fixtures.json."""
    bpg, W, ctx = eng
    ctxs = [ctx] + [ctx.shared() for _ in range(3)]
    texts = []
    for stem in G.STEMS:
        inst, wtns, gad = G.load(stem)
        texts.append((stem, inst, wtns, gad))
    out = bpg.prove_text_batch(ctxs, texts, [G.SEED_BLIND] * len(texts), [G.SEED_PROVE] * len(texts), verify=True)
    for stem, (status, proof, coms, accepted) in zip(G.STEMS, out):
        assert status == 0 and accepted, stem
        assert hashlib.sha256(proof).hexdigest() == FIX[stem]["proof_sha256"], stem
        assert hashlib.sha256(coms.encode()).hexdigest() == FIX[stem]["coms_sha256"], stem
    ver = bpg.verify_text_batch(ctxs, [(t[0], t[1], t[3], o[2], o[1]) for t, o in zip(texts, out)])
    assert ver == [(0, True)] * len(texts)
    bad = [(t[0], t[1], t[3], o[2], o[1][:40] + bytes([o[1][40] ^ 1]) + o[1][41:]) for t, o in zip(texts, out)]
    assert [a for _, a in bpg.verify_text_batch(ctxs, bad)] == [False] * len(texts)
    for c in ctxs[1:]:
        c.close()


def _bit_runs(st):
    """(runs, host_index): maximal runs of multipliers assigned (1 - b, b), as the reference's range proof allocates them."""
    runs, host = [], []
    cur = None
    for i in range(st.n):
        l = int.from_bytes(bytes(st.aL[32 * i: 32 * i + 32]), "little")
        r = int.from_bytes(bytes(st.aR[32 * i: 32 * i + 32]), "little")
        if l in (0, 1) and r in (0, 1) and l + r == 1:
            if cur is None or cur[0] + cur[1] != i or cur[1] == 256:
                cur = [i, 0, 0]
                runs.append(cur)
            cur[2] |= r << cur[1]
            cur[1] += 1
        else:
            host.append(i)
            cur = None
    return [tuple(x) for x in runs], host


def test_device_witness_generation_f3(eng):
    """range-proof bits built in HBM from the values (bpg_prover_load_cs_bits) give the same proof bytes as
    uploading every multiplier, and as the oracle; partial coverage (LESS_THAN: 378 bit multipliers + 1 product) and a
    statement without any bit run work; runs that do not partition the multipliers are rejected."""
    bpg, W, ctx = eng
    cases = [W.bounds_check_statement(5, max_bytes=8, label=b"f3-a"), W.bounds_check_statement(3, max_bytes=1, label=b"f3-b")]
    gad, inst, wtns = W.batch_texts(2)[0]
    cases.append(bpg.flatten_prover("f3-lt", inst, wtns, gad, b"\x04" * 32))
    for st in cases:
        runs, host = _bit_runs(st)
        assert sum(r[1] for r in runs) + len(host) == st.n and runs
        aLh = b"".join(bytes(st.aL[32 * i: 32 * i + 32]) for i in host)
        aRh = b"".join(bytes(st.aR[32 * i: 32 * i + 32]) for i in host)
        p = bpg.Prover(ctx, bpg.Transcript(st.label))
        coms = p.commit_batch_packed(st.v_bytes, st.vbl_bytes)
        p.load_cs_bits(st.n, runs, aLh, aRh, host, st.row_start, st.term_var, st.term_coef, st.q)
        proof = p.prove(b"\x07" * 32)
        want, coms_c = coracle.prove_flat(st, b"\x07" * 32)
        assert proof == want and coms == b"".join(coms_c)
        assert W.prove_statement(bpg, ctx, st, b"\x07" * 32)[0] == proof
    st = cases[0]
    runs, host = _bit_runs(st)
    for bad in (runs[:-1], runs + [runs[0]], [(runs[0][0], runs[0][1] + 1, runs[0][2])] + runs[1:]):
        p = bpg.Prover(ctx, bpg.Transcript(st.label))
        p.commit_batch_packed(st.v_bytes, st.vbl_bytes)
        with pytest.raises(bpg.BpgError):
            p.load_cs_bits(st.n, bad, b"", b"", [], st.row_start, st.term_var, st.term_coef, st.q)
