"""GPU parity at the BASELINE.json sizes.

config 2 (BOUND 64-bit x1024, n = 2^17): the oracle needs ~35 s for this statement, so the full size is checked through
size-independent properties (prove -> verify accepts; any tampering rejects; the resident-circuit path and the
host-buffer path give the same bytes; a second context gives the same bytes), and the same workload at x128
(n = 2^14) is compared byte for byte with the C oracle.  config 3 (Merkle depth 32) and config 4 (LESS_THAN /
SET_MEMBER) statements are flattened by the library's own front end and cross-checked with the oracle verifier."""
import pytest

from oracle import coracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(ctx):
    import bulletproof_gadgets_b200 as bpg
    from bulletproof_gadgets_b200 import workloads as W
    return bpg, W, ctx


def test_config2_x128_bytes_identical_to_oracle(eng):
    bpg, W, ctx = eng
    st = W.bounds_check_statement(128)
    proof, coms = W.prove_statement(bpg, ctx, st, b"\x07" * 32)
    p_c, coms_c = coracle.prove_flat(st, b"\x07" * 32)
    assert coms == coms_c and proof == p_c
    assert W.verify_statement(bpg, ctx, st, proof, coms) is True
    assert coracle.verify_flat(st, coms, proof, b"\x09" * 32) is True


def test_config2_full_size_properties(eng):
    bpg, W, ctx = eng
    st = W.bounds_check_statement(1024)
    assert (st.n, st.m, st.q) == (1 << 17, 3072, 265216)
    proof, coms = W.prove_statement(bpg, ctx, st, b"\x07" * 32)            # host-buffer path (bulk loaders)
    assert len(proof) == 1 + 11 * 32 + (2 * 17 + 2) * 32
    assert W.verify_statement(bpg, ctx, st, proof, coms) is True
    # resident circuit on a second context: same bytes
    c2 = ctx.shared()
    circ = bpg.Circuit(c2, st.n, st.m, st.row_start, st.term_var, st.term_coef, st.q).set_witness(st.aL, st.aR)
    p = bpg.Prover(c2, bpg.Transcript(st.label))
    coms2 = p.commit_batch_packed(st.v_bytes, st.vbl_bytes)
    p.attach(circ)
    assert p.prove(b"\x07" * 32) == proof and coms2 == b"".join(coms)
    vf = bpg.Verifier(c2, bpg.Transcript(st.label))
    vf.commit_batch(coms2)
    vf.attach(circ)
    assert vf.verify(proof, b"\x55" * 32) is True
    # a different transcript-rng seed gives a different, equally valid proof (blinding factors change)
    p = bpg.Prover(c2, bpg.Transcript(st.label))
    p.commit_batch_packed(st.v_bytes, st.vbl_bytes)
    p.attach(circ)
    other = p.prove(b"\x08" * 32)
    assert other != proof and W.verify_statement(bpg, ctx, st, other, coms) is True
    # tampering anywhere rejects: one flipped bit per 32-byte field, one wrong commitment, one wrong constraint system
    for off in (1, 33, 65, 97, 1 + 8 * 32, 1 + 9 * 32, 1 + 11 * 32, len(proof) - 64, len(proof) - 32):
        bad = bytearray(proof)
        bad[off + 3] ^= 0x10
        try:
            ok = W.verify_statement(bpg, ctx, st, bytes(bad), coms)
        except bpg.BpgError as e:
            assert e.code == -1          # a scalar pushed out of range is a FormatError
            ok = False
        assert ok is False, off
    assert W.verify_statement(bpg, ctx, st, proof, [coms[1], coms[0]] + coms[2:]) is False
    smaller = W.bounds_check_statement(1023)                               # a different constraint system
    assert W.verify_statement(bpg, ctx, smaller, proof, coms[: smaller.m]) is False
    del p, vf, circ
    c2.close()


def test_config3_merkle_depth32(eng):
    bpg, W, ctx = eng
    gad, inst, wtns = W.merkle_text(32)
    st = bpg.flatten_prover("merkle32", inst, wtns, gad, b"\x03" * 32)
    assert (st.n, st.m, st.q) == (63180, 4, 126363)
    proof, text, ncons = bpg.prove(ctx, "merkle32", inst, wtns, gad, blinding_seed=b"\x03" * 32, rng_seed=b"\x07" * 32)
    assert ncons == st.q and len(proof) == 1441
    assert bpg.verify(ctx, "merkle32", inst, proof, text, gad) is True
    vs = bpg.flatten_verifier("merkle32", inst, text, gad)
    assert coracle.verify_flat(vs, vs.V, proof, b"\x09" * 32) is True        # GPU proof, oracle verifier
    # a wrong root must not verify
    lines = inst.split("\n")
    root = bytearray(bytes.fromhex(lines[0].split("0x")[1]))
    root[-1] ^= 1
    bad_inst = "\n".join(["I0 = 0x" + root.hex()] + lines[1:])
    assert bpg.verify(ctx, "merkle32", bad_inst, proof, text, gad) is False


def test_config4_statements(eng):
    bpg, W, ctx = eng
    for k, (gad, inst, wtns) in enumerate(W.batch_texts(6)):
        proof, text, _ = bpg.prove(ctx, "b%d" % k, inst, wtns, gad, blinding_seed=bytes([k + 1]) * 32, rng_seed=b"\x07" * 32)
        assert bpg.verify(ctx, "b%d" % k, inst, proof, text, gad) is True
        st = bpg.flatten_prover("b%d" % k, inst, wtns, gad, bytes([k + 1]) * 32)
        assert st.n == (379 if k % 2 == 0 else 32)
        p_c, coms_c = coracle.prove_flat(st, b"\x07" * 32)
        assert p_c == proof
