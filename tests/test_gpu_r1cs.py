"""GPU parity: R1CS prove / verify through the C ABI vs the oracle (same seeded inputs).

Bar: proofs byte-identical to the oracle's; GPU proofs verify with the oracle verifier and oracle
proofs verify with the GPU verifier; tampered proofs / wrong statements reject on both."""
import pytest

from tests import circuits as C

pytestmark = pytest.mark.gpu

SMALL = [C.c_empty, C.c_mul3, lambda: C.c_range(8, 1), lambda: C.c_range(4, 3), C.c_unreduced, lambda: C.c_chain(5),
         lambda: C.c_chain(37)]


@pytest.fixture(scope="module")
def backends(ctx):
    return C.OracleBackend(), C.GpuBackend(ctx)


@pytest.mark.parametrize("mk", SMALL)
def test_proof_bytes_identical_and_cross_verify(backends, mk):
    ob, gb = backends
    circ = mk()
    p_o, coms_o = circ.prove(ob)
    p_g, coms_g = circ.prove(gb)
    assert coms_g == coms_o, "Pedersen commitments differ"
    assert p_g == p_o, "proof bytes differ from the oracle"
    n = len(p_g)
    assert (n - 1) % 32 == 0
    assert circ.verify(gb, p_o, coms_o) is True          # oracle-made proof, GPU verifier
    assert circ.verify(ob, p_g, coms_g) is True          # GPU-made proof, oracle verifier
    assert circ.verify(gb, p_g, coms_g, seed=b"\x55" * 32) is True


@pytest.mark.parametrize("mk", [C.c_mul3, lambda: C.c_range(8, 1), C.c_empty])
def test_rejections_match_oracle(backends, mk):
    ob, gb = backends
    circ = mk()
    proof, coms = circ.prove(gb)
    # wrong transcript label
    assert circ.verify(gb, proof, coms, label=b"other") is False
    assert circ.verify(ob, proof, coms, label=b"other") is False
    # tamper each 32-byte field once (flip a low bit): both must agree on reject / format error
    for off in range(1, len(proof), 32):
        bad = bytearray(proof)
        bad[off] ^= 1
        bad = bytes(bad)
        r_o, r_g = circ.verify(ob, bad, coms), circ.verify(gb, bad, coms)
        assert r_g == r_o and r_g is not True, (off, r_o, r_g)
    # wrong commitment
    if len(coms) > 1:
        assert circ.verify(gb, proof, [coms[1], coms[0]] + coms[2:]) is False
    # malformed encodings
    for bad in (b"", proof[:-1], proof[:33], b"\x02" + proof[1:], proof + b"\x00" * 32):
        r_o, r_g = circ.verify(ob, bad, coms), circ.verify(gb, bad, coms)
        assert r_g == r_o and r_g is not True, (len(bad), r_o, r_g)
    # non-canonical scalar in the proof (t_x += l) -> FormatError on both
    from oracle.pyref.merlin import L
    tx_off = 1 + 8 * 32
    tx = int.from_bytes(proof[tx_off:tx_off + 32], "little") + L
    if tx < (1 << 256):
        bad = proof[:tx_off] + tx.to_bytes(32, "little") + proof[tx_off + 32:]
        assert circ.verify(gb, bad, coms) == "format" and circ.verify(ob, bad, coms) == "format"


def test_false_statement_cannot_verify(backends):
    ob, gb = backends
    circ = C.c_mul3()
    circ.values[1] = (circ.values[1] + 1) % (1 << 252)   # y is wrong
    proof, coms = circ.prove(gb)
    assert circ.verify(gb, proof, coms) is False
    assert circ.verify(ob, proof, coms) is False
