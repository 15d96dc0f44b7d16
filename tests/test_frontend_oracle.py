"""The oracle's statement front end (oracle/pyref/frontend.py) against the reference's own known answers and
fixtures: MiMC images (/root/reference/src/mimc_hash/mimc.rs:104-143), scalar byte order
(/root/reference/src/conversions.rs:114-150), the hash images / Merkle roots inside the fixture `.inst` files,
the circuit sizes of SURVEY.md appendix B, accept/reject outcomes of the 13 CI fixtures
(.github/workflows/integration_tests.yml:20-58) and of the negative unit cases (equality_gadget.rs:122,158,
less_than_gadget.rs:171-333, set_membership_gadget.rs:283-363, utils.rs:89, inequality_gadget.rs:335)."""
import hashlib
import json
import os

import pytest

from oracle import coracle
from oracle.pyref import frontend as F
from tests import frontend_glue as G

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "fixtures.json")))

BYTES_1 = bytes([0x7b, 0x24, 0x60, 0xbe, 0x18, 0x05, 0x44, 0xcd, 0x18, 0xe3, 0xe7, 0xe2, 0x73, 0x30, 0xce, 0xc9,
                 0x51, 0x7a, 0x31, 0x4a, 0xcb, 0xd4, 0xa0, 0x11, 0xd2, 0x73, 0xa5, 0x9b, 0x48, 0x0c, 0x1e, 0x00])
BYTES_2 = bytes([0x7b, 0x98, 0x7c, 0xf9, 0x7a, 0x9f, 0x1b, 0xd5, 0x49, 0x23, 0x47, 0xd6, 0xf4, 0xe5, 0x50, 0xae,
                 0x29, 0x49, 0xa5, 0x13, 0xde, 0x92, 0xfe, 0x50, 0x65, 0x35, 0x0e, 0xbc, 0xd5, 0x1d, 0xb6, 0x04])


def test_conversions_kats():
    """conversions.rs:114-150 -- `as_bytes()` of the results."""
    s = F.le_to_scalars(BYTES_1 + BYTES_2)
    assert [F.sc_bytes(x) for x in s] == [BYTES_1, BYTES_2]
    assert F.sc_bytes(F.le_to_scalar(BYTES_1)) == BYTES_1
    assert F.sc_bytes(F.be_to_scalar(BYTES_1)) == BYTES_1[::-1]
    s = F.be_to_scalars(BYTES_1 + BYTES_2)
    assert [F.sc_bytes(x) for x in s] == [BYTES_2[::-1], BYTES_1[::-1]]
    # Scalar::from_bits: bit 255 cleared, no reduction
    assert F.be_to_scalar(b"\xff" * 32) == (1 << 255) - 1
    assert F.scalar_to_be(F.be_to_scalar(b"\x01\x02")) == bytes(30) + b"\x01\x02"


def test_mimc_hash_kats():
    """mimc.rs:104-143."""
    pre1 = bytes([0x38, 0x53, 0x54, 0x50, 0x43, 0x30, 0x43, 0x54, 0x6f, 0x31, 0x38, 0x77, 0x61, 0x5a, 0x6a, 0x42, 0x36, 0x63])
    assert F.scalar_to_be(F.mimc_hash(pre1)).hex() == "0d2203069ac15f58172bae1b3af98d8982deef9df37482c1a920b8832ee813a4"
    pre2 = b"The quick brown fox jumps over t"
    assert len(pre2) == 32
    assert F.scalar_to_be(F.mimc_hash(pre2)).hex() == "01245409f28ae2f076077d4a40bd91551b3a03b1ad8adb2b1da116d29c60a85c"


def _vars(text):
    return dict(F.parse_var_line(line[0], line) for line in text.splitlines())


def test_fixture_hash_images_and_merkle_roots():
    """example.wtns:3 = mimc_hash(W1); example.inst:3 (I2) = H(H(W1), H(I3)); I7 = H(H(I6), H(W4));
    I5 = H(I2, I7) -- the reference's implicit known answers."""
    inst, wtns, _ = G.load("example")
    I, W = _vars(inst), _vars(wtns)
    h = F.mimc_hash
    assert F.scalar_to_be(h(W["W1"])) == W["W2"]
    n1 = F.mimc_sponge([h(W["W1"]), h(I["I3"])])
    n2 = F.mimc_sponge([h(I["I6"]), h(W["W4"])])
    assert F.scalar_to_be(n1) == I["I2"] and F.scalar_to_be(n2) == I["I7"]
    assert F.scalar_to_be(F.mimc_sponge([n1, n2])) == I["I5"]


def test_sub_matches_dalek_wrap():
    """Scalar52::sub adds l once; for unreduced operands far below zero dalek's result picks up 2^260."""
    big = (1 << 255) - 1
    assert F.sc_sub(5, 3) == 2 and F.sc_sub(3, 5) == F.L - 2
    assert F.sc_sub(0, big) == (-big + F.L + (1 << 260)) % F.L
    assert F.sc_sub(big, 1) == (big - 1) % F.L


@pytest.mark.parametrize("stem", G.STEMS)
def test_fixture_sizes_match_survey_table(stem):
    """Flat circuit sizes (n, m, q, nnz) -- SURVEY.md appendix B, computed there by an independent simulation."""
    inst, wtns, gad = G.load(stem)
    st = F.compile_prover(stem, inst, wtns, gad, G.blinding())
    g = GOLD[stem]
    assert (st.n, st.m, st.q, st.nnz) == (g["n"], g["m"], g["q"], g["nnz"])
    n_pad = 1
    while n_pad < st.n:
        n_pad *= 2
    lg = n_pad.bit_length() - 1
    assert g["proof_len"] == 1 + 11 * 32 + (2 * lg + 2) * 32


SURVEY_TABLE = {"example": (14988, 33, 30007, 89638), "bounds_check": (1440, 9, 2889, 7215), "equality": (0, 12, 9, 18),
                "inequality": (24, 36, 63, 162), "less_than": (1137, 12, 2289, 5706), "merkle_tree": (27216, 31, 54454, 163446),
                "mimc_hash": (18468, 32, 36946, 110904), "set_membership": (15600, 72, 31256, 93728),
                "or": (9452, 27, 19180, 56821), "or2": (17236, 38, 34758, 103568), "or3": (3, 3, 9, 21),
                "or4": (22561, 46, 49594, 137660), "or5": (4721, 29, 10531, 27325)}


def test_golden_sizes_equal_survey():
    for stem, tup in SURVEY_TABLE.items():
        g = GOLD[stem]
        assert (g["n"], g["m"], g["q"], g["nnz"]) == tup


def _roundtrip(stem, inst, wtns, gad, label=None):
    st = F.compile_prover(label or stem, inst, wtns, gad, G.blinding())
    proof, coms = coracle.prove_flat(st, G.SEED_PROVE)
    text = st.coms_text(coms)
    vs = F.compile_verifier(label or stem, inst, text, gad)
    assert (vs.n, vs.q, vs.nnz) == (st.n, st.q, st.nnz)
    return proof, text, coracle.verify_flat(vs, vs.V, proof, G.SEED_VERIFY)


@pytest.mark.parametrize("stem", ["equality", "or3", "inequality", "bounds_check", "less_than", "or5", "or", "example"])
def test_fixtures_prove_and_verify_with_the_oracle(stem):
    inst, wtns, gad = G.load(stem)
    proof, text, ok = _roundtrip(stem, inst, wtns, gad)
    assert ok is True
    g = GOLD[stem]
    assert hashlib.sha256(proof).hexdigest() == g["proof_sha256"]
    assert hashlib.sha256(text.encode()).hexdigest() == g["coms_sha256"]
    # prover and verifier must be invoked with the same name (it labels the transcript)
    vs = F.compile_verifier("other-name", inst, text, gad)
    assert coracle.verify_flat(vs, vs.V, proof, G.SEED_VERIFY) is False


NEGATIVE = [
    # (gadgets, inst, wtns) -- statements that are false: a proof can be produced but must not verify
    ("EQUALS W0 I0", "I0 = 0x05", "W0 = 0x06"),                                    # equality_gadget.rs:122
    ("EQUALS W0 W1", "", "W0 = 0x%s\nW1 = 0x07" % ("11" * 40)),                   # limb count mismatch (equality_gadget.rs:158)
    ("LESS_THAN W0 W1", "", "W0 = 0x09\nW1 = 0x08"),                              # left > right (less_than_gadget.rs:171)
    ("LESS_THAN W0 W1", "", "W0 = 0x08\nW1 = 0x08"),                              # left = right (less_than_gadget.rs:255)
    ("SET_MEMBER W0 I0 I1", "I0 = 0x01\nI1 = 0x02", "W0 = 0x03"),                # not a member (set_membership_gadget.rs:283)
    ("BOUND W0 I0 I1", "I0 = 0x10\nI1 = 0x20", "W0 = 0x21"),                     # above max (utils.rs:89: out of range)
    ("BOUND W0 I0 I1", "I0 = 0x10\nI1 = 0x20", "W0 = 0x0f"),                     # below min
    ("UNEQUAL W0 I0", "I0 = 0x2a", "W0 = 0x2a"),                                  # equal values (inequality_gadget.rs:335)
    ("HASH I0 W0", "I0 = 0x0d2203069ac15f58172bae1b3af98d8982deef9df37482c1a920b8832ee813a5", "W0 = 0x385354504330435%s" % "46f313877615a6a423663"),
]
POSITIVE = [
    ("EQUALS W0 I0", "I0 = 0x05", "W0 = 0x05"),
    ("LESS_THAN W0 W1", "", "W0 = 0x07\nW1 = 0x08"),
    ("SET_MEMBER W0 I0 W1 I1", "I0 = 0x01\nI1 = 0x02", "W0 = 0x03\nW1 = 0x03"),
    ("BOUND W0 I0 I1", "I0 = 0x10\nI1 = 0x20", "W0 = 0x20"),
    ("UNEQUAL W0 I0", "I0 = 0x2a", "W0 = 0x2b"),
    ("UNEQUAL W0 I0", "I0 = 0x2a", "W0 = 0x%s" % ("ff" * 32)),                  # from_bits value >= l (inequality_gadget.rs:264-302)
    ("UNEQUAL W0 I0", "I0 = 0x%s" % ("2a" * 33), "W0 = 0x2b"),                  # limb mismatch -> constrain(0): always true (inequality_gadget.rs:53-55)
    ("HASH I0 W0", "I0 = 0x0d2203069ac15f58172bae1b3af98d8982deef9df37482c1a920b8832ee813a4", "W0 = 0x385354504330435%s" % "46f313877615a6a423663"),
    ("OR\n[\n{\nEQUALS W0 I0\n}\n{\nEQUALS W0 I1\n}\n]", "I0 = 0x01\nI1 = 0x02", "W0 = 0x02"),
]


@pytest.mark.parametrize("case", range(len(NEGATIVE)))
def test_false_statements_do_not_verify(case):
    gad, inst, wtns = NEGATIVE[case]
    _, _, ok = _roundtrip("neg-%d" % case, inst, wtns, gad)
    assert ok is False


@pytest.mark.parametrize("case", range(len(POSITIVE)))
def test_true_statements_verify(case):
    gad, inst, wtns = POSITIVE[case]
    _, _, ok = _roundtrip("pos-%d" % case, inst, wtns, gad)
    assert ok is True


def test_or_of_two_false_clauses_does_not_verify():
    gad = "OR\n[\n{\nEQUALS W0 I0\n}\n{\nEQUALS W0 I1\n}\n]"
    _, _, ok = _roundtrip("or-neg", "I0 = 0x01\nI1 = 0x02", "W0 = 0x03", gad)
    assert ok is False


def test_front_end_panics_like_the_reference():
    with pytest.raises(F.FrontendPanic):
        F.compile_prover("x", "", "W0 = 0x01", "FROBNICATE W0", G.blinding())          # unknown gadget
    with pytest.raises(F.FrontendPanic):
        F.compile_prover("x", "", "W0 = 0x01", "EQUALS W0 I9", G.blinding())           # missing instance var
    with pytest.raises(F.FrontendPanic):
        F.compile_prover("x", "I0 = 0x00\nI1 = 0xff", "W0 = 0x%s" % ("01" * 33), "BOUND W0 I0 I1", G.blinding())  # > 32 bytes
    with pytest.raises(F.FrontendPanic):
        F.compile_prover("x", "", "W0 = 0x1", "EQUALS W0 W0", G.blinding())            # odd hex
    with pytest.raises(F.FrontendPanic):
        F.compile_verifier("x", "I0 = 0x00\nI1 = 0xff", "", "BOUND W0 I0 I1")              # missing commitment C0-0
    # EQUALS against a witness without commitments is NOT a panic: limb-count mismatch -> constrain(1)
    vs = F.compile_verifier("x", "I0 = 0x01", "", "EQUALS W0 I0")
    assert (vs.n, vs.q) == (0, 1)
