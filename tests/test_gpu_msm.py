"""GPU parity: generators, Pedersen commitments, fixed-base and arbitrary-point MSMs through the C ABI
against the oracle (bit-exact 32-byte ristretto encodings), plus size-independent properties at full size."""
import hashlib
import json
import os
import random

import numpy as np
import pytest

from oracle import coracle
from oracle.pyref import ed, r1cs as O
from oracle.pyref.merlin import L

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "r1cs_small.json")))


def sc(tag, i):
    return int.from_bytes(hashlib.shake_256(b"%s-%d" % (tag, i)).digest(64), "little") % L


def test_generators_match_golden_and_oracle(ctx):
    ctx.gens_ensure(4096)
    g = GOLDEN["generators"]
    assert [x.hex() for x in ctx.gens_compressed("G", 0, 16)] == g["G"]
    assert [x.hex() for x in ctx.gens_compressed("H", 0, 16)] == g["H"]
    assert ctx.gens_compressed("B")[0].hex() == g["B"]
    assert ctx.gens_compressed("B_blinding")[0].hex() == g["B_blinding"]
    # deep into the chain (SHAKE256 stream positions 64*i): against the C restatement
    assert ctx.gens_compressed("G", 4000, 96) == coracle.gens("G", 4000, 96)
    assert ctx.gens_compressed("H", 4000, 96) == coracle.gens("H", 4000, 96)


def test_field_selftest_on_device(ctx):
    """fe_sqr (36 IMAD.WIDE schedule, tools/gen_fe_sqr.py) == fe_mul(a, a) on 4 M biased values, on the device."""
    import ctypes
    import bulletproof_gadgets_b200 as bpg
    bad = ctypes.c_uint64(99)
    for seed in (1, 2):
        assert bpg.lib().bpg_selftest_field(ctx._h, 1 << 21, seed, ctypes.byref(bad)) == 0
        assert bad.value == 0


EDGE_CASES = [
    ([], [], None, None), ([1], [], None, None), ([], [], 1, None), ([], [], None, 5), ([0, 0], [0], 0, 0),
    ([L - 1], [2], 3, L - 4), ([1] * 64, [1] * 64, None, None), ([2**255 - 1], [2**252], None, None),
    ([32768, 32769, 65535, 65536, 65537, 2**16 * 32768, 2**240], [], None, None),
    ([i & 1 for i in range(64)], [1 - (i & 1) for i in range(64)], None, 12345),
]


@pytest.mark.parametrize("case", range(len(EDGE_CASES)))
def test_msm_gens_edge_cases_vs_bigint(ctx, case):
    sG, sH, sB, sBb = EDGE_CASES[case]
    bg, pg = O.BulletproofGens(64), O.PedersenGens()
    acc = ed.msm([s % L for s in sG], bg.G[:len(sG)]) + ed.msm([s % L for s in sH], bg.H[:len(sH)])
    if sB is not None:
        acc = acc + pg.B * sB
    if sBb is not None:
        acc = acc + pg.B_blinding * sBb
    assert ctx.msm_gens(sG, sH, sB, sBb) == acc.compress()


@pytest.mark.parametrize("n", [1, 3, 190, 500, 800, 4096, 16384])
def test_msm_gens_matches_c_oracle(ctx, n):
    """uniform scalars (verifier / S-commitment shape) and 0/1 scalars (a_L / a_R shape)."""
    rnd = random.Random(n)
    sG = b"".join(rnd.randrange(L).to_bytes(32, "little") for _ in range(n))
    sH = b"".join(rnd.randrange(L).to_bytes(32, "little") for _ in range(n - n // 3))
    sB, sBb = rnd.randrange(L).to_bytes(32, "little"), rnd.randrange(L).to_bytes(32, "little")
    assert ctx.msm_gens_bytes(sG, sH, sB, sBb) == coracle.msm_gens(sG, sH, sB, sBb)
    bits = b"".join(bytes([rnd.randrange(2)]) + bytes(31) for _ in range(n))
    assert ctx.msm_gens_bytes(bits, bits, None, sBb) == coracle.msm_gens(bits, bits, None, sBb)


@pytest.mark.parametrize("n", [31, 32, 33, 64, 511, 1024])
def test_small_msm_kernel_structured_scalars_vs_c_oracle(ctx, n):
    """k_msm_small (one launch, <= 2050 points): slices of 32 scalars, buckets with more than 64 entries of one slice shared by
    the whole CTA -- scalars whose 32 byte digits all fall into ONE bucket, -1 (a_R of a range proof), and runs of equal values."""
    rnd = random.Random(1000 + n)
    one_bucket = bytes([1] * 31 + [0])                       # digit 1 in windows 0..30
    minus_one = (L - 1).to_bytes(32, "little")
    top = (L - 1 - 2**200).to_bytes(32, "little")
    pick = [one_bucket, minus_one, top, (2**252).to_bytes(32, "little"), bytes(32)]
    sG = b"".join(one_bucket for _ in range(n))
    sH = b"".join(minus_one if i % 3 else one_bucket for i in range(n))
    sBb = rnd.randrange(L).to_bytes(32, "little")
    assert ctx.msm_gens_bytes(sG, sH, None, sBb) == coracle.msm_gens(sG, sH, None, sBb)
    sG = b"".join(pick[rnd.randrange(len(pick))] for _ in range(n))
    sH = b"".join(pick[(i // 7) % len(pick)] for i in range(n - 1))
    assert ctx.msm_gens_bytes(sG, sH, sBb, None) == coracle.msm_gens(sG, sH, sBb, None)


def test_msm_arbitrary_points_vs_oracle(ctx):
    rnd = random.Random(77)
    pts = [ed.from_uniform_bytes(hashlib.shake_256(b"dyn-%d" % i).digest(64)).compress() for i in range(40)]
    for n in (0, 1, 2, 17, 40):
        s = [rnd.randrange(L) for _ in range(n)]
        want = coracle.msm(s, pts[:n]) if n else bytes(32)
        assert ctx.msm(s, pts[:n]) == want
    import bulletproof_gadgets_b200 as bpg
    with pytest.raises(bpg.BpgError) as e:      # undecodable point -> VerificationError, like dalek's optional_multiscalar_mul
        ctx.msm([1, 2], [pts[0], bytes.fromhex("00" + "ff" * 31)])
    assert e.value.code == -2


@pytest.mark.parametrize("n", [512, 777, 4096, 20000])
def test_msm_variable_base_pippenger_vs_oracle(ctx, n):
    """bpg_msm above 512 points takes the variable-base Pippenger path (window sums over per-window bucket sets, combined
    with doublings): same bytes as the C oracle's vartime Pippenger and as the thread-per-point path; zero scalars, the
    scalar l - 1, repeated points and the identity among the inputs."""
    import bulletproof_gadgets_b200 as bpg
    rnd = random.Random(n)
    base = [ed.from_uniform_bytes(hashlib.shake_256(b"vb-%d" % i).digest(64)).compress() for i in range(64)]
    pts = [base[rnd.randrange(64)] for _ in range(n)]
    pts[3] = bytes(32)                                   # the identity encodes as 32 zero bytes
    s = [rnd.randrange(L) for _ in range(n)]
    s[0], s[1], s[2], s[5] = 0, L - 1, 1, 2**252
    want = coracle.msm(s, pts)
    assert ctx.msm(s, pts) == want
    bits = [rnd.randrange(2) for _ in range(n)]          # 0/1 scalars: one heavy bucket in window 0
    assert ctx.msm(bits, pts) == coracle.msm(bits, pts)
    with pytest.raises(bpg.BpgError) as e:
        ctx.msm(s, pts[:7] + [bytes.fromhex("01" + "00" * 31)] + pts[8:])
    assert e.value.code == -2


def test_pedersen_batch_vs_oracle(ctx):
    pg = O.PedersenGens()
    vs = [0, 1, L - 1, 2**255 - 1, 2**64 - 1] + [sc(b"pv", i) for i in range(60)]
    rs = [0, 0, 1, 7, L - 2] + [sc(b"pr", i) for i in range(60)]
    got = ctx.pedersen_commit_batch(vs, rs)
    for i in (0, 1, 2, 3, 4, 5, 33, 64):
        assert got[i] == pg.commit(vs[i], rs[i]).compress(), i
    want = [coracle.msm_gens(b"", b"", (v % L).to_bytes(32, "little"), (r % L).to_bytes(32, "little")) for v, r in zip(vs, rs)]
    assert got == want


@pytest.mark.parametrize("lg", [17, 18])
def test_full_size_linearity(ctx, lg):
    """At BASELINE sizes the oracle is too slow: check msm(s) + msm(t) == msm(s + t) instead (2^17 / 2^18 points)."""
    n = 1 << (lg - 1)
    rng = np.random.default_rng(lg)

    def rand_scalars():
        a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        a[:, 31] &= 0x0F
        return a

    def addmod(a, b):
        out = bytearray()
        ab, bb = a.tobytes(), b.tobytes()
        for i in range(n):
            out += ((int.from_bytes(ab[32 * i: 32 * i + 32], "little") + int.from_bytes(bb[32 * i: 32 * i + 32], "little")) % L).to_bytes(32, "little")
        return bytes(out)

    s1g, s1h, s2g, s2h = (rand_scalars() for _ in range(4))
    r1 = ctx.msm_gens_bytes(s1g.tobytes(), s1h.tobytes())
    r2 = ctx.msm_gens_bytes(s2g.tobytes(), s2h.tobytes())
    r3 = ctx.msm_gens_bytes(addmod(s1g, s2g), addmod(s1h, s2h))
    assert (ed.decompress(r1) + ed.decompress(r2)).compress() == r3
    assert ctx.msm_gens_bytes(s1g.tobytes(), s1h.tobytes()) == r1       # idempotent (work buffers reused)


def test_shared_contexts_agree(ctx):
    """A second context on the same GPU (own stream, shared generator tables) computes the same bytes."""
    c2 = ctx.shared()
    s = [sc(b"sh", i) for i in range(300)]
    assert c2.msm_gens(s, s[:100], 5, 6) == ctx.msm_gens(s, s[:100], 5, 6)
    c2.gens_ensure(8192)                       # growth through the child is visible to the parent
    assert ctx.msm_gens(s, s, None, None) == c2.msm_gens(s, s, None, None)
    c2.close()


def test_point_range_sharded_msm(ctx):
    """One MSM split by contiguous point ranges over several contexts; partial points added on the host."""
    from bulletproof_gadgets_b200 import sharding
    rnd = random.Random(8)
    sG = b"".join(rnd.randrange(L).to_bytes(32, "little") for _ in range(5000))
    sH = b"".join(rnd.randrange(L).to_bytes(32, "little") for _ in range(3001))
    sB, sBb = rnd.randrange(L).to_bytes(32, "little"), rnd.randrange(L).to_bytes(32, "little")
    whole = ctx.msm_gens_bytes(sG, sH, sB, sBb)
    assert whole == coracle.msm_gens(sG, sH, sB, sBb)
    for parts in (1, 2, 3, 8):
        ctxs = [ctx] + [ctx.shared() for _ in range(parts - 1)]
        assert sharding.msm_gens_sharded(ctxs, sG, sH, sB, sBb) == whole
        for c in ctxs[1:]:
            c.close()
