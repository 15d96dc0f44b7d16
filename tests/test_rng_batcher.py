"""Host logic: the Merlin TranscriptRng stream of Prover::prove (2n x fill_bytes(64)) through the cross-proof
batcher (AVX-512 Keccak-f x8, csrc/keccak_x8_native.cpp + merlin.cpp) must produce exactly the bytes of the
oracle's one-stream-at-a-time STROBE, whether a stream runs alone or in a batch of 2..8 with unequal lengths."""
import ctypes
import threading

import pytest

import bulletproof_gadgets_b200 as bpg
from bulletproof_gadgets_b200 import build
from oracle.pyref.merlin import Transcript as OTranscript


@pytest.fixture(scope="module")
def lib():
    build.build_lib()
    return bpg.lib()


def oracle_stream(label, witnesses, seed, warm, n):
    T = OTranscript(label)
    T.append_message(b"dom-sep", b"r1cs v1")
    b = T.build_rng()
    for w in witnesses:
        b.rekey_with_witness_bytes(b"v_blinding", w)
    rng = b.finalize(seed)
    for _ in range(warm):
        rng.fill_bytes(64)
    return b"".join(rng.fill_bytes(64) for _ in range(n))


def product_stream(lib, label, witnesses, seed, warm, n):
    T = bpg.Transcript(label)
    T.append_message(b"dom-sep", b"r1cs v1")
    out = ctypes.create_string_buffer(64 * max(n, 1))
    rc = lib.bpg_transcript_rng_fill64(T._h, b"".join(witnesses) or None, len(witnesses), seed, warm, out, n)
    assert rc == 0
    return out.raw[: 64 * n]


def test_single_stream_matches_oracle(lib):
    for warm, n in ((3, 200), (0, 70), (1, 1), (3, 0), (3, 63), (3, 64)):
        wit = [bytes([i + 1]) * 32 for i in range(3)]
        assert product_stream(lib, b"solo", wit, b"\x05" * 32, warm, n) == oracle_stream(b"solo", wit, b"\x05" * 32, warm, n)


def test_x8_kernel_on_short_unequal_streams(lib):
    """Concurrent short streams of unequal length (the kernel hands each state back when its stream ends):
    whatever batches happen to form, the bytes equal the oracle's."""
    n = 8
    counts = [64 + 37 * i for i in range(n)]
    want = [oracle_stream(b"short-%d" % i, [bytes([i]) * 32], bytes([0x40 + i]) * 32, 3, counts[i]) for i in range(n)]
    got = [None] * n

    def work(i):
        got[i] = product_stream(lib, b"short-%d" % i, [bytes([i]) * 32], bytes([0x40 + i]) * 32, 3, counts[i])

    for rep in range(10):
        ts = [threading.Thread(target=work, args=(i,)) for i in range(n)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        assert got == want


@pytest.mark.parametrize("nthreads", [2, 5, 8, 11])
def test_concurrent_long_streams_are_batched_and_exact(lib, nthreads):
    """Long streams (a proof's worth of draws overlaps in time): sequential calls take the one-stream path (checked
    against the oracle on its first draws), concurrent calls must form vector batches and return the same bytes."""
    counts = [20000 + 1500 * i for i in range(nthreads)]
    args = [(b"long-%d" % i, [bytes([i]) * 32, bytes([i + 1]) * 32], bytes([0x60 + i]) * 32, 3, counts[i]) for i in range(nthreads)]
    want = [product_stream(lib, *a) for a in args]                       # one at a time: scalar path
    assert want[0][: 64 * 40] == oracle_stream(*args[0][:4], 40)
    before = lib.bpg_rng_batcher_stat(1)
    got = [None] * nthreads
    gate = threading.Barrier(nthreads)

    def work(i):
        gate.wait()
        got[i] = product_stream(lib, *args[i])

    formed = False
    for rep in range(12):
        ts = [threading.Thread(target=work, args=(i,)) for i in range(nthreads)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        assert got == want
        formed = lib.bpg_rng_batcher_stat(1) > before
        if formed and rep >= 1:
            break
    if nthreads >= 5 and "avx512f" in open("/proc/cpuinfo").read():      # pairs may legitimately run alone (low-latency policy)
        assert formed, "no vector batch was formed on an AVX-512 host"


def test_scalar_keccak_path_in_a_subprocess():
    """BPG_KECCAK=scalar disables both AVX-512 paths (single-state permutation and the x8 batcher): the portable code
    must give the same transcript / rng bytes (this is what a host without AVX-512 runs)."""
    import os
    import subprocess
    import sys
    code = (
        "import ctypes, sys; sys.path.insert(0, %r)\n"
        "import bulletproof_gadgets_b200 as bpg\n"
        "T = bpg.Transcript(b'test protocol'); T.append_message(b'some label', b'some data')\n"
        "print(T.challenge_bytes(b'challenge', 32).hex())\n"
        "T = bpg.Transcript(b'solo'); T.append_message(b'dom-sep', b'r1cs v1')\n"
        "out = ctypes.create_string_buffer(64 * 300)\n"
        "assert bpg.lib().bpg_transcript_rng_fill64(T._h, b'\\x01' * 32, 1, b'\\x05' * 32, 3, out, 300) == 0\n"
        "print(out.raw.hex())\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    outs = []
    for mode in ("scalar", "auto"):
        env = dict(os.environ, BPG_KECCAK=mode)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr
        outs.append(r.stdout.split())
    assert outs[0] == outs[1]
    assert outs[0][0] == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"
    assert bytes.fromhex(outs[0][1]) == oracle_stream(b"solo", [b"\x01" * 32], b"\x05" * 32, 3, 300)
