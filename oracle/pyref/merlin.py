"""ORACLE (test infrastructure, never shipped) -- Merlin 2.0.1 transcripts over STROBE-128.

Restates merlin 2.0.1 (pinned /root/reference/Cargo.lock:403-405; source not vendored):
strobe.rs (Strobe128: new128, run_f, absorb/overwrite/squeeze, begin_op, meta_ad/ad/prf/key)
and transcript.rs (Transcript, TranscriptRngBuilder, TranscriptRng), plus the bulletproofs
TranscriptProtocol extension trait (transcript.rs of bulletproofs 2.1.0 fork,
/root/reference/Cargo.lock:78-80).

Reference call sites: Transcript::new  /root/reference/src/prove.rs:45,55,193 and
/root/reference/src/verify.rs:44,49,135.

Pinned by: merlin's own "test protocol" known-answer (tests/test_oracle_anchors.py).
"""

L = 2**252 + 27742317777372353535851937790883648493

_RC = [
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000,
    0x000000000000808B, 0x0000000080000001, 0x8000000080008081, 0x8000000000008009,
    0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
    0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003,
    0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
    0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008,
]
_ROT = [
    [0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14],
]
_M64 = (1 << 64) - 1


def _rol(x, n):
    n %= 64
    return ((x << n) | (x >> (64 - n))) & _M64 if n else x


def keccak_f1600(state: bytearray):
    A = [[int.from_bytes(state[8 * (x + 5 * y): 8 * (x + 5 * y) + 8], "little") for y in range(5)] for x in range(5)]
    for rnd in range(24):
        C = [A[x][0] ^ A[x][1] ^ A[x][2] ^ A[x][3] ^ A[x][4] for x in range(5)]
        Dv = [C[(x - 1) % 5] ^ _rol(C[(x + 1) % 5], 1) for x in range(5)]
        A = [[A[x][y] ^ Dv[x] for y in range(5)] for x in range(5)]
        B = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                B[y][(2 * x + 3 * y) % 5] = _rol(A[x][y], _ROT[x][y])
        A = [[B[x][y] ^ ((~B[(x + 1) % 5][y]) & B[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        A[0][0] ^= _RC[rnd]
    for x in range(5):
        for y in range(5):
            state[8 * (x + 5 * y): 8 * (x + 5 * y) + 8] = (A[x][y] & _M64).to_bytes(8, "little")


STROBE_R = 166
FLAG_I, FLAG_A, FLAG_C, FLAG_T, FLAG_M, FLAG_K = 1, 2, 4, 8, 16, 32


class Strobe128:
    def __init__(self, protocol_label: bytes = None, _clone=None):
        if _clone is not None:
            self.state = bytearray(_clone.state)
            self.pos, self.pos_begin, self.cur_flags = _clone.pos, _clone.pos_begin, _clone.cur_flags
            return
        st = bytearray(200)
        st[0:6] = bytes([1, STROBE_R + 2, 1, 0, 1, 96])
        st[6:18] = b"STROBEv1.0.2"
        keccak_f1600(st)
        self.state = st
        self.pos = self.pos_begin = self.cur_flags = 0
        self.meta_ad(protocol_label, False)

    def clone(self):
        return Strobe128(_clone=self)

    def _run_f(self):
        self.state[self.pos] ^= self.pos_begin
        self.state[self.pos + 1] ^= 0x04
        self.state[STROBE_R + 1] ^= 0x80
        keccak_f1600(self.state)
        self.pos = self.pos_begin = 0

    def _absorb(self, data):
        for b in data:
            self.state[self.pos] ^= b
            self.pos += 1
            if self.pos == STROBE_R:
                self._run_f()

    def _overwrite(self, data):
        for b in data:
            self.state[self.pos] = b
            self.pos += 1
            if self.pos == STROBE_R:
                self._run_f()

    def _squeeze(self, n):
        out = bytearray(n)
        for i in range(n):
            out[i] = self.state[self.pos]
            self.state[self.pos] = 0
            self.pos += 1
            if self.pos == STROBE_R:
                self._run_f()
        return bytes(out)

    def _begin_op(self, flags, more):
        if more:
            assert self.cur_flags == flags
            return
        assert flags & FLAG_T == 0
        old_begin = self.pos_begin
        self.pos_begin = self.pos + 1
        self.cur_flags = flags
        self._absorb(bytes([old_begin, flags]))
        if (flags & (FLAG_C | FLAG_K)) and self.pos != 0:
            self._run_f()

    def meta_ad(self, data, more):
        self._begin_op(FLAG_M | FLAG_A, more)
        self._absorb(data)

    def ad(self, data, more):
        self._begin_op(FLAG_A, more)
        self._absorb(data)

    def prf(self, n, more=False):
        self._begin_op(FLAG_I | FLAG_A | FLAG_C, more)
        return self._squeeze(n)

    def key(self, data, more):
        self._begin_op(FLAG_A | FLAG_C, more)
        self._overwrite(data)


def _le32(n):
    return int(n).to_bytes(4, "little")


class Transcript:
    def __init__(self, label: bytes = None, _strobe=None):
        if _strobe is not None:
            self.strobe = _strobe
            return
        self.strobe = Strobe128(b"Merlin v1.0")
        self.append_message(b"dom-sep", label)

    def clone(self):
        return Transcript(_strobe=self.strobe.clone())

    def append_message(self, label, message):
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(_le32(len(message)), True)
        self.strobe.ad(message, False)

    def append_u64(self, label, x):
        self.append_message(label, int(x).to_bytes(8, "little"))

    def challenge_bytes(self, label, n):
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(_le32(n), True)
        return self.strobe.prf(n, False)

    # ---- bulletproofs TranscriptProtocol ----
    def r1cs_domain_sep(self):
        self.append_message(b"dom-sep", b"r1cs v1")

    def r1cs_1phase_domain_sep(self):
        self.append_message(b"dom-sep", b"r1cs-1phase")

    def innerproduct_domain_sep(self, n):
        self.append_message(b"dom-sep", b"ipp v1")
        self.append_u64(b"n", n)

    def append_scalar(self, label, s_bytes):
        assert len(s_bytes) == 32
        self.append_message(label, s_bytes)

    def append_point(self, label, p_bytes):
        assert len(p_bytes) == 32
        self.append_message(label, p_bytes)

    def validate_and_append_point(self, label, p_bytes):
        if p_bytes == bytes(32):
            raise VerificationError("identity point in transcript")
        self.append_message(label, p_bytes)

    def challenge_scalar(self, label):
        return int.from_bytes(self.challenge_bytes(label, 64), "little") % L

    def build_rng(self):
        return TranscriptRngBuilder(self.strobe.clone())


class VerificationError(Exception):
    pass


class TranscriptRngBuilder:
    def __init__(self, strobe):
        self.strobe = strobe

    def rekey_with_witness_bytes(self, label, witness):
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(_le32(len(witness)), True)
        self.strobe.key(witness, False)
        return self

    def finalize(self, external32: bytes):
        """`external32` replaces the 32 bytes merlin draws from the caller's RNG (thread_rng in
        the reference: bulletproofs r1cs/prover.rs `builder.finalize(&mut thread_rng())`)."""
        assert len(external32) == 32
        self.strobe.meta_ad(b"rng", False)
        self.strobe.key(external32, False)
        return TranscriptRng(self.strobe)


class TranscriptRng:
    def __init__(self, strobe):
        self.strobe = strobe

    def fill_bytes(self, n):
        self.strobe.meta_ad(_le32(n), False)
        return self.strobe.prf(n, False)

    def random_scalar(self):
        """Scalar::random: 64 bytes -> from_bytes_mod_order_wide."""
        return int.from_bytes(self.fill_bytes(64), "little") % L
