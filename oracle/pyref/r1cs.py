"""ORACLE (test infrastructure, never shipped) -- Bulletproofs R1CS prover / verifier / IPP.

Big-int restatement of the FairAds fork of dalek `bulletproofs` 2.1.0
(/root/reference/Cargo.lock:78-80, git 3c00b01e...; source NOT vendored -> restated from the
published upstream 2.0.0 algorithm: r1cs/{prover,verifier,proof,linear_combination}.rs,
inner_product_proof.rs, generators.rs, util.rs).  1-phase proofs only: the reference never
calls specify_randomized_constraints.

Reference call sites:
  Prover::new/commit/prove        /root/reference/src/prove.rs:47,79  /root/reference/src/gadget.rs:32
  Verifier::new/commit/verify     /root/reference/src/verify.rs:46,71 /root/reference/src/lalrpop/assignment_parser.rs:138
  BulletproofGens::new(n',1)      /root/reference/src/prove.rs:78     /root/reference/src/verify.rs:70
  R1CSProof::{to_bytes,from_bytes} /root/reference/src/prove.rs:81    /root/reference/src/verify.rs:53
  ConstraintSystem::{multiply,allocate_multiplier,constrain} /root/reference/src/cs_buffer.rs:89-116

Parity with dalek bytes: UNPINNED (no reference test pins proof bytes; no Rust toolchain here).
Pinned pieces: B_blinding constant, Merlin KAT, RFC 9496 vectors, accept/reject outcomes.

Scalars are Python ints.  Values that came from Scalar::from_bits may be >= L (never reduced
until arithmetic touches them), exactly like dalek.
"""
import hashlib

from . import ed
from .ed import Point
from .merlin import L, Transcript, VerificationError

# ---------------------------------------------------------------------------------- generators


class PedersenGens:
    """PedersenGens::default(): B = ristretto basepoint, B_blinding = hash_from_bytes::<Sha3_512>(B)."""

    def __init__(self):
        self.B = ed.BASEPOINT
        self.B_blinding = ed.from_uniform_bytes(hashlib.sha3_512(ed.BASEPOINT_COMPRESSED).digest())

    def commit(self, v, r):
        return self.B * (v % L) + self.B_blinding * (r % L)


_GENS_CACHE = {}


def generators_chain(label: bytes, n: int):
    """GeneratorsChain::new(label).take(n): SHAKE256("GeneratorsChain"||label) 64 B per point."""
    key = (label, n)
    if key not in _GENS_CACHE:
        stream = hashlib.shake_256(b"GeneratorsChain" + label).digest(64 * n)
        _GENS_CACHE[key] = [ed.from_uniform_bytes(stream[64 * i: 64 * i + 64]) for i in range(n)]
    return _GENS_CACHE[key]


class BulletproofGens:
    """BulletproofGens::new(gens_capacity, 1) -- party 0 only."""

    def __init__(self, gens_capacity):
        self.gens_capacity = gens_capacity
        self.G = generators_chain(b"G" + (0).to_bytes(4, "little"), gens_capacity)
        self.H = generators_chain(b"H" + (0).to_bytes(4, "little"), gens_capacity)


def round_pow2(num):
    """/root/reference/src/prove.rs:33-35 (ceil(log2) via f64; 0 -> 1)."""
    n = 1
    while n < num:
        n *= 2
    return n


# ---------------------------------------------------------------------------------- variables

COMMITTED, MUL_LEFT, MUL_RIGHT, MUL_OUT, ONE = 0, 1, 2, 3, 4


class Variable(tuple):
    """(kind, index) like bulletproofs::r1cs::Variable."""

    def __new__(cls, kind, idx=0):
        return tuple.__new__(cls, (kind, idx))

    kind = property(lambda s: s[0])
    idx = property(lambda s: s[1])


def One():
    return Variable(ONE, 0)


class LC:
    """LinearCombination: plain term list, +/- concatenate, no merging."""

    __slots__ = ("terms",)

    def __init__(self, terms=None):
        self.terms = list(terms) if terms else []

    @staticmethod
    def of(x):
        if isinstance(x, LC):
            return x
        if isinstance(x, Variable):
            return LC([(x, 1)])
        if isinstance(x, int):
            return LC([(One(), x)])  # Scalar -> LC; keeps raw (possibly unreduced) value
        raise TypeError(x)

    def __add__(self, o):
        return LC(self.terms + LC.of(o).terms)

    def __sub__(self, o):
        return LC(self.terms + [(v, (-c) % L) for v, c in LC.of(o).terms])

    def __neg__(self):
        return LC([(v, (-c) % L) for v, c in self.terms])

    def scale(self, s):
        return LC([(v, (c * s) % L) for v, c in self.terms])


# ---------------------------------------------------------------------------------- helpers


def inner_product(a, b):
    return sum(x * y for x, y in zip(a, b)) % L


def inv(x):
    return pow(x % L, L - 2, L)


def sbytes(x):
    return int(x).to_bytes(32, "little")


class R1CSError(Exception):
    pass


class FormatError(R1CSError):
    pass


# ---------------------------------------------------------------------------------- IPP


def ipp_create(transcript, Q, G_factors, H_factors, G, H, a, b, trace=None):
    """InnerProductProof::create (inner_product_proof.rs)."""
    n = len(G)
    assert n == len(H) == len(a) == len(b) == len(G_factors) == len(H_factors)
    assert n & (n - 1) == 0 and n >= 1
    transcript.innerproduct_domain_sep(n)
    G = list(G)
    H = list(H)
    a = list(a)
    b = list(b)
    Lv, Rv = [], []
    first = True
    while n != 1:
        n //= 2
        aL, aR, bL, bR = a[:n], a[n:], b[:n], b[n:]
        GL, GR, HL, HR = G[:n], G[n:], H[:n], H[n:]
        cL = inner_product(aL, bR)
        cR = inner_product(aR, bL)
        if first:
            gfL, gfR, hfL, hfR = G_factors[:n], G_factors[n:], H_factors[:n], H_factors[n:]
        else:
            gfL = gfR = hfL = hfR = [1] * n
        Lp = ed.msm([x * g % L for x, g in zip(aL, gfR)] + [x * h % L for x, h in zip(bR, hfL)] + [cL], GR + HL + [Q])
        Rp = ed.msm([x * g % L for x, g in zip(aR, gfL)] + [x * h % L for x, h in zip(bL, hfR)] + [cR], GL + HR + [Q])
        Lc, Rc = Lp.compress(), Rp.compress()
        Lv.append(Lc)
        Rv.append(Rc)
        transcript.append_point(b"L", Lc)
        transcript.append_point(b"R", Rc)
        u = transcript.challenge_scalar(b"u")
        ui = inv(u)
        if trace is not None:
            trace.setdefault("ipp_u", []).append(u)
        a = [(aL[i] * u + ui * aR[i]) % L for i in range(n)]
        b = [(bL[i] * ui + u * bR[i]) % L for i in range(n)]
        G = [GL[i] * (ui * gfL[i] % L) + GR[i] * (u * gfR[i] % L) for i in range(n)]
        H = [HL[i] * (u * hfL[i] % L) + HR[i] * (ui * hfR[i] % L) for i in range(n)]
        first = False
    return Lv, Rv, a[0] % L, b[0] % L


def ipp_verification_scalars(Lv, Rv, n, transcript):
    lg_n = len(Lv)
    if lg_n >= 32 or n != (1 << lg_n):
        raise VerificationError("ipp size")
    transcript.innerproduct_domain_sep(n)
    ch = []
    for Lc, Rc in zip(Lv, Rv):
        transcript.validate_and_append_point(b"L", Lc)
        transcript.validate_and_append_point(b"R", Rc)
        ch.append(transcript.challenge_scalar(b"u"))
    ch_inv = [inv(c) for c in ch]
    allinv = 1
    for c in ch_inv:
        allinv = allinv * c % L
    ch_sq = [c * c % L for c in ch]
    ch_inv_sq = [c * c % L for c in ch_inv]
    s = [allinv]
    for i in range(1, n):
        lg_i = i.bit_length() - 1
        k = 1 << lg_i
        s.append(s[i - k] * ch_sq[(lg_n - 1) - lg_i] % L)
    return ch_sq, ch_inv_sq, s


# ---------------------------------------------------------------------------------- proof codec


class R1CSProof:
    FIELDS_P = ("A_I1", "A_O1", "S1", "A_I2", "A_O2", "S2", "T_1", "T_3", "T_4", "T_5", "T_6")

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def to_bytes(self):
        out = bytearray()
        if self.A_I2 == bytes(32) and self.A_O2 == bytes(32) and self.S2 == bytes(32):
            out.append(0)
            out += self.A_I1 + self.A_O1 + self.S1
        else:
            out.append(1)
            out += self.A_I1 + self.A_O1 + self.S1 + self.A_I2 + self.A_O2 + self.S2
        out += self.T_1 + self.T_3 + self.T_4 + self.T_5 + self.T_6
        out += sbytes(self.t_x) + sbytes(self.t_x_blinding) + sbytes(self.e_blinding)
        for Lc, Rc in zip(self.L_vec, self.R_vec):
            out += Lc + Rc
        out += sbytes(self.a) + sbytes(self.b)
        return bytes(out)

    @staticmethod
    def from_bytes(buf):
        if len(buf) == 0:
            raise FormatError("empty")
        version, body = buf[0], buf[1:]
        if len(body) % 32 != 0:
            raise FormatError("length")
        if version == 0:
            minlen = 11 * 32
        elif version == 1:
            minlen = 14 * 32
        else:
            raise FormatError("version")
        if len(body) < minlen:
            raise FormatError("short")
        pos = [0]

        def rd():
            c = body[pos[0]: pos[0] + 32]
            pos[0] += 32
            return bytes(c)

        def rds():
            v = int.from_bytes(rd(), "little")
            if v >= L:
                raise FormatError("non-canonical scalar")
            return v

        A_I1, A_O1, S1 = rd(), rd(), rd()
        if version == 0:
            A_I2 = A_O2 = S2 = bytes(32)
        else:
            A_I2, A_O2, S2 = rd(), rd(), rd()
        T_1, T_3, T_4, T_5, T_6 = rd(), rd(), rd(), rd(), rd()
        t_x, t_x_blinding, e_blinding = rds(), rds(), rds()
        rest = body[pos[0]:]
        ne = len(rest) // 32
        if ne < 2 or (ne - 2) % 2 != 0:
            raise FormatError("ipp length")
        lg_n = (ne - 2) // 2
        if lg_n >= 32:
            raise FormatError("ipp too large")
        L_vec, R_vec = [], []
        for _ in range(lg_n):
            L_vec.append(rd())
            R_vec.append(rd())
        a, b = rds(), rds()
        return R1CSProof(A_I1=A_I1, A_O1=A_O1, S1=S1, A_I2=A_I2, A_O2=A_O2, S2=S2, T_1=T_1, T_3=T_3, T_4=T_4,
                         T_5=T_5, T_6=T_6, t_x=t_x, t_x_blinding=t_x_blinding, e_blinding=e_blinding,
                         L_vec=L_vec, R_vec=R_vec, a=a, b=b)


# ---------------------------------------------------------------------------------- constraint system


class _CS:
    def __init__(self):
        self.constraints = []  # list of term lists
        self.n = 0

    def _flatten(self, z, n, m, verifier):
        wL, wR, wO, wV = [0] * n, [0] * n, [0] * n, [0] * m
        wc = 0
        exp_z = z
        for terms in self.constraints:
            for var, coeff in terms:
                k, i = var
                if k == MUL_LEFT:
                    wL[i] = (wL[i] + exp_z * coeff) % L
                elif k == MUL_RIGHT:
                    wR[i] = (wR[i] + exp_z * coeff) % L
                elif k == MUL_OUT:
                    wO[i] = (wO[i] + exp_z * coeff) % L
                elif k == COMMITTED:
                    wV[i] = (wV[i] - exp_z * coeff) % L
                elif verifier:
                    wc = (wc - exp_z * coeff) % L
            exp_z = exp_z * z % L
        return wL, wR, wO, wV, wc

    def constrain(self, lc):
        self.constraints.append(list(LC.of(lc).terms))

    def num_constraints(self):
        return len(self.constraints)


class Prover(_CS):
    def __init__(self, pc_gens, transcript):
        super().__init__()
        self.pc_gens = pc_gens
        self.transcript = transcript
        transcript.r1cs_domain_sep()
        self.a_L, self.a_R, self.a_O, self.v, self.v_blinding = [], [], [], [], []

    def commit(self, v, v_blinding):
        i = len(self.v)
        self.v.append(v)
        self.v_blinding.append(v_blinding)
        V = self.pc_gens.commit(v, v_blinding).compress()
        self.transcript.append_point(b"V", V)
        return V, Variable(COMMITTED, i)

    def eval(self, lc):
        acc = 0
        for (k, i), c in LC.of(lc).terms:
            if k == MUL_LEFT:
                val = self.a_L[i]
            elif k == MUL_RIGHT:
                val = self.a_R[i]
            elif k == MUL_OUT:
                val = self.a_O[i]
            elif k == COMMITTED:
                val = self.v[i]
            else:
                val = 1
            acc += c * val
        return acc % L

    def multiply(self, left, right):
        left, right = LC.of(left), LC.of(right)
        l, r = self.eval(left), self.eval(right)
        o = l * r % L
        i = len(self.a_L)
        self.a_L.append(l)
        self.a_R.append(r)
        self.a_O.append(o)
        lv, rv, ov = Variable(MUL_LEFT, i), Variable(MUL_RIGHT, i), Variable(MUL_OUT, i)
        self.constrain(left - lv)
        self.constrain(right - rv)
        return lv, rv, ov

    def allocate_multiplier(self, assignment):
        if assignment is None:
            raise R1CSError("MissingAssignment")
        l, r = assignment
        o = l * r % L
        i = len(self.a_L)
        self.a_L.append(l)  # raw (from_bits values stay unreduced)
        self.a_R.append(r)
        self.a_O.append(o)
        return Variable(MUL_LEFT, i), Variable(MUL_RIGHT, i), Variable(MUL_OUT, i)

    def num_multipliers(self):
        return len(self.a_L)

    def prove(self, bp_gens, external32=bytes(32), trace=None):
        """Prover::prove.  `external32` = the 32 bytes TranscriptRngBuilder::finalize pulls from
        thread_rng() in the reference."""
        T = self.transcript
        T.append_u64(b"m", len(self.v))
        builder = T.build_rng()
        for vb in self.v_blinding:
            builder.rekey_with_witness_bytes(b"v_blinding", sbytes(vb))
        rng = builder.finalize(external32)

        n = len(self.a_L)
        if bp_gens.gens_capacity < n:
            raise R1CSError("InvalidGeneratorsLength")
        G, H = bp_gens.G, bp_gens.H
        B, Bb = self.pc_gens.B, self.pc_gens.B_blinding

        i_bl = rng.random_scalar()
        o_bl = rng.random_scalar()
        s_bl = rng.random_scalar()
        s_L = [rng.random_scalar() for _ in range(n)]
        s_R = [rng.random_scalar() for _ in range(n)]

        A_I1 = ed.msm([i_bl] + self.a_L + self.a_R, [Bb] + G[:n] + H[:n]).compress()
        A_O1 = ed.msm([o_bl] + self.a_O, [Bb] + G[:n]).compress()
        S1 = ed.msm([s_bl] + s_L + s_R, [Bb] + G[:n] + H[:n]).compress()
        T.append_point(b"A_I1", A_I1)
        T.append_point(b"A_O1", A_O1)
        T.append_point(b"S1", S1)
        T.r1cs_1phase_domain_sep()

        padded_n = round_pow2(n) if n > 0 else 1
        # usize::next_power_of_two(0) == 1
        pad = padded_n - n
        if bp_gens.gens_capacity < padded_n:
            raise R1CSError("InvalidGeneratorsLength")
        Z32 = bytes(32)
        T.append_point(b"A_I2", Z32)
        T.append_point(b"A_O2", Z32)
        T.append_point(b"S2", Z32)

        y = T.challenge_scalar(b"y")
        z = T.challenge_scalar(b"z")
        wL, wR, wO, wV, _ = self._flatten(z, n, len(self.v), verifier=False)

        y_inv = inv(y)
        exp_y_inv = [1] * padded_n
        for i in range(1, padded_n):
            exp_y_inv[i] = exp_y_inv[i - 1] * y_inv % L
        l1, l2, l3 = [0] * n, [0] * n, [0] * n
        r0, r1, r3 = [0] * n, [0] * n, [0] * n
        exp_y = 1
        for i in range(n):
            l1[i] = (self.a_L[i] + exp_y_inv[i] * wR[i]) % L
            l2[i] = self.a_O[i] % L
            l3[i] = s_L[i]
            r0[i] = (wO[i] - exp_y) % L
            r1[i] = (exp_y * self.a_R[i] + wL[i]) % L
            r3[i] = exp_y * s_R[i] % L
            exp_y = exp_y * y % L

        # VecPoly3::special_inner_product (l0 = 0, r2 = 0)
        t1 = inner_product(l1, r0)
        t2 = (inner_product(l1, r1) + inner_product(l2, r0)) % L
        t3 = (inner_product(l2, r1) + inner_product(l3, r0)) % L
        t4 = (inner_product(l1, r3) + inner_product(l3, r1)) % L
        t5 = inner_product(l2, r3)
        t6 = inner_product(l3, r3)

        tb1 = rng.random_scalar()
        tb3 = rng.random_scalar()
        tb4 = rng.random_scalar()
        tb5 = rng.random_scalar()
        tb6 = rng.random_scalar()
        T_1 = self.pc_gens.commit(t1, tb1).compress()
        T_3 = self.pc_gens.commit(t3, tb3).compress()
        T_4 = self.pc_gens.commit(t4, tb4).compress()
        T_5 = self.pc_gens.commit(t5, tb5).compress()
        T_6 = self.pc_gens.commit(t6, tb6).compress()
        for lab, pt in ((b"T_1", T_1), (b"T_3", T_3), (b"T_4", T_4), (b"T_5", T_5), (b"T_6", T_6)):
            T.append_point(lab, pt)

        u = T.challenge_scalar(b"u")
        x = T.challenge_scalar(b"x")
        tb2 = sum(c * vb for c, vb in zip(wV, self.v_blinding)) % L

        def poly6(c1, c2, c3, c4, c5, c6):
            return x * (c1 + x * (c2 + x * (c3 + x * (c4 + x * (c5 + x * c6))))) % L

        t_x = poly6(t1, t2, t3, t4, t5, t6)
        t_x_blinding = poly6(tb1, tb2, tb3, tb4, tb5, tb6)
        l_vec = [(x * (l1[i] + x * (l2[i] + x * l3[i]))) % L for i in range(n)] + [0] * pad
        r_vec = [(r0[i] + x * (r1[i] + x * (x * r3[i]))) % L for i in range(n)] + [0] * pad
        for i in range(n, padded_n):
            r_vec[i] = (-exp_y) % L
            exp_y = exp_y * y % L
        e_blinding = x * (i_bl + x * (o_bl + x * s_bl)) % L

        T.append_scalar(b"t_x", sbytes(t_x))
        T.append_scalar(b"t_x_blinding", sbytes(t_x_blinding))
        T.append_scalar(b"e_blinding", sbytes(e_blinding))
        w = T.challenge_scalar(b"w")
        Q = B * w
        G_factors = [1] * n + [u] * pad
        H_factors = [exp_y_inv[i] * G_factors[i] % L for i in range(padded_n)]
        if trace is not None:
            trace.update(y=y, z=z, u=u, x=x, w=w, t=(t1, t2, t3, t4, t5, t6), l_vec=l_vec, r_vec=r_vec,
                         wL=wL, wR=wR, wO=wO, wV=wV, i_bl=i_bl, o_bl=o_bl, s_bl=s_bl)
        Lv, Rv, a, b = ipp_create(T, Q, G_factors, H_factors, G[:padded_n], H[:padded_n], l_vec, r_vec, trace)
        return R1CSProof(A_I1=A_I1, A_O1=A_O1, S1=S1, A_I2=Z32, A_O2=Z32, S2=Z32, T_1=T_1, T_3=T_3, T_4=T_4,
                         T_5=T_5, T_6=T_6, t_x=t_x, t_x_blinding=t_x_blinding, e_blinding=e_blinding,
                         L_vec=Lv, R_vec=Rv, a=a, b=b)


class Verifier(_CS):
    def __init__(self, transcript):
        super().__init__()
        self.transcript = transcript
        transcript.r1cs_domain_sep()
        self.V = []
        self.num_vars = 0

    def commit(self, V_bytes):
        i = len(self.V)
        self.V.append(bytes(V_bytes))
        self.transcript.append_point(b"V", bytes(V_bytes))
        return Variable(COMMITTED, i)

    def _alloc(self):
        i = self.num_vars
        self.num_vars += 1
        return Variable(MUL_LEFT, i), Variable(MUL_RIGHT, i), Variable(MUL_OUT, i)

    def multiply(self, left, right):
        left, right = LC.of(left), LC.of(right)
        lv, rv, ov = self._alloc()
        self.constrain(left - lv)
        self.constrain(right - rv)
        return lv, rv, ov

    def allocate_multiplier(self, assignment=None):
        return self._alloc()

    def num_multipliers(self):
        return self.num_vars

    def mega_msm_terms(self, proof, pc_gens, bp_gens, external32=bytes(32)):
        """Returns (scalars, points or None) of the single verification MSM (r1cs/verifier.rs verify)."""
        T = self.transcript
        T.append_u64(b"m", len(self.V))
        n = self.num_vars
        T.validate_and_append_point(b"A_I1", proof.A_I1)
        T.validate_and_append_point(b"A_O1", proof.A_O1)
        T.validate_and_append_point(b"S1", proof.S1)
        T.r1cs_1phase_domain_sep()
        padded_n = round_pow2(n) if n > 0 else 1
        pad = padded_n - n
        if bp_gens.gens_capacity < padded_n:
            raise R1CSError("InvalidGeneratorsLength")
        T.append_point(b"A_I2", proof.A_I2)
        T.append_point(b"A_O2", proof.A_O2)
        T.append_point(b"S2", proof.S2)
        y = T.challenge_scalar(b"y")
        z = T.challenge_scalar(b"z")
        for lab, pt in ((b"T_1", proof.T_1), (b"T_3", proof.T_3), (b"T_4", proof.T_4), (b"T_5", proof.T_5),
                        (b"T_6", proof.T_6)):
            T.validate_and_append_point(lab, pt)
        u = T.challenge_scalar(b"u")
        x = T.challenge_scalar(b"x")
        T.append_scalar(b"t_x", sbytes(proof.t_x))
        T.append_scalar(b"t_x_blinding", sbytes(proof.t_x_blinding))
        T.append_scalar(b"e_blinding", sbytes(proof.e_blinding))
        w = T.challenge_scalar(b"w")
        wL, wR, wO, wV, wc = self._flatten(z, n, len(self.V), verifier=True)
        u_sq, u_inv_sq, s = ipp_verification_scalars(proof.L_vec, proof.R_vec, padded_n, T)
        a, b = proof.a, proof.b
        y_inv = inv(y)
        y_inv_vec = [1] * padded_n
        for i in range(1, padded_n):
            y_inv_vec[i] = y_inv_vec[i - 1] * y_inv % L
        yneg_wR = [wR[i] * y_inv_vec[i] % L for i in range(n)] + [0] * pad
        delta = inner_product(yneg_wR[:n], wL)
        uf = [1] * n + [u] * pad
        g_scalars = [uf[i] * (x * yneg_wR[i] - a * s[i]) % L for i in range(padded_n)]
        wLp = wL + [0] * pad
        wOp = wO + [0] * pad
        h_scalars = [uf[i] * (y_inv_vec[i] * (x * wLp[i] + wOp[i] - b * s[padded_n - 1 - i]) - 1) % L
                     for i in range(padded_n)]
        rng = T.build_rng().finalize(external32)
        r = rng.random_scalar()
        xx = x * x % L
        rxx = r * xx % L
        xxx = x * xx % L
        T_scalars = [r * x % L, rxx * x % L, rxx * xx % L, rxx * xxx % L, rxx * xx % L * xx % L]
        scalars = ([x, xx, xxx, u * x % L, u * xx % L, u * xxx % L] + [wVi * rxx % L for wVi in wV] + T_scalars
                   + [(w * (proof.t_x - a * b) + r * (xx * (wc + delta) - proof.t_x)) % L,
                      (-proof.e_blinding - r * proof.t_x_blinding) % L]
                   + g_scalars + h_scalars + u_sq + u_inv_sq)
        enc = ([proof.A_I1, proof.A_O1, proof.S1, proof.A_I2, proof.A_O2, proof.S2] + self.V
               + [proof.T_1, proof.T_3, proof.T_4, proof.T_5, proof.T_6])
        dyn = [ed.decompress(e) for e in enc]
        tail = [ed.decompress(e) for e in proof.L_vec + proof.R_vec]
        if any(p is None for p in dyn + tail):
            raise VerificationError("undecodable point")
        points = dyn + [pc_gens.B, pc_gens.B_blinding] + bp_gens.G[:padded_n] + bp_gens.H[:padded_n] + tail
        return scalars, points

    def verify(self, proof, pc_gens, bp_gens, external32=bytes(32)):
        scalars, points = self.mega_msm_terms(proof, pc_gens, bp_gens, external32)
        if not ed.msm(scalars, points).is_identity():
            raise VerificationError("mega check")
        return True
