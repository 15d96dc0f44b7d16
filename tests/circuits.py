"""Small constraint systems written once against a backend-neutral interface, so the same circuit
can be driven through the oracle (oracle/pyref/r1cs.py) and through the CUDA library (C ABI)."""
import hashlib

from oracle.pyref import r1cs as O
from oracle.pyref.merlin import L, Transcript as OTranscript


def det_scalar(tag, i=0, bits=None):
    v = int.from_bytes(hashlib.shake_256(b"%s/%d" % (tag if isinstance(tag, bytes) else tag.encode(), i)).digest(64), "little")
    return v % L if bits is None else v % (1 << bits)


class OracleBackend:
    name = "oracle"

    def __init__(self):
        self.pc = O.PedersenGens()

    def transcript(self, label):
        return OTranscript(label)

    def prover(self, T):
        return O.Prover(self.pc, T)

    def verifier(self, T):
        return O.Verifier(T)

    lc = staticmethod(O.LC.of)
    one = staticmethod(O.One)

    def prove(self, prover, seed):
        n = prover.num_multipliers()
        return prover.prove(O.BulletproofGens(O.round_pow2(n) if n else 1), seed).to_bytes()

    def verify(self, verifier, proof, seed):
        n = verifier.num_multipliers()
        try:
            pr = O.R1CSProof.from_bytes(proof)
        except O.FormatError:
            return "format"
        try:
            verifier.verify(pr, self.pc, O.BulletproofGens(O.round_pow2(n) if n else 1), seed)
            return True
        except Exception:
            return False


class GpuBackend:
    name = "gpu"

    def __init__(self, ctx):
        import bulletproof_gadgets_b200 as bpg
        self.bpg, self.ctx = bpg, ctx

    def transcript(self, label):
        return self.bpg.Transcript(label)

    def prover(self, T):
        return self.bpg.Prover(self.ctx, T)

    def verifier(self, T):
        return self.bpg.Verifier(self.ctx, T)

    def lc(self, x):
        return self.bpg.LinearCombination.of(x)

    def one(self):
        return self.bpg.Variable.One()

    def prove(self, prover, seed):
        return prover.prove(seed)

    def verify(self, verifier, proof, seed):
        try:
            return verifier.verify(proof, seed)
        except self.bpg.BpgError as e:
            if e.code == -1:
                return "format"
            raise


# ---------------------------------------------------------------------------------------- circuits
# each circuit: build(be, cs, committed_vars, witness_or_None)


def range_proof(be, cs, v_lc, value, n_bits):
    """Mirror of /root/reference/src/utils.rs:5-35 (range_proof)."""
    exp2 = 1
    acc = be.lc(v_lc)
    for i in range(n_bits):
        assign = None
        if value is not None:
            bit = (value >> i) & 1
            assign = ((1 - bit) % L, bit)
        a, b, o = cs.allocate_multiplier(assign)
        cs.constrain(be.lc(o))
        cs.constrain(be.lc(a) + be.lc(b) - be.lc(1))
        acc = acc - be.lc(b).scale(exp2)
        exp2 = exp2 * 2 % L
    cs.constrain(acc)


class Circuit:
    """A statement: committed values + a builder.  `values` are the committed scalars (prover)."""

    def __init__(self, label, values, builder, blind_tag=b"blind"):
        self.label, self.values, self.builder, self.blind_tag = label, values, builder, blind_tag

    def prove(self, be, seed=b"\x07" * 32):
        T = be.transcript(self.label)
        p = be.prover(T)
        coms, vars_ = [], []
        for i, v in enumerate(self.values):
            V, var = p.commit(v, det_scalar(self.blind_tag, i))
            coms.append(V)
            vars_.append(var)
        self.builder(be, p, vars_, self.values)
        return be.prove(p, seed), coms

    def verify(self, be, proof, coms, seed=b"\x09" * 32, label=None):
        T = be.transcript(label or self.label)
        v = be.verifier(T)
        vars_ = [v.commit(c) for c in coms]
        self.builder(be, v, vars_, None)
        return be.verify(v, proof, seed)


def c_empty():
    """n = 0 multipliers (EQUALS-only statements; /root/reference/src/equality/equality_gadget.rs:68)."""
    def build(be, cs, vars_, vals):
        cs.constrain(be.lc(vars_[0]) - be.lc(vars_[1]))
    return Circuit(b"empty", [1234567, 1234567], build)


def c_mul3():
    """3 multipliers -> padded to 4 (non power of two n)."""
    def build(be, cs, vars_, vals):
        x = vals[0] if vals else None
        l, r, o = cs.multiply(be.lc(vars_[0]), be.lc(vars_[0]))            # x^2
        l2, r2, o2 = cs.multiply(be.lc(o), be.lc(vars_[0]))                 # x^3
        a = cs.allocate_multiplier(((x * 5) % L, 7) if vals else None)
        cs.constrain(be.lc(a[0]) - be.lc(vars_[0]).scale(5))
        cs.constrain(be.lc(a[1]) - be.lc(7))
        cs.constrain(be.lc(o2) + be.lc(a[2]) - be.lc(vars_[1]))             # x^3 + 35x = y
    x = det_scalar(b"mul3")
    return Circuit(b"mul3", [x, (pow(x, 3, L) + 35 * x) % L], build)


def c_range(n_bits=8, count=1, tag=b"range"):
    """`count` range proofs of n_bits each: n = count*n_bits multipliers, 0/1-valued a_L/a_R."""
    def build(be, cs, vars_, vals):
        for i, var in enumerate(vars_):
            range_proof(be, cs, be.lc(var), vals[i] if vals else None, n_bits)
    vals = [det_scalar(tag, i, bits=n_bits) for i in range(count)]
    return Circuit(b"range-%d-%d" % (n_bits, count), vals, build)


def c_unreduced():
    """Witness / constant of 32 x 0xff with the top bit cleared: a Scalar::from_bits value >= l
    (/root/reference/src/inequality/inequality_gadget.rs:264-302)."""
    big = (1 << 255) - 1
    def build(be, cs, vars_, vals):
        a = cs.allocate_multiplier((big, 3) if vals else None)
        cs.constrain(be.lc(a[0]) - be.lc(vars_[0]))
        cs.constrain(be.lc(a[2]) - be.lc(big).scale(3))
    return Circuit(b"unreduced", [big], build)


def c_chain(n=37, tag=b"chain"):
    """x_{k+1} = x_k^2 + c_k chain through `multiply`: uniformly random a_L/a_R/a_O (MiMC-like)."""
    consts = [det_scalar(tag + b"-c", k) for k in range(n)]
    x0 = det_scalar(tag)
    def build(be, cs, vars_, vals):
        cur = be.lc(vars_[0])
        for k in range(n):
            _, _, o = cs.multiply(cur, cur)
            cur = be.lc(o) + be.lc(consts[k])
        cs.constrain(cur - be.lc(vars_[1]))
    x = x0
    for k in range(n):
        x = (x * x + consts[k]) % L
    return Circuit(b"chain-%d" % n, [x0, x], build)
