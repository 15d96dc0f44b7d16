"""ORACLE (test infrastructure, never shipped) -- the statement front end of the reference.

Restates, in plain python, what sits between the five file formats and dalek's constraint system:

  conversions            /root/reference/src/conversions.rs:6-76
  commitments / .coms    /root/reference/src/commitments.rs:9-48, /root/reference/src/lalrpop/assignment_parser.rs:15-217
  grammars               /root/reference/src/lalrpop/gadget_grammar.lalrpop:6-85, var_grammar.lalrpop:3-29, ast.rs:3-86
  cs buffers             /root/reference/src/cs_buffer.rs:6-199
  prove() / verify()     /root/reference/src/prove.rs:37-514, /root/reference/src/verify.rs:36-415
  Gadget::setup          /root/reference/src/gadget.rs:19-39
  range_proof            /root/reference/src/utils.rs:5-35
  gadgets                bounds_check_gadget.rs:13-64, equality_gadget.rs:10-40, inequality_gadget.rs:11-114,
                         less_than_gadget.rs:15-83, set_membership_gadget.rs:12-132, mimc_hash_gadget.rs:7-151,
                         mimc.rs:7-97, merkle_tree_gadget.rs:39-114, or_conjunction.rs:4-67

The output is the FLAT statement the real prover / verifier end up holding after `assign_buffer`
(committed values in commit order, multiplier assignments, constraints in order with the two implicit
constraints of every `multiply`), which any engine (python oracle, C oracle, CUDA library) can prove.

Scalars are python ints holding dalek's RAW 32 bytes: `Scalar::from_bits` values may be >= l and stay
unreduced until arithmetic touches them; `==` compares raw bytes; +, -, *, invert return canonical values
(sub reproduces dalek's Scalar52::sub wrap for differences below -l).

Pinned by the reference's own known answers: MiMC images (mimc.rs:104-143), byte-order KATs
(conversions.rs:114-150), and the hash images / Merkle roots embedded in the fixtures
(tests/test_frontend_oracle.py).  Randomness (commitment blindings) is injected by the caller.
"""
import re

from .merlin import L
from .mimc_consts import ROUND_CONSTANTS_769

COMMITTED, MUL_LEFT, MUL_RIGHT, MUL_OUT, ONE = 0, 1, 2, 3, 4
MASK255 = (1 << 255) - 1
ONE_VAR = (ONE, 0)


class FrontendPanic(Exception):
    """Where the reference panics (expect/unwrap/assert): malformed input, missing variables, ..."""


# ---------------------------------------------------------------------------------- dalek Scalar semantics
def sc_add(a, b):
    return (a + b) % L


def sc_sub(a, b):
    """dalek 3.2 `&Scalar - &Scalar`: Scalar52::sub adds l once on underflow, then reduces."""
    d = a - b
    if d < 0:
        d += L
        if d < 0:
            d += 1 << 260
    return d % L


def sc_mul(a, b):
    return a * b % L


def sc_neg(a):
    return (-(a % L)) % L


def sc_invert(a):
    return pow(a % L, L - 2, L)


def sc_bytes(a):
    return int(a).to_bytes(32, "little")


# ---------------------------------------------------------------------------------- conversions.rs
def le_to_scalars(b):
    b = bytes(b)
    if len(b) % 32:
        b += bytes(32 - len(b) % 32)
    return [int.from_bytes(b[i: i + 32], "little") & MASK255 for i in range(0, len(b), 32)]


def be_to_scalars(b):
    return le_to_scalars(bytes(b)[::-1])


def le_to_scalar(b):
    b = bytes(b)
    if len(b) > 32:
        raise FrontendPanic("the given vector is longer than 32 bytes")
    if len(b) % 32:
        b += bytes(32 - len(b) % 32)
    if len(b) < 32:
        raise FrontendPanic("empty byte string")  # slice [0..32] out of range in the reference
    return int.from_bytes(b[:32], "little") & MASK255


def be_to_scalar(b):
    return le_to_scalar(bytes(b)[::-1])


def scalar_to_be(s):
    return sc_bytes(s)[::-1]


# ---------------------------------------------------------------------------------- mimc.rs
ROUNDS = 486


def _mimc_encrypt(p, k):
    state = p
    for i in range(ROUNDS):
        tmp = sc_add(state, sc_add(k, ROUND_CONSTANTS_769[i]))
        state = sc_mul(sc_mul(tmp, tmp), tmp)
    return sc_add(state, k)


def mimc_sponge(preimage):
    state = 0
    for x in preimage:
        state = sc_add(state, x)
        state = _mimc_encrypt(state, 0)
    return state


def _pkcs7_block(last_block):
    """(padded block, was_padded): mimc.rs:77-97 / mimc_hash_gadget.rs:15-37."""
    le = sc_bytes(last_block).rstrip(b"\x00")
    if len(le) < 32:
        pad = 32 - len(le)
        return le_to_scalar(le + bytes([pad]) * pad), True
    return le_to_scalar(bytes([32]) * 32), False


def mimc_hash(preimage_bytes):
    pre = be_to_scalars(preimage_bytes)
    if not pre:
        raise FrontendPanic("mimc_hash of an empty preimage")
    padded, replaced = _pkcs7_block(pre[-1])
    if replaced:
        pre.pop()
    pre.append(padded)
    return mimc_sponge(pre)


# ---------------------------------------------------------------------------------- linear combinations
class LC:
    """dalek LinearCombination: a plain term list; +/- concatenate, nothing is merged."""

    __slots__ = ("terms",)

    def __init__(self, terms=()):
        self.terms = list(terms)

    @staticmethod
    def var(v):
        return LC([(v, 1)])

    @staticmethod
    def const(s):
        return LC([(ONE_VAR, s)])  # Scalar -> LC keeps the raw scalar

    def __add__(self, o):
        return LC(self.terms + o.terms)

    def __sub__(self, o):
        return LC(self.terms + [(v, sc_neg(c)) for v, c in o.terms])

    def scale(self, s):
        return LC([(v, sc_mul(c, s)) for v, c in self.terms])


# ---------------------------------------------------------------------------------- cs_buffer.rs
class Buffer:
    """ProverBuffer / VerifierBuffer: records operations; its throw-away inner prover only matters through the
    multiplier counter that numbers the variables handed back to the gadgets."""

    def __init__(self, proving):
        self.proving = proving
        self.ops = []
        self.cache = []
        self.n_mult = 0

    def _alloc(self):
        i = self.n_mult
        self.n_mult += 1
        return (MUL_LEFT, i), (MUL_RIGHT, i), (MUL_OUT, i)

    def multiply(self, left, right):
        self.ops.append(("mul", left, right))
        return self._alloc()

    def allocate_multiplier(self, assignment):
        if self.proving:
            if assignment is None:
                raise FrontendPanic("MissingAssignment")
            self.ops.append(("alloc", assignment))
        else:
            self.ops.append(("alloc", None))
        return self._alloc()

    def constrain(self, lc):
        self.ops.append(("con", lc))

    def commit_drvd(self, scalars):
        self.ops.append(("commit", list(scalars)))

    def initialize_from(self, initialization):
        for ops in initialization:
            for op in ops:
                if op[0] in ("mul", "alloc"):
                    self.n_mult += 1

    def rewind(self):
        self.cache.append(self.ops)
        self.ops = []


# ---------------------------------------------------------------------------------- gadgets (assemble / preprocess)
def range_proof(cs, x, n, x_assignment):
    """utils.rs:5-35.  x_assignment: raw scalar or None (verifier)."""
    exp_2 = 1
    xb = None if x_assignment is None else sc_bytes(x_assignment)
    for i in range(n):
        assign = None
        if xb is not None:
            bit = (xb[i // 8] >> (i % 8)) & 1
            assign = (1 - bit, bit)
        a, b, o = cs.allocate_multiplier(assign)
        cs.constrain(LC.var(o))
        cs.constrain(LC.var(a) + (LC.var(b) - LC.const(1)))
        x = x - LC.var(b).scale(exp_2)
        exp_2 = sc_add(exp_2, exp_2)
    cs.constrain(x)


class BoundsCheck:
    def __init__(self, min_bytes, max_bytes):
        self.n = (len(max_bytes) * 8) & 0xFF  # `as u8`
        self.min, self.max = be_to_scalar(min_bytes), be_to_scalar(max_bytes)

    def preprocess(self, witnesses):
        v = witnesses[0]
        return [sc_sub(v, self.min), sc_sub(self.max, v)]

    def assemble(self, cs, _witnesses, derived):
        (a_val, a), (b_val, b) = derived[0], derived[1]
        a_lc, b_lc = LC.var(a), LC.var(b)
        cs.constrain((a_lc + b_lc) - LC.const(sc_sub(self.max, self.min)))
        range_proof(cs, a_lc, self.n, a_val)
        range_proof(cs, b_lc, self.n, b_val)


class MimcHash256:
    def __init__(self, image_lc=None):
        self.image = image_lc if image_lc is not None else LC.const(0)

    def preprocess(self, witnesses):
        last = witnesses[-1]
        padded, replaced = _pkcs7_block(last)
        return [padded, sc_sub(padded, last)] if replaced else [padded]

    def assemble(self, cs, witnesses, derived):
        coms = list(witnesses)
        padded_block = derived[0][1]
        if len(derived) == 2:
            padding = derived[1][1]
            last = LC.var(coms.pop())
            cs.constrain((last + LC.var(padding)) - LC.var(padded_block))
        coms.append(padded_block)
        h = self.mimc_sponge(cs, [LC.var(v) for v in coms])
        cs.constrain(h - self.image)

    def mimc_sponge(self, cs, preimage):
        key_zero = LC.const(0)
        state = LC.const(0)
        for variable in preimage:
            state = state + variable
            state = self._encrypt(cs, state, key_zero)
        return state

    @staticmethod
    def _encrypt(cs, p, k):
        for i in range(ROUNDS):
            t = (p + k) + LC.const(ROUND_CONSTANTS_769[i])
            x_k_ci, _, sqr = cs.multiply(t, t)
            _, _, cube = cs.multiply(LC.var(sqr), LC.var(x_k_ci))
            p = LC.var(cube)
        return p + k


class MerkleTree256:
    """pattern: 'W', 'I' or (left, right)."""

    def __init__(self, root_lc, instance_lcs, witness_lcs, pattern):
        self.root, self.i_vals, self.w_vals, self.pattern = root_lc, list(instance_lcs), list(witness_lcs), pattern
        self.gadget = MimcHash256()

    def preprocess(self, _):
        return []

    def assemble(self, cs, _w, _d):
        w, i = list(self.w_vals), list(self.i_vals)
        h = self._parse(cs, w, i, self.pattern)
        cs.constrain(h - self.root)

    def _next(self, vals):
        if not vals:
            raise FrontendPanic("too few variables provided to satisfy the given pattern")
        return vals.pop(0)

    def _parse(self, cs, w, i, pat):
        if pat == "W":
            pre = [self._next(w)]
        elif pat == "I":
            pre = [self._next(i)]
        else:
            pre = []
            for side in pat:  # left first, then right (evaluation order of the vec![..] in the reference)
                if side == "W":
                    pre.append(self._next(w))
                elif side == "I":
                    pre.append(self._next(i))
                else:
                    pre.append(self._parse(cs, w, i, side))
        return self.gadget.mimc_sponge(cs, pre)


class Equality:
    def __init__(self, right_lcs):
        self.right = right_lcs

    def preprocess(self, _):
        return []

    def assemble(self, cs, left_vars, _d):
        if len(self.right) != len(left_vars):
            return cs.constrain(LC.const(1))
        for r, l in zip(self.right, left_vars):
            cs.constrain(r - LC.var(l))


class Inequality:
    def __init__(self, right_lcs, right_assignment):
        self.right, self.right_assignment = right_lcs, right_assignment

    @staticmethod
    def compare(left, right):
        lb, rb = sc_bytes(left), sc_bytes(right)
        for i in range(31, -1, -1):
            if lb[i] > rb[i]:
                return True
            if lb[i] < rb[i]:
                return False
        return True

    def preprocess(self, left_hand):
        if self.right_assignment is None:
            raise FrontendPanic("missing right hand assignment")
        out, total = [], 0
        for i, left in enumerate(left_hand):
            right = self.right_assignment[i] if i < len(self.right_assignment) else 0
            delta = sc_sub(left, right) if self.compare(left, right) else sc_sub(right, left)
            out.append(delta)
            if delta == 0:
                out.append(0)
            else:
                inv = sc_invert(delta)
                out.append(inv)
                total = sc_add(total, sc_mul(delta, inv))
        out.append(sc_invert(total))
        return out

    def assemble(self, cs, left_vars, derived):
        if len(self.right) != len(left_vars):
            return cs.constrain(LC.const(0))
        total = LC.const(0)
        for i, lv in enumerate(left_vars):
            right_lc, left_lc = self.right[i], LC.var(lv)
            delta, delta_inv = LC.var(derived[2 * i][1]), LC.var(derived[2 * i + 1][1])
            left = (left_lc - right_lc) - delta
            right = (right_lc - left_lc) - delta
            _, _, zero = cs.multiply(left, right)
            cs.constrain(LC.var(zero))
            _, _, zero_or_one = cs.multiply(delta, delta_inv)
            total = total + LC.var(zero_or_one)
        _, _, one = cs.multiply(total, LC.var(derived[-1][1]))
        cs.constrain(LC.const(1) - LC.var(one))


class LessThan:
    def __init__(self, left_lc, left_val, right_lc, right_val):
        self.left, self.left_val, self.right, self.right_val = left_lc, left_val, right_lc, right_val

    def preprocess(self, _):
        if self.left_val is None or self.right_val is None:
            raise FrontendPanic("missing right hand assignment")
        delta = sc_sub(self.right_val, self.left_val)
        return [delta, 0 if delta == 0 else sc_invert(delta)]

    def assemble(self, cs, _w, derived):
        delta_val, delta = derived[0]
        delta_lc, delta_inv_lc = LC.var(delta), LC.var(derived[1][1])
        n = 126
        range_proof(cs, self.left, n, self.left_val)
        range_proof(cs, self.right, n, self.right_val)
        range_proof(cs, delta_lc, n, delta_val)
        _, _, one = cs.multiply(delta_lc, delta_inv_lc)
        cs.constrain(LC.const(1) - LC.var(one))
        cs.constrain((self.right - self.left) - delta_lc)


class SetMembership:
    def __init__(self, value_lc, value_val, instance_lcs, instance_vals):
        self.value, self.value_val, self.instance_lcs, self.instance_vals = value_lc, value_val, instance_lcs, instance_vals

    def preprocess(self, witnesses):
        if self.value_val is None or self.instance_vals is None:
            raise FrontendPanic("missing value assignment")
        return [1 if e == self.value_val else 0 for e in list(witnesses) + list(self.instance_vals)]

    def assemble(self, cs, witnesses, derived):
        one_hot = []
        for _, bit in derived:
            bit_lc = LC.var(bit)
            _, _, zero = cs.multiply(LC.const(1) - bit_lc, bit_lc)
            cs.constrain(LC.var(zero))
            one_hot.append(bit_lc)
        total = LC.const(0)
        for b in one_hot:
            total = total + b
        cs.constrain(LC.const(1) - total)
        elems = [LC.var(w) for w in witnesses] + list(self.instance_lcs)
        if len(one_hot) != len(elems):
            return cs.constrain(LC.const(1))
        prod = LC.const(0)
        for b, e in zip(one_hot, elems):
            _, _, p = cs.multiply(b, e)
            prod = prod + LC.var(p)
        cs.constrain(self.value - prod)


def or_combine(main, inner):
    """or_conjunction.rs:4-38."""
    per_clause = []
    for ops in inner.cache:
        cons = []
        for op in ops:
            if op[0] == "mul":
                main.multiply(op[1], op[2])
            elif op[0] == "alloc":
                main.allocate_multiplier(op[1])
            elif op[0] == "con":
                cons.append(op[1])
        per_clause.append(cons)
    if not per_clause:
        return
    combos = [[c] for c in per_clause[0]]
    for lst in per_clause[1:]:
        combos = [xs + [y] for xs in combos for y in lst]
    for combo in combos:
        prod = combo[0]
        for c in combo[1:]:
            _, _, p = main.multiply(prod, c)
            prod = LC.var(p)
        main.constrain(prod)


# ---------------------------------------------------------------------------------- grammars
_OPS = {"OR", "HASH", "]", "BOUND", "[", "MERKLE", "}", "EQUALS", "{", "UNEQUAL", "LESS_THAN", "SET_MEMBER"}
_VAR_LINE = {k: re.compile(r"^\s*(%s)\s*=\s*0[xX]([0-9a-fA-F]+)\s*$" % pat)
             for k, pat in (("I", r"I\d+"), ("W", r"W\d+"), ("C", r"[C|D]\d+-\d+(?:-\d+)?"))}


def gadget_op(line):
    tok = line.split()
    op = tok[0] if tok else ""
    if op not in _OPS:
        raise FrontendPanic("unknown gadget: %s" % op)
    return op


def parse_var_line(kind, line):
    m = _VAR_LINE[kind].match(line)
    if not m:
        raise FrontendPanic("unparsable %s line: %r" % (kind, line))
    h = m.group(2)
    if len(h) % 2:
        raise FrontendPanic("odd number of hex digits: %r" % line)  # hex::decode(..).unwrap()
    return m.group(1), bytes.fromhex(h)


def _tokens(line):
    return re.findall(r"[()]|[A-Za-z_]+\d*|\S", line)


def _is(tok, kind):
    return re.fullmatch(kind + r"\d+", tok) is not None


def parse_two(line, op, allowed):
    """`OP a b` with the (kind, kind) pairs in `allowed`; returns tokens."""
    t = line.split()
    if len(t) != 3 or t[0] != op:
        raise FrontendPanic("cannot parse %r" % line)
    for ka, kb in allowed:
        if _is(t[1], ka) and _is(t[2], kb):
            return t[1], t[2]
    raise FrontendPanic("cannot parse %r" % line)


def parse_tree(line):
    """MERKLE root tree -> (root, instance_vars, witness_vars, pattern); variables in left-to-right order."""
    t = _tokens(line)
    if len(t) < 3 or t[0] != "MERKLE" or not (_is(t[1], "I") or _is(t[1], "W")):
        raise FrontendPanic("cannot parse %r" % line)
    pos = [2]
    inst, wtns = [], []

    def node():
        if pos[0] >= len(t) or t[pos[0]] != "(":
            raise FrontendPanic("cannot parse %r" % line)
        pos[0] += 1
        sides = []
        for _ in range(2):
            if pos[0] >= len(t):
                raise FrontendPanic("cannot parse %r" % line)
            tok = t[pos[0]]
            if tok == "(":
                sides.append(node())
            elif _is(tok, "W"):
                wtns.append(tok)
                sides.append("W")
                pos[0] += 1
            elif _is(tok, "I"):
                inst.append(tok)
                sides.append("I")
                pos[0] += 1
            else:
                raise FrontendPanic("cannot parse %r" % line)
        if pos[0] >= len(t) or t[pos[0]] != ")":
            raise FrontendPanic("cannot parse %r" % line)
        pos[0] += 1
        return tuple(sides)

    pat = node()
    if pos[0] != len(t):
        raise FrontendPanic("cannot parse %r" % line)
    return t[1], inst, wtns, pat


# ---------------------------------------------------------------------------------- flat statement
def _tag(var):
    return (var[0] << 29) | var[1]


class FlatStatement:
    """What the real Prover / Verifier hold after assign_buffer.  Same attribute names as
    bulletproof_gadgets_b200.workloads.FlatStatement (duck-typed by the engines)."""

    def __init__(self, label):
        self.label = label
        self.v, self.vbl = [], []
        self.V = []                       # verifier: compressed commitments
        self.com_names = []               # .coms line names, commit order
        self._aL, self._aR = [], []
        self.n = 0
        self._rows = [0]
        self._tvar, self._tcoef = [], bytearray()

    # -- real constraint system (dalek Prover / Verifier semantics)
    def _eval(self, lc):
        acc = 0
        for (k, i), c in lc.terms:
            if k == MUL_LEFT:
                val = self._aL[i]
            elif k == MUL_RIGHT:
                val = self._aR[i]
            elif k == MUL_OUT:
                val = sc_mul(self._aL[i], self._aR[i])
            elif k == COMMITTED:
                val = self.v[i]
            else:
                val = 1
            acc += c * val
        return acc % L

    def constrain(self, lc):
        for var, c in lc.terms:
            self._tvar.append(_tag(var))
            self._tcoef += sc_bytes(c)
        self._rows.append(len(self._tvar))

    def replay(self, ops, proving):
        for op in ops:
            if op[0] == "mul":
                i = self.n
                if proving:
                    self._aL.append(self._eval(op[1]))
                    self._aR.append(self._eval(op[2]))
                self.n += 1
                self.constrain(op[1] - LC.var((MUL_LEFT, i)))
                self.constrain(op[2] - LC.var((MUL_RIGHT, i)))
            elif op[0] == "alloc":
                if proving:
                    self._aL.append(op[1][0])
                    self._aR.append(op[1][1])
                self.n += 1
            elif op[0] == "con":
                self.constrain(op[1])

    def finish(self):
        import numpy as np
        self.m = len(self.v) if self.v else len(self.V)
        self.q = len(self._rows) - 1
        self.row_start = np.asarray(self._rows, dtype=np.uint32)
        self.term_var = np.asarray(self._tvar if self._tvar else [0], dtype=np.uint32)
        self.term_coef = bytes(self._tcoef) if self._tcoef else bytes(32)
        self.aL = b"".join(sc_bytes(x) for x in self._aL)
        self.aR = b"".join(sc_bytes(x) for x in self._aR)
        self.v_bytes = b"".join(sc_bytes(x) for x in self.v)
        self.vbl_bytes = b"".join(sc_bytes(x) for x in self.vbl)
        return self

    @property
    def nnz(self):
        return int(self.row_start[-1])

    def coms_text(self, coms):
        """.coms file for commitments `coms` (32-byte strings, commit order)."""
        return "".join("%s = 0x%s\n" % (nm, c.hex()) for nm, c in zip(self.com_names, coms))


# ---------------------------------------------------------------------------------- prove.rs
class _ProverSide:
    def __init__(self, name, blinding):
        self.st = FlatStatement(name.encode() if isinstance(name, str) else name)
        self.blinding = blinding
        self.instance = {}
        self.witness = {}   # name -> (scalars, vars, bytes)

    def commit(self, scalar, com_name):
        st = self.st
        var = (COMMITTED, len(st.v))
        st.vbl.append(self.blinding(len(st.v)))
        st.v.append(scalar)
        st.com_names.append(com_name)
        return var

    def setup(self, gadget, witnesses, index, subroutine, first=0):
        """Gadget::setup + parse_derived_witness naming: D<index>-<subroutine>-<k>."""
        derived = []
        for k, s in enumerate(gadget.preprocess(witnesses)):
            derived.append((s, self.commit(s, "D%d-%d-%d" % (index, subroutine, first + k))))
        return derived

    def get_witness(self, name, single=False):
        if name not in self.witness:
            raise FrontendPanic("missing witness var %s" % name)
        w = self.witness[name]
        if single and len(w[0]) != 1:
            raise FrontendPanic("witness var %s is longer than 32 bytes" % name)
        return w

    def get_instance(self, name, max32=False):
        if name not in self.instance:
            raise FrontendPanic("missing instance var %s" % name)
        b = self.instance[name]
        if max32 and len(b) > 32:
            raise FrontendPanic("instance var %s is longer than 32 bytes" % name)
        return b

    # hash_witness (prove.rs:142-172): image commitment is D<index>-<sub>-0, the HASH-derived ones follow
    def hash_witness(self, buf, name, index, subroutine):
        scalars, vars_, raw = self.get_witness(name)
        image = mimc_hash(raw)
        image_scalar = be_to_scalar(scalar_to_be(image))
        image_var = self.commit(image_scalar, "D%d-%d-0" % (index, subroutine))
        buf.commit_drvd([image_scalar])
        g = MimcHash256(LC.var(image_var))
        derived = self.setup(g, scalars, index, subroutine, first=1)
        buf.commit_drvd([s for s, _ in derived])
        g.assemble(buf, vars_, derived)
        return image_scalar, image_var

    def gadget_line(self, line, buf, index):
        op = gadget_op(line)
        if op == "BOUND":
            w, lo, hi = _parse_bound(line)
            scalars, vars_, _ = self.get_witness(w, single=True)
            g = BoundsCheck(self.get_instance(lo, True), self.get_instance(hi, True))
            derived = self.setup(g, scalars, index, 0)
            buf.commit_drvd([s for s, _ in derived])
            g.assemble(buf, vars_, derived)
        elif op == "HASH":
            img, pre = parse_two(line, "HASH", (("W", "W"), ("I", "W")))
            image = self._single_lc(img)
            scalars, vars_, _ = self.get_witness(pre)
            g = MimcHash256(image)
            derived = self.setup(g, scalars, index, 0)
            buf.commit_drvd([s for s, _ in derived])
            g.assemble(buf, vars_, derived)
        elif op == "MERKLE":
            root, inst, wtns, pat = parse_tree(line)
            root_lc = self._single_lc(root)
            inst_lcs = [LC.const(mimc_hash(self.get_instance(i))) for i in inst]
            w_lcs = [LC.var(self.hash_witness(buf, w, index, k)[1]) for k, w in enumerate(wtns)]
            MerkleTree256(root_lc, inst_lcs, w_lcs, pat).assemble(buf, [], [])
        elif op == "EQUALS":
            left, right = _normalise(parse_two(line, "EQUALS", (("W", "I"), ("I", "W"), ("W", "W"))))
            _, left_vars, _ = self.get_witness(left)
            Equality(self._multi(right)[1]).assemble(buf, left_vars, [])
        elif op == "LESS_THAN":
            l, r = parse_two(line, "LESS_THAN", (("W", "W"),))
            ls, lv, _ = self.get_witness(l, single=True)
            rs, rv, _ = self.get_witness(r, single=True)
            g = LessThan(LC.var(lv[0]), ls[0], LC.var(rv[0]), rs[0])
            derived = self.setup(g, [], index, 0)
            buf.commit_drvd([s for s, _ in derived])
            g.assemble(buf, [], derived)
        elif op == "UNEQUAL":
            left, right = _normalise(parse_two(line, "UNEQUAL", (("W", "I"), ("I", "W"), ("W", "W"))))
            lsc, lvars, _ = self.get_witness(left)
            rsc, rlcs = self._multi(right)
            g = Inequality(rlcs, rsc)
            derived = self.setup(g, lsc, index, 0)
            buf.commit_drvd([s for s, _ in derived])
            g.assemble(buf, lvars, derived)
        elif op == "SET_MEMBER":
            self._set_member(line, buf, index)

    def _single_lc(self, name):
        if name[0] == "W":
            return LC.var(self.get_witness(name, single=True)[1][0])
        return LC.const(be_to_scalar(self.get_instance(name, True)))

    def _multi(self, name):
        """(scalars, lcs) of a witness or instance variable, one per 32-byte limb."""
        if name[0] == "W":
            sc, vars_, _ = self.get_witness(name)
            return sc, [LC.var(v) for v in vars_]
        sc = be_to_scalars(self.get_instance(name))
        return sc, [LC.const(s) for s in sc]

    def _set_member(self, line, buf, index):
        t = line.split()
        if len(t) < 3 or not all(_is(x, "W") or _is(x, "I") for x in t[1:]):
            raise FrontendPanic("cannot parse %r" % line)
        member, elems = t[1], t[2:]
        m_sc, m_lcs = self._multi(member)
        if not m_sc:
            raise FrontendPanic("empty member")
        member_scalar, member_lc = m_sc[0], m_lcs[0]
        hashing = len(m_sc) > 1
        w_vars, w_sc, i_lcs, i_sc = [], [], [], []
        if not hashing:
            for e in elems:
                sc, lcs = self._multi(e)
                if len(sc) == 1:
                    if e[0] == "W":
                        w_sc.append(sc[0])
                        w_vars.append(self.get_witness(e)[1][0])
                    else:
                        i_sc.append(sc[0])
                        i_lcs.append(lcs[0])
                else:
                    hashing = True
        if hashing:
            sub = 1
            if member[0] == "W":
                member_scalar, mv = self.hash_witness(buf, member, index, sub)
                member_lc = LC.var(mv)
                sub += 1
            else:
                member_scalar = mimc_hash(self.get_instance(member))
                member_lc = LC.const(member_scalar)
            w_vars, w_sc, i_lcs, i_sc = [], [], [], []
            for e in elems:
                if e[0] == "W":
                    s, v = self.hash_witness(buf, e, index, sub)
                    sub += 1
                    w_vars.append(v)
                    w_sc.append(s)
                else:
                    s = mimc_hash(self.get_instance(e))
                    i_lcs.append(LC.const(s))
                    i_sc.append(s)
        g = SetMembership(member_lc, member_scalar, i_lcs, i_sc)
        derived = self.setup(g, w_sc, index, 0)
        buf.commit_drvd([s for s, _ in derived])
        g.assemble(buf, w_vars, derived)


def _parse_bound(line):
    t = line.split()
    if len(t) != 4 or t[0] != "BOUND" or not (_is(t[1], "W") and _is(t[2], "I") and _is(t[3], "I")):
        raise FrontendPanic("cannot parse %r" % line)
    return t[1], t[2], t[3]


def _normalise(pair):
    a, b = pair
    return (b, a) if a[0] == "I" else (a, b)  # grammar: (Witness, Instance|Witness)


def _run_lines(side, lines, top):
    """The `.gadgets` walk shared by prove.rs:62-70 / verify.rs:57-65 incl. OR blocks (prove.rs:184-220)."""
    it = iter(enumerate(lines))
    peeked = []

    def nxt():
        if peeked:
            return peeked.pop()
        return next(it, None)

    def conjunction(parent, initialization):
        inner = Buffer(parent.proving)
        inner.initialize_from(initialization)
        item = nxt()
        if item is None:
            raise FrontendPanic("unexpected end of input")
        while item is not None:
            idx, line = item
            op = gadget_op(line)
            if op == "]":
                break
            if op == "}":
                inner.rewind()
            else:
                local = list(initialization) + [list(inner.ops)]
                if op == "OR":
                    conjunction(inner, local)
                side.gadget_line(line, inner, idx)
            item = nxt()
        for ops in inner.cache:  # add_commitments_to_parent (prover only; Commit ops never reach the real cs)
            for o in ops:
                if o[0] == "commit":
                    parent.commit_drvd(o[1])
        or_combine(parent, inner)

    item = nxt()
    while item is not None:
        idx, line = item
        op = gadget_op(line)
        if op == "OR":
            conjunction(top, [list(top.ops)])
        side.gadget_line(line, top, idx)
        item = nxt()


def compile_prover(name, instance, witness, gadgets, blinding):
    """prove() of prove.rs:37-82 up to (not including) Prover::prove.  `blinding(k)` = blinding factor of the k-th
    commitment (the reference draws Scalar::random(thread_rng()))."""
    side = _ProverSide(name, blinding)
    for line in instance.splitlines():
        nm, b = parse_var_line("I", line)
        side.instance[nm] = b
    for line in witness.splitlines():
        nm, b = parse_var_line("W", line)
        scalars = be_to_scalars(b)
        vars_ = [side.commit(s, "C%s-%d" % (nm[1:], k)) for k, s in enumerate(scalars)]
        side.witness[nm] = (scalars, vars_, b)
    top = Buffer(True)
    _run_lines(side, gadgets.splitlines(), top)
    side.st.replay(top.ops, True)
    return side.st.finish()


# ---------------------------------------------------------------------------------- verify.rs
class _VerifierSide:
    def __init__(self, name):
        self.st = FlatStatement(name.encode() if isinstance(name, str) else name)
        self.instance = {}
        self.coms = {}

    def get_instance(self, name, max32=False):
        if name not in self.instance:
            raise FrontendPanic("missing instance var %s" % name)
        b = self.instance[name]
        if max32 and len(b) > 32:
            raise FrontendPanic("instance var %s is longer than 32 bytes" % name)
        return b

    def get_commitment(self, name, index):
        key = "C%s-%d" % (name[1:], index)
        if key not in self.coms:
            raise FrontendPanic("missing commitment %s" % key)
        return self.coms[key]

    def all_commitments(self, name):
        out, i = [], 0
        while "C%s-%d" % (name[1:], i) in self.coms:
            out.append(self.coms["C%s-%d" % (name[1:], i)])
            i += 1
        return out

    def derived(self, gadget, index, subroutine, optional=False):
        key = "D%d-%d-%d" % (gadget, subroutine, index)
        if key not in self.coms:
            if optional:
                return None
            raise FrontendPanic("missing commitment %s" % key)
        return self.coms[key]

    def _single_lc(self, name):
        if name[0] == "W":
            return LC.var(self.get_commitment(name, 0))
        return LC.const(be_to_scalar(self.get_instance(name, True)))

    def _multi_lcs(self, name):
        if name[0] == "W":
            return [LC.var(v) for v in self.all_commitments(name)]
        return [LC.const(s) for s in be_to_scalars(self.get_instance(name))]

    def hash_witness(self, buf, name, index, subroutine):
        pre = self.all_commitments(name)
        image = self.derived(index, 0, subroutine)
        d = [(None, self.derived(index, 1, subroutine))]
        d2 = self.derived(index, 2, subroutine, optional=True)
        if d2 is not None:
            d.append((None, d2))
        MimcHash256(LC.var(image)).assemble(buf, pre, d)
        return image

    def gadget_line(self, line, buf, index):
        op = gadget_op(line)
        if op == "BOUND":
            w, lo, hi = _parse_bound(line)
            var = self.get_commitment(w, 0)
            g = BoundsCheck(self.get_instance(lo, True), self.get_instance(hi, True))
            g.assemble(buf, [var], [(None, self.derived(index, 0, 0)), (None, self.derived(index, 1, 0))])
        elif op == "HASH":
            img, pre = parse_two(line, "HASH", (("W", "W"), ("I", "W")))
            image = self._single_lc(img)
            pre_vars = self.all_commitments(pre)
            d = [(None, self.derived(index, 0, 0))]
            d2 = self.derived(index, 1, 0, optional=True)
            if d2 is not None:
                d.append((None, d2))
            MimcHash256(image).assemble(buf, pre_vars, d)
        elif op == "MERKLE":
            root, inst, wtns, pat = parse_tree(line)
            root_lc = self._single_lc(root)
            inst_lcs = [LC.const(mimc_hash(self.get_instance(i))) for i in inst]
            w_lcs = [LC.var(self.hash_witness(buf, w, index, k)) for k, w in enumerate(wtns)]
            MerkleTree256(root_lc, inst_lcs, w_lcs, pat).assemble(buf, [], [])
        elif op == "EQUALS":
            left, right = _normalise(parse_two(line, "EQUALS", (("W", "I"), ("I", "W"), ("W", "W"))))
            Equality(self._multi_lcs(right)).assemble(buf, self.all_commitments(left), [])
        elif op == "LESS_THAN":
            l, r = parse_two(line, "LESS_THAN", (("W", "W"),))
            lv, rv = self.get_commitment(l, 0), self.get_commitment(r, 0)
            d = [(None, self.derived(index, 0, 0)), (None, self.derived(index, 1, 0))]
            LessThan(LC.var(lv), None, LC.var(rv), None).assemble(buf, [], d)
        elif op == "UNEQUAL":
            left, right = _normalise(parse_two(line, "UNEQUAL", (("W", "I"), ("I", "W"), ("W", "W"))))
            lvars = self.all_commitments(left)
            d = [(None, self.derived(index, i, 0)) for i in range(2 * len(lvars))]
            d.append((None, self.derived(index, 2 * len(lvars), 0)))
            Inequality(self._multi_lcs(right), None).assemble(buf, lvars, d)
        elif op == "SET_MEMBER":
            self._set_member(line, buf, index)

    def _set_member(self, line, buf, index):
        t = line.split()
        if len(t) < 3 or not all(_is(x, "W") or _is(x, "I") for x in t[1:]):
            raise FrontendPanic("cannot parse %r" % line)
        member, elems = t[1], t[2:]
        m_lcs = self._multi_lcs(member)
        if not m_lcs:
            raise FrontendPanic("index out of bounds")  # member_lcs[0]
        member_lc = m_lcs[0]
        hashing = False
        w_vars, i_lcs = [], []
        for e in elems:
            lcs = self._multi_lcs(e)
            if len(lcs) == 1:
                if e[0] == "W":
                    w_vars.append(self.get_commitment(e, 0))
                else:
                    i_lcs.append(lcs[0])
            else:
                hashing = True
        if len(m_lcs) > 1:
            hashing = True
        derived = [(None, self.derived(index, k, 0)) for k in range(len(elems))]
        if hashing:
            sub = 1
            if member[0] == "W":
                member_lc = LC.var(self.hash_witness(buf, member, index, sub))
                sub += 1
            else:
                member_lc = LC.const(mimc_hash(self.get_instance(member)))
            w_vars, i_lcs = [], []
            for e in elems:
                if e[0] == "W":
                    w_vars.append(self.hash_witness(buf, e, index, sub))
                    sub += 1
                else:
                    i_lcs.append(LC.const(mimc_hash(self.get_instance(e))))
        SetMembership(member_lc, None, i_lcs, None).assemble(buf, w_vars, derived)


def compile_verifier(name, instance, commitments, gadgets):
    """verify() of verify.rs:36-73 up to (not including) Verifier::verify; the proof bytes are checked by the engine."""
    side = _VerifierSide(name)
    for line in instance.splitlines():
        nm, b = parse_var_line("I", line)
        side.instance[nm] = b
    for line in commitments.splitlines():
        nm, b = parse_var_line("C", line)
        if len(b) != 32:
            raise FrontendPanic("commitment %s is not 32 bytes" % nm)  # CompressedRistretto::from_slice
        side.coms[nm] = (COMMITTED, len(side.st.V))
        side.st.V.append(b)
        side.st.com_names.append(nm)
    top = Buffer(False)
    _run_lines(side, gadgets.splitlines(), top)
    side.st.replay(top.ops, False)
    return side.st.finish()
