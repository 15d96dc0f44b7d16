"""Synthetic statements of the BASELINE.json configs, flattened to the arrays the C ABI bulk loaders
take (bpg_prover_load_cs / bpg_verifier_load_cs).  Pure host-side data preparation (numpy); the
circuits are what the reference's gadgets emit:

  bounds_check  BOUND v in [min, max], two n-bit range proofs per statement
                /root/reference/src/bounds_check/bounds_check_gadget.rs:13-48, /root/reference/src/utils.rs:5-35
"""
import numpy as np

L_ORDER = 2**252 + 27742317777372353535851937790883648493
COMMITTED, MUL_LEFT, MUL_RIGHT, MUL_OUT, ONE = 0, 1, 2, 3, 4


def _tag(kind, idx):
    return (kind << 29) | idx


class FlatStatement:
    """committed values + blindings + multiplier assignments + CSR constraints."""

    def __init__(self, label, v, vbl, aL, aR, row_start, term_var, term_coef, n, q):
        self.label = label
        self.v, self.vbl = v, vbl                      # lists of ints
        self.aL, self.aR = aL, aR                      # bytes, n*32
        self.row_start, self.term_var = row_start, term_var
        self.term_coef = term_coef                     # bytes, nnz*32
        self.n, self.q = n, q
        self.m = len(v)
        # packed forms for the batch entry points (k*32 bytes, little-endian)
        self.v_bytes = b"".join(int(x).to_bytes(32, "little") for x in v)
        self.vbl_bytes = b"".join(int(x).to_bytes(32, "little") for x in vbl)

    def pin(self, bpg):
        """Moves the bulk arrays into page-locked host memory (bpg_host_alloc) so that the C ABI's uploads are
        asynchronous DMA; returns self."""
        import numpy as np
        self.aL = bpg.pinned_copy(self.aL, np.uint8)
        self.aR = bpg.pinned_copy(self.aR, np.uint8)
        self.row_start = bpg.pinned_copy(self.row_start)
        self.term_var = bpg.pinned_copy(self.term_var)
        self.term_coef = bpg.pinned_copy(self.term_coef, np.uint8)
        return self

    @property
    def nnz(self):
        return int(self.row_start[-1])


def _scalar_bytes(x):
    return int(x % L_ORDER).to_bytes(32, "little")


def bounds_check_statement(count=1024, max_bytes=8, seed=20261018, label=b"bench-bound"):
    """`count` x  BOUND W_i I_min I_max  with I_min = 0, I_max = 2^(8*max_bytes)-1.
    Committed variables in the reference's order: the `count` witnesses W_i first (.wtns replay,
    /root/reference/src/lalrpop/assignment_parser.rs), then per statement the derived a = v - min, b = max - v
    (/root/reference/src/gadget.rs:28-36).  Equal, array for array, to the library front end's flattening of
    bounds_check_text(count) -- tests/test_host_abi.py::test_bench_statement_is_the_reference_gadget."""
    n_bits = 8 * max_bytes
    vmin, vmax = 0, (1 << n_bits) - 1
    rng = np.random.default_rng(seed)
    vals = [int.from_bytes(rng.integers(0, 256, size=max_bytes, dtype=np.uint8).tobytes(), "big") for _ in range(count)]
    v, vbl = [], []
    brng = np.random.default_rng(seed + 1)
    n = count * 2 * n_bits
    aL = np.zeros((n, 32), dtype=np.uint8)
    aR = np.zeros((n, 32), dtype=np.uint8)
    row_start = [0]
    tvar, tcoef = [], []
    one_b, minus_one_b = _scalar_bytes(1), _scalar_bytes(-1)
    pow2_neg = [_scalar_bytes(-(1 << i)) for i in range(n_bits)]
    mult = 0
    def blind():
        return int.from_bytes(brng.integers(0, 256, size=32, dtype=np.uint8).tobytes(), "little") % L_ORDER

    for val in vals:
        v.append(val % L_ORDER)
        vbl.append(blind())
    for s, val in enumerate(vals):
        a, b = val - vmin, vmax - val
        for x in (a, b):
            v.append(x % L_ORDER)
            vbl.append(blind())
        va, vb = _tag(COMMITTED, count + 2 * s), _tag(COMMITTED, count + 2 * s + 1)
        # (a + b) - (max - min) = 0
        tvar += [va, vb, _tag(ONE, 0)]
        tcoef += [one_b, one_b, _scalar_bytes(-(vmax - vmin))]
        row_start.append(len(tvar))
        for var, x in ((va, a), (vb, b)):
            acc_vars, acc_coef = [var], [one_b]
            for i in range(n_bits):
                bit = (x >> i) & 1
                aL[mult, 0] = 1 - bit
                aR[mult, 0] = bit
                tvar.append(_tag(MUL_OUT, mult))          # o = 0
                tcoef.append(one_b)
                row_start.append(len(tvar))
                tvar += [_tag(MUL_LEFT, mult), _tag(MUL_RIGHT, mult), _tag(ONE, 0)]   # a + (b - 1) = 0
                tcoef += [one_b, one_b, minus_one_b]
                row_start.append(len(tvar))
                acc_vars.append(_tag(MUL_RIGHT, mult))
                acc_coef.append(pow2_neg[i])
                mult += 1
            tvar += acc_vars                               # x - sum b_i 2^i = 0
            tcoef += acc_coef
            row_start.append(len(tvar))
    assert mult == n
    return FlatStatement(label, v, vbl, aL.tobytes(), aR.tobytes(), np.asarray(row_start, dtype=np.uint32),
                         np.asarray(tvar, dtype=np.uint32), b"".join(tcoef), n, len(row_start) - 1)


def _hex(b):
    return "0x" + bytes(b).hex()


def bounds_check_text(count=1024, max_bytes=8, seed=20261018):
    """The same workload as bounds_check_statement in the reference's text formats: `count` lines  BOUND W<i> I0 I1
    (/root/reference/src/bounds_check/bounds_check_gadget.rs:13-64).  Returns (gadgets, inst, wtns)."""
    rng = np.random.default_rng(seed)
    vals = [rng.integers(0, 256, size=max_bytes, dtype=np.uint8).tobytes() for _ in range(count)]
    inst = "I0 = 0x00\nI1 = 0x%s\n" % ("ff" * max_bytes)
    wtns = "".join("W%d = %s\n" % (i, _hex(v)) for i, v in enumerate(vals))
    gadgets = "".join("BOUND W%d I0 I1\n" % i for i in range(count))
    return gadgets, inst, wtns


def merkle_text(depth=32, seed=20261018, witness_siblings=False, hashers=None):
    """BASELINE config 3: `MERKLE I0 (((..(W0 I1) I2)..) I<depth>)` -- membership of leaf W0 under root I0 with MiMC
    (/root/reference/src/merkle_tree/merkle_tree_gadget.rs:39-114, /root/reference/src/prove.rs:289-321).
    Leaves are 4..32 random bytes; the root is computed with the library's mimc_hash / sponge.  With
    witness_siblings the siblings are W1..W<depth> (each hashed in-circuit).  `hashers` = (mimc_hash, mimc_sponge)
    callables returning ints (default: the library's host routines).  Returns (gadgets, inst, wtns) text."""
    if hashers is None:
        from . import api
        hashers = (api.mimc_hash, api.mimc_sponge)
    mimc_hash, mimc_sponge = hashers
    rng = np.random.default_rng(seed)
    leaves = [rng.integers(0, 256, size=int(rng.integers(4, 33)), dtype=np.uint8).tobytes() for _ in range(depth + 1)]
    leaves = [b if b.strip(b"\0") else b"\x01" + b[1:] for b in leaves]
    node = mimc_hash(leaves[0])
    for k in range(1, depth + 1):
        node = mimc_sponge([node, mimc_hash(leaves[k])])
    root = node.to_bytes(32, "big")
    sib = "W" if witness_siblings else "I"
    tree = "(W0 %s1)" % sib
    for k in range(2, depth + 1):
        tree = "(%s %s%d)" % (tree, sib, k)
    gadgets = "MERKLE I0 " + tree
    inst = ["I0 = " + _hex(root)]
    wtns = ["W0 = " + _hex(leaves[0])]
    for k in range(1, depth + 1):
        (wtns if witness_siblings else inst).append("%s%d = %s" % (sib, k, _hex(leaves[k])))
    return gadgets, "\n".join(inst), "\n".join(wtns)


def batch_texts(count=4096, seed=4096):
    """BASELINE config 4: independent proofs, alternating `LESS_THAN W0 W1` (15-byte values, W0 < W1; n = 379) and
    `SET_MEMBER W0 I0..I15` (member at a random index; n = 32).  Returns a list of (gadgets, inst, wtns)."""
    out = []
    for i in range(count):
        rng = np.random.default_rng(seed + i)
        if i % 2 == 0:
            a = int.from_bytes(rng.integers(0, 256, size=15, dtype=np.uint8).tobytes(), "big") >> 1
            b = a + 1 + (int.from_bytes(rng.integers(0, 256, size=14, dtype=np.uint8).tobytes(), "big"))
            out.append(("LESS_THAN W0 W1", "", "W0 = 0x%030x\nW1 = 0x%030x" % (a, b)))
        else:
            elems = [rng.integers(0, 256, size=15, dtype=np.uint8).tobytes() for _ in range(16)]
            member = elems[int(rng.integers(0, 16))]
            inst = "\n".join("I%d = %s" % (k, _hex(e)) for k, e in enumerate(elems))
            out.append(("SET_MEMBER W0 " + " ".join("I%d" % k for k in range(16)), inst, "W0 = " + _hex(member)))
    return out


def prove_statement(bpg, ctx, st, seed=b"\x07" * 32):
    """Drives one statement through the C ABI prover: returns (proof bytes, commitments)."""
    T = bpg.Transcript(st.label)
    p = bpg.Prover(ctx, T)
    packed = p.commit_batch_packed(st.v_bytes, st.vbl_bytes)
    coms = [packed[32 * i: 32 * i + 32] for i in range(st.m)]
    p.load_cs(st.aL, st.aR, st.row_start, st.term_var, st.term_coef, st.q)
    return p.prove(seed), coms


def verify_statement(bpg, ctx, st, proof, coms, seed=b"\x09" * 32):
    T = bpg.Transcript(st.label)
    vf = bpg.Verifier(ctx, T)
    vf.commit_batch(coms)
    vf.load_cs(st.n, st.row_start, st.term_var, st.term_coef, st.q)
    return vf.verify(proof, seed)
