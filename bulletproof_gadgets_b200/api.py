"""Python mirror of the reference-facing API, bound to the C ABI (include/bpg.h).

Names follow dalek `bulletproofs::r1cs` as used by the reference's gadgets
(/root/reference/src/gadget.rs:7-60): Prover.commit / multiply / allocate_multiplier / constrain /
prove, Verifier.commit / ... / verify, LinearCombination, Variable, merlin Transcript.
Scalars are ints (or 32-byte little-endian `bytes`); compressed points are 32-byte `bytes`.
"""
import ctypes
from ctypes import byref, c_int, c_size_t, c_uint32, c_void_p

from . import _capi
from ._capi import BpgError, check, lib

L_ORDER = 2**252 + 27742317777372353535851937790883648493


def _sb(x):
    """Scalar -> 32 bytes LE.  ints are taken as given when < 2^255 (Scalar::from_bits semantics)."""
    if isinstance(x, (bytes, bytearray)):
        assert len(x) == 32
        return bytes(x)
    x = int(x)
    if x < 0:
        x %= L_ORDER
    return x.to_bytes(32, "little")


def _buf(x):
    """bytes / bytearray / numpy array -> something ctypes passes as a pointer (None for empty)."""
    if x is None:
        return None
    if isinstance(x, (bytes, bytearray)):
        return bytes(x) or None
    return x.ctypes.data if x.size else None          # numpy (e.g. pinned_empty)


def _nbytes(x):
    return len(x) if isinstance(x, (bytes, bytearray)) else x.nbytes


def _scalars(xs):
    """list of ints / 32-byte strings, or an already packed buffer -> (pointer-able, count)."""
    if isinstance(xs, (list, tuple)):
        b = b"".join(_sb(x) for x in xs)
        return b or None, len(xs)
    return _buf(xs), _nbytes(xs) // 32


def pinned_empty(nbytes):
    """numpy uint8 array over page-locked host memory (bpg_host_alloc): buffers handed to the bulk
    loaders from here are uploaded by asynchronous DMA."""
    import numpy as np
    import weakref
    nbytes = int(nbytes)
    p = lib().bpg_host_alloc(nbytes)
    if not p:
        raise BpgError(_capi.E_CUDA, (lib().bpg_last_error() or b"").decode())
    buf = (ctypes.c_uint8 * max(nbytes, 1)).from_address(p)
    weakref.finalize(buf, lib().bpg_host_free, p)   # `buf` is the base of every numpy view taken below
    arr = np.frombuffer(buf, dtype=np.uint8)[:nbytes]
    return arr


def pinned_copy(data, dtype=None):
    """Copies bytes / a numpy array into pinned memory; returns a numpy view of the same dtype."""
    import numpy as np
    src = np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray)) else np.ascontiguousarray(data)
    out = pinned_empty(src.nbytes)
    out[:] = src.view(np.uint8).reshape(-1)
    return out.view(dtype or src.dtype)


class Context:
    """One per GPU: device streams, cached generator tables, work buffers (bpg_ctx)."""

    def __init__(self, device=0, _h=None):
        self._h = c_void_p() if _h is None else _h
        if _h is None:
            check(lib().bpg_ctx_create(device, byref(self._h)))

    def shared(self):
        """Another context on the same GPU (own stream + work buffers, generator tables shared)."""
        h = c_void_p()
        check(lib().bpg_ctx_create_shared(self._h, byref(h)))
        return Context(_h=h)

    def close(self):
        if self._h:
            lib().bpg_ctx_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set(self, key, value):
        check(lib().bpg_ctx_set(self._h, key.encode(), int(value)))

    def get(self, key):
        return lib().bpg_ctx_get(self._h, key.encode())

    def gens_ensure(self, capacity):
        """BulletproofGens::new(capacity, 1) + PedersenGens::default()."""
        check(lib().bpg_gens_ensure(self._h, capacity))

    def gens_compressed(self, which, start=0, count=1):
        idx = {"G": 0, "H": 1, "B": 2, "B_blinding": 3}[which]
        buf = ctypes.create_string_buffer(32 * count)
        check(lib().bpg_gens_compressed(self._h, idx, start, count, buf))
        return [buf.raw[32 * i: 32 * i + 32] for i in range(count)]

    def msm_gens(self, sG=(), sH=(), sB=None, sBb=None):
        """sum sG_i G_i + sum sH_i H_i + sB B + sBb B_blinding -> compressed ristretto."""
        bG = b"".join(_sb(s) for s in sG)
        bH = b"".join(_sb(s) for s in sH)
        out = ctypes.create_string_buffer(32)
        check(lib().bpg_msm_gens(self._h, bG or None, len(bG) // 32, bH or None, len(bH) // 32,
                                 _sb(sB) if sB is not None else None, _sb(sBb) if sBb is not None else None, out))
        return out.raw

    def msm_gens_bytes(self, bG=b"", bH=b"", sB=None, sBb=None):
        out = ctypes.create_string_buffer(32)
        check(lib().bpg_msm_gens(self._h, bG or None, len(bG) // 32, bH or None, len(bH) // 32, sB, sBb, out))
        return out.raw

    def msm_gens_dev(self, d_sG, nG, d_sH, nH, d_sB=None, d_sBb=None):
        """Device-pointer flavour (ints from torch .data_ptr())."""
        out = ctypes.create_string_buffer(32)
        check(lib().bpg_msm_gens_dev(self._h, d_sG, nG, d_sH, nH, d_sB, d_sBb, out))
        return out.raw

    def msm_gens_range_dev(self, d_sG, g_start, nG, d_sH, h_start, nH):
        """Point-range form over device-resident scalars (one shard of a large MSM)."""
        out = ctypes.create_string_buffer(32)
        check(lib().bpg_msm_gens_range_dev(self._h, d_sG, g_start, nG, d_sH, h_start, nH, None, None, out))
        return out.raw

    def msm(self, scalars, points):
        """vartime_multiscalar_mul over arbitrary compressed points."""
        bs = b"".join(_sb(s) for s in scalars)
        bp = b"".join(points)
        assert len(bs) == len(bp)
        out = ctypes.create_string_buffer(32)
        check(lib().bpg_msm(self._h, bs, bp, len(bs) // 32, out))
        return out.raw

    def pedersen_commit_batch(self, vs, rs):
        k = len(vs)
        out = ctypes.create_string_buffer(32 * max(k, 1))
        check(lib().bpg_pedersen_commit_batch(self._h, b"".join(_sb(v) for v in vs), b"".join(_sb(r) for r in rs), k,
                                              out))
        return [out.raw[32 * i: 32 * i + 32] for i in range(k)]


class Transcript:
    """merlin::Transcript (host side)."""

    def __init__(self, label=None, _h=None):
        self._h = _h if _h is not None else c_void_p(lib().bpg_transcript_new(label, len(label)))

    def clone(self):
        return Transcript(_h=c_void_p(lib().bpg_transcript_clone(self._h)))

    def append_message(self, label, msg):
        lib().bpg_transcript_append_message(self._h, label, len(label), msg, len(msg))

    def append_u64(self, label, x):
        self.append_message(label, int(x).to_bytes(8, "little"))

    def challenge_bytes(self, label, n):
        buf = ctypes.create_string_buffer(n)
        lib().bpg_transcript_challenge_bytes(self._h, label, len(label), buf, n)
        return buf.raw

    def __del__(self):
        try:
            if self._h:
                lib().bpg_transcript_free(self._h)
                self._h = None
        except Exception:
            pass


COMMITTED, MUL_LEFT, MUL_RIGHT, MUL_OUT, ONE = 0, 1, 2, 3, 4


class Variable(int):
    """uint32 tag kind<<29 | index (dalek r1cs::Variable)."""

    @staticmethod
    def make(kind, idx=0):
        return Variable((kind << 29) | idx)

    @staticmethod
    def One():
        return Variable(ONE << 29)

    kind = property(lambda s: int(s) >> 29)
    idx = property(lambda s: int(s) & ((1 << 29) - 1))


class LinearCombination:
    """Term list; + and - concatenate (no merging), like dalek's LinearCombination."""

    __slots__ = ("terms",)

    def __init__(self, terms=None):
        self.terms = list(terms) if terms else []

    @staticmethod
    def of(x):
        if isinstance(x, LinearCombination):
            return x
        if isinstance(x, Variable):
            return LinearCombination([(x, 1)])
        if isinstance(x, int):
            return LinearCombination([(Variable.One(), x)])
        raise TypeError(type(x))

    def __add__(self, o):
        return LinearCombination(self.terms + LinearCombination.of(o).terms)

    def __sub__(self, o):
        return LinearCombination(self.terms + [(v, (-c) % L_ORDER) for v, c in LinearCombination.of(o).terms])

    def __neg__(self):
        return LinearCombination([(v, (-c) % L_ORDER) for v, c in self.terms])

    def scale(self, s):
        return LinearCombination([(v, (c * s) % L_ORDER) for v, c in self.terms])

    def _pack(self):
        n = len(self.terms)
        vars_ = (c_uint32 * max(n, 1))(*[int(v) for v, _ in self.terms])
        coef = b"".join(_sb(c) for _, c in self.terms)
        return vars_, coef, n


class _CS:
    def _vars3(self, arr):
        return Variable(arr[0]), Variable(arr[1]), Variable(arr[2])


class Circuit:
    """Device-resident flattened constraint system (+ optional witness): bpg_circuit."""

    def __init__(self, ctx, n, m, row_start, term_var, term_coef, q):
        self.ctx, self.n, self.m, self.q = ctx, n, m, q
        self._h = c_void_p()
        check(lib().bpg_circuit_create(ctx._h, n, m, row_start.ctypes.data, term_var.ctypes.data, _buf(term_coef), q,
                                       byref(self._h)))

    def set_witness(self, aL, aR):
        assert _nbytes(aL) == _nbytes(aR) == 32 * self.n
        check(lib().bpg_circuit_set_witness(self._h, _buf(aL), _buf(aR)))
        return self

    def __del__(self):
        try:
            if self._h:
                lib().bpg_circuit_free(self._h)
                self._h = None
        except Exception:
            pass


class Prover(_CS):
    """bulletproofs::r1cs::Prover (GPU-backed)."""

    def attach(self, circuit):
        check(lib().bpg_prover_attach(self._h, circuit._h))
        self._circuit = circuit

    def __init__(self, ctx, transcript):
        self.ctx, self.transcript = ctx, transcript
        self._h = c_void_p()
        check(lib().bpg_prover_new(ctx._h, transcript._h, byref(self._h)))

    def __del__(self):
        try:
            if self._h:
                lib().bpg_prover_free(self._h)
                self._h = None
        except Exception:
            pass

    def commit(self, v, v_blinding):
        out = ctypes.create_string_buffer(32)
        var = c_uint32()
        check(lib().bpg_prover_commit(self._h, _sb(v), _sb(v_blinding), out, byref(var)))
        return out.raw, Variable(var.value)

    def commit_batch(self, vs, blindings):
        """k commitments in one launch.  vs / blindings: lists of scalars or packed k*32-byte buffers."""
        bv, k = _scalars(vs)
        bb, kb = _scalars(blindings)
        assert k == kb
        out = ctypes.create_string_buffer(32 * max(k, 1))
        first = c_uint32()
        check(lib().bpg_prover_commit_batch(self._h, bv, bb, k, out, byref(first)))
        raw = out.raw
        return [(raw[32 * i: 32 * i + 32], Variable(first.value + i)) for i in range(k)]

    def commit_batch_packed(self, vs, blindings):
        """Same, returning the k commitments as one k*32-byte string."""
        bv, k = _scalars(vs)
        bb, _ = _scalars(blindings)
        out = ctypes.create_string_buffer(32 * max(k, 1))
        check(lib().bpg_prover_commit_batch(self._h, bv, bb, k, out, None))
        return out.raw[: 32 * k]

    def allocate_multiplier(self, assignment):
        if assignment is None:
            raise BpgError(_capi.E_MISSING_ASSIGNMENT, "allocate_multiplier(None) on a prover")
        arr = (c_uint32 * 3)()
        check(lib().bpg_prover_allocate_multiplier(self._h, _sb(assignment[0]), _sb(assignment[1]), arr))
        return self._vars3(arr)

    def multiply(self, left, right):
        lv, lc, ln = LinearCombination.of(left)._pack()
        rv, rc, rn = LinearCombination.of(right)._pack()
        arr = (c_uint32 * 3)()
        check(lib().bpg_prover_multiply(self._h, lv, lc, ln, rv, rc, rn, arr))
        return self._vars3(arr)

    def constrain(self, lc):
        v, c, n = LinearCombination.of(lc)._pack()
        check(lib().bpg_prover_constrain(self._h, v, c, n))

    def load_cs(self, aL, aR, row_start, term_var, term_coef, q):
        """Bulk allocate_multiplier + constrain: aL/aR are n*32 bytes, constraints a CSR term list
        (numpy uint32 row_start[q+1], term_var[nnz]; term_coef nnz*32 bytes)."""
        assert _nbytes(aL) == _nbytes(aR) and _nbytes(aL) % 32 == 0
        check(lib().bpg_prover_load_cs(self._h, _buf(aL), _buf(aR), _nbytes(aL) // 32, row_start.ctypes.data,
                                       term_var.ctypes.data, _buf(term_coef), q))

    def load_cs_bits(self, n, runs, aLh, aRh, host_index, row_start, term_var, term_coef, q):
        """bpg_prover_load_cs_bits: `runs` = [(first, nbits, value int)], the other multipliers as compact arrays (bytes, 32 B
        each) with their multiplier indices.  Witness bits are generated on the device."""
        import numpy as np
        arr = (_capi.BitRun * max(len(runs), 1))()
        for k, (first, nbits, value) in enumerate(runs):
            arr[k].first, arr[k].nbits = first, nbits
            arr[k].value[:] = list(int(value).to_bytes(32, "little"))
        idx = np.asarray(host_index, dtype=np.uint32)
        check(lib().bpg_prover_load_cs_bits(self._h, n, arr, len(runs), _buf(aLh), _buf(aRh), idx.ctypes.data if idx.size else None,
                                            idx.size, row_start.ctypes.data, term_var.ctypes.data, _buf(term_coef), q))

    def num_constraints(self):
        return lib().bpg_prover_num_constraints(self._h)

    def num_multipliers(self):
        return lib().bpg_prover_num_multipliers(self._h)

    def prove(self, rng_seed=None):
        """Prover::prove(&bp_gens).to_bytes(); rng_seed (32 bytes) pins merlin's external entropy."""
        cap = 1 + 14 * 32 + (2 * 32 + 2) * 32
        buf = ctypes.create_string_buffer(cap)
        n = c_size_t()
        check(lib().bpg_prover_prove(self._h, rng_seed, buf, cap, byref(n)))
        return buf.raw[: n.value]


class Verifier(_CS):
    """bulletproofs::r1cs::Verifier (GPU-backed)."""

    def attach(self, circuit):
        check(lib().bpg_verifier_attach(self._h, circuit._h))
        self._circuit = circuit

    def __init__(self, ctx, transcript):
        self.ctx, self.transcript = ctx, transcript
        self._h = c_void_p()
        check(lib().bpg_verifier_new(ctx._h, transcript._h, byref(self._h)))

    def __del__(self):
        try:
            if self._h:
                lib().bpg_verifier_free(self._h)
                self._h = None
        except Exception:
            pass

    def commit(self, V):
        var = c_uint32()
        check(lib().bpg_verifier_commit(self._h, bytes(V), byref(var)))
        return Variable(var.value)

    def commit_batch(self, Vs):
        """Vs: list of 32-byte commitments or one packed k*32-byte string."""
        if isinstance(Vs, (list, tuple)):
            Vs = b"".join(Vs)
        k = len(Vs) // 32
        first = c_uint32()
        check(lib().bpg_verifier_commit_batch(self._h, Vs or None, k, byref(first)))
        return [Variable(first.value + i) for i in range(k)]

    def allocate_multiplier(self, assignment=None):
        arr = (c_uint32 * 3)()
        check(lib().bpg_verifier_allocate_multiplier(self._h, arr))
        return self._vars3(arr)

    def multiply(self, left, right):
        lv, lc, ln = LinearCombination.of(left)._pack()
        rv, rc, rn = LinearCombination.of(right)._pack()
        arr = (c_uint32 * 3)()
        check(lib().bpg_verifier_multiply(self._h, lv, lc, ln, rv, rc, rn, arr))
        return self._vars3(arr)

    def constrain(self, lc):
        v, c, n = LinearCombination.of(lc)._pack()
        check(lib().bpg_verifier_constrain(self._h, v, c, n))

    def load_cs(self, n, row_start, term_var, term_coef, q):
        check(lib().bpg_verifier_load_cs(self._h, n, row_start.ctypes.data, term_var.ctypes.data, _buf(term_coef), q))

    def num_vars(self):
        return lib().bpg_verifier_num_vars(self._h)

    def verify(self, proof, rng_seed=None):
        """True = accepted, False = rejected (VerificationError); malformed proofs raise FormatError."""
        rc = lib().bpg_verifier_verify(self._h, proof, len(proof), rng_seed)
        if rc == _capi.OK:
            return True
        if rc == _capi.E_VERIFY:
            return False
        check(rc)


def prove(ctx, name, instance, witness, gadgets, blinding_seed=None, rng_seed=None):
    """prove() of /root/reference/src/prove.rs:37-43 -> (proof bytes, commitments text, #constraints)."""
    out = ctypes.POINTER(_capi.ProofArtifacts)()
    check(lib().bpg_prove(ctx._h, name.encode(), instance.encode(), witness.encode(), gadgets.encode(),
                          blinding_seed, rng_seed, byref(out)))
    try:
        art = out.contents
        proof = bytes(art.proof[: art.proof_len])
        return proof, art.commitments.decode(), art.num_constraints
    finally:
        lib().bpg_free_proof(out)


def verify(ctx, name, instance, proof, commitments, gadgets, rng_seed=None):
    """verify() of /root/reference/src/verify.rs:36-42 -> bool."""
    acc = c_int()
    check(lib().bpg_verify(ctx._h, name.encode(), instance.encode(), gadgets.encode(), commitments.encode(), proof,
                           len(proof), rng_seed, byref(acc)))
    return bool(acc.value)


def mimc_hash(data):
    """mimc_hash of /root/reference/src/mimc_hash/mimc.rs:61-75 -> scalar (int)."""
    out = ctypes.create_string_buffer(32)
    check(lib().bpg_mimc_hash(bytes(data), len(data), out))
    return int.from_bytes(out.raw, "little")


def mimc_sponge(scalars):
    """Un-padded MiMC sponge over scalars (Merkle inner nodes; mimc.rs:24-40) -> scalar (int)."""
    out = ctypes.create_string_buffer(32)
    check(lib().bpg_mimc_sponge(b"".join(_sb(s) for s in scalars), len(scalars), out))
    return int.from_bytes(out.raw, "little")


def _take_flat(out, proving, label):
    import numpy as np
    from .workloads import FlatStatement
    f = out.contents
    n, m, q, nnz = f.n, f.m, f.q, f.nnz

    def get(p, k):
        return bytes(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8 * k)).contents) if k and p else b""

    st = FlatStatement.__new__(FlatStatement)
    st.label, st.n, st.q, st.m = label, n, q, m
    st.v_bytes, st.vbl_bytes = (get(f.v32m, 32 * m), get(f.vbl32m, 32 * m)) if proving else (b"", b"")
    st.v = [int.from_bytes(st.v_bytes[32 * i: 32 * i + 32], "little") for i in range(m)] if proving else []
    st.vbl = [int.from_bytes(st.vbl_bytes[32 * i: 32 * i + 32], "little") for i in range(m)] if proving else []
    st.V = [get(f.V32m, 32 * m)[32 * i: 32 * i + 32] for i in range(m)] if not proving else []
    st.aL, st.aR = (get(f.aL32n, 32 * n), get(f.aR32n, 32 * n)) if proving else (b"", b"")
    st.row_start = np.ctypeslib.as_array(f.row_start, shape=(q + 1,)).copy()      # one copy out of the C arrays
    st.term_var = np.ctypeslib.as_array(f.term_var, shape=(nnz,)).copy() if nnz else np.zeros(1, dtype=np.uint32)
    st.term_coef = get(f.term_coef32, 32 * nnz) or bytes(32)
    st.com_names = f.com_names.decode().split("\n")[:-1]
    lib().bpg_flat_statement_free(out)
    return st


def flatten_prover(name, instance, witness, gadgets, blinding_seed=None):
    """The flat statement the real prover holds after assign_buffer (/root/reference/src/prove.rs:84-99), built by the
    library's own front end from the text formats.  Host only."""
    out = ctypes.POINTER(_capi.FlatStatementC)()
    check(lib().bpg_frontend_flatten_prover(name.encode(), instance.encode(), witness.encode(), gadgets.encode(),
                                            blinding_seed, byref(out)))
    return _take_flat(out, True, name.encode())


def flatten_verifier(name, instance, commitments, gadgets):
    out = ctypes.POINTER(_capi.FlatStatementC)()
    check(lib().bpg_frontend_flatten_verifier(name.encode(), instance.encode(), commitments.encode(), gadgets.encode(),
                                              byref(out)))
    return _take_flat(out, False, name.encode())


# ------------------------------------------------------------------------------------------------
# batches (bpg_r1cs_prove_batch / bpg_r1cs_verify_batch / bpg_prove_batch / bpg_verify_batch): the library owns the
# host threads, one per context
# ------------------------------------------------------------------------------------------------
def _ctx_array(ctxs):
    arr = (c_void_p * len(ctxs))(*[c._h for c in ctxs])
    return ctypes.cast(arr, ctypes.POINTER(c_void_p)), arr


def _ptr(x, keep):
    """address of bytes / numpy data, keeping the object alive in `keep`."""
    if x is None:
        return None
    if isinstance(x, (bytes, bytearray)):
        if not x:
            return None
        b = ctypes.create_string_buffer(bytes(x), len(x)) if isinstance(x, bytearray) else x
        keep.append(b)
        return ctypes.cast(ctypes.c_char_p(b) if isinstance(b, bytes) else b, c_void_p).value
    keep.append(x)
    return x.ctypes.data if x.size else None


def prove_batch(ctxs, statements, seeds, circuits=None, verify=False, verify_seeds=None):
    """N independent flat statements (workloads.FlatStatement-like: label, v_bytes, vbl_bytes, m, aL, aR, n, row_start,
    term_var, term_coef, q) through bpg_r1cs_prove_batch.  `circuits`: one resident Circuit (with witness) for all, or a
    list.  Returns a list of (status, proof bytes, commitments bytes m*32)."""
    n = len(statements)
    jobs = (_capi.ProveJob * n)()
    keep, outs = [], []
    for k, st in enumerate(statements):
        j = jobs[k]
        j.label, j.label_len = _ptr(st.label, keep), len(st.label)
        j.v32m, j.vbl32m, j.m = _ptr(st.v_bytes, keep), _ptr(st.vbl_bytes, keep), st.m
        circ = circuits[k] if isinstance(circuits, (list, tuple)) else circuits
        if circ is not None:
            j.circuit = circ._h
            keep.append(circ)
        else:
            j.aL32n, j.aR32n = _ptr(st.aL, keep), _ptr(st.aR, keep)
        j.n, j.q = st.n, st.q
        j.row_start, j.term_var, j.term_coef32 = _ptr(st.row_start, keep), _ptr(st.term_var, keep), _ptr(st.term_coef, keep)
        j.rng_seed32 = _ptr(seeds[k] if seeds else None, keep)
        j.verify_seed32 = _ptr(verify_seeds[k] if verify_seeds else None, keep)
        j.flags = _capi.JOB_VERIFY if verify else 0
        V = ctypes.create_string_buffer(32 * max(st.m, 1))
        P = ctypes.create_string_buffer(1 + 14 * 32 + 66 * 32)
        outs.append((V, P))
        j.V_out32m, j.proof_out, j.proof_cap = ctypes.addressof(V), ctypes.addressof(P), len(P)
    cp, arr = _ctx_array(ctxs)
    rc = lib().bpg_r1cs_prove_batch(cp, len(ctxs), jobs, n)
    if rc < 0:
        check(rc)
    return [(jobs[k].status, outs[k][1].raw[:jobs[k].proof_len], outs[k][0].raw[:32 * statements[k].m]) for k in range(n)]


def verify_batch(ctxs, statements, coms, proofs, seeds=None, circuits=None):
    """bpg_r1cs_verify_batch: coms[k] = m*32 bytes.  Returns the list of statuses (0 accepted, -2 rejected, -1 malformed)."""
    n = len(statements)
    jobs = (_capi.VerifyJob * n)()
    keep = []
    for k, st in enumerate(statements):
        j = jobs[k]
        j.label, j.label_len = _ptr(st.label, keep), len(st.label)
        c = coms[k] if isinstance(coms[k], (bytes, bytearray)) else b"".join(coms[k])
        j.V32m, j.m = _ptr(c, keep), len(c) // 32
        circ = circuits[k] if isinstance(circuits, (list, tuple)) else circuits
        if circ is not None:
            j.circuit = circ._h
            keep.append(circ)
        j.n, j.q = st.n, st.q
        j.row_start, j.term_var, j.term_coef32 = _ptr(st.row_start, keep), _ptr(st.term_var, keep), _ptr(st.term_coef, keep)
        j.proof, j.proof_len = _ptr(proofs[k], keep), len(proofs[k])
        j.rng_seed32 = _ptr(seeds[k] if seeds else None, keep)
    cp, arr = _ctx_array(ctxs)
    rc = lib().bpg_r1cs_verify_batch(cp, len(ctxs), jobs, n)
    if rc < 0:
        check(rc)
    return [jobs[k].status for k in range(n)]


def prove_text_batch(ctxs, texts, blinding_seeds=None, rng_seeds=None, verify=False):
    """bpg_prove_batch over (name, instance, witness, gadgets) tuples.  Returns [(status, proof, commitments text,
    accepted)]; `accepted` is meaningful with verify=True."""
    n = len(texts)
    jobs = (_capi.TextJob * n)()
    keep = []
    for k, (name, inst, wtns, gad) in enumerate(texts):
        j = jobs[k]
        j.name, j.instance, j.witness, j.gadgets = name.encode(), inst.encode(), wtns.encode(), gad.encode()
        j.blinding_seed32 = _ptr(blinding_seeds[k] if blinding_seeds else None, keep)
        j.rng_seed32 = _ptr(rng_seeds[k] if rng_seeds else None, keep)
        j.verify_seed32 = j.rng_seed32
        j.flags = _capi.JOB_VERIFY if verify else 0
    cp, arr = _ctx_array(ctxs)
    rc = lib().bpg_prove_batch(cp, len(ctxs), jobs, n)
    if rc < 0:
        check(rc)
    out = []
    for k in range(n):
        j = jobs[k]
        if j.artifacts:
            a = j.artifacts.contents
            out.append((j.status, ctypes.string_at(a.proof, a.proof_len), a.commitments.decode(), bool(j.accepted)))
            lib().bpg_free_proof(j.artifacts)
        else:
            out.append((j.status, b"", "", False))
    return out


def verify_text_batch(ctxs, texts, seeds=None):
    """bpg_verify_batch over (name, instance, gadgets, commitments text, proof) tuples -> [(status, accepted)]."""
    n = len(texts)
    jobs = (_capi.TextJob * n)()
    keep = []
    for k, (name, inst, gad, coms, proof) in enumerate(texts):
        j = jobs[k]
        j.name, j.instance, j.gadgets, j.commitments = name.encode(), inst.encode(), gad.encode(), coms.encode()
        j.proof, j.proof_len = _ptr(proof, keep), len(proof)
        j.verify_seed32 = _ptr(seeds[k] if seeds else None, keep)
    cp, arr = _ctx_array(ctxs)
    rc = lib().bpg_verify_batch(cp, len(ctxs), jobs, n)
    if rc < 0:
        check(rc)
    return [(jobs[k].status, bool(jobs[k].accepted)) for k in range(n)]


def c_prove(name, instance, witness, gadgets):
    """The reference's own C entry point (include/bulletproofs_gadgets.h): -> (proof bytes, commitments text) or None."""
    a = lib().c_prove(name.encode(), instance.encode(), witness.encode(), gadgets.encode())
    if not a:
        return None
    try:
        c = a.contents
        assert c.proof_cap >= c.proof_len
        return ctypes.string_at(c.proof, c.proof_len), c.commitments.decode()
    finally:
        lib().free_proof(a)


def c_verify(name, instance, gadgets, commitments, proof):
    return bool(lib().c_verify(name.encode(), instance.encode(), gadgets.encode(), commitments.encode(), proof, len(proof)))
