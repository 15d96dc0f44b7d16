"""ctypes loader for libbpg.so and the prototypes of include/bpg.h."""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_int, c_int64, c_size_t, c_uint8, c_uint32, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbpg.so")

OK, E_FORMAT, E_VERIFY, E_GENS_LEN, E_MISSING_ASSIGNMENT, E_CUDA, E_ARG, E_GADGET = 0, -1, -2, -3, -4, -5, -6, -7
_NAMES = {E_FORMAT: "FormatError", E_VERIFY: "VerificationError", E_GENS_LEN: "InvalidGeneratorsLength",
          E_MISSING_ASSIGNMENT: "MissingAssignment", E_CUDA: "CudaError", E_ARG: "ArgumentError",
          E_GADGET: "GadgetError"}


class BpgError(RuntimeError):
    def __init__(self, code, msg=""):
        self.code = code
        super().__init__("%s (%d): %s" % (_NAMES.get(code, "Error"), code, msg))


class ProofArtifacts(ctypes.Structure):
    _fields_ = [("commitments", c_char_p), ("proof", POINTER(c_uint8)), ("proof_len", c_size_t),
                ("num_constraints", c_uint64)]


class RefProofArtifacts(ctypes.Structure):
    """struct ProofArtifacts of include/bulletproofs_gadgets.h (/root/reference/interfaces/ios/src/lib.rs:11-19)."""
    _fields_ = [("commitments", c_char_p), ("proof", POINTER(c_uint8)), ("proof_len", c_size_t), ("proof_cap", c_size_t)]


class ProveJob(ctypes.Structure):
    _fields_ = [("label", c_void_p), ("label_len", c_size_t), ("v32m", c_void_p), ("vbl32m", c_void_p), ("m", c_uint64),
                ("circuit", c_void_p), ("aL32n", c_void_p), ("aR32n", c_void_p), ("n", c_uint64), ("row_start", c_void_p),
                ("term_var", c_void_p), ("term_coef32", c_void_p), ("q", c_uint64), ("rng_seed32", c_void_p),
                ("verify_seed32", c_void_p), ("flags", c_uint32), ("V_out32m", c_void_p), ("proof_out", c_void_p),
                ("proof_cap", c_size_t), ("proof_len", c_size_t), ("status", c_int)]


class VerifyJob(ctypes.Structure):
    _fields_ = [("label", c_void_p), ("label_len", c_size_t), ("V32m", c_void_p), ("m", c_uint64), ("circuit", c_void_p),
                ("n", c_uint64), ("row_start", c_void_p), ("term_var", c_void_p), ("term_coef32", c_void_p), ("q", c_uint64),
                ("proof", c_void_p), ("proof_len", c_size_t), ("rng_seed32", c_void_p), ("status", c_int)]


class TextJob(ctypes.Structure):
    _fields_ = [("name", c_char_p), ("instance", c_char_p), ("witness", c_char_p), ("gadgets", c_char_p),
                ("commitments", c_char_p), ("proof", c_void_p), ("proof_len", c_size_t), ("blinding_seed32", c_void_p),
                ("rng_seed32", c_void_p), ("verify_seed32", c_void_p), ("flags", c_uint32),
                ("artifacts", POINTER(ProofArtifacts)), ("accepted", c_int), ("status", c_int)]


JOB_VERIFY = 1


class BitRun(ctypes.Structure):
    _fields_ = [("first", c_uint64), ("nbits", c_uint32), ("reserved", c_uint32), ("value", c_uint8 * 32)]


class FlatStatementC(ctypes.Structure):
    _fields_ = [("n", c_uint64), ("m", c_uint64), ("q", c_uint64), ("nnz", c_uint64), ("v32m", POINTER(c_uint8)),
                ("vbl32m", POINTER(c_uint8)), ("V32m", POINTER(c_uint8)), ("aL32n", POINTER(c_uint8)),
                ("aR32n", POINTER(c_uint8)), ("row_start", POINTER(c_uint32)), ("term_var", POINTER(c_uint32)),
                ("term_coef32", POINTER(c_uint8)), ("com_names", c_char_p)]


_lib = None

_u8p = POINTER(c_uint8)
_u32p = POINTER(c_uint32)

PROTOTYPES = {
    "bpg_last_error": (c_char_p, []),
    "bpg_ctx_create": (c_int, [c_int, POINTER(c_void_p)]),
    "bpg_ctx_create_shared": (c_int, [c_void_p, POINTER(c_void_p)]),
    "bpg_ctx_destroy": (None, [c_void_p]),
    "bpg_host_alloc": (c_void_p, [c_size_t]),
    "bpg_host_free": (None, [c_void_p]),
    "bpg_ctx_set": (c_int, [c_void_p, c_char_p, c_int64]),
    "bpg_ctx_get": (c_int64, [c_void_p, c_char_p]),
    "bpg_selftest_field": (c_int, [c_void_p, c_uint64, c_uint64, POINTER(c_uint64)]),
    "bpg_msm_gens_range_dev": (c_int, [c_void_p, c_void_p, c_uint64, c_uint64, c_void_p, c_uint64, c_uint64, c_void_p, c_void_p,
                                       c_char_p]),
    "bpg_gens_ensure": (c_int, [c_void_p, c_uint64]),
    "bpg_gens_compressed": (c_int, [c_void_p, c_int, c_uint64, c_uint64, c_char_p]),
    "bpg_msm_gens": (c_int, [c_void_p, c_char_p, c_uint64, c_char_p, c_uint64, c_char_p, c_char_p, c_char_p]),
    "bpg_msm_gens_range": (c_int, [c_void_p, c_void_p, c_uint64, c_uint64, c_void_p, c_uint64, c_uint64, c_char_p, c_char_p,
                                   c_char_p]),
    "bpg_point_sum": (c_int, [c_char_p, c_uint64, c_char_p]),
    "bpg_msm_gens_dev": (c_int, [c_void_p, c_void_p, c_uint64, c_void_p, c_uint64, c_void_p, c_void_p, c_char_p]),
    "bpg_msm": (c_int, [c_void_p, c_char_p, c_char_p, c_uint64, c_char_p]),
    "bpg_pedersen_commit_batch": (c_int, [c_void_p, c_char_p, c_char_p, c_uint64, c_char_p]),
    "bpg_transcript_new": (c_void_p, [c_char_p, c_size_t]),
    "bpg_transcript_clone": (c_void_p, [c_void_p]),
    "bpg_transcript_free": (None, [c_void_p]),
    "bpg_transcript_append_message": (None, [c_void_p, c_char_p, c_size_t, c_char_p, c_size_t]),
    "bpg_transcript_challenge_bytes": (None, [c_void_p, c_char_p, c_size_t, c_char_p, c_size_t]),
    "bpg_transcript_rng_fill64": (c_int, [c_void_p, c_char_p, c_size_t, c_char_p, c_size_t, c_char_p, c_size_t]),
    "bpg_rng_batcher_stat": (c_int64, [c_int]),
    "bpg_prover_new": (c_int, [c_void_p, c_void_p, POINTER(c_void_p)]),
    "bpg_prover_free": (None, [c_void_p]),
    "bpg_prover_commit": (c_int, [c_void_p, c_char_p, c_char_p, c_char_p, _u32p]),
    "bpg_prover_commit_batch": (c_int, [c_void_p, c_void_p, c_void_p, c_uint64, c_char_p, _u32p]),
    "bpg_prover_allocate_multiplier": (c_int, [c_void_p, c_char_p, c_char_p, _u32p]),
    "bpg_prover_multiply": (c_int, [c_void_p, _u32p, c_char_p, c_size_t, _u32p, c_char_p, c_size_t, _u32p]),
    "bpg_prover_constrain": (c_int, [c_void_p, _u32p, c_char_p, c_size_t]),
    "bpg_prover_load_cs": (c_int, [c_void_p, c_void_p, c_void_p, c_uint64, c_void_p, c_void_p, c_void_p, c_uint64]),
    "bpg_prover_load_cs_bits": (c_int, [c_void_p, c_uint64, POINTER(BitRun), c_uint64, c_void_p, c_void_p, c_void_p, c_uint64,
                                        c_void_p, c_void_p, c_void_p, c_uint64]),
    "bpg_prover_num_constraints": (c_uint64, [c_void_p]),
    "bpg_prover_num_multipliers": (c_uint64, [c_void_p]),
    "bpg_prover_prove": (c_int, [c_void_p, c_char_p, c_char_p, c_size_t, POINTER(c_size_t)]),
    "bpg_circuit_create": (c_int, [c_void_p, c_uint64, c_uint64, c_void_p, c_void_p, c_void_p, c_uint64,
                                   POINTER(c_void_p)]),
    "bpg_circuit_set_witness": (c_int, [c_void_p, c_void_p, c_void_p]),
    "bpg_circuit_free": (None, [c_void_p]),
    "bpg_prover_attach": (c_int, [c_void_p, c_void_p]),
    "bpg_verifier_attach": (c_int, [c_void_p, c_void_p]),
    "bpg_verifier_new": (c_int, [c_void_p, c_void_p, POINTER(c_void_p)]),
    "bpg_verifier_free": (None, [c_void_p]),
    "bpg_verifier_commit": (c_int, [c_void_p, c_char_p, _u32p]),
    "bpg_verifier_commit_batch": (c_int, [c_void_p, c_void_p, c_uint64, _u32p]),
    "bpg_verifier_allocate_multiplier": (c_int, [c_void_p, _u32p]),
    "bpg_verifier_multiply": (c_int, [c_void_p, _u32p, c_char_p, c_size_t, _u32p, c_char_p, c_size_t, _u32p]),
    "bpg_verifier_constrain": (c_int, [c_void_p, _u32p, c_char_p, c_size_t]),
    "bpg_verifier_load_cs": (c_int, [c_void_p, c_uint64, c_void_p, c_void_p, c_void_p, c_uint64]),
    "bpg_verifier_num_vars": (c_uint64, [c_void_p]),
    "bpg_verifier_verify": (c_int, [c_void_p, c_char_p, c_size_t, c_char_p]),
    "bpg_prove": (c_int, [c_void_p, c_char_p, c_char_p, c_char_p, c_char_p, c_char_p, c_char_p,
                          POINTER(POINTER(ProofArtifacts))]),
    "bpg_verify": (c_int, [c_void_p, c_char_p, c_char_p, c_char_p, c_char_p, c_char_p, c_size_t, c_char_p,
                           POINTER(c_int)]),
    "bpg_free_proof": (None, [POINTER(ProofArtifacts)]),
    "bpg_frontend_flatten_prover": (c_int, [c_char_p, c_char_p, c_char_p, c_char_p, c_char_p,
                                            POINTER(POINTER(FlatStatementC))]),
    "bpg_frontend_flatten_verifier": (c_int, [c_char_p, c_char_p, c_char_p, c_char_p, POINTER(POINTER(FlatStatementC))]),
    "bpg_flat_statement_free": (None, [POINTER(FlatStatementC)]),
    "bpg_mimc_hash": (c_int, [c_char_p, c_size_t, c_char_p]),
    "bpg_mimc_sponge": (c_int, [c_char_p, c_size_t, c_char_p]),
    "bpg_r1cs_prove_batch": (c_int, [POINTER(c_void_p), c_size_t, POINTER(ProveJob), c_size_t]),
    "bpg_r1cs_verify_batch": (c_int, [POINTER(c_void_p), c_size_t, POINTER(VerifyJob), c_size_t]),
    "bpg_prove_batch": (c_int, [POINTER(c_void_p), c_size_t, POINTER(TextJob), c_size_t]),
    "bpg_verify_batch": (c_int, [POINTER(c_void_p), c_size_t, POINTER(TextJob), c_size_t]),
}
# include/bulletproofs_gadgets.h: the reference's own C ABI (same names as /root/reference/interfaces/ios/src/lib.rs)
REF_PROTOTYPES = {
    "c_prove": (POINTER(RefProofArtifacts), [c_char_p, c_char_p, c_char_p, c_char_p]),
    "c_verify": (ctypes.c_bool, [c_char_p, c_char_p, c_char_p, c_char_p, c_char_p, c_size_t]),
    "free_proof": (None, [POINTER(RefProofArtifacts)]),
}


def lib():
    """Loads libbpg.so (built in-tree by bulletproof_gadgets_b200.build).  Fails loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BpgError(E_CUDA, "libbpg.so is not built (run `python -m bulletproof_gadgets_b200.build`); "
                                   "there is no CPU fallback")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in list(PROTOTYPES.items()) + list(REF_PROTOTYPES.items()):
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != OK:
        raise BpgError(rc, (lib().bpg_last_error() or b"").decode("utf-8", "replace"))
    return rc
