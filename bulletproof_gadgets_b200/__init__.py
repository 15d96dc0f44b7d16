"""bulletproof_gadgets_b200 -- B200-native Bulletproofs R1CS prove/verify hot path.

Thin ctypes binding over the C ABI declared in include/bpg.h (libbpg.so, hand-written sm_100a
CUDA).  The Python names mirror dalek's `bulletproofs::r1cs` API that the reference's gadgets are
written against (/root/reference/src/gadget.rs:7-60, /root/reference/src/cs_buffer.rs:89-116).
There is no CPU fallback: importing works anywhere, but every arithmetic call raises BpgError when
the CUDA library or an sm_100a device is missing.
"""
import ctypes
import os

from . import _capi
from ._capi import BpgError, lib  # noqa: F401
from .api import (  # noqa: F401
    Circuit,
    Context,
    LinearCombination,
    Prover,
    Transcript,
    Variable,
    Verifier,
    c_prove,
    c_verify,
    flatten_prover,
    flatten_verifier,
    mimc_hash,
    mimc_sponge,
    pinned_copy,
    pinned_empty,
    prove,
    prove_batch,
    prove_text_batch,
    verify,
    verify_batch,
    verify_text_batch,
)

__all__ = ["Context", "Transcript", "Prover", "Verifier", "LinearCombination", "Variable", "BpgError", "prove",
           "verify", "lib"]
