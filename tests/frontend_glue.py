"""Test glue: deterministic blindings + fixture loading for the statement front end tests."""
import hashlib
import os

from oracle.pyref.merlin import L

FIXDIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fixtures")
STEMS = ["example", "bounds_check", "equality", "inequality", "less_than", "merkle_tree", "mimc_hash", "set_membership",
         "or", "or2", "or3", "or4", "or5"]
SEED_BLIND, SEED_PROVE, SEED_VERIFY = b"fixture-blindings--bpg-golden-01", b"\x07" * 32, b"\x09" * 32
assert len(SEED_BLIND) == 32  # also handed to the C ABI (bpg_prove blinding_seed32)


def blinding(seed=SEED_BLIND):
    """k-th commitment blinding = SHAKE256(seed || LE64(k)) 64 bytes, reduced mod l (the reference draws
    Scalar::random(thread_rng()); /root/reference/src/commitments.rs:28,40, /root/reference/src/gadget.rs:32)."""
    return lambda k: int.from_bytes(hashlib.shake_256(seed + k.to_bytes(8, "little")).digest(64), "little") % L


def load(stem):
    return tuple(open(os.path.join(FIXDIR, stem + ext)).read() for ext in (".inst", ".wtns", ".gadgets"))
