#!/usr/bin/env python3
"""Regenerates tests/golden/baseline_sizes.json: SHA-256 of the proof / commitments / MSM outputs the ORACLE produces
at the full BASELINE.json sizes, so that `pytest -m gpu` can compare the CUDA path byte for byte where the oracle
itself is too slow to run inside the test (config 2 x1024: ~35 s, config 3: ~25-40 s, MSM 2^20: ~15 s).

    python tests/golden/make_baseline_golden.py          (about 5 minutes of CPU)

ORACLE-generated: oracle/pyref front end (python) flattens the text statements, oracle/c (restatement of dalek's
CPU algorithms) proves and verifies.  The Rust reference cannot run in this image, so these pin the CUDA path to the
oracle, not to dalek.  Inputs are rebuilt from fixed seeds by the same workload functions the tests
call (bulletproof_gadgets_b200/workloads.py builds inputs only -- no product code produces any byte hashed here)."""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bulletproof_gadgets_b200 import workloads as W  # noqa: E402  (input builders only)
from oracle import coracle  # noqa: E402
from oracle.pyref import frontend as F  # noqa: E402
from tests import frontend_glue as G  # noqa: E402

SEED_P, SEED_V = b"\x07" * 32, b"\x09" * 32
sha = lambda b: hashlib.sha256(b).hexdigest()
out = {"seeds": {"prove": SEED_P.hex(), "verify": SEED_V.hex()}}


def msm_scalars(lg, kind, seed):
    """2^lg scalars: first half multiplies G_0.., second half H_0..  kind 'uniform' (< 2^252) or 'bits' (0/1)."""
    n = 1 << lg
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        a[:, 31] &= 0x0F
    else:
        a = np.zeros((n, 32), dtype=np.uint8)
        a[:, 0] = rng.integers(0, 2, size=n, dtype=np.uint8)
    return a


def record_flat(key, st, extra=None):
    t0 = time.time()
    proof, coms = coracle.prove_flat(st, SEED_P)
    assert coracle.verify_flat(st, coms, proof, SEED_V) is True
    out[key] = {"n": st.n, "m": st.m, "q": st.q, "proof_len": len(proof), "proof_sha256": sha(proof),
                "coms_sha256": sha(b"".join(coms)), "proof_head": proof[:33].hex()}
    out[key].update(extra or {})
    print("%-28s n=%-7d %s  (%.1f s)" % (key, st.n, out[key]["proof_sha256"][:16], time.time() - t0), flush=True)


if __name__ == "__main__":
    # config 5: raw fixed-base MSM (verifier mega-MSM shape) -- compressed ristretto output
    out["msm"] = {}
    for lg in (17, 18, 20):
        for kind in ("uniform", "bits"):
            t0 = time.time()
            a = msm_scalars(lg, kind, 1000 + lg)
            h = (1 << lg) // 2
            r = coracle.msm_gens(a[:h].tobytes(), a[h:].tobytes(), None, None)
            out["msm"]["%s_2^%d" % (kind, lg)] = r.hex()
            print("msm %-8s 2^%d %s  (%.1f s)" % (kind, lg, r.hex()[:16], time.time() - t0), flush=True)
    # config 2: BOUND 64-bit x1024 in one proof (the statement bench.py times, default seed / label)
    record_flat("config2_bound_x1024", W.bounds_check_statement(1024))
    # config 3: Merkle membership depth 32 with MiMC, siblings as instance values (n' = 2^16) and as witnesses (n' = 2^17)
    hashers = (F.mimc_hash, F.mimc_sponge)
    for wit in (False, True):
        gad, inst, wtns = W.merkle_text(32, witness_siblings=wit, hashers=hashers)
        st = F.compile_prover("merkle32", inst, wtns, gad, G.blinding(b"\x03" * 32))
        record_flat("config3_merkle32_%s" % ("witness" if wit else "instance"), st)
    # config 4: the first 8 statements of the 4096-proof batch (LESS_THAN / SET_MEMBER alternating)
    for k, (gad, inst, wtns) in enumerate(W.batch_texts(8)):
        st = F.compile_prover("b%d" % k, inst, wtns, gad, G.blinding(bytes([k + 1]) * 32))
        record_flat("config4_stmt%d" % k, st)
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "baseline_sizes.json"), "w"), indent=1)
    print("written")
