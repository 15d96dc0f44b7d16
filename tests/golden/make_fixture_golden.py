#!/usr/bin/env python3
"""Regenerates tests/golden/fixtures.json: for each of the reference's 13 fixture statements the flat circuit
sizes, the .coms text and the SHA-256 of the proof bytes, as produced by the oracle (python front end +
C restatement of dalek's prover) under the seeded randomness of tests/frontend_glue.py.

    python tests/golden/make_fixture_golden.py          (about 90 s of CPU)

ORACLE-generated (the Rust reference cannot run in this image); every proof is checked with the oracle verifier."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import coracle  # noqa: E402
from oracle.pyref import frontend as F  # noqa: E402
from tests import frontend_glue as G  # noqa: E402

out = {}
for stem in G.STEMS:
    inst, wtns, gad = G.load(stem)
    st = F.compile_prover(stem, inst, wtns, gad, G.blinding())
    proof, coms = coracle.prove_flat(st, G.SEED_PROVE)
    text = st.coms_text(coms)
    vs = F.compile_verifier(stem, inst, text, gad)
    assert coracle.verify_flat(vs, vs.V, proof, G.SEED_VERIFY) is True
    out[stem] = {"n": st.n, "m": st.m, "q": st.q, "nnz": st.nnz, "proof_len": len(proof),
                 "proof_sha256": hashlib.sha256(proof).hexdigest(), "coms_sha256": hashlib.sha256(text.encode()).hexdigest(),
                 "coms_first": text.splitlines()[0] if text else "", "proof_head": proof[:65].hex()}
    print(stem, out[stem]["n"], out[stem]["proof_sha256"][:16])
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures.json"), "w"), indent=1)
