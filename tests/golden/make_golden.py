#!/usr/bin/env python3
"""Regenerates tests/golden/r1cs_small.json from the big-int python oracle (oracle/pyref).

    python tests/golden/make_golden.py

The reference itself cannot run in this image (pure Rust, no cargo/rustc), so these are
ORACLE-generated goldens: they freeze the oracle's bytes (generators, commitments, proofs under a seeded
transcript rng) so that the C restatement, the CUDA library and any future dalek dump can be compared
against one committed set of vectors.  Circuits: tests/circuits.py."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.pyref import r1cs as O  # noqa: E402
from tests import circuits as C  # noqa: E402

SEED_PROVE = b"\x07" * 32

CASES = {
    "empty": "c_empty()",
    "mul3": "c_mul3()",
    "range-8-1": "c_range(8, 1)",
    "range-4-3": "c_range(4, 3)",
    "unreduced": "c_unreduced()",
    "chain-5": "c_chain(5)",
    "chain-37": "c_chain(37)",
}


def main():
    ob = C.OracleBackend()
    out = {"seed_prove": SEED_PROVE.hex(), "cases": {}}
    gens = O.BulletproofGens(16)
    pc = O.PedersenGens()
    out["generators"] = {"G": [p.compress().hex() for p in gens.G], "H": [p.compress().hex() for p in gens.H],
                         "B": pc.B.compress().hex(), "B_blinding": pc.B_blinding.compress().hex()}
    for name, expr in CASES.items():
        circ = eval("C." + expr)
        proof, coms = circ.prove(ob, SEED_PROVE)
        assert circ.verify(ob, proof, coms) is True
        out["cases"][name] = {"make": expr, "label": circ.label.decode(), "commitments": [c.hex() for c in coms],
                              "proof": proof.hex()}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "r1cs_small.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
