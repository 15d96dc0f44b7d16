"""GPU parity on randomly generated constraint systems (seeded): random multipliers, random linear combinations over
every variable kind (committed, left, right, output, One) with random coefficients, each constraint completed so that
the witness satisfies it.  The CUDA library (op-by-op C ABI and bulk loader), the C oracle and the python oracle must
produce the same proof bytes; the proofs verify; perturbing one coefficient makes them fail."""
import random

import pytest

from oracle import coracle
from oracle.pyref import r1cs as O
from oracle.pyref.merlin import L, Transcript as OTranscript

pytestmark = pytest.mark.gpu
K_COMMITTED, K_LEFT, K_RIGHT, K_OUT, K_ONE = 0, 1, 2, 3, 4


def random_system(seed, n, m, q, max_terms=6):
    rnd = random.Random(seed)
    small = lambda: rnd.choice([0, 1, 2, L - 1, rnd.randrange(1 << 16), rnd.randrange(L), (1 << 255) - 1 - rnd.randrange(5)])
    v = [small() % L if rnd.random() < 0.8 else (1 << 255) - 19 for _ in range(m)]       # some Scalar::from_bits values >= l
    vbl = [rnd.randrange(L) for _ in range(m)]
    aL = [small() for _ in range(n)]
    aR = [small() for _ in range(n)]
    aO = [(l % L) * (r % L) % L for l, r in zip(aL, aR)]
    values = {(K_COMMITTED, i): v[i] % L for i in range(m)}
    values.update({(K_LEFT, i): aL[i] % L for i in range(n)})
    values.update({(K_RIGHT, i): aR[i] % L for i in range(n)})
    values.update({(K_OUT, i): aO[i] for i in range(n)})
    keys = list(values)
    cons = []
    for _ in range(q):
        terms = []
        for _ in range(rnd.randrange(0, max_terms + 1)):
            var = rnd.choice(keys) if keys else (K_ONE, 0)
            terms.append((var, rnd.choice([1, L - 1, rnd.randrange(L), 2**255 - 1 - rnd.randrange(3)])))
        total = sum((c % L) * values[var] for var, c in terms) % L
        terms.append(((K_ONE, 0), (-total) % L))               # complete the constraint: lc - value = 0
        if rnd.random() < 0.3:
            terms.append(((K_ONE, 0), 0))                      # Scalar::zero().into() terms occur in the gadgets
        rnd.shuffle(terms)
        cons.append(terms)
    return v, vbl, aL, aR, aO, cons


def gpu_prove_ops(bpg, ctx, label, v, vbl, aL, aR, cons, seed):
    p = bpg.Prover(ctx, bpg.Transcript(label))
    coms = [p.commit(x, b)[0] for x, b in zip(v, vbl)]
    for l, r in zip(aL, aR):
        p.allocate_multiplier((l, r))
    for terms in cons:
        p.constrain(bpg.LinearCombination([(bpg.Variable.make(k, i), c) for (k, i), c in terms]))
    return p.prove(seed), coms


def gpu_verify_ops(bpg, ctx, label, coms, n, cons, proof, seed=b"\x09" * 32):
    vf = bpg.Verifier(ctx, bpg.Transcript(label))
    for c in coms:
        vf.commit(c)
    for _ in range(n):
        vf.allocate_multiplier()
    for terms in cons:
        vf.constrain(bpg.LinearCombination([(bpg.Variable.make(k, i), c) for (k, i), c in terms]))
    return vf.verify(proof, seed)


@pytest.mark.parametrize("seed,n,m,q", [(1, 0, 1, 3), (2, 1, 0, 2), (3, 5, 2, 9), (4, 8, 3, 20), (5, 13, 4, 40), (6, 33, 1, 70),
                                        (7, 64, 6, 10), (8, 3, 0, 0), (9, 100, 9, 300)])
def test_random_system_bytes_identical(ctx, seed, n, m, q):
    import bulletproof_gadgets_b200 as bpg
    v, vbl, aL, aR, aO, cons = random_system(seed, n, m, q)
    label = b"random-%d" % seed
    prove_seed = bytes([seed]) * 32
    p_c, coms_c = coracle.prove(label, v, vbl, aL, aR, aO, cons, prove_seed)
    p_g, coms_g = gpu_prove_ops(bpg, ctx, label, v, vbl, aL, aR, cons, prove_seed)
    assert coms_g == coms_c
    assert p_g == p_c
    assert gpu_verify_ops(bpg, ctx, label, coms_g, n, cons, p_g) is True
    assert coracle.verify(label, coms_g, n, cons, p_g, b"\x09" * 32) is True
    if n <= 13:                                                # the big-int python oracle as third implementation
        T = OTranscript(label)
        po = O.Prover(O.PedersenGens(), T)
        for x, b in zip(v, vbl):
            po.commit(x, b)
        for l, r in zip(aL, aR):
            po.allocate_multiplier((l, r))
        for terms in cons:
            po.constrain(O.LC([(O.Variable(k, i), c) for (k, i), c in terms]))
        assert po.prove(O.BulletproofGens(O.round_pow2(n) if n else 1), prove_seed).to_bytes() == p_g
    # one perturbed coefficient on a non-constant term: the statement is (almost surely) false now
    cand = [(j, t) for j in range(q) for t in range(len(cons[j])) if cons[j][t][0][0] != K_ONE]
    if cand:
        j, t = cand[len(cand) // 2]
        bad = [list(x) for x in cons]
        var, c = bad[j][t]
        bad[j][t] = (var, (c + 1) % L)
        p_bad, coms_bad = gpu_prove_ops(bpg, ctx, label, v, vbl, aL, aR, bad, prove_seed)
        assert p_bad != p_g                                     # w_L / w_R / w_O / w_V changed, hence T_1.. and everything after
        r_g = gpu_verify_ops(bpg, ctx, label, coms_bad, n, bad, p_bad)
        r_c = coracle.verify(label, coms_bad, n, bad, p_bad, b"\x09" * 32)
        assert r_g == r_c                                       # accepts only if the perturbed variable's value is 0
        values_zero = {K_COMMITTED: v, K_LEFT: aL, K_RIGHT: aR, K_OUT: aO}[var[0]][var[1]] % L == 0
        assert r_g is values_zero
