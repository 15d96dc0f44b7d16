#!/usr/bin/env python3
"""Generates (and checks) the carry-chain schedule of the dedicated field squaring in csrc/fe25519.cuh.

a^2 = sum_i a_i^2 2^(64 i) + 2 * sum_{i<j} a_i a_j 2^(32 (i+j)):  28 cross products + 8 squares = 36 IMAD.WIDE instead
of the 64 of the general product.  The cross products use the same even/odd column accumulators as fe_mul_wide: for a
fixed multiplier a_i, the partners j = i+1, i+3, .. land on consecutive (lo, hi) limb pairs of the odd array O, and
j = i+2, i+4, .. on the even array E, so every row is one `mad.lo.cc / madc.hi.cc` chain.

    python tools/gen_fe_sqr.py            simulates the schedule limb by limb on random inputs (asserting that no carry-out
                                          limb overflows) and prints the C code of fe_sqr_wide's device branch."""
import random

M = (1 << 32) - 1


def schedule():
    """list of (array, start, i, [j..])"""
    rows = []
    for i in range(7):
        odd = list(range(i + 1, 8, 2))      # columns 2i+1, 2i+3, ..  -> O[2i ..]
        even = list(range(i + 2, 8, 2))     # columns 2i+2, ..        -> E[2i+2 ..]
        if odd:
            rows.append(("O", 2 * i, i, odd))
        if even:
            rows.append(("E", 2 * i + 2, i, even))
    return rows


def simulate(a):
    E, O = [0] * 17, [0] * 16
    for arr, start, i, js in schedule():
        acc = E if arr == "E" else O
        carry, k = 0, start
        for j in js:
            p = a[i] * a[j]
            s = acc[k] + (p & M) + carry
            acc[k], carry = s & M, s >> 32
            s = acc[k + 1] + (p >> 32) + carry
            acc[k + 1], carry = s & M, s >> 32
            k += 2
        s = acc[k] + carry
        assert s <= M, "carry-out limb overflow"
        acc[k] = s
    # C = E + (O << 32)
    r = [E[0]] + [0] * 15
    carry = 0
    for k in range(1, 16):
        s = E[k] + O[k - 1] + carry
        r[k], carry = s & M, s >> 32
    assert carry == 0 and E[16] == 0 and O[15] == 0
    # 2C
    top = r[15] >> 31
    assert top == 0
    for k in range(15, 0, -1):
        r[k] = ((r[k] << 1) | (r[k - 1] >> 31)) & M
    r[0] = (r[0] << 1) & M
    # + squares, one carry chain over the 16 limbs
    carry = 0
    for i in range(8):
        p = a[i] * a[i]
        s = r[2 * i] + (p & M) + carry
        r[2 * i], carry = s & M, s >> 32
        s = r[2 * i + 1] + (p >> 32) + carry
        r[2 * i + 1], carry = s & M, s >> 32
    assert carry == 0
    return r


def check(n=20000):
    rnd = random.Random(1)
    specials = [[M] * 8, [0] * 8, [1] + [0] * 7, [M] + [0] * 7, [0] * 7 + [M]]
    for t in range(n):
        a = specials[t] if t < len(specials) else [rnd.choice([0, 1, M, M - 1, rnd.getrandbits(32)]) for _ in range(8)]
        r = simulate(a)
        want = sum(x << (32 * i) for i, x in enumerate(a)) ** 2
        got = sum(x << (32 * i) for i, x in enumerate(r))
        assert got == want, a


def emit():
    macro = {1: "FE_MADROW3", 2: "FE_MADROW5", 3: "FE_MADROW7", 4: "FE_MADROW9"}
    out = []
    for arr, start, i, js in schedule():
        args = ", ".join("a.v[%d]" % j for j in js)
        out.append("    %s(%s + %d, %s, a.v[%d]);" % (macro[len(js)], arr, start, args, i))
    return "\n".join(out)


if __name__ == "__main__":
    check()
    print("// schedule verified on 20000 inputs")
    print(emit())
