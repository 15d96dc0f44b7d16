// Host-side scalar helpers (mod l) built on the portable branch of sc25519.cuh.  Used by the
// protocol driver for the handful of per-proof challenge computations; vectors live on the GPU.
#pragma once
#include <string.h>

#include "sc25519.cuh"

namespace bpg {

struct Scalar {
    sc s;
    Scalar() { s = sc_zero(); }
    explicit Scalar(uint64_t x) {
        s = sc_zero();
        s.v[0] = (uint32_t)x;
        s.v[1] = (uint32_t)(x >> 32);
    }
    static Scalar from_sc(const sc& x) {
        Scalar r;
        r.s = x;
        return r;
    }
    // raw 32 bytes, no reduction (Scalar::from_bits semantics; top bit must be clear)
    static Scalar from_bytes_raw(const uint8_t b[32]) {
        Scalar r;
        memcpy(r.s.v, b, 32);
        return r;
    }
    // canonical representative of a 256-bit value
    static Scalar from_bytes_mod_order(const uint8_t b[32]) {
        Scalar r = from_bytes_raw(b);
        if (!r.is_canonical()) r.s = sc_reduce(r.s);  // almost every input already is: skip the two Montgomery products
        return r;
    }
    // Scalar::from_bytes_mod_order_wide
    static Scalar from_bytes_wide(const uint8_t b[64]) {
        sc lo, hi;
        memcpy(lo.v, b, 32);
        memcpy(hi.v, b + 32, 32);
        const uint32_t Rl[8] = SC_R_LIMBS;
        Scalar r;
        r.s = sc_add(sc_reduce(lo), sc_mul(hi, sc_const(Rl)));
        return r;
    }
    bool is_canonical() const {
        sc t;
        return sc_sub_raw(&t, s, sc_L()) != 0;  // s < l
    }
    void to_bytes(uint8_t out[32]) const { memcpy(out, s.v, 32); }
    bool is_zero() const { return sc_is_zero(s); }
    Scalar operator+(const Scalar& o) const { return from_sc(sc_add(s, o.s)); }
    Scalar operator-(const Scalar& o) const { return from_sc(sc_sub(s, o.s)); }
    Scalar operator*(const Scalar& o) const { return from_sc(sc_mul(s, o.s)); }
    Scalar operator-() const { return from_sc(sc_neg(s)); }
    // s^(l-2) (dalek Scalar::invert; 0 -> 0).  Stays in Montgomery form for the whole ladder (one Montgomery product
    // per step instead of two) with a 4-bit fixed window: 252 squarings + 63 + 14 products.
    Scalar invert() const {
        const uint32_t Ll[8] = SC_L_LIMBS;
        uint32_t e[8];
        for (int i = 0; i < 8; i++) e[i] = Ll[i];
        e[0] -= 2;  // l - 2 (no borrow: low limb of l is 0x5cf5d3ed)
        const sc rr = sc_RR();
        sc tbl[16];                                   // tbl[k] = s^k * R
        tbl[0] = sc_montmul(sc_one(), rr);            // R mod l
        tbl[1] = sc_montmul(sc_reduce(s), rr);        // s * R
        for (int k = 2; k < 16; k++) tbl[k] = sc_montmul(tbl[k - 1], tbl[1]);
        sc acc = tbl[(e[7] >> 28) & 15];
        for (int nib = 62; nib >= 0; nib--) {
            acc = sc_montmul(acc, acc);
            acc = sc_montmul(acc, acc);
            acc = sc_montmul(acc, acc);
            acc = sc_montmul(acc, acc);
            const uint32_t d = (e[nib >> 3] >> (4 * (nib & 7))) & 15;
            if (d) acc = sc_montmul(acc, tbl[d]);
        }
        return from_sc(sc_montmul(acc, sc_one()));    // out of Montgomery form
    }
};

}  // namespace bpg
