// Scalar field mod l = 2^252 + 27742317777372353535851937790883648493 on 8 x 32-bit limbs.
//
// Replaces curve25519-dalek 3.2.0 `Scalar` / `Scalar52` (scalar.rs, backend/serial/u64/scalar.rs:
// montgomery_mul, add, sub, invert; /root/reference/Cargo.lock:155-157, not vendored) --
// SURVEY.md row K2.  Same strategy as dalek: a*b = montmul(montmul(a,b), R^2), R = 2^256.
// Inputs may be any value < 2^255 (Scalar::from_bits, /root/reference/src/conversions.rs:18);
// outputs are canonical (< l).
#pragma once
#include "consts.cuh"
#include "fe25519.cuh"

struct sc {
    uint32_t v[8];
};

BPG_HD sc sc_zero() {
    sc r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
}
BPG_HD sc sc_one() {
    sc r = sc_zero();
    r.v[0] = 1;
    return r;
}
BPG_HD sc sc_const(const uint32_t (&l)[8]) {
    sc r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = l[i];
    return r;
}
BPG_HD sc sc_L() {
    const uint32_t l[8] = SC_L_LIMBS;
    return sc_const(l);
}
BPG_HD sc sc_RR() {
    const uint32_t l[8] = SC_RR_LIMBS;
    return sc_const(l);
}

// r = a - b, returns borrow (0/1)
BPG_HD uint32_t sc_sub_raw(sc* r, const sc& a, const sc& b) {
    int64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += (int64_t)a.v[i] - (int64_t)b.v[i];
        r->v[i] = (uint32_t)c;
        c >>= 32;
    }
    return (uint32_t)(c & 1);
}
BPG_HD uint32_t sc_add_raw(sc* r, const sc& a, const sc& b) {
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a.v[i] + b.v[i];
        r->v[i] = (uint32_t)c;
        c >>= 32;
    }
    return (uint32_t)c;
}
BPG_HD sc sc_select(bool c, const sc& a, const sc& b) {
    sc r;
    uint32_t m = 0u - (uint32_t)c;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = (a.v[i] & m) | (b.v[i] & ~m);
    return r;
}
// x < 2l  ->  x mod l
BPG_HD sc sc_csub_l(const sc& x) {
    sc t;
    uint32_t bw = sc_sub_raw(&t, x, sc_L());
    return sc_select(bw != 0, x, t);
}

// Montgomery product a*b/R mod l, canonical provided a*b < R*l
BPG_HD sc sc_montmul(const sc& a, const sc& b) {
    uint32_t t[17];
    {
        fe fa, fb;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            fa.v[i] = a.v[i];
            fb.v[i] = b.v[i];
        }
        fe_mul_wide(t, fa, fb);
    }
    t[16] = 0;
    const uint32_t Ll[8] = SC_L_LIMBS;
    uint32_t hc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t m = t[i] * SC_LFACTOR;
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (Ll[j] == 0) {  // limbs 4..6 of l are zero: only the carry moves
                c += t[i + j];
            } else {
                c += (uint64_t)m * Ll[j] + t[i + j];
            }
            t[i + j] = (uint32_t)c;
            c >>= 32;
        }
        c += (uint64_t)t[i + 8] + hc;
        t[i + 8] = (uint32_t)c;
        hc = (uint32_t)(c >> 32);
    }
    sc r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[8 + i];
    return sc_csub_l(r);
}

BPG_HD sc sc_mul(const sc& a, const sc& b) { return sc_montmul(sc_montmul(a, b), sc_RR()); }
// canonical representative of any x < 2^256
BPG_HD sc sc_reduce(const sc& x) { return sc_mul(x, sc_one()); }

BPG_HD sc sc_add(const sc& a, const sc& b) {  // canonical inputs
    sc r;
    sc_add_raw(&r, a, b);
    return sc_csub_l(r);
}
BPG_HD sc sc_sub(const sc& a, const sc& b) {  // canonical inputs
    sc r, t;
    uint32_t bw = sc_sub_raw(&r, a, b);
    sc_add_raw(&t, r, sc_L());
    return sc_select(bw != 0, t, r);
}
BPG_HD sc sc_neg(const sc& a) { return sc_sub(sc_zero(), a); }
BPG_HD bool sc_is_zero(const sc& a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= a.v[i];
    return o == 0;
}
