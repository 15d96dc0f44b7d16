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
#if !defined(__CUDA_ARCH__) && defined(__SIZEOF_INT128__) && !defined(BPG_SC_PORTABLE)
    unsigned __int128 bw = 0;  // host: four 64-bit limbs
    for (int i = 0; i < 4; i++) {
        const uint64_t x = (uint64_t)a.v[2 * i] | ((uint64_t)a.v[2 * i + 1] << 32), y = (uint64_t)b.v[2 * i] | ((uint64_t)b.v[2 * i + 1] << 32);
        const unsigned __int128 d = (unsigned __int128)x - y - bw;
        r->v[2 * i] = (uint32_t)d;
        r->v[2 * i + 1] = (uint32_t)((uint64_t)d >> 32);
        bw = (d >> 64) & 1;
    }
    return (uint32_t)bw;
#else
    int64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += (int64_t)a.v[i] - (int64_t)b.v[i];
        r->v[i] = (uint32_t)c;
        c >>= 32;
    }
    return (uint32_t)(c & 1);
#endif
}
BPG_HD uint32_t sc_add_raw(sc* r, const sc& a, const sc& b) {
#if !defined(__CUDA_ARCH__) && defined(__SIZEOF_INT128__) && !defined(BPG_SC_PORTABLE)
    unsigned __int128 cy = 0;
    for (int i = 0; i < 4; i++) {
        const uint64_t x = (uint64_t)a.v[2 * i] | ((uint64_t)a.v[2 * i + 1] << 32), y = (uint64_t)b.v[2 * i] | ((uint64_t)b.v[2 * i + 1] << 32);
        cy += (unsigned __int128)x + y;
        r->v[2 * i] = (uint32_t)cy;
        r->v[2 * i + 1] = (uint32_t)((uint64_t)cy >> 32);
        cy >>= 64;
    }
    return (uint32_t)cy;
#else
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a.v[i] + b.v[i];
        r->v[i] = (uint32_t)c;
        c >>= 32;
    }
    return (uint32_t)c;
#endif
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

#if !defined(__CUDA_ARCH__) && defined(__SIZEOF_INT128__)
// Host-side Montgomery product on 4 x 64-bit limbs: the same integer (T + m*l) / 2^256 as the 32-bit limb loop below
// (Montgomery reduction does not depend on the limb size), about 15x faster with 64x64->128 multiplies.  The statement
// front end spends most of its time here (MiMC witnesses: ~10^5 products for a depth-32 Merkle statement).
constexpr uint64_t sc_lfactor64() {  // -l^{-1} mod 2^64 by Newton iteration from the 32-bit constant's source
    const uint32_t Ll[8] = SC_L_LIMBS;
    const uint64_t l0 = (uint64_t)Ll[0] | ((uint64_t)Ll[1] << 32);
    uint64_t inv = l0;  // correct to 3 bits
    for (int k = 0; k < 6; k++) inv *= 2 - l0 * inv;
    return 0 - inv;
}
inline sc sc_montmul_host64(const sc& a, const sc& b) {
    typedef unsigned __int128 u128;
    const uint32_t Ll[8] = SC_L_LIMBS;
    const uint64_t L0 = (uint64_t)Ll[0] | ((uint64_t)Ll[1] << 32), L1 = (uint64_t)Ll[2] | ((uint64_t)Ll[3] << 32);
    const uint64_t L3 = (uint64_t)Ll[6] | ((uint64_t)Ll[7] << 32);  // limb 2 of l is zero
    constexpr uint64_t LF = sc_lfactor64();
    uint64_t A[4], B[4], t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        A[i] = (uint64_t)a.v[2 * i] | ((uint64_t)a.v[2 * i + 1] << 32);
        B[i] = (uint64_t)b.v[2 * i] | ((uint64_t)b.v[2 * i + 1] << 32);
    }
    for (int i = 0; i < 4; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 4; j++) {
            const u128 x = (u128)A[i] * B[j] + t[i + j] + c;
            t[i + j] = (uint64_t)x;
            c = (uint64_t)(x >> 64);
        }
        t[i + 4] = c;
    }
    uint64_t hc = 0;
    for (int i = 0; i < 4; i++) {
        const uint64_t m = t[i] * LF;
        u128 x = (u128)m * L0 + t[i];
        uint64_t c = (uint64_t)(x >> 64);
        x = (u128)m * L1 + t[i + 1] + c;
        t[i + 1] = (uint64_t)x;
        c = (uint64_t)(x >> 64);
        x = (u128)t[i + 2] + c;
        t[i + 2] = (uint64_t)x;
        c = (uint64_t)(x >> 64);
        x = (u128)m * L3 + t[i + 3] + c;
        t[i + 3] = (uint64_t)x;
        c = (uint64_t)(x >> 64);
        x = (u128)t[i + 4] + c + hc;
        t[i + 4] = (uint64_t)x;
        hc = (uint64_t)(x >> 64);
    }
    sc r;
    for (int i = 0; i < 4; i++) {
        r.v[2 * i] = (uint32_t)t[4 + i];
        r.v[2 * i + 1] = (uint32_t)(t[4 + i] >> 32);
    }
    return r;  // caller subtracts l once
}
#define BPG_SC_HOST64 1
#endif

// Montgomery product a*b/R mod l, canonical provided a*b < R*l
BPG_HD sc sc_montmul(const sc& a, const sc& b) {
#if defined(BPG_SC_HOST64) && !defined(__CUDA_ARCH__) && !defined(BPG_SC_PORTABLE)
    return sc_csub_l(sc_montmul_host64(a, b));
#else
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
#endif
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
