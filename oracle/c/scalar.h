/* ORACLE (test infrastructure / CPU baseline; never linked into the product).
 * Scalars mod l and the digit recodings of curve25519-dalek 3.2.0 scalar.rs
 * (/root/reference/Cargo.lock:155-157, not vendored): Montgomery multiplication (dalek uses
 * 5 x 52-bit limbs; this uses 4 x 64 with u128 -- same values), from_bytes_mod_order_wide,
 * invert, to_radix_16, non_adjacent_form(w), to_radix_2w(w).
 */
#ifndef BPO_SCALAR_H
#define BPO_SCALAR_H
#include <stdint.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t v[4]; } scl;

static const scl SC_L = {{0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0x0000000000000000ULL, 0x1000000000000000ULL}};
static const scl SC_RR = {{0xa40611e3449c0f01ULL, 0xd00e1ba768859347ULL, 0xceec73d217f5be65ULL, 0x0399411b7c309a3dULL}};
static const scl SC_R = {{0xd6ec31748d98951dULL, 0xc6ef5bf4737dcf70ULL, 0xfffffffffffffffeULL, 0x0fffffffffffffffULL}};
#define SC_LFACTOR 0xd2b51da312547e1bULL
static const scl SC_ZERO = {{0, 0, 0, 0}};
static const scl SC_ONE = {{1, 0, 0, 0}};

static inline uint64_t scl_sub_raw(scl* r, scl a, scl b) {
    u128 bw = 0;
    for (int i = 0; i < 4; i++) {
        u128 t = (u128)a.v[i] - b.v[i] - (uint64_t)bw;
        r->v[i] = (uint64_t)t;
        bw = (t >> 64) & 1;
    }
    return (uint64_t)bw;
}
static inline uint64_t scl_add_raw(scl* r, scl a, scl b) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (u128)a.v[i] + b.v[i];
        r->v[i] = (uint64_t)c;
        c >>= 64;
    }
    return (uint64_t)c;
}
static inline scl scl_csub(scl x) {
    scl t;
    return scl_sub_raw(&t, x, SC_L) ? x : t;
}
static inline scl scl_montmul(scl a, scl b) {
    uint64_t t[9] = {0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a.v[j] * b.v[i] + t[i + j];
            t[i + j] = (uint64_t)c;
            c >>= 64;
        }
        t[i + 4] = (uint64_t)c;
    }
    uint64_t hc = 0;
    for (int i = 0; i < 4; i++) {
        uint64_t m = t[i] * SC_LFACTOR;
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)m * SC_L.v[j] + t[i + j];
            t[i + j] = (uint64_t)c;
            c >>= 64;
        }
        c += (u128)t[i + 4] + hc;
        t[i + 4] = (uint64_t)c;
        hc = (uint64_t)(c >> 64);
    }
    scl r = {{t[4], t[5], t[6], t[7]}};
    return scl_csub(r);
}
static inline scl scl_mul(scl a, scl b) { return scl_montmul(scl_montmul(a, b), SC_RR); }
static inline scl scl_reduce(scl a) { return scl_mul(a, SC_ONE); }
static inline scl scl_add(scl a, scl b) {
    scl r;
    scl_add_raw(&r, a, b);
    return scl_csub(r);
}
static inline scl scl_sub(scl a, scl b) {
    scl r, t;
    if (scl_sub_raw(&r, a, b)) {
        scl_add_raw(&t, r, SC_L);
        return t;
    }
    return r;
}
static inline scl scl_neg(scl a) { return scl_sub(SC_ZERO, a); }
static inline scl scl_from_bytes(const uint8_t* b) {
    scl r;
    memcpy(r.v, b, 32);
    return r;
}
static inline void scl_to_bytes(uint8_t* b, scl a) { memcpy(b, a.v, 32); }
static inline scl scl_from_wide(const uint8_t* b) {
    scl lo, hi;
    memcpy(lo.v, b, 32);
    memcpy(hi.v, b + 32, 32);
    return scl_add(scl_reduce(lo), scl_mul(hi, SC_R));
}
static inline int scl_is_canonical(scl a) {
    scl t;
    return scl_sub_raw(&t, a, SC_L) != 0;
}
static inline int scl_is_zero(scl a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3]) == 0; }
static inline scl scl_invert(scl a) {
    scl e = SC_L;
    e.v[0] -= 2;
    scl acc = SC_ONE;
    for (int i = 255; i >= 0; i--) {
        acc = scl_mul(acc, acc);
        if ((e.v[i >> 6] >> (i & 63)) & 1) acc = scl_mul(acc, a);
    }
    return acc;
}
static inline scl scl_from_u64(uint64_t x) {
    scl r = {{x, 0, 0, 0}};
    return r;
}

/* Scalar::to_radix_16: 64 digits in [-8, 8) (top digit up to 8) */
static inline void scl_to_radix_16(int8_t out[64], const uint8_t s[32]) {
    for (int i = 0; i < 32; i++) {
        out[2 * i] = s[i] & 15;
        out[2 * i + 1] = (s[i] >> 4) & 15;
    }
    for (int i = 0; i < 63; i++) {
        int8_t carry = (int8_t)((out[i] + 8) >> 4);
        out[i] -= (int8_t)(carry << 4);
        out[i + 1] += carry;
    }
}
/* Scalar::non_adjacent_form(w): 256 digits, odd, |d| < 2^(w-1) */
static inline void scl_naf(int8_t naf[256], const uint8_t s[32], int w) {
    memset(naf, 0, 256);
    uint64_t x[5] = {0};
    memcpy(x, s, 32);
    const uint64_t width = 1ULL << w, window_mask = width - 1;
    int pos = 0;
    uint64_t carry = 0;
    while (pos < 256) {
        int idx = pos / 64, bit = pos % 64;
        uint64_t bit_buf = bit < 64 - w ? x[idx] >> bit : (x[idx] >> bit) | (x[idx + 1] << (64 - bit));
        uint64_t window = carry + (bit_buf & window_mask);
        if ((window & 1) == 0) {
            pos += 1;
            continue;
        }
        if (window < width / 2) {
            carry = 0;
            naf[pos] = (int8_t)window;
        } else {
            carry = 1;
            naf[pos] = (int8_t)((int64_t)window - (int64_t)width);
        }
        pos += w;
    }
}
/* Scalar::to_radix_2w(w), w in 6..8: digits in [-2^w/2, 2^w/2) */
static inline int scl_radix_2w_size(int w) { return w == 8 ? 33 : (256 + w - 1) / w; }
static inline void scl_to_radix_2w(int8_t digits[43], const uint8_t s[32], int w) {
    uint64_t x[5] = {0};
    memcpy(x, s, 32);
    const uint64_t radix = 1ULL << w, window_mask = radix - 1;
    uint64_t carry = 0;
    memset(digits, 0, 43);
    const int digits_count = (256 + w - 1) / w;
    for (int i = 0; i < digits_count; i++) {
        int bit_offset = i * w, idx = bit_offset / 64, bit = bit_offset % 64;
        uint64_t bit_buf = (bit < 64 - w || idx == 3) ? x[idx] >> bit : (x[idx] >> bit) | (x[idx + 1] << (64 - bit));
        uint64_t coef = carry + (bit_buf & window_mask);
        carry = (coef + radix / 2) >> w;
        digits[i] = (int8_t)((int64_t)coef - (int64_t)(carry << w));
    }
    if (w == 8) digits[digits_count] += (int8_t)carry;
    else digits[digits_count - 1] += (int8_t)(carry << w);
}
#endif
