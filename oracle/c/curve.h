/* ORACLE (test infrastructure / CPU baseline; never linked into the product).
 *
 * CPU restatement of curve25519-dalek 3.2.0's SERIAL u64 backend (the backend the reference ships
 * with: feature avx2_backend is on at /root/reference/Cargo.toml:22 but no target-feature flag is
 * set anywhere) -- crate pinned at /root/reference/Cargo.lock:155-157, source NOT
 * vendored, so this follows the published algorithms:
 *   backend/serial/u64/field.rs   FieldElement51 (5 x 51-bit limbs, u128 products)
 *   backend/serial/curve_models   EdwardsPoint / ProjectivePoint / CompletedPoint / (Projective|Affine)Niels
 *   ristretto.rs                  compress / decompress / elligator (RFC 9496)
 *   scalar.rs                     Scalar mod l, radix-16 / NAF / radix-2^w recodings
 * Parity: UNPINNED against dalek itself; pinned against RFC 9496 vectors and the python oracle
 * (tests/test_oracle_c.py).
 */
#ifndef BPO_CURVE_H
#define BPO_CURVE_H
#include <stdint.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t v[5]; } fe51;

#define MASK51 ((1ULL << 51) - 1)

static const fe51 FE_ZERO = {{0, 0, 0, 0, 0}};
static const fe51 FE_ONE = {{1, 0, 0, 0, 0}};
static const fe51 FE_D = {{929955233495203ULL, 466365720129213ULL, 1662059464998953ULL, 2033849074728123ULL, 1442794654840575ULL}};
static const fe51 FE_D2 = {{1859910466990425ULL, 932731440258426ULL, 1072319116312658ULL, 1815898335770999ULL, 633789495995903ULL}};
static const fe51 FE_SQRT_M1 = {{1718705420411056ULL, 234908883556509ULL, 2233514472574048ULL, 2117202627021982ULL, 765476049583133ULL}};
static const fe51 FE_SQRT_AD_MINUS_ONE = {{2241493124984347ULL, 425987919032274ULL, 2207028919301688ULL, 1220490630685848ULL, 974799131293748ULL}};
static const fe51 FE_INVSQRT_A_MINUS_D = {{278908739862762ULL, 821645201101625ULL, 8113234426968ULL, 1777959178193151ULL, 2118520810568447ULL}};
static const fe51 FE_ONE_MINUS_D_SQ = {{1136626929484150ULL, 1998550399581263ULL, 496427632559748ULL, 118527312129759ULL, 45110755273534ULL}};
static const fe51 FE_D_MINUS_ONE_SQ = {{1507062230895904ULL, 1572317787530805ULL, 683053064812840ULL, 317374165784489ULL, 1572899562415810ULL}};

static inline fe51 fe_add(fe51 a, fe51 b) {
    fe51 r;
    for (int i = 0; i < 5; i++) r.v[i] = a.v[i] + b.v[i];
    return r;
}
static inline fe51 fe_weak_reduce(fe51 a) {
    uint64_t c0 = a.v[0] >> 51, c1 = a.v[1] >> 51, c2 = a.v[2] >> 51, c3 = a.v[3] >> 51, c4 = a.v[4] >> 51;
    fe51 r;
    r.v[0] = (a.v[0] & MASK51) + c4 * 19;
    r.v[1] = (a.v[1] & MASK51) + c0;
    r.v[2] = (a.v[2] & MASK51) + c1;
    r.v[3] = (a.v[3] & MASK51) + c2;
    r.v[4] = (a.v[4] & MASK51) + c3;
    return r;
}
/* a - b: add 16p first so limbs stay positive (field.rs Sub) */
static inline fe51 fe_sub(fe51 a, fe51 b) {
    fe51 r;
    r.v[0] = (a.v[0] + 36028797018963664ULL) - b.v[0];
    r.v[1] = (a.v[1] + 36028797018963952ULL) - b.v[1];
    r.v[2] = (a.v[2] + 36028797018963952ULL) - b.v[2];
    r.v[3] = (a.v[3] + 36028797018963952ULL) - b.v[3];
    r.v[4] = (a.v[4] + 36028797018963952ULL) - b.v[4];
    return fe_weak_reduce(r);
}
static inline fe51 fe_neg(fe51 a) { return fe_sub(FE_ZERO, a); }

static inline fe51 fe_mul(fe51 x, fe51 y) {
    const uint64_t *a = x.v, *b = y.v;
    uint64_t b1_19 = b[1] * 19, b2_19 = b[2] * 19, b3_19 = b[3] * 19, b4_19 = b[4] * 19;
    u128 c0 = (u128)a[0] * b[0] + (u128)a[4] * b1_19 + (u128)a[3] * b2_19 + (u128)a[2] * b3_19 + (u128)a[1] * b4_19;
    u128 c1 = (u128)a[1] * b[0] + (u128)a[0] * b[1] + (u128)a[4] * b2_19 + (u128)a[3] * b3_19 + (u128)a[2] * b4_19;
    u128 c2 = (u128)a[2] * b[0] + (u128)a[1] * b[1] + (u128)a[0] * b[2] + (u128)a[4] * b3_19 + (u128)a[3] * b4_19;
    u128 c3 = (u128)a[3] * b[0] + (u128)a[2] * b[1] + (u128)a[1] * b[2] + (u128)a[0] * b[3] + (u128)a[4] * b4_19;
    u128 c4 = (u128)a[4] * b[0] + (u128)a[3] * b[1] + (u128)a[2] * b[2] + (u128)a[1] * b[3] + (u128)a[0] * b[4];
    fe51 r;
    c1 += (uint64_t)(c0 >> 51); r.v[0] = (uint64_t)c0 & MASK51;
    c2 += (uint64_t)(c1 >> 51); r.v[1] = (uint64_t)c1 & MASK51;
    c3 += (uint64_t)(c2 >> 51); r.v[2] = (uint64_t)c2 & MASK51;
    c4 += (uint64_t)(c3 >> 51); r.v[3] = (uint64_t)c3 & MASK51;
    uint64_t carry = (uint64_t)(c4 >> 51); r.v[4] = (uint64_t)c4 & MASK51;
    r.v[0] += carry * 19;
    r.v[1] += r.v[0] >> 51;
    r.v[0] &= MASK51;
    return r;
}
static inline fe51 fe_sq(fe51 x) {
    const uint64_t* a = x.v;
    uint64_t a3_19 = 19 * a[3], a4_19 = 19 * a[4];
    u128 c0 = (u128)a[0] * a[0] + 2 * ((u128)a[1] * a4_19 + (u128)a[2] * a3_19);
    u128 c1 = (u128)a[3] * a3_19 + 2 * ((u128)a[0] * a[1] + (u128)a[2] * a4_19);
    u128 c2 = (u128)a[1] * a[1] + 2 * ((u128)a[0] * a[2] + (u128)a[4] * a3_19);
    u128 c3 = (u128)a[4] * a4_19 + 2 * ((u128)a[0] * a[3] + (u128)a[1] * a[2]);
    u128 c4 = (u128)a[2] * a[2] + 2 * ((u128)a[0] * a[4] + (u128)a[1] * a[3]);
    fe51 r;
    c1 += (uint64_t)(c0 >> 51); r.v[0] = (uint64_t)c0 & MASK51;
    c2 += (uint64_t)(c1 >> 51); r.v[1] = (uint64_t)c1 & MASK51;
    c3 += (uint64_t)(c2 >> 51); r.v[2] = (uint64_t)c2 & MASK51;
    c4 += (uint64_t)(c3 >> 51); r.v[3] = (uint64_t)c3 & MASK51;
    uint64_t carry = (uint64_t)(c4 >> 51); r.v[4] = (uint64_t)c4 & MASK51;
    r.v[0] += carry * 19;
    r.v[1] += r.v[0] >> 51;
    r.v[0] &= MASK51;
    return r;
}
static inline fe51 fe_sq2(fe51 x) {
    fe51 r = fe_sq(x);
    for (int i = 0; i < 5; i++) r.v[i] *= 2;
    return r;
}
static inline fe51 fe_pow2k(fe51 x, int k) {
    for (int i = 0; i < k; i++) x = fe_sq(x);
    return x;
}
static inline fe51 fe_from_bytes(const uint8_t* s) {
    uint64_t w[4];
    memcpy(w, s, 32);
    fe51 r;
    r.v[0] = w[0] & MASK51;
    r.v[1] = ((w[0] >> 51) | (w[1] << 13)) & MASK51;
    r.v[2] = ((w[1] >> 38) | (w[2] << 26)) & MASK51;
    r.v[3] = ((w[2] >> 25) | (w[3] << 39)) & MASK51;
    r.v[4] = (w[3] >> 12) & MASK51; /* top bit ignored */
    return r;
}
static inline void fe_to_bytes(uint8_t* s, fe51 a) {
    fe51 h = fe_weak_reduce(a);
    uint64_t q = (h.v[0] + 19) >> 51;
    q = (h.v[1] + q) >> 51;
    q = (h.v[2] + q) >> 51;
    q = (h.v[3] + q) >> 51;
    q = (h.v[4] + q) >> 51;
    h.v[0] += 19 * q;
    h.v[1] += h.v[0] >> 51; h.v[0] &= MASK51;
    h.v[2] += h.v[1] >> 51; h.v[1] &= MASK51;
    h.v[3] += h.v[2] >> 51; h.v[2] &= MASK51;
    h.v[4] += h.v[3] >> 51; h.v[3] &= MASK51;
    h.v[4] &= MASK51;
    uint64_t w[4];
    w[0] = h.v[0] | (h.v[1] << 51);
    w[1] = (h.v[1] >> 13) | (h.v[2] << 38);
    w[2] = (h.v[2] >> 26) | (h.v[3] << 25);
    w[3] = (h.v[3] >> 39) | (h.v[4] << 12);
    memcpy(s, w, 32);
}
static inline int fe_is_zero(fe51 a) {
    uint8_t b[32];
    fe_to_bytes(b, a);
    uint8_t o = 0;
    for (int i = 0; i < 32; i++) o |= b[i];
    return o == 0;
}
static inline int fe_is_neg(fe51 a) {
    uint8_t b[32];
    fe_to_bytes(b, a);
    return b[0] & 1;
}
static inline int fe_eq(fe51 a, fe51 b) {
    uint8_t x[32], y[32];
    fe_to_bytes(x, a);
    fe_to_bytes(y, b);
    return memcmp(x, y, 32) == 0;
}
static inline fe51 fe_cneg(fe51 a, int c) { return c ? fe_neg(a) : a; }
static inline fe51 fe_abs(fe51 a) { return fe_cneg(a, fe_is_neg(a)); }

/* (z^(2^250-1), z^11) -- field.rs pow22501 */
static inline fe51 fe_pow22501(fe51 z, fe51* z11) {
    fe51 t0 = fe_sq(z);
    fe51 t1 = fe_sq(fe_sq(t0));
    fe51 t2 = fe_mul(z, t1);
    fe51 t3 = fe_mul(t0, t2);
    fe51 t4 = fe_sq(t3);
    fe51 t5 = fe_mul(t2, t4);
    fe51 t6 = fe_pow2k(t5, 5);
    fe51 t7 = fe_mul(t6, t5);
    fe51 t8 = fe_pow2k(t7, 10);
    fe51 t9 = fe_mul(t8, t7);
    fe51 t10 = fe_pow2k(t9, 20);
    fe51 t11 = fe_mul(t10, t9);
    fe51 t12 = fe_pow2k(t11, 10);
    fe51 t13 = fe_mul(t12, t7);
    fe51 t14 = fe_pow2k(t13, 50);
    fe51 t15 = fe_mul(t14, t13);
    fe51 t16 = fe_pow2k(t15, 100);
    fe51 t17 = fe_mul(t16, t15);
    fe51 t18 = fe_pow2k(t17, 50);
    *z11 = t3;
    return fe_mul(t18, t13);
}
static inline fe51 fe_invert(fe51 z) {
    fe51 z11;
    fe51 t19 = fe_pow22501(z, &z11);
    return fe_mul(fe_pow2k(t19, 5), z11);
}
static inline fe51 fe_pow_p58(fe51 z) {
    fe51 z11;
    fe51 t19 = fe_pow22501(z, &z11);
    return fe_mul(fe_pow2k(t19, 2), z);
}
/* FieldElement::sqrt_ratio_i */
static inline int fe_sqrt_ratio_i(fe51* out, fe51 u, fe51 v) {
    fe51 v3 = fe_mul(fe_sq(v), v);
    fe51 v7 = fe_mul(fe_sq(v3), v);
    fe51 r = fe_mul(fe_mul(u, v3), fe_pow_p58(fe_mul(u, v7)));
    fe51 check = fe_mul(v, fe_sq(r));
    fe51 nu = fe_neg(u);
    int correct = fe_eq(check, u), flipped = fe_eq(check, nu), flipped_i = fe_eq(check, fe_mul(nu, FE_SQRT_M1));
    if (flipped || flipped_i) r = fe_mul(r, FE_SQRT_M1);
    *out = fe_abs(r);
    return correct || flipped;
}

/* ---------------------------------------------------------------------------- curve models */
typedef struct { fe51 X, Y, Z, T; } ge_p3;       /* EdwardsPoint */
typedef struct { fe51 X, Y, Z; } ge_p2;          /* ProjectivePoint */
typedef struct { fe51 X, Y, Z, T; } ge_p1p1;     /* CompletedPoint */
typedef struct { fe51 YpX, YmX, Z, T2d; } ge_pniels; /* ProjectiveNielsPoint */

static inline ge_p3 ge_identity(void) {
    ge_p3 r = {FE_ZERO, FE_ONE, FE_ONE, FE_ZERO};
    return r;
}
static inline ge_p2 ge_p2_identity(void) {
    ge_p2 r = {FE_ZERO, FE_ONE, FE_ONE};
    return r;
}
static inline ge_p2 p3_to_p2(const ge_p3* p) {
    ge_p2 r = {p->X, p->Y, p->Z};
    return r;
}
static inline ge_pniels p3_to_pniels(const ge_p3* p) {
    ge_pniels r = {fe_add(p->Y, p->X), fe_sub(p->Y, p->X), p->Z, fe_mul(p->T, FE_D2)};
    return r;
}
static inline ge_p3 p1p1_to_p3(const ge_p1p1* p) {
    ge_p3 r = {fe_mul(p->X, p->T), fe_mul(p->Y, p->Z), fe_mul(p->Z, p->T), fe_mul(p->X, p->Y)};
    return r;
}
static inline ge_p2 p1p1_to_p2(const ge_p1p1* p) {
    ge_p2 r = {fe_mul(p->X, p->T), fe_mul(p->Y, p->Z), fe_mul(p->Z, p->T)};
    return r;
}
static inline ge_p1p1 ge_add_pniels(const ge_p3* p, const ge_pniels* q) {
    fe51 PP = fe_mul(fe_add(p->Y, p->X), q->YpX), MM = fe_mul(fe_sub(p->Y, p->X), q->YmX);
    fe51 TT2d = fe_mul(p->T, q->T2d), ZZ = fe_mul(p->Z, q->Z);
    fe51 ZZ2 = fe_add(ZZ, ZZ);
    ge_p1p1 r = {fe_sub(PP, MM), fe_add(PP, MM), fe_add(ZZ2, TT2d), fe_sub(ZZ2, TT2d)};
    return r;
}
static inline ge_p1p1 ge_sub_pniels(const ge_p3* p, const ge_pniels* q) {
    fe51 PM = fe_mul(fe_add(p->Y, p->X), q->YmX), MP = fe_mul(fe_sub(p->Y, p->X), q->YpX);
    fe51 TT2d = fe_mul(p->T, q->T2d), ZZ = fe_mul(p->Z, q->Z);
    fe51 ZZ2 = fe_add(ZZ, ZZ);
    ge_p1p1 r = {fe_sub(PM, MP), fe_add(PM, MP), fe_sub(ZZ2, TT2d), fe_add(ZZ2, TT2d)};
    return r;
}
static inline ge_p1p1 ge_p2_dbl(const ge_p2* p) {
    fe51 XX = fe_sq(p->X), YY = fe_sq(p->Y), ZZ2 = fe_sq2(p->Z);
    fe51 XpY_sq = fe_sq(fe_add(p->X, p->Y));
    fe51 YYpXX = fe_add(YY, XX), YYmXX = fe_sub(YY, XX);
    ge_p1p1 r = {fe_sub(XpY_sq, YYpXX), YYpXX, YYmXX, fe_sub(ZZ2, YYmXX)};
    return r;
}
static inline ge_p3 ge_add(const ge_p3* p, const ge_p3* q) {
    ge_pniels n = p3_to_pniels(q);
    ge_p1p1 c = ge_add_pniels(p, &n);
    return p1p1_to_p3(&c);
}
static inline ge_p3 ge_dbl(const ge_p3* p) {
    ge_p2 a = p3_to_p2(p);
    ge_p1p1 c = ge_p2_dbl(&a);
    return p1p1_to_p3(&c);
}
/* EdwardsPoint::mul_by_pow_2 */
static inline ge_p3 ge_mul_by_pow_2(const ge_p3* p, int k) {
    ge_p2 s = p3_to_p2(p);
    ge_p1p1 r;
    for (int i = 0; i < k - 1; i++) {
        r = ge_p2_dbl(&s);
        s = p1p1_to_p2(&r);
    }
    r = ge_p2_dbl(&s);
    return p1p1_to_p3(&r);
}
static inline ge_p3 ge_neg(const ge_p3* p) {
    ge_p3 r = {fe_neg(p->X), p->Y, p->Z, fe_neg(p->T)};
    return r;
}

/* ---------------------------------------------------------------------------- ristretto255 */
static inline void ristretto_compress(uint8_t out[32], const ge_p3* p) {
    fe51 u1 = fe_mul(fe_add(p->Z, p->Y), fe_sub(p->Z, p->Y));
    fe51 u2 = fe_mul(p->X, p->Y);
    fe51 invsqrt;
    fe_sqrt_ratio_i(&invsqrt, FE_ONE, fe_mul(u1, fe_sq(u2)));
    fe51 i1 = fe_mul(invsqrt, u1), i2 = fe_mul(invsqrt, u2);
    fe51 z_inv = fe_mul(i1, fe_mul(i2, p->T));
    fe51 den_inv = i2;
    fe51 iX = fe_mul(p->X, FE_SQRT_M1), iY = fe_mul(p->Y, FE_SQRT_M1);
    fe51 ench = fe_mul(i1, FE_INVSQRT_A_MINUS_D);
    int rotate = fe_is_neg(fe_mul(p->T, z_inv));
    fe51 X = rotate ? iY : p->X, Y = rotate ? iX : p->Y;
    if (rotate) den_inv = ench;
    Y = fe_cneg(Y, fe_is_neg(fe_mul(X, z_inv)));
    fe51 s = fe_abs(fe_mul(den_inv, fe_sub(p->Z, Y)));
    fe_to_bytes(out, s);
}
static inline int ristretto_decompress(ge_p3* out, const uint8_t in[32]) {
    fe51 s = fe_from_bytes(in);
    uint8_t chk[32];
    fe_to_bytes(chk, s);
    if (memcmp(chk, in, 32) != 0 || (in[0] & 1)) return 0;
    fe51 ss = fe_sq(s);
    fe51 u1 = fe_sub(FE_ONE, ss), u2 = fe_add(FE_ONE, ss);
    fe51 u2_sqr = fe_sq(u2);
    fe51 v = fe_sub(fe_neg(fe_mul(FE_D, fe_sq(u1))), u2_sqr);
    fe51 I;
    int ok = fe_sqrt_ratio_i(&I, FE_ONE, fe_mul(v, u2_sqr));
    fe51 Dx = fe_mul(I, u2), Dy = fe_mul(I, fe_mul(Dx, v));
    fe51 x = fe_abs(fe_mul(fe_add(s, s), Dx));
    fe51 y = fe_mul(u1, Dy);
    fe51 t = fe_mul(x, y);
    if (!ok || fe_is_neg(t) || fe_is_zero(y)) return 0;
    out->X = x; out->Y = y; out->Z = FE_ONE; out->T = t;
    return 1;
}
static inline ge_p3 ristretto_elligator(fe51 r_0) {
    fe51 r = fe_mul(FE_SQRT_M1, fe_sq(r_0));
    fe51 Ns = fe_mul(fe_add(r, FE_ONE), FE_ONE_MINUS_D_SQ);
    fe51 c = fe_neg(FE_ONE);
    fe51 Dv = fe_mul(fe_sub(c, fe_mul(FE_D, r)), fe_add(r, FE_D));
    fe51 s;
    int sq = fe_sqrt_ratio_i(&s, Ns, Dv);
    fe51 s_prime = fe_neg(fe_abs(fe_mul(s, r_0)));
    if (!sq) { s = s_prime; c = r; }
    fe51 Nt = fe_sub(fe_mul(fe_mul(c, fe_sub(r, FE_ONE)), FE_D_MINUS_ONE_SQ), Dv);
    fe51 s_sq = fe_sq(s);
    ge_p1p1 cp = {fe_mul(fe_add(s, s), Dv), fe_sub(FE_ONE, s_sq), fe_mul(Nt, FE_SQRT_AD_MINUS_ONE), fe_add(FE_ONE, s_sq)};
    return p1p1_to_p3(&cp);
}
static inline ge_p3 ristretto_from_uniform_bytes(const uint8_t b[64]) {
    ge_p3 a = ristretto_elligator(fe_from_bytes(b)), c = ristretto_elligator(fe_from_bytes(b + 32));
    return ge_add(&a, &c);
}
static inline int ristretto_is_identity(const ge_p3* p) { return fe_is_zero(p->X) || fe_is_zero(p->Y); }

#endif
