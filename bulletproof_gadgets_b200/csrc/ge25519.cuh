// Edwards25519 group arithmetic (extended coordinates, a = -1) and the ristretto255 codec.
//
// Replaces curve25519-dalek 3.2.0 `EdwardsPoint` / `RistrettoPoint` / `CompressedRistretto`
// (edwards.rs, ristretto.rs, backend/serial/curve_models; /root/reference/Cargo.lock:155-157,
// not vendored) -- SURVEY.md rows K8 / a13.  Encoding, decoding and the Elligator map follow
// RFC 9496 sections 4.3.1, 4.3.2, 4.3.4; reference call sites:
//   CompressedRistretto::from_slice  /root/reference/src/lalrpop/assignment_parser.rs:137
//   Prover::commit -> compress       /root/reference/src/gadget.rs:32
#pragma once
#include "consts.cuh"
#include "fe25519.cuh"

struct alignas(16) ge_ext {  // (X:Y:Z:T), x = X/Z, y = Y/Z, xy = T/Z
    fe X, Y, Z, T;
};
struct alignas(16) ge_niels {  // affine: (y+x, y-x, 2dxy); identity = (1, 1, 0)
    fe yp, ym, t2d;
};

BPG_HD fe fe_const(const uint32_t (&l)[8]) {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = l[i];
    return r;
}
#define FE_CONST(NAME)                         \
    BPG_HD fe fe_##NAME() {                    \
        const uint32_t l[8] = FE_##NAME##_LIMBS; \
        return fe_const(l);                    \
    }
FE_CONST(D)
FE_CONST(2D)
FE_CONST(SQRT_M1)
FE_CONST(SQRT_AD_MINUS_ONE)
FE_CONST(INVSQRT_A_MINUS_D)
FE_CONST(ONE_MINUS_D_SQ)
FE_CONST(D_MINUS_ONE_SQ)

BPG_HD ge_ext ge_identity() {
    ge_ext r;
    r.X = fe_zero();
    r.Y = fe_one();
    r.Z = fe_one();
    r.T = fe_zero();
    return r;
}
BPG_HD ge_niels ge_niels_identity() {
    ge_niels r;
    r.yp = fe_one();
    r.ym = fe_one();
    r.t2d = fe_zero();
    return r;
}

// r = p + q (unified, complete for a=-1, d non-square): add-2008-hwcd-3, 9M
BPG_HD ge_ext ge_add(const ge_ext& p, const ge_ext& q) {
    fe A = fe_mul(fe_sub(p.Y, p.X), fe_sub(q.Y, q.X));
    fe B = fe_mul(fe_add(p.Y, p.X), fe_add(q.Y, q.X));
    fe C = fe_mul(fe_mul(p.T, fe_2D()), q.T);
    fe Dd = fe_mul(p.Z, q.Z);
    Dd = fe_add(Dd, Dd);
    fe E = fe_sub(B, A), F = fe_sub(Dd, C), G = fe_add(Dd, C), H = fe_add(B, A);
    ge_ext r;
    r.X = fe_mul(E, F);
    r.Y = fe_mul(G, H);
    r.Z = fe_mul(F, G);
    r.T = fe_mul(E, H);
    return r;
}

// r = p +/- q with q affine Niels: 7M.  neg selects subtraction (digit sign in the MSM).
BPG_HD ge_ext ge_madd(const ge_ext& p, const ge_niels& q, bool neg) {
    fe qa = fe_select(neg, q.yp, q.ym);
    fe qb = fe_select(neg, q.ym, q.yp);
    fe A = fe_mul(fe_sub(p.Y, p.X), qa);
    fe B = fe_mul(fe_add(p.Y, p.X), qb);
    fe C = fe_mul(p.T, q.t2d);
    fe Dd = fe_add(p.Z, p.Z);
    fe E = fe_sub(B, A), H = fe_add(B, A);
    fe DmC = fe_sub(Dd, C), DpC = fe_add(Dd, C);
    fe F = fe_select(neg, DpC, DmC), G = fe_select(neg, DmC, DpC);
    ge_ext r;
    r.X = fe_mul(E, F);
    r.Y = fe_mul(G, H);
    r.Z = fe_mul(F, G);
    r.T = fe_mul(E, H);
    return r;
}

// The same with the sign already applied to the (y+x, y-x) pair by the caller (qa = neg ? y+x : y-x, qb = the other one --
// the bucket kernel swaps the two load addresses instead of selecting 16 limbs); only F / G still depend on the sign.
BPG_HD ge_ext ge_madd_swapped(const ge_ext& p, const fe& qa, const fe& qb, const fe& t2d, bool neg) {
    fe A = fe_mul(fe_sub(p.Y, p.X), qa);
    fe B = fe_mul(fe_add(p.Y, p.X), qb);
    fe C = fe_mul(p.T, t2d);
    fe Dd = fe_add(p.Z, p.Z);
    fe E = fe_sub(B, A), H = fe_add(B, A);
    fe DmC = fe_sub(Dd, C), DpC = fe_add(Dd, C);
    fe F = fe_select(neg, DpC, DmC), G = fe_select(neg, DmC, DpC);
    ge_ext r;
    r.X = fe_mul(E, F);
    r.Y = fe_mul(G, H);
    r.Z = fe_mul(F, G);
    r.T = fe_mul(E, H);
    return r;
}

// r = 2p: dbl-2008-hwcd (a=-1), 4M + 4S
BPG_HD ge_ext ge_dbl(const ge_ext& p) {
    fe A = fe_sqr(p.X), B = fe_sqr(p.Y);
    fe C = fe_sqr(p.Z);
    C = fe_add(C, C);
    fe Dn = fe_neg(A);               // a*A
    fe xy = fe_add(p.X, p.Y);
    fe E = fe_sub(fe_sub(fe_sqr(xy), A), B);
    fe G = fe_add(Dn, B);
    fe F = fe_sub(G, C);
    fe H = fe_sub(Dn, B);
    ge_ext r;
    r.X = fe_mul(E, F);
    r.Y = fe_mul(G, H);
    r.Z = fe_mul(F, G);
    r.T = fe_mul(E, H);
    return r;
}

// 2p without the T coordinate (3M + 4S): for runs of doublings, where only the last one needs T (ge_dbl reads X, Y, Z only)
BPG_HD ge_ext ge_dbl_not(const ge_ext& p) {
    fe A = fe_sqr(p.X), B = fe_sqr(p.Y);
    fe C = fe_sqr(p.Z);
    C = fe_add(C, C);
    fe Dn = fe_neg(A);
    fe xy = fe_add(p.X, p.Y);
    fe E = fe_sub(fe_sub(fe_sqr(xy), A), B);
    fe G = fe_add(Dn, B);
    fe F = fe_sub(G, C);
    fe H = fe_sub(Dn, B);
    ge_ext r;
    r.X = fe_mul(E, F);
    r.Y = fe_mul(G, H);
    r.Z = fe_mul(F, G);
    r.T = p.T;  // not maintained
    return r;
}

BPG_HD ge_ext ge_neg(const ge_ext& p) {
    ge_ext r = p;
    r.X = fe_neg(p.X);
    r.T = fe_neg(p.T);
    return r;
}

// affine Niels form of p given zinv = 1/Z
BPG_HD ge_niels ge_to_niels(const ge_ext& p, const fe& zinv) {
    fe x = fe_mul(p.X, zinv), y = fe_mul(p.Y, zinv);
    ge_niels r;
    r.yp = fe_add(y, x);
    r.ym = fe_sub(y, x);
    r.t2d = fe_mul(fe_mul(x, y), fe_2D());
    return r;
}
BPG_HD ge_ext ge_from_niels(const ge_niels& q) { return ge_madd(ge_identity(), q, false); }

// ristretto equality with the identity: X == 0 or Y == 0 (RFC 9496 4.3.3 with (0,1))
BPG_HD bool ge_is_ristretto_identity(const ge_ext& p) { return fe_is_zero(p.X) || fe_is_zero(p.Y); }

// ------------------------------------------------------------------------------------------
// ristretto255
// ------------------------------------------------------------------------------------------
// RFC 9496 4.2 SQRT_RATIO_M1
BPG_HD bool fe_sqrt_ratio_m1(fe* out, const fe& u, const fe& v) {
    fe v3 = fe_mul(fe_sqr(v), v);
    fe v7 = fe_mul(fe_sqr(v3), v);
    fe r = fe_mul(fe_mul(u, v3), fe_pow_p58(fe_mul(u, v7)));
    fe check = fe_mul(v, fe_sqr(r));
    fe neg_u = fe_neg(u);
    bool correct = fe_eq(check, u);
    bool flipped = fe_eq(check, neg_u);
    bool flipped_i = fe_eq(check, fe_mul(neg_u, fe_SQRT_M1()));
    fe r_prime = fe_mul(r, fe_SQRT_M1());
    r = fe_select(flipped || flipped_i, r_prime, r);
    *out = fe_abs(r);
    return correct || flipped;
}

// RFC 9496 4.3.2 Encode
BPG_HD void ge_ristretto_compress(uint8_t out[32], const ge_ext& p) {
    fe u1 = fe_mul(fe_add(p.Z, p.Y), fe_sub(p.Z, p.Y));
    fe u2 = fe_mul(p.X, p.Y);
    fe invsqrt;
    fe_sqrt_ratio_m1(&invsqrt, fe_one(), fe_mul(u1, fe_sqr(u2)));
    fe den1 = fe_mul(invsqrt, u1), den2 = fe_mul(invsqrt, u2);
    fe z_inv = fe_mul(fe_mul(den1, den2), p.T);
    fe ix0 = fe_mul(p.X, fe_SQRT_M1()), iy0 = fe_mul(p.Y, fe_SQRT_M1());
    fe ench = fe_mul(den1, fe_INVSQRT_A_MINUS_D());
    bool rotate = fe_is_neg(fe_mul(p.T, z_inv));
    fe x = fe_select(rotate, iy0, p.X);
    fe y = fe_select(rotate, ix0, p.Y);
    fe den_inv = fe_select(rotate, ench, den2);
    y = fe_cneg(y, fe_is_neg(fe_mul(x, z_inv)));
    fe s = fe_abs(fe_mul(den_inv, fe_sub(p.Z, y)));
    fe_to_bytes(out, s);
}

// RFC 9496 4.3.1 Decode; false for a non-canonical / invalid encoding.  Result has Z = 1.
BPG_HD bool ge_ristretto_decompress(ge_ext* out, const uint8_t in[32]) {
    fe s = fe_from_bytes(in);
    uint8_t chk[32];
    fe_to_bytes(chk, s);
    bool canonical = true;
    for (int i = 0; i < 32; i++) canonical = canonical && (chk[i] == in[i]);
    if (!canonical || (in[0] & 1)) return false;
    fe ss = fe_sqr(s);
    fe u1 = fe_sub(fe_one(), ss), u2 = fe_add(fe_one(), ss);
    fe u2_sqr = fe_sqr(u2);
    fe v = fe_sub(fe_neg(fe_mul(fe_D(), fe_sqr(u1))), u2_sqr);
    fe invsqrt;
    bool was_square = fe_sqrt_ratio_m1(&invsqrt, fe_one(), fe_mul(v, u2_sqr));
    fe den_x = fe_mul(invsqrt, u2);
    fe den_y = fe_mul(fe_mul(invsqrt, den_x), v);
    fe x = fe_abs(fe_mul(fe_add(s, s), den_x));
    fe y = fe_mul(u1, den_y);
    fe t = fe_mul(x, y);
    if (!was_square || fe_is_neg(t) || fe_is_zero(y)) return false;
    out->X = x;
    out->Y = y;
    out->Z = fe_one();
    out->T = t;
    return true;
}

// RFC 9496 4.3.4 MAP (Elligator 2)
BPG_HD ge_ext ge_elligator(const fe& t) {
    fe r = fe_mul(fe_SQRT_M1(), fe_sqr(t));
    fe u = fe_mul(fe_add(r, fe_one()), fe_ONE_MINUS_D_SQ());
    fe minus_one = fe_neg(fe_one());
    fe v = fe_mul(fe_sub(minus_one, fe_mul(r, fe_D())), fe_add(r, fe_D()));
    fe s;
    bool was_square = fe_sqrt_ratio_m1(&s, u, v);
    fe s_prime = fe_neg(fe_abs(fe_mul(s, t)));
    s = fe_select(was_square, s, s_prime);
    fe c = fe_select(was_square, minus_one, r);
    fe N = fe_sub(fe_mul(fe_mul(c, fe_sub(r, fe_one())), fe_D_MINUS_ONE_SQ()), v);
    fe sv = fe_mul(s, v);
    fe w0 = fe_add(sv, sv);
    fe w1 = fe_mul(N, fe_SQRT_AD_MINUS_ONE());
    fe s2 = fe_sqr(s);
    fe w2 = fe_sub(fe_one(), s2), w3 = fe_add(fe_one(), s2);
    ge_ext p;
    p.X = fe_mul(w0, w3);
    p.Y = fe_mul(w2, w1);
    p.Z = fe_mul(w1, w3);
    p.T = fe_mul(w0, w2);
    return p;
}

// RistrettoPoint::from_uniform_bytes: two Elligator maps on the 255-bit halves, summed
BPG_HD ge_ext ge_from_uniform_bytes(const uint8_t b[64]) {
    fe r0 = fe_from_bytes(b), r1 = fe_from_bytes(b + 32);
    r0.v[7] &= 0x7fffffffu;
    r1.v[7] &= 0x7fffffffu;
    return ge_add(ge_elligator(r0), ge_elligator(r1));
}
