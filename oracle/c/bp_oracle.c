/* ORACLE (test infrastructure / CPU baseline; never linked into the product).
 *
 * Single-threaded C restatement of the hot path exactly as the reference executes it on the CPU:
 *   curve25519-dalek 3.2.0   backend/serial/scalar_mul/{straus,pippenger}.rs  (const-time radix-16
 *                            Straus; vartime NAF-5 Straus below 190 points, vartime Pippenger w=6/7/8 above)
 *   merlin 2.0.1             strobe.rs, transcript.rs
 *   bulletproofs 2.1.0 fork  generators.rs, r1cs/{prover,verifier,proof}.rs, inner_product_proof.rs
 * (pins: /root/reference/Cargo.lock:78-80,155-157,403-405; sources NOT vendored -> restated from the
 * published algorithms).  Reference call sites: /root/reference/src/prove.rs:46-47,78-81,
 * /root/reference/src/verify.rs:44-46,53,70-71, /root/reference/src/gadget.rs:32.
 *
 * Parity with dalek's bytes: UNPINNED (no Rust toolchain, no reference test pins proof bytes).
 * Pinned against: RFC 9496 vectors, Merlin's KAT, B_blinding, and the big-int python oracle
 * (tests/test_oracle_c.py).  Timed by bench.py as the CPU baseline ("port", 1 core).
 */
#include <stdio.h>
#include <stdlib.h>

#include "curve.h"
#include "scalar.h"

/* ============================================================================ keccak / merlin */
static inline uint64_t rol64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }
static void keccak_f1600(uint64_t s[25]) {
    static const uint64_t RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
        0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
        0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
        0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
        0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    static const int RHO[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
    static const int PI[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
    for (int round = 0; round < 24; round++) {
        uint64_t bc[5];
        for (int i = 0; i < 5; i++) bc[i] = s[i] ^ s[i + 5] ^ s[i + 10] ^ s[i + 15] ^ s[i + 20];
        for (int i = 0; i < 5; i++) {
            uint64_t t = bc[(i + 4) % 5] ^ rol64(bc[(i + 1) % 5], 1);
            for (int j = 0; j < 25; j += 5) s[j + i] ^= t;
        }
        uint64_t t = s[1];
        for (int i = 0; i < 24; i++) {
            int j = PI[i];
            uint64_t b = s[j];
            s[j] = rol64(t, RHO[i]);
            t = b;
        }
        for (int j = 0; j < 25; j += 5) {
            for (int i = 0; i < 5; i++) bc[i] = s[j + i];
            for (int i = 0; i < 5; i++) s[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
        }
        s[0] ^= RC[round];
    }
}
typedef struct { uint64_t st[25]; size_t rate, pos; int squeezing; uint8_t suffix; } sponge_t;
static void sponge_init(sponge_t* s, size_t rate, uint8_t suffix) { memset(s, 0, sizeof *s); s->rate = rate; s->suffix = suffix; }
static void sponge_absorb(sponge_t* s, const uint8_t* d, size_t n) {
    uint8_t* b = (uint8_t*)s->st;
    for (size_t i = 0; i < n; i++) { b[s->pos++] ^= d[i]; if (s->pos == s->rate) { keccak_f1600(s->st); s->pos = 0; } }
}
static void sponge_squeeze(sponge_t* s, uint8_t* out, size_t n) {
    uint8_t* b = (uint8_t*)s->st;
    if (!s->squeezing) { b[s->pos] ^= s->suffix; b[s->rate - 1] ^= 0x80; keccak_f1600(s->st); s->pos = 0; s->squeezing = 1; }
    for (size_t i = 0; i < n; i++) { if (s->pos == s->rate) { keccak_f1600(s->st); s->pos = 0; } out[i] = b[s->pos++]; }
}

#define STROBE_R 166
enum { FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32 };
typedef struct { uint64_t st64[25]; uint8_t pos, pos_begin, cur_flags; } strobe_t;
#define ST(s) ((uint8_t*)(s)->st64)
static void strobe_run_f(strobe_t* s) {
    ST(s)[s->pos] ^= s->pos_begin; ST(s)[s->pos + 1] ^= 0x04; ST(s)[STROBE_R + 1] ^= 0x80;
    keccak_f1600(s->st64); s->pos = 0; s->pos_begin = 0;
}
static void strobe_absorb(strobe_t* s, const uint8_t* d, size_t n) { for (size_t i = 0; i < n; i++) { ST(s)[s->pos++] ^= d[i]; if (s->pos == STROBE_R) strobe_run_f(s); } }
static void strobe_overwrite(strobe_t* s, const uint8_t* d, size_t n) { for (size_t i = 0; i < n; i++) { ST(s)[s->pos++] = d[i]; if (s->pos == STROBE_R) strobe_run_f(s); } }
static void strobe_squeeze(strobe_t* s, uint8_t* d, size_t n) { for (size_t i = 0; i < n; i++) { d[i] = ST(s)[s->pos]; ST(s)[s->pos++] = 0; if (s->pos == STROBE_R) strobe_run_f(s); } }
static void strobe_begin_op(strobe_t* s, uint8_t flags, int more) {
    if (more) return;
    uint8_t old = s->pos_begin;
    s->pos_begin = s->pos + 1; s->cur_flags = flags;
    uint8_t hdr[2] = {old, flags};
    strobe_absorb(s, hdr, 2);
    if ((flags & (FLAG_C | FLAG_K)) && s->pos != 0) strobe_run_f(s);
}
static void strobe_meta_ad(strobe_t* s, const void* d, size_t n, int more) { strobe_begin_op(s, FLAG_M | FLAG_A, more); strobe_absorb(s, d, n); }
static void strobe_ad(strobe_t* s, const void* d, size_t n, int more) { strobe_begin_op(s, FLAG_A, more); strobe_absorb(s, d, n); }
static void strobe_prf(strobe_t* s, uint8_t* d, size_t n) { strobe_begin_op(s, FLAG_I | FLAG_A | FLAG_C, 0); strobe_squeeze(s, d, n); }
static void strobe_key(strobe_t* s, const void* d, size_t n) { strobe_begin_op(s, FLAG_A | FLAG_C, 0); strobe_overwrite(s, d, n); }
static void strobe_new(strobe_t* s, const char* label) {
    memset(s, 0, sizeof *s);
    const uint8_t init[6] = {1, STROBE_R + 2, 1, 0, 1, 96};
    memcpy(ST(s), init, 6); memcpy(ST(s) + 6, "STROBEv1.0.2", 12);
    keccak_f1600(s->st64);
    strobe_meta_ad(s, label, strlen(label), 0);
}
typedef strobe_t transcript_t;
static void t_append(transcript_t* t, const char* label, const void* m, size_t n) {
    uint32_t len = (uint32_t)n;
    strobe_meta_ad(t, label, strlen(label), 0); strobe_meta_ad(t, &len, 4, 1); strobe_ad(t, m, n, 0);
}
static void t_append_u64(transcript_t* t, const char* label, uint64_t x) { t_append(t, label, &x, 8); }
static void t_challenge(transcript_t* t, const char* label, uint8_t* out, size_t n) {
    uint32_t len = (uint32_t)n;
    strobe_meta_ad(t, label, strlen(label), 0); strobe_meta_ad(t, &len, 4, 1); strobe_prf(t, out, n);
}
static scl t_challenge_scalar(transcript_t* t, const char* label) { uint8_t b[64]; t_challenge(t, label, b, 64); return scl_from_wide(b); }
static void t_new(transcript_t* t, const uint8_t* label, size_t n) { strobe_new(t, "Merlin v1.0"); t_append(t, "dom-sep", label, n); }
static void t_append_scalar(transcript_t* t, const char* label, scl s) { uint8_t b[32]; scl_to_bytes(b, s); t_append(t, label, b, 32); }
static void rng_fill(strobe_t* r, uint8_t* out, size_t n) { uint32_t len = (uint32_t)n; strobe_meta_ad(r, &len, 4, 0); strobe_prf(r, out, n); }
static scl rng_scalar(strobe_t* r) { uint8_t b[64]; rng_fill(r, b, 64); return scl_from_wide(b); }

/* ============================================================================ MSM algorithms */
/* LookupTable<ProjectiveNielsPoint>: [P, 2P, .., 8P]; constant-time select */
static void lookup_from(ge_pniels tbl[8], const ge_p3* P) {
    tbl[0] = p3_to_pniels(P);
    for (int j = 0; j < 7; j++) { ge_p1p1 c = ge_add_pniels(P, &tbl[j]); ge_p3 e = p1p1_to_p3(&c); tbl[j + 1] = p3_to_pniels(&e); }
}
static inline void fe_cmov(fe51* d, const fe51* s, uint64_t mask) { for (int i = 0; i < 5; i++) d->v[i] ^= mask & (d->v[i] ^ s->v[i]); }
static ge_pniels lookup_select(const ge_pniels tbl[8], int8_t x) {
    int xmask = x >> 7; int xabs = (x + xmask) ^ xmask;
    ge_pniels t = {FE_ONE, FE_ONE, FE_ONE, FE_ZERO};
    for (int j = 1; j < 9; j++) {
        uint64_t m = (uint64_t)0 - (uint64_t)(xabs == j);
        fe_cmov(&t.YpX, &tbl[j - 1].YpX, m); fe_cmov(&t.YmX, &tbl[j - 1].YmX, m); fe_cmov(&t.Z, &tbl[j - 1].Z, m); fe_cmov(&t.T2d, &tbl[j - 1].T2d, m);
    }
    if (xmask) { fe51 tmp = t.YpX; t.YpX = t.YmX; t.YmX = tmp; t.T2d = fe_neg(t.T2d); }
    return t;
}
/* Straus::multiscalar_mul (constant time) */
static ge_p3 msm_straus_ct(const scl* scalars, const ge_p3* points, size_t n) {
    ge_pniels* tables = malloc(sizeof(ge_pniels) * 8 * (n ? n : 1));
    int8_t* digits = malloc(64 * (n ? n : 1));
    for (size_t i = 0; i < n; i++) { lookup_from(tables + 8 * i, &points[i]); uint8_t b[32]; scl_to_bytes(b, scalars[i]); scl_to_radix_16(digits + 64 * i, b); }
    ge_p3 Q = ge_identity();
    for (int j = 63; j >= 0; j--) {
        Q = ge_mul_by_pow_2(&Q, 4);
        for (size_t i = 0; i < n; i++) { ge_pniels R = lookup_select(tables + 8 * i, digits[64 * i + j]); ge_p1p1 c = ge_add_pniels(&Q, &R); Q = p1p1_to_p3(&c); }
    }
    free(tables); free(digits);
    return Q;
}
/* NafLookupTable5: [P, 3P, .., 15P] */
static void naf_table(ge_pniels tbl[8], const ge_p3* A) {
    tbl[0] = p3_to_pniels(A);
    ge_p3 A2 = ge_dbl(A);
    for (int i = 0; i < 7; i++) { ge_p1p1 c = ge_add_pniels(&A2, &tbl[i]); ge_p3 e = p1p1_to_p3(&c); tbl[i + 1] = p3_to_pniels(&e); }
}
/* Straus::optional_multiscalar_mul (variable time) */
static ge_p3 msm_straus_vt(const scl* scalars, const ge_p3* points, size_t n) {
    ge_pniels* tables = malloc(sizeof(ge_pniels) * 8 * (n ? n : 1));
    int8_t* nafs = malloc(256 * (n ? n : 1));
    for (size_t i = 0; i < n; i++) { naf_table(tables + 8 * i, &points[i]); uint8_t b[32]; scl_to_bytes(b, scalars[i]); scl_naf(nafs + 256 * i, b, 5); }
    ge_p2 r = ge_p2_identity();
    for (int i = 255; i >= 0; i--) {
        ge_p1p1 t = ge_p2_dbl(&r);
        for (size_t k = 0; k < n; k++) {
            int8_t d = nafs[256 * k + i];
            if (d > 0) { ge_p3 e = p1p1_to_p3(&t); t = ge_add_pniels(&e, &tables[8 * k + d / 2]); }
            else if (d < 0) { ge_p3 e = p1p1_to_p3(&t); t = ge_sub_pniels(&e, &tables[8 * k + (-d) / 2]); }
        }
        r = p1p1_to_p2(&t);
    }
    free(tables); free(nafs);
    /* ProjectivePoint::to_extended: (XZ, YZ, Z^2, XY) */
    ge_p3 out = {fe_mul(r.X, r.Z), fe_mul(r.Y, r.Z), fe_sq(r.Z), fe_mul(r.X, r.Y)};
    return out;
}
/* Pippenger::optional_multiscalar_mul (variable time) */
static ge_p3 msm_pippenger_vt(const scl* scalars, const ge_p3* points, size_t n) {
    const int w = n < 500 ? 6 : (n < 800 ? 7 : 8);
    const int max_digit = 1 << w, digits_count = scl_radix_2w_size(w), buckets_count = max_digit / 2;
    int8_t* digits = malloc(43 * n);
    ge_pniels* pts = malloc(sizeof(ge_pniels) * n);
    for (size_t i = 0; i < n; i++) { uint8_t b[32]; scl_to_bytes(b, scalars[i]); scl_to_radix_2w(digits + 43 * i, b, w); pts[i] = p3_to_pniels(&points[i]); }
    ge_p3* buckets = malloc(sizeof(ge_p3) * buckets_count);
    ge_p3 total = ge_identity();
    for (int di = digits_count - 1; di >= 0; di--) {
        for (int b = 0; b < buckets_count; b++) buckets[b] = ge_identity();
        for (size_t i = 0; i < n; i++) {
            int d = digits[43 * i + di];
            if (d > 0) { ge_p1p1 c = ge_add_pniels(&buckets[d - 1], &pts[i]); buckets[d - 1] = p1p1_to_p3(&c); }
            else if (d < 0) { ge_p1p1 c = ge_sub_pniels(&buckets[-d - 1], &pts[i]); buckets[-d - 1] = p1p1_to_p3(&c); }
        }
        ge_p3 inter = buckets[buckets_count - 1], sum = buckets[buckets_count - 1];
        for (int b = buckets_count - 2; b >= 0; b--) { inter = ge_add(&inter, &buckets[b]); sum = ge_add(&sum, &inter); }
        if (di == digits_count - 1) total = sum;
        else { total = ge_mul_by_pow_2(&total, w); total = ge_add(&total, &sum); }
    }
    free(digits); free(pts); free(buckets);
    return total;
}
/* RistrettoPoint::vartime_multiscalar_mul dispatch: Straus below 190 points, Pippenger otherwise */
static ge_p3 msm_vartime(const scl* scalars, const ge_p3* points, size_t n) {
    return n < 190 ? msm_straus_vt(scalars, points, n) : msm_pippenger_vt(scalars, points, n);
}

/* ============================================================================ generators */
static const uint8_t BASEPOINT_COMPRESSED[32] = {0xe2, 0xf2, 0xae, 0x0a, 0x6a, 0xbc, 0x4e, 0x71, 0xa8, 0x84, 0xa9, 0x61, 0xc5, 0x00, 0x51, 0x5f,
                                                 0x58, 0xe3, 0x0b, 0x6a, 0xa5, 0x82, 0xdd, 0x8d, 0xb6, 0xa6, 0x59, 0x45, 0xe0, 0x8d, 0x2d, 0x76};
typedef struct { ge_p3 B, Bb; ge_p3 *G, *H; size_t cap; } gens_t;
static void gens_free(gens_t* g) { free(g->G); free(g->H); g->G = g->H = NULL; g->cap = 0; }
static void gens_new(gens_t* g, size_t cap) {
    memset(g, 0, sizeof *g);
    ristretto_decompress(&g->B, BASEPOINT_COMPRESSED);
    sponge_t s; sponge_init(&s, 72, 0x06); sponge_absorb(&s, BASEPOINT_COMPRESSED, 32);
    uint8_t h[64]; sponge_squeeze(&s, h, 64);
    g->Bb = ristretto_from_uniform_bytes(h);
    g->cap = cap;
    g->G = malloc(sizeof(ge_p3) * (cap ? cap : 1)); g->H = malloc(sizeof(ge_p3) * (cap ? cap : 1));
    for (int which = 0; which < 2; which++) {
        sponge_t sh; sponge_init(&sh, 136, 0x1f);
        const uint8_t label[5] = {(uint8_t)(which ? 'H' : 'G'), 0, 0, 0, 0};
        sponge_absorb(&sh, (const uint8_t*)"GeneratorsChain", 15); sponge_absorb(&sh, label, 5);
        for (size_t i = 0; i < cap; i++) { uint8_t u[64]; sponge_squeeze(&sh, u, 64); (which ? g->H : g->G)[i] = ristretto_from_uniform_bytes(u); }
    }
}
static ge_p3 pedersen_commit(const gens_t* g, scl v, scl r) { scl s[2] = {v, r}; ge_p3 p[2] = {g->B, g->Bb}; return msm_straus_ct(s, p, 2); }

/* ============================================================================ constraint system */
enum { V_COMMITTED = 0, V_LEFT = 1, V_RIGHT = 2, V_OUT = 3, V_ONE = 4 };
typedef struct { const uint32_t* row_start; const uint32_t* term_var; const scl* term_coef; size_t q; } cs_t;
static void flatten(const cs_t* cs, scl z, size_t n, size_t m, scl* wL, scl* wR, scl* wO, scl* wV, scl* wc) {
    for (size_t i = 0; i < n; i++) wL[i] = wR[i] = wO[i] = SC_ZERO;
    for (size_t i = 0; i < m; i++) wV[i] = SC_ZERO;
    if (wc) *wc = SC_ZERO;
    scl exp_z = z;
    for (size_t j = 0; j < cs->q; j++) {
        for (uint32_t e = cs->row_start[j]; e < cs->row_start[j + 1]; e++) {
            uint32_t k = cs->term_var[e] >> 29, i = cs->term_var[e] & ((1u << 29) - 1);
            scl t = scl_mul(exp_z, cs->term_coef[e]);
            switch (k) {
                case V_LEFT: wL[i] = scl_add(wL[i], t); break;
                case V_RIGHT: wR[i] = scl_add(wR[i], t); break;
                case V_OUT: wO[i] = scl_add(wO[i], t); break;
                case V_COMMITTED: wV[i] = scl_sub(wV[i], t); break;
                case V_ONE: if (wc) *wc = scl_sub(*wc, t); break;
            }
        }
        exp_z = scl_mul(exp_z, z);
    }
}
static scl inner_product(const scl* a, const scl* b, size_t n) { scl acc = SC_ZERO; for (size_t i = 0; i < n; i++) acc = scl_add(acc, scl_mul(a[i], b[i])); return acc; }
static size_t next_pow2(size_t n) { size_t p = 1; while (p < n) p <<= 1; return p; }

/* InnerProductProof::create; L/R written interleaved to lr (64 bytes per round) */
static void ipp_create(transcript_t* T, const ge_p3* Q, const scl* G_factors, const scl* H_factors, ge_p3* G, ge_p3* H, scl* a, scl* b, size_t n, uint8_t* lr, scl* a_out, scl* b_out) {
    t_append(T, "dom-sep", "ipp v1", 6); t_append_u64(T, "n", n);
    int first = 1;
    scl* sc_buf = malloc(sizeof(scl) * (n + 1));
    ge_p3* pt_buf = malloc(sizeof(ge_p3) * (n + 1));
    while (n != 1) {
        n /= 2;
        scl *aL = a, *aR = a + n, *bL = b, *bR = b + n;
        ge_p3 *GL = G, *GR = G + n, *HL = H, *HR = H + n;
        scl cL = inner_product(aL, bR, n), cR = inner_product(aR, bL, n);
        for (int side = 0; side < 2; side++) {
            for (size_t i = 0; i < n; i++) {
                if (side == 0) { sc_buf[i] = first ? scl_mul(aL[i], G_factors[n + i]) : aL[i]; sc_buf[n + i] = first ? scl_mul(bR[i], H_factors[i]) : bR[i]; pt_buf[i] = GR[i]; pt_buf[n + i] = HL[i]; }
                else { sc_buf[i] = first ? scl_mul(aR[i], G_factors[i]) : aR[i]; sc_buf[n + i] = first ? scl_mul(bL[i], H_factors[n + i]) : bL[i]; pt_buf[i] = GL[i]; pt_buf[n + i] = HR[i]; }
            }
            sc_buf[2 * n] = side == 0 ? cL : cR; pt_buf[2 * n] = *Q;
            ge_p3 P = msm_vartime(sc_buf, pt_buf, 2 * n + 1);
            ristretto_compress(lr + 32 * side, &P);
        }
        t_append(T, "L", lr, 32); t_append(T, "R", lr + 32, 32);
        lr += 64;
        scl u = t_challenge_scalar(T, "u"), ui = scl_invert(u);
        for (size_t i = 0; i < n; i++) {
            aL[i] = scl_add(scl_mul(aL[i], u), scl_mul(ui, aR[i]));
            bL[i] = scl_add(scl_mul(bL[i], ui), scl_mul(u, bR[i]));
            scl s2[2]; ge_p3 p2[2];
            s2[0] = first ? scl_mul(ui, G_factors[i]) : ui; s2[1] = first ? scl_mul(u, G_factors[n + i]) : u; p2[0] = GL[i]; p2[1] = GR[i];
            GL[i] = msm_vartime(s2, p2, 2);
            s2[0] = first ? scl_mul(u, H_factors[i]) : u; s2[1] = first ? scl_mul(ui, H_factors[n + i]) : ui; p2[0] = HL[i]; p2[1] = HR[i];
            HL[i] = msm_vartime(s2, p2, 2);
        }
        first = 0;
    }
    *a_out = a[0]; *b_out = b[0];
    free(sc_buf); free(pt_buf);
}

/* ============================================================================ exported API */
static gens_t g_cache; /* BulletproofGens are rebuilt by the reference on every run; the cache is only
                          used when the caller asks for it (bench times both ways) */
static const gens_t* get_gens(size_t cap, int use_cache, gens_t* local) {
    if (use_cache) { if (g_cache.cap < cap || !g_cache.G) { gens_free(&g_cache); gens_new(&g_cache, cap); } return &g_cache; }
    gens_new(local, cap); return local;
}

/* Transcript::new(label) + Prover::new + m x Prover::commit + Prover::prove -> R1CSProof::to_bytes.
 * returns proof length, or -1 on error.  V_out (32*m) receives the commitments. */
long bpo_prove(const uint8_t* label, size_t label_len, const uint8_t* v32, const uint8_t* vbl32, size_t m,
               const uint8_t* aL32, const uint8_t* aR32, const uint8_t* aO32, size_t n,
               const uint32_t* row_start, const uint32_t* term_var, const uint8_t* term_coef32, size_t q,
               const uint8_t seed[32], int cache_gens, uint8_t* V_out, uint8_t* proof_out, size_t proof_cap) {
    transcript_t T; t_new(&T, label, label_len);
    t_append(&T, "dom-sep", "r1cs v1", 7);
    const size_t npad = next_pow2(n ? n : 1);
    size_t lg = 0; while (((size_t)1 << lg) < npad) lg++;
    const size_t plen = 1 + 13 * 32 + 64 * lg;
    if (plen > proof_cap) return -1;
    gens_t local; const gens_t* g = get_gens(npad, cache_gens, &local);
    scl* vbl = malloc(sizeof(scl) * (m + 1));
    for (size_t i = 0; i < m; i++) {
        vbl[i] = scl_reduce(scl_from_bytes(vbl32 + 32 * i));
        ge_p3 V = pedersen_commit(g, scl_reduce(scl_from_bytes(v32 + 32 * i)), vbl[i]);
        ristretto_compress(V_out + 32 * i, &V);
        t_append(&T, "V", V_out + 32 * i, 32);
    }
    const size_t nterms = row_start[q];
    scl* coef = malloc(sizeof(scl) * (nterms + 1));
    for (size_t e = 0; e < nterms; e++) coef[e] = scl_reduce(scl_from_bytes(term_coef32 + 32 * e));
    cs_t cs = {row_start, term_var, coef, q};
    scl *aL = malloc(sizeof(scl) * npad), *aR = malloc(sizeof(scl) * npad), *aO = malloc(sizeof(scl) * npad);
    for (size_t i = 0; i < n; i++) { aL[i] = scl_reduce(scl_from_bytes(aL32 + 32 * i)); aR[i] = scl_reduce(scl_from_bytes(aR32 + 32 * i)); aO[i] = scl_reduce(scl_from_bytes(aO32 + 32 * i)); }

    /* ---- Prover::prove ---- */
    t_append_u64(&T, "m", m);
    strobe_t rng = T;
    for (size_t i = 0; i < m; i++) { uint32_t len = 32; strobe_meta_ad(&rng, "v_blinding", 10, 0); strobe_meta_ad(&rng, &len, 4, 1); strobe_key(&rng, vbl32 + 32 * i, 32); }
    strobe_meta_ad(&rng, "rng", 3, 0); strobe_key(&rng, seed, 32);
    scl i_bl = rng_scalar(&rng), o_bl = rng_scalar(&rng), s_bl = rng_scalar(&rng);
    scl *sL = malloc(sizeof(scl) * npad), *sR = malloc(sizeof(scl) * npad);
    for (size_t i = 0; i < n; i++) sL[i] = rng_scalar(&rng);
    for (size_t i = 0; i < n; i++) sR[i] = rng_scalar(&rng);
    scl* sbuf = malloc(sizeof(scl) * (2 * npad + 2)); ge_p3* pbuf = malloc(sizeof(ge_p3) * (2 * npad + 2));
    uint8_t A_I1[32], A_O1[32], S1[32];
    sbuf[0] = i_bl; pbuf[0] = g->Bb; for (size_t i = 0; i < n; i++) { sbuf[1 + i] = aL[i]; pbuf[1 + i] = g->G[i]; sbuf[1 + n + i] = aR[i]; pbuf[1 + n + i] = g->H[i]; }
    { ge_p3 P = msm_straus_ct(sbuf, pbuf, 2 * n + 1); ristretto_compress(A_I1, &P); }
    sbuf[0] = o_bl; for (size_t i = 0; i < n; i++) sbuf[1 + i] = aO[i];
    { ge_p3 P = msm_straus_ct(sbuf, pbuf, n + 1); ristretto_compress(A_O1, &P); }
    sbuf[0] = s_bl; for (size_t i = 0; i < n; i++) { sbuf[1 + i] = sL[i]; sbuf[1 + n + i] = sR[i]; }
    { ge_p3 P = msm_straus_ct(sbuf, pbuf, 2 * n + 1); ristretto_compress(S1, &P); }
    static const uint8_t ZERO32[32] = {0};
    t_append(&T, "A_I1", A_I1, 32); t_append(&T, "A_O1", A_O1, 32); t_append(&T, "S1", S1, 32);
    t_append(&T, "dom-sep", "r1cs-1phase", 11);
    t_append(&T, "A_I2", ZERO32, 32); t_append(&T, "A_O2", ZERO32, 32); t_append(&T, "S2", ZERO32, 32);
    scl y = t_challenge_scalar(&T, "y"), z = t_challenge_scalar(&T, "z");
    scl *wL = malloc(sizeof(scl) * npad), *wR = malloc(sizeof(scl) * npad), *wO = malloc(sizeof(scl) * npad), *wV = malloc(sizeof(scl) * (m + 1));
    flatten(&cs, z, n, m, wL, wR, wO, wV, NULL);
    scl y_inv = scl_invert(y);
    scl* exp_y_inv = malloc(sizeof(scl) * npad);
    exp_y_inv[0] = SC_ONE; for (size_t i = 1; i < npad; i++) exp_y_inv[i] = scl_mul(exp_y_inv[i - 1], y_inv);
    scl *l1 = malloc(sizeof(scl) * npad), *r0 = malloc(sizeof(scl) * npad), *r1 = malloc(sizeof(scl) * npad), *r3 = malloc(sizeof(scl) * npad);
    scl exp_y = SC_ONE;
    for (size_t i = 0; i < n; i++) {
        l1[i] = scl_add(aL[i], scl_mul(exp_y_inv[i], wR[i]));
        r0[i] = scl_sub(wO[i], exp_y);
        r1[i] = scl_add(scl_mul(exp_y, aR[i]), wL[i]);
        r3[i] = scl_mul(exp_y, sR[i]);
        exp_y = scl_mul(exp_y, y);
    }
    scl t1 = inner_product(l1, r0, n);
    scl t2 = scl_add(inner_product(l1, r1, n), inner_product(aO, r0, n));
    scl t3 = scl_add(inner_product(aO, r1, n), inner_product(sL, r0, n));
    scl t4 = scl_add(inner_product(l1, r3, n), inner_product(sL, r1, n));
    scl t5 = inner_product(aO, r3, n), t6 = inner_product(sL, r3, n);
    scl tb1 = rng_scalar(&rng), tb3 = rng_scalar(&rng), tb4 = rng_scalar(&rng), tb5 = rng_scalar(&rng), tb6 = rng_scalar(&rng);
    uint8_t Tc[5][32];
    { scl tv[5] = {t1, t3, t4, t5, t6}, tr[5] = {tb1, tb3, tb4, tb5, tb6}; const char* lab[5] = {"T_1", "T_3", "T_4", "T_5", "T_6"};
      for (int k = 0; k < 5; k++) { ge_p3 P = pedersen_commit(g, tv[k], tr[k]); ristretto_compress(Tc[k], &P); }
      for (int k = 0; k < 5; k++) t_append(&T, lab[k], Tc[k], 32); }
    scl u = t_challenge_scalar(&T, "u"), x = t_challenge_scalar(&T, "x");
    scl tb2 = inner_product(wV, vbl, m);
#define POLY6(c1, c2, c3, c4, c5, c6) scl_mul(x, scl_add(c1, scl_mul(x, scl_add(c2, scl_mul(x, scl_add(c3, scl_mul(x, scl_add(c4, scl_mul(x, scl_add(c5, scl_mul(x, c6)))))))))))
    scl t_x = POLY6(t1, t2, t3, t4, t5, t6), t_xb = POLY6(tb1, tb2, tb3, tb4, tb5, tb6);
    scl *lv = malloc(sizeof(scl) * npad), *rv = malloc(sizeof(scl) * npad);
    for (size_t i = 0; i < n; i++) {
        lv[i] = scl_mul(x, scl_add(l1[i], scl_mul(x, scl_add(aO[i], scl_mul(x, sL[i])))));
        rv[i] = scl_add(r0[i], scl_mul(x, scl_add(r1[i], scl_mul(x, scl_mul(x, r3[i])))));
    }
    for (size_t i = n; i < npad; i++) { lv[i] = SC_ZERO; rv[i] = scl_neg(exp_y); exp_y = scl_mul(exp_y, y); }
    scl e_bl = scl_mul(x, scl_add(i_bl, scl_mul(x, scl_add(o_bl, scl_mul(x, s_bl)))));
    t_append_scalar(&T, "t_x", t_x); t_append_scalar(&T, "t_x_blinding", t_xb); t_append_scalar(&T, "e_blinding", e_bl);
    scl w = t_challenge_scalar(&T, "w");
    ge_p3 Q; { scl s1[1] = {w}; ge_p3 p1[1] = {g->B}; Q = msm_straus_ct(s1, p1, 1); }
    scl *Gf = malloc(sizeof(scl) * npad), *Hf = malloc(sizeof(scl) * npad);
    for (size_t i = 0; i < npad; i++) { Gf[i] = i < n ? SC_ONE : u; Hf[i] = scl_mul(exp_y_inv[i], Gf[i]); }
    ge_p3 *Gc = malloc(sizeof(ge_p3) * npad), *Hc = malloc(sizeof(ge_p3) * npad);
    memcpy(Gc, g->G, sizeof(ge_p3) * npad); memcpy(Hc, g->H, sizeof(ge_p3) * npad);
    uint8_t* o = proof_out; *o++ = 0;
    memcpy(o, A_I1, 32); o += 32; memcpy(o, A_O1, 32); o += 32; memcpy(o, S1, 32); o += 32;
    for (int k = 0; k < 5; k++) { memcpy(o, Tc[k], 32); o += 32; }
    scl_to_bytes(o, t_x); o += 32; scl_to_bytes(o, t_xb); o += 32; scl_to_bytes(o, e_bl); o += 32;
    scl a_fin, b_fin;
    ipp_create(&T, &Q, Gf, Hf, Gc, Hc, lv, rv, npad, o, &a_fin, &b_fin);
    o += 64 * lg;
    scl_to_bytes(o, a_fin); o += 32; scl_to_bytes(o, b_fin); o += 32;
    free(vbl); free(coef); free(aL); free(aR); free(aO); free(sL); free(sR); free(sbuf); free(pbuf); free(wL); free(wR); free(wO); free(wV);
    free(exp_y_inv); free(l1); free(r0); free(r1); free(r3); free(lv); free(rv); free(Gf); free(Hf); free(Gc); free(Hc);
    if (!cache_gens) gens_free(&local);
    return (long)(o - proof_out);
}

/* Transcript::new + Verifier::new + m x commit + R1CSProof::from_bytes + Verifier::verify.
 * returns 1 accepted, 0 rejected (VerificationError), -1 malformed proof (FormatError). */
int bpo_verify(const uint8_t* label, size_t label_len, const uint8_t* V32, size_t m, size_t n,
               const uint32_t* row_start, const uint32_t* term_var, const uint8_t* term_coef32, size_t q,
               const uint8_t* proof, size_t proof_len, const uint8_t seed[32], int cache_gens) {
    if (proof_len == 0) return -1;
    const uint8_t version = proof[0]; const uint8_t* body = proof + 1; size_t blen = proof_len - 1;
    if (blen % 32) return -1;
    size_t minlen; if (version == 0) minlen = 11 * 32; else if (version == 1) minlen = 14 * 32; else return -1;
    if (blen < minlen) return -1;
    const uint8_t *A_I1 = body, *A_O1 = body + 32, *S1 = body + 64; size_t pos = 96;
    static const uint8_t ZERO32[32] = {0};
    const uint8_t *A_I2 = ZERO32, *A_O2 = ZERO32, *S2 = ZERO32;
    if (version == 1) { A_I2 = body + pos; A_O2 = body + pos + 32; S2 = body + pos + 64; pos += 96; }
    const uint8_t* Tp[5]; for (int k = 0; k < 5; k++) { Tp[k] = body + pos; pos += 32; }
    scl t_x = scl_from_bytes(body + pos), t_xb = scl_from_bytes(body + pos + 32), e_bl = scl_from_bytes(body + pos + 64); pos += 96;
    if (!scl_is_canonical(t_x) || !scl_is_canonical(t_xb) || !scl_is_canonical(e_bl)) return -1;
    size_t ne = (blen - pos) / 32; if (ne < 2 || (ne - 2) % 2) return -1;
    size_t lg = (ne - 2) / 2; if (lg >= 32) return -1;
    const uint8_t* lr = body + pos; pos += 64 * lg;
    scl a = scl_from_bytes(body + pos), b = scl_from_bytes(body + pos + 32);
    if (!scl_is_canonical(a) || !scl_is_canonical(b)) return -1;

    transcript_t T; t_new(&T, label, label_len);
    t_append(&T, "dom-sep", "r1cs v1", 7);
    for (size_t i = 0; i < m; i++) t_append(&T, "V", V32 + 32 * i, 32);
    t_append_u64(&T, "m", m);
#define VAL_APPEND(lab, p) do { if (memcmp(p, ZERO32, 32) == 0) return 0; t_append(&T, lab, p, 32); } while (0)
    VAL_APPEND("A_I1", A_I1); VAL_APPEND("A_O1", A_O1); VAL_APPEND("S1", S1);
    t_append(&T, "dom-sep", "r1cs-1phase", 11);
    const size_t npad = next_pow2(n ? n : 1);
    t_append(&T, "A_I2", A_I2, 32); t_append(&T, "A_O2", A_O2, 32); t_append(&T, "S2", S2, 32);
    scl y = t_challenge_scalar(&T, "y"), z = t_challenge_scalar(&T, "z");
    VAL_APPEND("T_1", Tp[0]); VAL_APPEND("T_3", Tp[1]); VAL_APPEND("T_4", Tp[2]); VAL_APPEND("T_5", Tp[3]); VAL_APPEND("T_6", Tp[4]);
    scl u = t_challenge_scalar(&T, "u"), x = t_challenge_scalar(&T, "x");
    t_append_scalar(&T, "t_x", t_x); t_append_scalar(&T, "t_x_blinding", t_xb); t_append_scalar(&T, "e_blinding", e_bl);
    scl w = t_challenge_scalar(&T, "w");
    const size_t nterms = row_start[q];
    scl* coef = malloc(sizeof(scl) * (nterms + 1));
    for (size_t e = 0; e < nterms; e++) coef[e] = scl_reduce(scl_from_bytes(term_coef32 + 32 * e));
    cs_t cs = {row_start, term_var, coef, q};
    scl *wL = malloc(sizeof(scl) * npad), *wR = malloc(sizeof(scl) * npad), *wO = malloc(sizeof(scl) * npad), *wV = malloc(sizeof(scl) * (m + 1)), wc;
    flatten(&cs, z, n, m, wL, wR, wO, wV, &wc);
    int result = 0;
    scl *s = NULL, *scalars = NULL; ge_p3* points = NULL; scl* yiv = NULL; gens_t local; int have_local = 0;
    if (((size_t)1 << lg) != npad) goto done;
    t_append(&T, "dom-sep", "ipp v1", 6); t_append_u64(&T, "n", npad);
    scl ch[32], chi[32];
    for (size_t k = 0; k < lg; k++) {
        if (memcmp(lr + 64 * k, ZERO32, 32) == 0 || memcmp(lr + 64 * k + 32, ZERO32, 32) == 0) goto done;
        t_append(&T, "L", lr + 64 * k, 32); t_append(&T, "R", lr + 64 * k + 32, 32);
        ch[k] = t_challenge_scalar(&T, "u");
    }
    scl allinv = SC_ONE;
    for (size_t k = 0; k < lg; k++) { chi[k] = scl_invert(ch[k]); allinv = scl_mul(allinv, chi[k]); }
    scl ch_sq[32], chi_sq[32];
    for (size_t k = 0; k < lg; k++) { ch_sq[k] = scl_mul(ch[k], ch[k]); chi_sq[k] = scl_mul(chi[k], chi[k]); }
    s = malloc(sizeof(scl) * npad);
    s[0] = allinv;
    for (size_t i = 1; i < npad; i++) { int lg_i = 63 - __builtin_clzll(i); size_t k = (size_t)1 << lg_i; s[i] = scl_mul(s[i - k], ch_sq[(lg - 1) - lg_i]); }
    scl y_inv = scl_invert(y);
    yiv = malloc(sizeof(scl) * npad);
    yiv[0] = SC_ONE; for (size_t i = 1; i < npad; i++) yiv[i] = scl_mul(yiv[i - 1], y_inv);
    scl delta = SC_ZERO;
    const size_t np = 6 + m + 5 + 2 + 2 * npad + 2 * lg;
    scalars = malloc(sizeof(scl) * np); points = malloc(sizeof(ge_p3) * np);
    strobe_t rng = T; strobe_meta_ad(&rng, "rng", 3, 0); strobe_key(&rng, seed, 32);
    scl r = rng_scalar(&rng);
    scl xx = scl_mul(x, x), rxx = scl_mul(r, xx), xxx = scl_mul(x, xx);
    const gens_t* g = get_gens(npad, cache_gens, &local); have_local = !cache_gens;
    size_t k = 0;
    const uint8_t* encs[6] = {A_I1, A_O1, S1, A_I2, A_O2, S2};
    scl hs[6] = {x, xx, xxx, scl_mul(u, x), scl_mul(u, xx), scl_mul(u, xxx)};
    for (int i = 0; i < 6; i++) { if (!ristretto_decompress(&points[k], encs[i])) goto done; scalars[k++] = hs[i]; }
    for (size_t j = 0; j < m; j++) { if (!ristretto_decompress(&points[k], V32 + 32 * j)) goto done; scalars[k++] = scl_mul(wV[j], rxx); }
    scl Ts[5] = {scl_mul(r, x), scl_mul(rxx, x), scl_mul(rxx, xx), scl_mul(rxx, xxx), scl_mul(scl_mul(rxx, xx), xx)};
    for (int i = 0; i < 5; i++) { if (!ristretto_decompress(&points[k], Tp[i])) goto done; scalars[k++] = Ts[i]; }
    for (size_t i = 0; i < n; i++) delta = scl_add(delta, scl_mul(scl_mul(wR[i], yiv[i]), wL[i]));
    scalars[k] = scl_add(scl_mul(w, scl_sub(t_x, scl_mul(a, b))), scl_mul(r, scl_sub(scl_mul(xx, scl_add(wc, delta)), t_x))); points[k++] = g->B;
    scalars[k] = scl_sub(scl_neg(e_bl), scl_mul(r, t_xb)); points[k++] = g->Bb;
    for (size_t i = 0; i < npad; i++) {
        scl uf = i < n ? SC_ONE : u, ywr = i < n ? scl_mul(wR[i], yiv[i]) : SC_ZERO;
        scalars[k] = scl_mul(uf, scl_sub(scl_mul(x, ywr), scl_mul(a, s[i]))); points[k++] = g->G[i];
    }
    for (size_t i = 0; i < npad; i++) {
        scl uf = i < n ? SC_ONE : u, wl = i < n ? wL[i] : SC_ZERO, wo = i < n ? wO[i] : SC_ZERO;
        scalars[k] = scl_mul(uf, scl_sub(scl_mul(yiv[i], scl_sub(scl_add(scl_mul(x, wl), wo), scl_mul(b, s[npad - 1 - i]))), SC_ONE)); points[k++] = g->H[i];
    }
    for (size_t i = 0; i < lg; i++) { if (!ristretto_decompress(&points[k], lr + 64 * i)) goto done; scalars[k++] = ch_sq[i]; }
    for (size_t i = 0; i < lg; i++) { if (!ristretto_decompress(&points[k], lr + 64 * i + 32)) goto done; scalars[k++] = chi_sq[i]; }
    { ge_p3 mega = msm_vartime(scalars, points, k); result = ristretto_is_identity(&mega); }
done:
    free(coef); free(wL); free(wR); free(wO); free(wV); free(s); free(scalars); free(points); free(yiv);
    if (have_local) gens_free(&local);
    return result;
}

/* ---- primitives exposed for pinning and for the raw-MSM baseline ---- */
/* vartime_multiscalar_mul over compressed points; 0 on success, -1 if a point fails to decode */
int bpo_msm(const uint8_t* scalars32, const uint8_t* points32, size_t n, int constant_time, uint8_t out[32]) {
    scl* s = malloc(sizeof(scl) * (n + 1)); ge_p3* p = malloc(sizeof(ge_p3) * (n + 1));
    int rc = 0;
    for (size_t i = 0; i < n && rc == 0; i++) { s[i] = scl_from_bytes(scalars32 + 32 * i); if (!ristretto_decompress(&p[i], points32 + 32 * i)) rc = -1; }
    if (rc == 0) { ge_p3 r = constant_time ? msm_straus_ct(s, p, n) : msm_vartime(s, p, n); ristretto_compress(out, &r); }
    free(s); free(p);
    return rc;
}
/* sum sG_i G_i + sum sH_i H_i + sB B + sBb B_blinding with dalek's vartime dispatch */
int bpo_msm_gens(const uint8_t* sG, size_t nG, const uint8_t* sH, size_t nH, const uint8_t* sB, const uint8_t* sBb, int constant_time, uint8_t out[32]) {
    size_t cap = nG > nH ? nG : nH;
    const gens_t* g = get_gens(cap ? cap : 1, 1, NULL);
    size_t n = nG + nH + 2, k = 0;
    scl* s = malloc(sizeof(scl) * n); ge_p3* p = malloc(sizeof(ge_p3) * n);
    for (size_t i = 0; i < nG; i++) { s[k] = scl_from_bytes(sG + 32 * i); p[k++] = g->G[i]; }
    for (size_t i = 0; i < nH; i++) { s[k] = scl_from_bytes(sH + 32 * i); p[k++] = g->H[i]; }
    if (sB) { s[k] = scl_from_bytes(sB); p[k++] = g->B; }
    if (sBb) { s[k] = scl_from_bytes(sBb); p[k++] = g->Bb; }
    ge_p3 r = constant_time ? msm_straus_ct(s, p, k) : msm_vartime(s, p, k);
    ristretto_compress(out, &r);
    free(s); free(p);
    return 0;
}
void bpo_gens(int which, size_t start, size_t count, uint8_t* out) {
    const gens_t* g = get_gens(start + count ? start + count : 1, 1, NULL);
    for (size_t i = 0; i < count; i++) {
        const ge_p3* p = which == 0 ? &g->G[start + i] : which == 1 ? &g->H[start + i] : which == 2 ? &g->B : &g->Bb;
        ristretto_compress(out + 32 * i, p);
    }
}
void bpo_from_uniform(const uint8_t in[64], uint8_t out[32]) { ge_p3 p = ristretto_from_uniform_bytes(in); ristretto_compress(out, &p); }
int bpo_decompress_compress(const uint8_t in[32], uint8_t out[32]) { ge_p3 p; if (!ristretto_decompress(&p, in)) return -1; ristretto_compress(out, &p); return 0; }
void bpo_scalar_mul(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) { scl_to_bytes(out, scl_mul(scl_from_bytes(a), scl_from_bytes(b))); }
void bpo_scalar_from_wide(const uint8_t in[64], uint8_t out[32]) { scl_to_bytes(out, scl_from_wide(in)); }
void bpo_scalar_invert(const uint8_t a[32], uint8_t out[32]) { scl_to_bytes(out, scl_invert(scl_from_bytes(a))); }
void bpo_merlin_kat(const uint8_t* label, size_t ll, const uint8_t* mlabel, const uint8_t* msg, size_t ml, const uint8_t* clabel, uint8_t* out, size_t n) {
    transcript_t T; t_new(&T, label, ll); t_append(&T, (const char*)mlabel, msg, ml); t_challenge(&T, (const char*)clabel, out, n);
}
void bpo_free_cache(void) { gens_free(&g_cache); }
