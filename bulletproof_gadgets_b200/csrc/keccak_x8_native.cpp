// Eight Merlin TranscriptRng streams in lockstep: Keccak-f[1600] x8 with AVX-512 (one 64-bit lane of eight
// independent states per zmm register).
//
// Why: Prover::prove draws s_L, s_R from merlin's TranscriptRng (bulletproofs r1cs/prover.rs; reached from
// /root/reference/src/prove.rs:79): 2n sequential 64-byte `fill_bytes` calls = 2n dependent Keccak-f
// permutations per proof, a strict hash chain that must stay bit-exact.  One chain cannot be vectorised, but the
// chains of DIFFERENT proofs are independent, and every `fill_bytes(64)` is the same STROBE op sequence from the
// same position, so eight proofs' draws run as one SIMD program (merlin.cpp: RngBatcher).  Compiled by g++ directly
// (build.py); selected at run time when the CPU has AVX-512F.
#include <immintrin.h>
#include <stdint.h>
#include <string.h>

#define X8_TARGET __attribute__((target("avx512f")))

namespace bpg {

static const uint64_t RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};

bool keccak_x8_supported() {
    __builtin_cpu_init();
    return __builtin_cpu_supports("avx512f");
}

#define XOR3(a, b, c) _mm512_ternarylogic_epi64(a, b, c, 0x96)
#define CHI(a, b, c) _mm512_ternarylogic_epi64(a, b, c, 0xD2) /* a ^ (~b & c) */
#define ROL(a, n) _mm512_rol_epi64(a, n)

// one round: A -> E (theta, rho, pi, chi, iota), lane index = x + 5*y
#define KROUND(A, E, rc)                                                                                        \
    do {                                                                                                        \
        __m512i C0 = XOR3(XOR3(A[0], A[5], A[10]), A[15], A[20]);                                               \
        __m512i C1 = XOR3(XOR3(A[1], A[6], A[11]), A[16], A[21]);                                               \
        __m512i C2 = XOR3(XOR3(A[2], A[7], A[12]), A[17], A[22]);                                               \
        __m512i C3 = XOR3(XOR3(A[3], A[8], A[13]), A[18], A[23]);                                               \
        __m512i C4 = XOR3(XOR3(A[4], A[9], A[14]), A[19], A[24]);                                               \
        __m512i R0 = ROL(C1, 1), R1 = ROL(C2, 1), R2 = ROL(C3, 1), R3 = ROL(C4, 1), R4 = ROL(C0, 1);            \
        /* D[x] = C[x-1] ^ rol(C[x+1], 1);  B[y + 5*((2x+3y)%5)] = rol(A[x+5y] ^ D[x], r[x][y]) */              \
        __m512i B0, B1, B2, B3, B4;                                                                             \
        B0 = XOR3(A[0], C4, R0);                                                                                \
        B1 = ROL(XOR3(A[6], C0, R1), 44);                                                                       \
        B2 = ROL(XOR3(A[12], C1, R2), 43);                                                                      \
        B3 = ROL(XOR3(A[18], C2, R3), 21);                                                                      \
        B4 = ROL(XOR3(A[24], C3, R4), 14);                                                                      \
        E[0] = _mm512_xor_si512(CHI(B0, B1, B2), _mm512_set1_epi64((long long)(rc)));                           \
        E[1] = CHI(B1, B2, B3);                                                                                 \
        E[2] = CHI(B2, B3, B4);                                                                                 \
        E[3] = CHI(B3, B4, B0);                                                                                 \
        E[4] = CHI(B4, B0, B1);                                                                                 \
        B0 = ROL(XOR3(A[3], C2, R3), 28);                                                                       \
        B1 = ROL(XOR3(A[9], C3, R4), 20);                                                                       \
        B2 = ROL(XOR3(A[10], C4, R0), 3);                                                                       \
        B3 = ROL(XOR3(A[16], C0, R1), 45);                                                                      \
        B4 = ROL(XOR3(A[22], C1, R2), 61);                                                                      \
        E[5] = CHI(B0, B1, B2);                                                                                 \
        E[6] = CHI(B1, B2, B3);                                                                                 \
        E[7] = CHI(B2, B3, B4);                                                                                 \
        E[8] = CHI(B3, B4, B0);                                                                                 \
        E[9] = CHI(B4, B0, B1);                                                                                 \
        B0 = ROL(XOR3(A[1], C0, R1), 1);                                                                        \
        B1 = ROL(XOR3(A[7], C1, R2), 6);                                                                        \
        B2 = ROL(XOR3(A[13], C2, R3), 25);                                                                      \
        B3 = ROL(XOR3(A[19], C3, R4), 8);                                                                       \
        B4 = ROL(XOR3(A[20], C4, R0), 18);                                                                      \
        E[10] = CHI(B0, B1, B2);                                                                                \
        E[11] = CHI(B1, B2, B3);                                                                                \
        E[12] = CHI(B2, B3, B4);                                                                                \
        E[13] = CHI(B3, B4, B0);                                                                                \
        E[14] = CHI(B4, B0, B1);                                                                                \
        B0 = ROL(XOR3(A[4], C3, R4), 27);                                                                       \
        B1 = ROL(XOR3(A[5], C4, R0), 36);                                                                       \
        B2 = ROL(XOR3(A[11], C0, R1), 10);                                                                      \
        B3 = ROL(XOR3(A[17], C1, R2), 15);                                                                      \
        B4 = ROL(XOR3(A[23], C2, R3), 56);                                                                      \
        E[15] = CHI(B0, B1, B2);                                                                                \
        E[16] = CHI(B1, B2, B3);                                                                                \
        E[17] = CHI(B2, B3, B4);                                                                                \
        E[18] = CHI(B3, B4, B0);                                                                                \
        E[19] = CHI(B4, B0, B1);                                                                                \
        B0 = ROL(XOR3(A[2], C1, R2), 62);                                                                       \
        B1 = ROL(XOR3(A[8], C2, R3), 55);                                                                       \
        B2 = ROL(XOR3(A[14], C3, R4), 39);                                                                      \
        B3 = ROL(XOR3(A[15], C4, R0), 41);                                                                      \
        B4 = ROL(XOR3(A[21], C0, R1), 2);                                                                       \
        E[20] = CHI(B0, B1, B2);                                                                                \
        E[21] = CHI(B1, B2, B3);                                                                                \
        E[22] = CHI(B2, B3, B4);                                                                                \
        E[23] = CHI(B3, B4, B0);                                                                                \
        E[24] = CHI(B4, B0, B1);                                                                                \
    } while (0)

X8_TARGET static inline void permute_x8(__m512i A[25]) {
    __m512i E[25];
    for (int r = 0; r < 24; r += 2) {
        KROUND(A, E, RC[r]);
        KROUND(E, A, RC[r + 1]);
    }
}

// plain permutation of eight states (lane-interleaved u64[25][8]); used by the self-test
X8_TARGET void keccak_f1600_x8(uint64_t st[25][8]) {
    __m512i A[25];
    for (int i = 0; i < 25; i++) A[i] = _mm512_loadu_si512(st[i]);
    permute_x8(A);
    for (int i = 0; i < 25; i++) _mm512_storeu_si512(st[i], A[i]);
}

// 8x8 transpose of 64-bit words: r[k] (state lane k of the eight streams) -> r[p] (lanes 0..7 of stream p)
X8_TARGET static inline void transpose8(__m512i r[8]) {
    __m512i t[8], u[8];
    for (int i = 0; i < 4; i++) {
        t[2 * i] = _mm512_unpacklo_epi64(r[2 * i], r[2 * i + 1]);
        t[2 * i + 1] = _mm512_unpackhi_epi64(r[2 * i], r[2 * i + 1]);
    }
    // t[2i]   = a0 b0 a2 b2 a4 b4 a6 b6   (a = r[2i], b = r[2i+1]);  t[2i+1] = a1 b1 a3 b3 a5 b5 a7 b7
    u[0] = _mm512_shuffle_i64x2(t[0], t[2], 0x88);  // 128-bit blocks 0,2 of t0 | 0,2 of t2
    u[1] = _mm512_shuffle_i64x2(t[1], t[3], 0x88);
    u[2] = _mm512_shuffle_i64x2(t[0], t[2], 0xdd);  // blocks 1,3
    u[3] = _mm512_shuffle_i64x2(t[1], t[3], 0xdd);
    u[4] = _mm512_shuffle_i64x2(t[4], t[6], 0x88);
    u[5] = _mm512_shuffle_i64x2(t[5], t[7], 0x88);
    u[6] = _mm512_shuffle_i64x2(t[4], t[6], 0xdd);
    u[7] = _mm512_shuffle_i64x2(t[5], t[7], 0xdd);
    // u[0] = (r0r1 col0)(r0r1 col4)(r2r3 col0)(r2r3 col4) ...
    r[0] = _mm512_shuffle_i64x2(u[0], u[4], 0x88);
    r[4] = _mm512_shuffle_i64x2(u[0], u[4], 0xdd);
    r[1] = _mm512_shuffle_i64x2(u[1], u[5], 0x88);
    r[5] = _mm512_shuffle_i64x2(u[1], u[5], 0xdd);
    r[2] = _mm512_shuffle_i64x2(u[2], u[6], 0x88);
    r[6] = _mm512_shuffle_i64x2(u[2], u[6], 0xdd);
    r[3] = _mm512_shuffle_i64x2(u[3], u[7], 0x88);
    r[7] = _mm512_shuffle_i64x2(u[3], u[7], 0xdd);
}

// `count[p]` steady-state `TranscriptRng::fill_bytes(64)` draws for up to eight STROBE-128 states.
//   state[p]: 200-byte STROBE state with pos = 64, pos_begin = 0 (i.e. right after a previous 64-byte draw);
//   out[p]:   receives 64 * count[p] bytes.  Streams with p >= n are ignored.
// Every draw XORs the same framing into the state (meta-AD header + LE32(64), PRF header, run_f padding), permutes,
// copies bytes 0..63 out and zeroes them -- exactly Strobe128::{meta_ad, prf} of merlin 2.0.1 (see merlin.cpp).
X8_TARGET void strobe_rng_fill64_x8(uint8_t* const state[8], uint8_t* const out[8], const size_t count[8], int n) {
    uint64_t lanes[25][8];
    memset(lanes, 0, sizeof lanes);
    size_t maxc = 0;
    for (int p = 0; p < n; p++) {
        uint64_t s[25];
        memcpy(s, state[p], 200);
        for (int k = 0; k < 25; k++) lanes[k][p] = s[k];
        if (count[p] > maxc) maxc = count[p];
    }
    __m512i A[25];
    for (int k = 0; k < 25; k++) A[k] = _mm512_loadu_si512(lanes[k]);
    // bytes 64..73: [old_begin=0, M|A=0x12] [64,0,0,0] [old_begin=65, I|A|C=0x07] then run_f: [pos_begin=71] [0x04]; byte 167 ^= 0x80
    const __m512i F8 = _mm512_set1_epi64(0x0741000000401200LL), F9 = _mm512_set1_epi64(0x0447LL),
                  F20 = _mm512_set1_epi64((long long)0x8000000000000000ULL), Z = _mm512_setzero_si512();
    for (size_t i = 0; i < maxc; i++) {
        A[8] = _mm512_xor_si512(A[8], F8);
        A[9] = _mm512_xor_si512(A[9], F9);
        A[20] = _mm512_xor_si512(A[20], F20);
        permute_x8(A);
        __m512i r[8] = {A[0], A[1], A[2], A[3], A[4], A[5], A[6], A[7]};
        transpose8(r);
        for (int p = 0; p < n; p++)
            if (i < count[p]) _mm512_storeu_si512(out[p] + 64 * i, r[p]);
        for (int k = 0; k < 8; k++) A[k] = Z;
        for (int p = 0; p < n; p++)
            if (i + 1 == count[p]) {  // this stream is done: hand its state back before the others move on
                for (int k = 0; k < 25; k++) {
                    _mm512_storeu_si512(lanes[k], A[k]);
                }
                uint64_t s[25];
                for (int k = 0; k < 25; k++) s[k] = lanes[k][p];
                memcpy(state[p], s, 200);
            }
    }
}

}  // namespace bpg
