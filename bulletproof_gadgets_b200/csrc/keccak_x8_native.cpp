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
#include <stdlib.h>
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

bool keccak_x8_supported() {  // BPG_KECCAK=scalar forces the portable code (tests cover both paths on AVX-512 hosts)
    const char* e = getenv("BPG_KECCAK");
    if (e && !strcmp(e, "scalar")) return false;
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

// ------------------------------------------------------------------------------------------------------------
// ONE state, AVX-512: for a proof that runs alone the chain cannot be batched, so shorten the chain itself.
// Five zmm registers hold the five planes (element x of register y = lane (x, y)); theta and rho work plane-wise,
// pi is one in-register permutation per plane that leaves the state column-major, where chi needs no shuffles at all
// (its x+1, x+2 neighbours are other registers); a 5x5 transpose (14 two-source permutes) restores plane order.
// About 40 micro-ops per round against ~150 scalar ones.
X8_TARGET void keccak_f1600_avx512(uint64_t st[25]) {
    const __mmask8 M5 = 0x1f;
    __m512i P0 = _mm512_maskz_loadu_epi64(M5, st), P1 = _mm512_maskz_loadu_epi64(M5, st + 5),
            P2 = _mm512_maskz_loadu_epi64(M5, st + 10), P3 = _mm512_maskz_loadu_epi64(M5, st + 15),
            P4 = _mm512_maskz_loadu_epi64(M5, st + 20);
    const __m512i IDX_M1 = _mm512_setr_epi64(4, 0, 1, 2, 3, 5, 6, 7), IDX_P1 = _mm512_setr_epi64(1, 2, 3, 4, 0, 5, 6, 7);
    // rho offsets r[x][y], one vector per plane y
    const __m512i RHO0 = _mm512_setr_epi64(0, 1, 62, 28, 27, 0, 0, 0), RHO1 = _mm512_setr_epi64(36, 44, 6, 55, 20, 0, 0, 0),
                  RHO2 = _mm512_setr_epi64(3, 10, 43, 25, 39, 0, 0, 0), RHO3 = _mm512_setr_epi64(41, 45, 15, 21, 8, 0, 0, 0),
                  RHO4 = _mm512_setr_epi64(18, 2, 61, 56, 14, 0, 0, 0);
    // pi: F[s][y'] = E[s][(3 y' + s) mod 5]   (s = old y = new x)
    const __m512i PI0 = _mm512_setr_epi64(0, 3, 1, 4, 2, 5, 6, 7), PI1 = _mm512_setr_epi64(1, 4, 2, 0, 3, 5, 6, 7),
                  PI2 = _mm512_setr_epi64(2, 0, 3, 1, 4, 5, 6, 7), PI3 = _mm512_setr_epi64(3, 1, 4, 2, 0, 5, 6, 7),
                  PI4 = _mm512_setr_epi64(4, 2, 0, 3, 1, 5, 6, 7);
    // transpose helpers
    const __m512i TM = _mm512_setr_epi64(0, 8, 1, 9, 2, 10, 3, 11), TN = _mm512_setr_epi64(4, 12, 4, 12, 4, 12, 4, 12);
    const __m512i T0 = _mm512_setr_epi64(0, 1, 8, 9, 0, 0, 0, 0), T1 = _mm512_setr_epi64(2, 3, 10, 11, 0, 0, 0, 0),
                  T2 = _mm512_setr_epi64(4, 5, 12, 13, 0, 0, 0, 0), T3 = _mm512_setr_epi64(6, 7, 14, 15, 0, 0, 0, 0);
    const __m512i Y0 = _mm512_set1_epi64(0), Y1 = _mm512_set1_epi64(1), Y2 = _mm512_set1_epi64(2), Y3 = _mm512_set1_epi64(3),
                  Y4 = _mm512_set1_epi64(4);
    for (int r = 0; r < 24; r++) {
        // theta
        const __m512i C = XOR3(XOR3(P0, P1, P2), P3, P4);
        const __m512i Cm1 = _mm512_permutexvar_epi64(IDX_M1, C);
        const __m512i Cp1 = ROL(_mm512_permutexvar_epi64(IDX_P1, C), 1);
        // theta + rho + pi
        const __m512i F0 = _mm512_permutexvar_epi64(PI0, _mm512_rolv_epi64(XOR3(P0, Cm1, Cp1), RHO0));
        const __m512i F1 = _mm512_permutexvar_epi64(PI1, _mm512_rolv_epi64(XOR3(P1, Cm1, Cp1), RHO1));
        const __m512i F2 = _mm512_permutexvar_epi64(PI2, _mm512_rolv_epi64(XOR3(P2, Cm1, Cp1), RHO2));
        const __m512i F3 = _mm512_permutexvar_epi64(PI3, _mm512_rolv_epi64(XOR3(P3, Cm1, Cp1), RHO3));
        const __m512i F4 = _mm512_permutexvar_epi64(PI4, _mm512_rolv_epi64(XOR3(P4, Cm1, Cp1), RHO4));
        // chi (column-major: register = x, element = y) and iota on lane (0, 0)
        __m512i Q0 = CHI(F0, F1, F2);
        const __m512i Q1 = CHI(F1, F2, F3), Q2 = CHI(F2, F3, F4), Q3 = CHI(F3, F4, F0), Q4 = CHI(F4, F0, F1);
        Q0 = _mm512_xor_si512(Q0, _mm512_maskz_set1_epi64(0x01, (long long)RC[r]));
        // transpose back to plane-major
        const __m512i M01 = _mm512_permutex2var_epi64(Q0, TM, Q1), M23 = _mm512_permutex2var_epi64(Q2, TM, Q3);
        const __m512i N01 = _mm512_permutex2var_epi64(Q0, TN, Q1), N23 = _mm512_permutex2var_epi64(Q2, TN, Q3);
        P0 = _mm512_mask_permutexvar_epi64(_mm512_permutex2var_epi64(M01, T0, M23), 0x10, Y0, Q4);
        P1 = _mm512_mask_permutexvar_epi64(_mm512_permutex2var_epi64(M01, T1, M23), 0x10, Y1, Q4);
        P2 = _mm512_mask_permutexvar_epi64(_mm512_permutex2var_epi64(M01, T2, M23), 0x10, Y2, Q4);
        P3 = _mm512_mask_permutexvar_epi64(_mm512_permutex2var_epi64(M01, T3, M23), 0x10, Y3, Q4);
        P4 = _mm512_mask_permutexvar_epi64(_mm512_permutex2var_epi64(N01, T0, N23), 0x10, Y4, Q4);
    }
    _mm512_mask_storeu_epi64(st, M5, P0);
    _mm512_mask_storeu_epi64(st + 5, M5, P1);
    _mm512_mask_storeu_epi64(st + 10, M5, P2);
    _mm512_mask_storeu_epi64(st + 15, M5, P3);
    _mm512_mask_storeu_epi64(st + 20, M5, P4);
}

}  // namespace bpg
