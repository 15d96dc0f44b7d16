// See merlin.hpp.
#include "merlin.hpp"

namespace bpg {

static inline uint64_t rol(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

void keccak_f1600(uint64_t s[25]) {
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
            uint64_t t = bc[(i + 4) % 5] ^ rol(bc[(i + 1) % 5], 1);
            for (int j = 0; j < 25; j += 5) s[j + i] ^= t;
        }
        uint64_t t = s[1];
        for (int i = 0; i < 24; i++) {
            int j = PI[i];
            uint64_t b = s[j];
            s[j] = rol(t, RHO[i]);
            t = b;
        }
        for (int j = 0; j < 25; j += 5) {
            for (int i = 0; i < 5; i++) bc[i] = s[j + i];
            for (int i = 0; i < 5; i++) s[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
        }
        s[0] ^= RC[round];
    }
}

// little-endian hosts only (x86-64 / aarch64): state bytes alias the u64 lanes directly
void Sponge::absorb(const uint8_t* d, size_t n) {
    uint8_t* b = reinterpret_cast<uint8_t*>(st);
    for (size_t i = 0; i < n; i++) {
        b[pos++] ^= d[i];
        if (pos == rate) {
            keccak_f1600(st);
            pos = 0;
        }
    }
}
void Sponge::squeeze(uint8_t* out, size_t n) {
    uint8_t* b = reinterpret_cast<uint8_t*>(st);
    if (!squeezing) {
        b[pos] ^= suffix;
        b[rate - 1] ^= 0x80;
        keccak_f1600(st);
        pos = 0;
        squeezing = true;
    }
    for (size_t i = 0; i < n; i++) {
        if (pos == rate) {
            keccak_f1600(st);
            pos = 0;
        }
        out[i] = b[pos++];
    }
}
void sha3_512(const uint8_t* d, size_t n, uint8_t out[64]) {
    Sponge s(72, 0x06);
    s.absorb(d, n);
    s.squeeze(out, 64);
}

// ---------------------------------------------------------------------------- STROBE-128
static const uint8_t STROBE_R = 166;
enum { FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32 };

Strobe128::Strobe128(const char* protocol_label) {
    memset(state, 0, sizeof state);
    const uint8_t init[6] = {1, (uint8_t)(STROBE_R + 2), 1, 0, 1, 96};
    memcpy(state, init, 6);
    memcpy(state + 6, "STROBEv1.0.2", 12);
    keccak_f1600(reinterpret_cast<uint64_t*>(state));
    pos = pos_begin = cur_flags = 0;
    meta_ad(reinterpret_cast<const uint8_t*>(protocol_label), strlen(protocol_label), false);
}
void Strobe128::run_f() {
    state[pos] ^= pos_begin;
    state[pos + 1] ^= 0x04;
    state[STROBE_R + 1] ^= 0x80;
    keccak_f1600(reinterpret_cast<uint64_t*>(state));
    pos = 0;
    pos_begin = 0;
}
void Strobe128::absorb(const uint8_t* d, size_t n) {
    for (size_t i = 0; i < n; i++) {
        state[pos++] ^= d[i];
        if (pos == STROBE_R) run_f();
    }
}
void Strobe128::overwrite(const uint8_t* d, size_t n) {
    for (size_t i = 0; i < n; i++) {
        state[pos++] = d[i];
        if (pos == STROBE_R) run_f();
    }
}
void Strobe128::squeeze(uint8_t* d, size_t n) {
    for (size_t i = 0; i < n; i++) {
        d[i] = state[pos];
        state[pos++] = 0;
        if (pos == STROBE_R) run_f();
    }
}
void Strobe128::begin_op(uint8_t flags, bool more) {
    if (more) return;  // continuing the current operation
    uint8_t old_begin = pos_begin;
    pos_begin = pos + 1;
    cur_flags = flags;
    uint8_t hdr[2] = {old_begin, flags};
    absorb(hdr, 2);
    if ((flags & (FLAG_C | FLAG_K)) && pos != 0) run_f();
}
void Strobe128::meta_ad(const uint8_t* d, size_t n, bool more) {
    begin_op(FLAG_M | FLAG_A, more);
    absorb(d, n);
}
void Strobe128::ad(const uint8_t* d, size_t n, bool more) {
    begin_op(FLAG_A, more);
    absorb(d, n);
}
void Strobe128::prf(uint8_t* out, size_t n, bool more) {
    begin_op(FLAG_I | FLAG_A | FLAG_C, more);
    squeeze(out, n);
}
void Strobe128::key(const uint8_t* d, size_t n, bool more) {
    begin_op(FLAG_A | FLAG_C, more);
    overwrite(d, n);
}

// ---------------------------------------------------------------------------- Transcript
static inline void le32(uint8_t o[4], uint32_t x) {
    o[0] = (uint8_t)x;
    o[1] = (uint8_t)(x >> 8);
    o[2] = (uint8_t)(x >> 16);
    o[3] = (uint8_t)(x >> 24);
}

Transcript::Transcript(const uint8_t* label, size_t n) : strobe("Merlin v1.0") { append_message("dom-sep", label, n); }

void Transcript::append_message(const uint8_t* label, size_t ln, const uint8_t* msg, size_t n) {
    uint8_t len[4];
    le32(len, (uint32_t)n);
    strobe.meta_ad(label, ln, false);
    strobe.meta_ad(len, 4, true);
    strobe.ad(msg, n, false);
}
void Transcript::append_message(const char* label, const uint8_t* msg, size_t n) {
    append_message(reinterpret_cast<const uint8_t*>(label), strlen(label), msg, n);
}
void Transcript::append_u64(const char* label, uint64_t x) {
    uint8_t b[8];
    for (int i = 0; i < 8; i++) b[i] = (uint8_t)(x >> (8 * i));
    append_message(label, b, 8);
}
void Transcript::challenge_bytes(const uint8_t* label, size_t ln, uint8_t* out, size_t n) {
    uint8_t len[4];
    le32(len, (uint32_t)n);
    strobe.meta_ad(label, ln, false);
    strobe.meta_ad(len, 4, true);
    strobe.prf(out, n, false);
}
void Transcript::challenge_bytes(const char* label, uint8_t* out, size_t n) {
    challenge_bytes(reinterpret_cast<const uint8_t*>(label), strlen(label), out, n);
}
TranscriptRng Transcript::build_rng(const std::vector<const uint8_t*>& witness32, const uint8_t external32[32]) const {
    Strobe128 s = strobe;
    uint8_t len[4];
    le32(len, 32);
    for (const uint8_t* w : witness32) {
        s.meta_ad(reinterpret_cast<const uint8_t*>("v_blinding"), 10, false);
        s.meta_ad(len, 4, true);
        s.key(w, 32, false);
    }
    s.meta_ad(reinterpret_cast<const uint8_t*>("rng"), 3, false);
    s.key(external32, 32, false);
    return TranscriptRng(s);
}
void TranscriptRng::fill_bytes(uint8_t* out, size_t n) {
    uint8_t len[4];
    le32(len, (uint32_t)n);
    strobe.meta_ad(len, 4, false);
    strobe.prf(out, n, false);
}

}  // namespace bpg
