// Host side: Keccak-f[1600], SHAKE256 / SHA3-512, STROBE-128 and the Merlin transcript.
//
// The Fiat-Shamir transcript stays on the CPU (BASELINE.json north_star).  Restates merlin 2.0.1
// strobe.rs / transcript.rs (/root/reference/Cargo.lock:403-405, not vendored) and the
// bulletproofs `TranscriptProtocol` extension trait; reference call sites
// /root/reference/src/prove.rs:45 and /root/reference/src/verify.rs:44.
#pragma once
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

namespace bpg {

void keccak_f1600(uint64_t st[25]);

// incremental sponge for SHAKE256 (rate 136, suffix 0x1f) and SHA3-512 (rate 72, suffix 0x06)
struct Sponge {
    uint64_t st[25];
    size_t rate, pos;
    uint8_t suffix;
    bool squeezing;
    Sponge(size_t rate_, uint8_t suffix_) : rate(rate_), pos(0), suffix(suffix_), squeezing(false) {
        memset(st, 0, sizeof st);
    }
    void absorb(const uint8_t* d, size_t n);
    void squeeze(uint8_t* out, size_t n);
};
inline Sponge shake256() { return Sponge(136, 0x1f); }
void sha3_512(const uint8_t* d, size_t n, uint8_t out[64]);

struct Strobe128 {
    alignas(8) uint8_t state[200];
    uint8_t pos, pos_begin, cur_flags;
    explicit Strobe128(const char* protocol_label);
    void meta_ad(const uint8_t* d, size_t n, bool more);
    void ad(const uint8_t* d, size_t n, bool more);
    void prf(uint8_t* out, size_t n, bool more);
    void key(const uint8_t* d, size_t n, bool more);

   private:
    void run_f();
    void absorb(const uint8_t* d, size_t n);
    void overwrite(const uint8_t* d, size_t n);
    void squeeze(uint8_t* d, size_t n);
    void begin_op(uint8_t flags, bool more);
};

struct TranscriptRng {
    Strobe128 strobe;
    explicit TranscriptRng(const Strobe128& s) : strobe(s) {}
    void fill_bytes(uint8_t* out, size_t n);
    // `count` consecutive fill_bytes(64) calls (the s_L / s_R draws of Prover::prove).  Same bytes as the loop;
    // when several proofs are in flight in this process their streams are run eight at a time (keccak_x8_native.cpp).
    void fill_many64(uint8_t* out, size_t count);
};
// Marker of a prover that exists and has not produced its proof yet (from Prover::new to the end of Prove::prove):
// the batcher sizes its batches and its patience by how many of them this process has in flight.
int proving_now();  // Prover::prove calls executing in this process right now
struct ProvingScope {
    bool active = false;
    void enter();
    void leave();
    ~ProvingScope() { leave(); }
};
// statistics of the stream batcher: [0] streams served, [1] vector batches, [2] streams that ran alone (scalar)
void rng_batcher_stats(uint64_t out[3]);

struct Transcript {
    Strobe128 strobe;
    Transcript(const uint8_t* label, size_t n);
    void append_message(const char* label, const uint8_t* msg, size_t n);
    void append_message(const uint8_t* label, size_t ln, const uint8_t* msg, size_t n);
    void append_u64(const char* label, uint64_t x);
    void challenge_bytes(const char* label, uint8_t* out, size_t n);
    void challenge_bytes(const uint8_t* label, size_t ln, uint8_t* out, size_t n);
    // TranscriptRngBuilder: clone, rekey with witness bytes, finalize with 32 external bytes
    TranscriptRng build_rng(const std::vector<const uint8_t*>& witness32, const uint8_t external32[32]) const;
};

}  // namespace bpg

struct bpg_transcript {
    bpg::Transcript t;
    bpg_transcript(const uint8_t* label, size_t n) : t(label, n) {}
};
