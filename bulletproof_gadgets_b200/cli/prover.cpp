// `prover <stem>`: process-level drop-in for /root/reference/src/bin/prover.rs:16-30.
// Reads <stem>.inst / .wtns / .gadgets, writes <stem>.coms / .proof, prints the number of constraints
// (/root/reference/src/prove.rs:75).  The transcript label is the stem itself (prove.rs:45).
// Optional determinism hooks (the reference draws from thread_rng): BPG_BLINDING_SEED, BPG_RNG_SEED = 64 hex digits;
// BPG_DEVICE = CUDA device index.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <fstream>
#include <sstream>
#include <string>

#include "bpg.h"

static bool slurp(const std::string& path, std::string* out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::stringstream ss;
    ss << f.rdbuf();
    *out = ss.str();
    return true;
}
static const uint8_t* env_seed(const char* name, uint8_t buf[32]) {
    const char* s = getenv(name);
    if (!s || strlen(s) != 64) return nullptr;
    for (int i = 0; i < 32; i++) {
        unsigned x;
        if (sscanf(s + 2 * i, "%2x", &x) != 1) return nullptr;
        buf[i] = (uint8_t)x;
    }
    return buf;
}

int main(int argc, char** argv) {
    if (argc < 2) {
        fprintf(stderr, "missing argument\n");
        return 101;
    }
    const std::string stem = argv[1];
    std::string inst, wtns, gadgets;
    if (!slurp(stem + ".inst", &inst) || !slurp(stem + ".wtns", &wtns) || !slurp(stem + ".gadgets", &gadgets)) {
        fprintf(stderr, "unable to read instance file\n");
        return 101;
    }
    bpg_ctx* ctx = nullptr;
    const char* dev = getenv("BPG_DEVICE");
    if (bpg_ctx_create(dev ? atoi(dev) : 0, &ctx) != BPG_OK) {
        fprintf(stderr, "%s\n", bpg_last_error());
        return 101;
    }
    uint8_t b1[32], b2[32];
    bpg_proof_artifacts* art = nullptr;
    const int rc = bpg_prove(ctx, stem.c_str(), inst.c_str(), wtns.c_str(), gadgets.c_str(), env_seed("BPG_BLINDING_SEED", b1),
                             env_seed("BPG_RNG_SEED", b2), &art);
    if (rc != BPG_OK) {
        fprintf(stderr, "unable to generate proof from provided files: %s\n", bpg_last_error());
        bpg_ctx_destroy(ctx);
        return 101;  // the reference panics (exit code 101)
    }
    printf("%llu\n", (unsigned long long)art->num_constraints);
    std::ofstream(stem + ".coms", std::ios::binary) << art->commitments;
    std::ofstream(stem + ".proof", std::ios::binary).write(reinterpret_cast<const char*>(art->proof), (std::streamsize)art->proof_len);
    bpg_free_proof(art);
    bpg_ctx_destroy(ctx);
    return 0;
}
