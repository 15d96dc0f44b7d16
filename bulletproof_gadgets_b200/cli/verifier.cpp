// `verifier <stem>`: process-level drop-in for /root/reference/src/bin/verifier.rs:14-25.
// Reads <stem>.inst / .coms / .proof / .gadgets and prints `true` or `false`.
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

int main(int argc, char** argv) {
    if (argc < 2) {
        fprintf(stderr, "missing argument\n");
        return 101;
    }
    const std::string stem = argv[1];
    std::string inst, coms, proof, gadgets;
    if (!slurp(stem + ".inst", &inst) || !slurp(stem + ".coms", &coms) || !slurp(stem + ".proof", &proof) ||
        !slurp(stem + ".gadgets", &gadgets)) {
        fprintf(stderr, "unable to read input files\n");
        return 101;
    }
    bpg_ctx* ctx = nullptr;
    const char* dev = getenv("BPG_DEVICE");
    if (bpg_ctx_create(dev ? atoi(dev) : 0, &ctx) != BPG_OK) {
        fprintf(stderr, "%s\n", bpg_last_error());
        return 101;
    }
    int accepted = 0;
    const int rc = bpg_verify(ctx, stem.c_str(), inst.c_str(), gadgets.c_str(), coms.c_str(),
                              reinterpret_cast<const uint8_t*>(proof.data()), proof.size(), nullptr, &accepted);
    bpg_ctx_destroy(ctx);
    if (rc != BPG_OK) {  // malformed proof / inputs: the reference panics here (R1CSProof::from_bytes(..).unwrap(), verify.rs:53)
        fprintf(stderr, "unable to verify provided files: %s\n", bpg_last_error());
        return 101;
    }
    printf("%s\n", accepted ? "true" : "false");
    return 0;
}
