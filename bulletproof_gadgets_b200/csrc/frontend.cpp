// Statement-level front end: prove()/verify() over the .gadgets/.inst/.wtns/.coms text formats.
// (placeholder until the C++ parsers + gadgets land)
#include "ctx.hpp"
extern "C" {
int bpg_prove(bpg_ctx*, const char*, const char*, const char*, const char*, const uint8_t*, const uint8_t*,
              bpg_proof_artifacts** out) {
    if (out) *out = nullptr;
    bpg_set_error("bpg_prove: statement front end not built yet");
    return BPG_E_GADGET;
}
int bpg_verify(bpg_ctx*, const char*, const char*, const char*, const char*, const uint8_t*, size_t, const uint8_t*,
               int* accepted) {
    if (accepted) *accepted = 0;
    bpg_set_error("bpg_verify: statement front end not built yet");
    return BPG_E_GADGET;
}
void bpg_free_proof(bpg_proof_artifacts* a) {
    if (!a) return;
    free(a->commitments);
    free(a->proof);
    delete a;
}
}
