/* bulletproofs_gadgets.h -- the reference's own C ABI, served by libbpg.so.
 *
 * Same symbols, argument order and struct layout as the library the reference builds for its mobile bindings:
 *   /root/reference/interfaces/ios/src/lib.rs:11-19   #[repr(C)] struct ProofArtifacts
 *   /root/reference/interfaces/ios/src/lib.rs:21      c_prove
 *   /root/reference/interfaces/ios/src/lib.rs:45      c_verify
 *   /root/reference/interfaces/ios/src/lib.rs:55      free_proof
 *   /root/reference/interfaces/ios/src/bulletproofs_ios.h:4-13   the C declarations shipped with it
 * so a caller of that library links against libbpg.so unchanged.  proof_len / proof_cap are `usize` in the Rust
 * definition (what the compiled library really exports); the shipped header spells them `int`, which only agrees with
 * it in the low half on a 64-bit target -- this header follows the Rust layout.
 *
 * Differences in behaviour (deliberate): the reference panics (aborts) on malformed input or a failing prover; here
 * c_prove returns NULL and c_verify returns false, with the reason in bpg_last_error() (bpg.h).  Randomness
 * (commitment blindings, transcript rng seed) comes from the OS, as with the reference's thread_rng().  The calls run on
 * a process-global pool of GPU contexts (device BPG_DEVICE, default 0; up to BPG_GLOBAL_CONTEXTS = 8 concurrent calls);
 * there is no CPU fallback: without an sm_100a device c_prove returns NULL.
 */
#ifndef BULLETPROOFS_GADGETS_H
#define BULLETPROOFS_GADGETS_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

struct ProofArtifacts {
    const char* commitments; /* .coms text, NUL terminated */
    const uint8_t* proof;    /* R1CSProof::to_bytes() */
    size_t proof_len;
    size_t proof_cap;
};

/* prove() of /root/reference/src/prove.rs:37-43 over the text of the .inst / .wtns / .gadgets files */
struct ProofArtifacts* c_prove(const char* name, const char* instance, const char* witness, const char* gadgets);
/* verify() of /root/reference/src/verify.rs:36-42 */
bool c_verify(const char* name, const char* instance, const char* gadgets, const char* commitments, const uint8_t* proof,
              size_t proof_len);
void free_proof(struct ProofArtifacts* artifacts_pointer);

#ifdef __cplusplus
}
#endif
#endif /* BULLETPROOFS_GADGETS_H */
