/* bpg.h -- C ABI of the B200-native Bulletproofs R1CS hot path (libbpg.so).
 *
 * Drop-in boundary for the path that `bulletproofs_gadgets` drives through dalek's
 * `bulletproofs` crate.  Every entry point names the reference interface it replaces
 * (paths under /root/reference; "dalek:" = the un-vendored crates pinned in Cargo.lock:78-80,
 * 155-157, 403-405).  Plain C: pointers and sizes only, no exceptions; status codes mirror
 * dalek's `R1CSError` variants.  All scalars and compressed points are 32-byte little-endian
 * buffers owned by the caller.  One context per GPU; a context is not re-entrant.
 *
 * There is no CPU fallback: every call that needs arithmetic fails with BPG_E_CUDA when no
 * sm_100a device is usable.
 */
#ifndef BPG_H
#define BPG_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BPG_OK 0
#define BPG_E_FORMAT (-1)             /* dalek: R1CSError::FormatError                */
#define BPG_E_VERIFY (-2)             /* dalek: R1CSError::VerificationError          */
#define BPG_E_GENS_LEN (-3)           /* dalek: R1CSError::InvalidGeneratorsLength    */
#define BPG_E_MISSING_ASSIGNMENT (-4) /* dalek: R1CSError::MissingAssignment          */
#define BPG_E_CUDA (-5)               /* device / driver failure (no CPU fallback)    */
#define BPG_E_ARG (-6)                /* bad argument                                  */
#define BPG_E_GADGET (-7)             /* dalek: R1CSError::GadgetError / front-end panic */

typedef struct bpg_ctx bpg_ctx;
typedef struct bpg_transcript bpg_transcript;
typedef struct bpg_prover bpg_prover;
typedef struct bpg_verifier bpg_verifier;

/* thread-local message of the last failing call */
const char* bpg_last_error(void);

/* ---- context and generators ------------------------------------------------------------ */
/* One context per GPU.  Replaces nothing in the reference (it has no device state). */
int bpg_ctx_create(int device, bpg_ctx** out);
/* A second (third, ...) context on the same GPU as `parent`: its own stream and work buffers, the
 * generator tables shared.  One host thread per context lets several proofs be in flight on one GPU:
 * the sequential Merlin rng stream of one proof overlaps the MSMs of another. */
int bpg_ctx_create_shared(bpg_ctx* parent, bpg_ctx** out);
void bpg_ctx_destroy(bpg_ctx* ctx);
/* Page-locked host memory for buffers handed to the bulk loaders (uploads then run as async DMA). */
void* bpg_host_alloc(size_t bytes);
void bpg_host_free(void* p);
/* tuning knobs: "task_len" (entries per accumulate task), "window_bits" (0 = auto) */
int bpg_ctx_set(bpg_ctx* ctx, const char* key, int64_t value);
/* counters: "launches" (kernels launched so far), "accum_us" / "accum_entries" (last timed MSM) */
int64_t bpg_ctx_get(bpg_ctx* ctx, const char* key);

/* Device self-test of the field layer (no reference counterpart): n pseudo-random 256-bit values with limbs biased towards
 * 0 and 2^32-1; checks the dedicated squaring against the general product and (a+1)^2 - a^2 - 2a - 1 == 0. */
int bpg_selftest_field(bpg_ctx* ctx, uint64_t n, uint64_t seed, uint64_t* mismatches);

/* BulletproofGens::new(capacity, 1) + PedersenGens::default()
 *   -- /root/reference/src/prove.rs:46,78 ; /root/reference/src/verify.rs:45,70.
 * Derives G_i, H_i (SHAKE256 "GeneratorsChain" stream on the host, Elligator on the GPU) and
 * builds the device-resident fixed-base window tables.  Idempotent; grows when needed. */
int bpg_gens_ensure(bpg_ctx* ctx, uint64_t capacity);
/* Compressed generators for inspection: which = 0:G 1:H 2:B 3:B_blinding. */
int bpg_gens_compressed(bpg_ctx* ctx, int which, uint64_t start, uint64_t count, uint8_t* out32n);

/* ---- multiscalar multiplication ------------------------------------------------------- */
/* sum sG[i]*G_i + sum sH[i]*H_i + sB*B + sBb*B_blinding, compressed.
 * Replaces dalek RistrettoPoint::multiscalar_mul / vartime_multiscalar_mul over generator
 * slices (bulletproofs r1cs/prover.rs A_I1/A_O1/S1, inner_product_proof.rs L/R) -- reached from
 * /root/reference/src/prove.rs:79.  Host pointers; any of sG/sH/sB/sBb may be NULL. */
int bpg_msm_gens(bpg_ctx* ctx, const uint8_t* sG, uint64_t nG, const uint8_t* sH, uint64_t nH,
                 const uint8_t* sB, const uint8_t* sBb, uint8_t out32[32]);
/* Point-range form: sG multiplies G_{g_start}.., sH multiplies H_{h_start}..  One large MSM (dalek
 * vartime_multiscalar_mul over generator slices, e.g. the verifier's mega-MSM reached from
 * /root/reference/src/verify.rs:71) is split across GPUs by contiguous point ranges; every GPU returns one
 * compressed partial point and the host adds them with bpg_point_sum. */
int bpg_msm_gens_range(bpg_ctx* ctx, const uint8_t* sG, uint64_t g_start, uint64_t nG, const uint8_t* sH, uint64_t h_start,
                       uint64_t nH, const uint8_t* sB, const uint8_t* sBb, uint8_t out32[32]);
/* Host-only sum of n compressed ristretto points (<= 8 partial results of a sharded MSM).  BPG_E_VERIFY if a
 * point does not decode.  Needs no GPU. */
int bpg_point_sum(const uint8_t* points32n, uint64_t n, uint8_t out32[32]);
/* Same with DEVICE pointers (scalars already resident in HBM, canonical). */
int bpg_msm_gens_dev(bpg_ctx* ctx, const void* d_sG, uint64_t nG, const void* d_sH, uint64_t nH,
                     const void* d_sB, const void* d_sBb, uint8_t out32[32]);
int bpg_msm_gens_range_dev(bpg_ctx* ctx, const void* d_sG, uint64_t g_start, uint64_t nG, const void* d_sH, uint64_t h_start,
                           uint64_t nH, const void* d_sB, const void* d_sBb, uint8_t out32[32]);
/* sum s[i]*P_i over arbitrary compressed points (dalek vartime_multiscalar_mul /
 * optional_multiscalar_mul).  Returns BPG_E_VERIFY if a point does not decode. */
int bpg_msm(bpg_ctx* ctx, const uint8_t* scalars32n, const uint8_t* points32n, uint64_t n, uint8_t out32[32]);

/* PedersenGens::commit(v, r) = v*B + r*B_blinding for k pairs
 *   -- /root/reference/src/gadget.rs:32 ; /root/reference/src/commitments.rs:28,40. */
int bpg_pedersen_commit_batch(bpg_ctx* ctx, const uint8_t* v32k, const uint8_t* r32k, uint64_t k, uint8_t* out32k);

/* ---- Merlin transcript (host) ---------------------------------------------------------- */
/* merlin::Transcript::new -- /root/reference/src/prove.rs:45 ; /root/reference/src/verify.rs:44 */
bpg_transcript* bpg_transcript_new(const uint8_t* label, size_t len);
bpg_transcript* bpg_transcript_clone(const bpg_transcript* t);
void bpg_transcript_free(bpg_transcript* t);
void bpg_transcript_append_message(bpg_transcript* t, const uint8_t* label, size_t label_len, const uint8_t* msg, size_t msg_len);
void bpg_transcript_challenge_bytes(bpg_transcript* t, const uint8_t* label, size_t label_len, uint8_t* out, size_t out_len);

/* merlin TranscriptRng as Prover::prove uses it (bulletproofs r1cs/prover.rs): clone the transcript, rekey with k
 * 32-byte witnesses under the label "v_blinding", finalize with 32 external bytes, discard `warm` 64-byte draws, then
 * write n 64-byte draws.  Host only; concurrent callers' streams are run eight at a time (AVX-512). */
int bpg_transcript_rng_fill64(const bpg_transcript* t, const uint8_t* witness32k, size_t k, const uint8_t seed32[32],
                              size_t warm, uint8_t* out64n, size_t n);
/* 0: streams served, 1: vector batches, 2: streams that ran alone */
int64_t bpg_rng_batcher_stat(int which);

/* ---- R1CS constraint system: prover --------------------------------------------------- */
/* Variables are uint32 tags: kind<<29 | index, kind = 0 Committed, 1 MultiplierLeft,
 * 2 MultiplierRight, 3 MultiplierOutput, 4 One  (dalek r1cs::Variable). */
#define BPG_VAR_COMMITTED(i) ((uint32_t)(0u << 29) | (uint32_t)(i))
#define BPG_VAR_LEFT(i) ((uint32_t)(1u << 29) | (uint32_t)(i))
#define BPG_VAR_RIGHT(i) ((uint32_t)(2u << 29) | (uint32_t)(i))
#define BPG_VAR_OUT(i) ((uint32_t)(3u << 29) | (uint32_t)(i))
#define BPG_VAR_ONE ((uint32_t)(4u << 29))

/* Prover::new(&pc_gens, &mut transcript) -- /root/reference/src/prove.rs:47.  The prover borrows
 * the transcript until bpg_prover_free / bpg_prover_prove. */
int bpg_prover_new(bpg_ctx* ctx, bpg_transcript* t, bpg_prover** out);
void bpg_prover_free(bpg_prover* p);
/* Prover::commit(v, v_blinding) -> (CompressedRistretto, Variable) -- /root/reference/src/gadget.rs:32 */
int bpg_prover_commit(bpg_prover* p, const uint8_t v[32], const uint8_t v_blinding[32], uint8_t V_out[32], uint32_t* var_out);
/* k commitments in one device launch; transcript appends stay in order */
int bpg_prover_commit_batch(bpg_prover* p, const uint8_t* v32k, const uint8_t* vb32k, uint64_t k, uint8_t* V_out32k, uint32_t* first_var_out);
/* ConstraintSystem::allocate_multiplier(Some((l, r))) -- /root/reference/src/cs_buffer.rs:100-104 */
int bpg_prover_allocate_multiplier(bpg_prover* p, const uint8_t l[32], const uint8_t r[32], uint32_t vars_out[3]);
/* ConstraintSystem::multiply(left, right): LCs as parallel arrays (var tag, 32-byte coeff)
 *   -- /root/reference/src/cs_buffer.rs:94-97 */
int bpg_prover_multiply(bpg_prover* p, const uint32_t* lvars, const uint8_t* lcoef32, size_t ln,
                        const uint32_t* rvars, const uint8_t* rcoef32, size_t rn, uint32_t vars_out[3]);
/* ConstraintSystem::constrain(lc) -- /root/reference/src/cs_buffer.rs:112-115 */
int bpg_prover_constrain(bpg_prover* p, const uint32_t* vars, const uint8_t* coef32, size_t n);
/* Bulk form of the three calls above: appends n multipliers (allocate_multiplier semantics: no implicit
 * constraints; a_O is recomputed as a_L*a_R) and q constraints given as a CSR term list
 * (row_start[q+1], term_var[nnz], term_coef32[nnz]).  Variables inside the terms are absolute tags.
 * Replaces the op-by-op replay of /root/reference/src/prove.rs:84-99 (assign_buffer). */
int bpg_prover_load_cs(bpg_prover* p, const uint8_t* aL32n, const uint8_t* aR32n, uint64_t n,
                       const uint32_t* row_start, const uint32_t* term_var, const uint8_t* term_coef32, uint64_t q);
/* The same with WITNESS GENERATION ON THE DEVICE for the reference's range proof (/root/reference/src/utils.rs:13-31:
 * for i < n_bits, allocate_multiplier((1 - bit_i(x), bit_i(x)))): the caller passes the values, not the bits.  Multipliers
 * [first, first + nbits) of run r get a_L = 1 - bit_i(value), a_R = bit_i(value), a_O = 0 in HBM; the h multipliers no run
 * covers come as compact host arrays aL32h / aR32h with their multiplier indices.  Runs and indices must partition
 * [0, n).  A BOUND x1024 statement (n = 2^17) uploads 2048 values instead of 8 MB of bit scalars. */
typedef struct bpg_bit_run {
    uint64_t first;
    uint32_t nbits; /* <= 256 */
    uint32_t reserved;
    uint8_t value[32]; /* little-endian, the range proof's x_assignment */
} bpg_bit_run;
int bpg_prover_load_cs_bits(bpg_prover* p, uint64_t n, const bpg_bit_run* runs, uint64_t n_runs, const uint8_t* aL32h,
                            const uint8_t* aR32h, const uint32_t* host_index, uint64_t h, const uint32_t* row_start,
                            const uint32_t* term_var, const uint8_t* term_coef32, uint64_t q);
/* Prover::num_constraints / get_num_multiplications (FairAds fork) -- /root/reference/src/prove.rs:75,78 */
uint64_t bpg_prover_num_constraints(const bpg_prover* p);
uint64_t bpg_prover_num_multipliers(const bpg_prover* p);
/* Prover::prove(&bp_gens) -> R1CSProof::to_bytes() -- /root/reference/src/prove.rs:79-81.
 * rng_seed32 replaces the 32 bytes merlin's TranscriptRngBuilder::finalize draws from
 * thread_rng(); NULL = draw from the OS.  Generators for round_pow2(n) are ensured. */
int bpg_prover_prove(bpg_prover* p, const uint8_t* rng_seed32, uint8_t* proof_out, size_t proof_cap, size_t* proof_len);

/* ---- device-resident circuits ------------------------------------------------------------ */
/* A flattened constraint system (and optionally its witness) kept in HBM so that many proofs /
 * verifications of the same circuit pay the host->device upload once.  What the circuit holds is
 * exactly what assign_buffer replays (/root/reference/src/prove.rs:84-99, src/verify.rs:75-90):
 * n multipliers, q constraints over m committed variables.  a_O is computed on the device. */
typedef struct bpg_circuit bpg_circuit;
int bpg_circuit_create(bpg_ctx* ctx, uint64_t n, uint64_t m, const uint32_t* row_start, const uint32_t* term_var,
                       const uint8_t* term_coef32, uint64_t q, bpg_circuit** out);
int bpg_circuit_set_witness(bpg_circuit* c, const uint8_t* aL32n, const uint8_t* aR32n);
void bpg_circuit_free(bpg_circuit* c);
/* Use the resident circuit instead of multiply/allocate_multiplier/constrain calls (exclusive). */
int bpg_prover_attach(bpg_prover* p, const bpg_circuit* c);
int bpg_verifier_attach(bpg_verifier* v, const bpg_circuit* c);

/* ---- R1CS constraint system: verifier ------------------------------------------------- */
/* Verifier::new(&mut transcript) -- /root/reference/src/verify.rs:46 */
int bpg_verifier_new(bpg_ctx* ctx, bpg_transcript* t, bpg_verifier** out);
void bpg_verifier_free(bpg_verifier* v);
/* Verifier::commit(V) -> Variable -- /root/reference/src/lalrpop/assignment_parser.rs:138 */
int bpg_verifier_commit(bpg_verifier* v, const uint8_t V[32], uint32_t* var_out);
/* k commitments in file order (the .coms replay of /root/reference/src/lalrpop/assignment_parser.rs:133-141) */
int bpg_verifier_commit_batch(bpg_verifier* v, const uint8_t* V32k, uint64_t k, uint32_t* first_var_out);
/* allocate_multiplier(None) / multiply / constrain -- /root/reference/src/cs_buffer.rs:173-199 */
int bpg_verifier_allocate_multiplier(bpg_verifier* v, uint32_t vars_out[3]);
int bpg_verifier_multiply(bpg_verifier* v, const uint32_t* lvars, const uint8_t* lcoef32, size_t ln,
                          const uint32_t* rvars, const uint8_t* rcoef32, size_t rn, uint32_t vars_out[3]);
int bpg_verifier_constrain(bpg_verifier* v, const uint32_t* vars, const uint8_t* coef32, size_t n);
/* Bulk form: n multipliers without assignments and q constraints (CSR) -- the replay of
 * /root/reference/src/verify.rs:75-90 (assign_buffer). */
int bpg_verifier_load_cs(bpg_verifier* v, uint64_t n, const uint32_t* row_start, const uint32_t* term_var,
                         const uint8_t* term_coef32, uint64_t q);
uint64_t bpg_verifier_num_vars(const bpg_verifier* v); /* Verifier::get_num_vars -- /root/reference/src/verify.rs:70 */
/* R1CSProof::from_bytes + Verifier::verify(&proof, &pc_gens, &bp_gens) -- /root/reference/src/verify.rs:53,71.
 * BPG_OK = accepted, BPG_E_VERIFY = rejected, BPG_E_FORMAT = malformed proof bytes. */
int bpg_verifier_verify(bpg_verifier* v, const uint8_t* proof, size_t proof_len, const uint8_t* rng_seed32);

/* ---- statement-level front end ---------------------------------------------------------- */
/* prove()/verify() of /root/reference/src/prove.rs:37-43 and /root/reference/src/verify.rs:36-42 with an explicit
 * context and injectable randomness.  The reference's own C ABI -- c_prove / c_verify / free_proof and
 * struct ProofArtifacts, /root/reference/interfaces/ios/src/lib.rs:11-55 -- is exported under exactly those names too:
 * see include/bulletproofs_gadgets.h.
 * blinding_seed32 seeds the commitment blindings the reference draws from thread_rng()
 * (/root/reference/src/commitments.rs:28,40, /root/reference/src/gadget.rs:32); NULL = OS entropy. */
typedef struct bpg_proof_artifacts {
    char* commitments;   /* .coms text, NUL terminated */
    uint8_t* proof;      /* .proof bytes */
    size_t proof_len;
    uint64_t num_constraints; /* what the reference prover prints (src/prove.rs:75) */
} bpg_proof_artifacts;
int bpg_prove(bpg_ctx* ctx, const char* name, const char* instance, const char* witness, const char* gadgets,
              const uint8_t* blinding_seed32, const uint8_t* rng_seed32, bpg_proof_artifacts** out);
int bpg_verify(bpg_ctx* ctx, const char* name, const char* instance, const char* gadgets, const char* commitments,
               const uint8_t* proof, size_t proof_len, const uint8_t* rng_seed32, int* accepted);
void bpg_free_proof(bpg_proof_artifacts* a);

/* ---- batches: N independent statements over a set of contexts ---------------------------- */
/* What the reference does one process per statement (`prover <stem>` / `verifier <stem>`, /root/reference/src/bin/{prover,verifier}.rs),
 * as a throughput call: the library runs one host thread per context (contexts of one GPU made with
 * bpg_ctx_create_shared, or of several GPUs), each keeping one statement in flight on its stream, so the sequential Merlin
 * rng stream of one proof overlaps the MSMs of the others.  Every job gets its own status (BPG_OK, BPG_E_VERIFY, ...).
 * Return value: < 0 = the call itself was malformed, otherwise the number of jobs whose status is not BPG_OK. */
#define BPG_JOB_VERIFY 1u /* prove jobs: run Verifier::verify on the fresh proof on the same context (status covers both) */
typedef struct bpg_prove_job {
    /* in: Transcript::new(label) -- /root/reference/src/prove.rs:45 */
    const uint8_t* label;
    size_t label_len;
    /* in: m committed values and blindings (Prover::commit) */
    const uint8_t* v32m;
    const uint8_t* vbl32m;
    uint64_t m;
    /* in: the constraint system, either a resident circuit (with witness) or host arrays as for bpg_prover_load_cs */
    const bpg_circuit* circuit;
    const uint8_t* aL32n;
    const uint8_t* aR32n;
    uint64_t n;
    const uint32_t* row_start;
    const uint32_t* term_var;
    const uint8_t* term_coef32;
    uint64_t q;
    const uint8_t* rng_seed32;    /* NULL = OS entropy */
    const uint8_t* verify_seed32; /* used with BPG_JOB_VERIFY; NULL = OS entropy */
    uint32_t flags;
    /* out */
    uint8_t* V_out32m; /* m compressed commitments */
    uint8_t* proof_out;
    size_t proof_cap;
    size_t proof_len;
    int status;
} bpg_prove_job;
typedef struct bpg_verify_job {
    const uint8_t* label;
    size_t label_len;
    const uint8_t* V32m;
    uint64_t m;
    const bpg_circuit* circuit; /* or the host arrays below, as for bpg_verifier_load_cs */
    uint64_t n;
    const uint32_t* row_start;
    const uint32_t* term_var;
    const uint8_t* term_coef32;
    uint64_t q;
    const uint8_t* proof;
    size_t proof_len;
    const uint8_t* rng_seed32;
    int status; /* BPG_OK = accepted, BPG_E_VERIFY = rejected, BPG_E_FORMAT = malformed proof */
} bpg_verify_job;
int bpg_r1cs_prove_batch(bpg_ctx* const* ctxs, size_t n_ctx, bpg_prove_job* jobs, size_t n_jobs);
int bpg_r1cs_verify_batch(bpg_ctx* const* ctxs, size_t n_ctx, bpg_verify_job* jobs, size_t n_jobs);
/* The same for statements in the reference's text formats (bpg_prove / bpg_verify per job, front end included). */
typedef struct bpg_text_job {
    const char* name;
    const char* instance;
    const char* witness;     /* prove jobs */
    const char* gadgets;
    const char* commitments; /* verify jobs */
    const uint8_t* proof;    /* verify jobs */
    size_t proof_len;
    const uint8_t* blinding_seed32;
    const uint8_t* rng_seed32;
    const uint8_t* verify_seed32;
    uint32_t flags;          /* BPG_JOB_VERIFY on prove jobs */
    /* out */
    bpg_proof_artifacts* artifacts; /* prove jobs; release with bpg_free_proof */
    int accepted;                   /* verify jobs, and prove jobs with BPG_JOB_VERIFY */
    int status;
} bpg_text_job;
int bpg_prove_batch(bpg_ctx* const* ctxs, size_t n_ctx, bpg_text_job* jobs, size_t n_jobs);
int bpg_verify_batch(bpg_ctx* const* ctxs, size_t n_ctx, bpg_text_job* jobs, size_t n_jobs);

/* The flat statement the real Prover / Verifier hold after assign_buffer (/root/reference/src/prove.rs:84-99,
 * /root/reference/src/verify.rs:75-90), as produced by the front end from the text formats.  Pure host code (no
 * GPU needed): lets a caller inspect / cache the circuit, or feed the bulk loaders itself.
 * Prover side: v32m, vbl32m, aL32n, aR32n set, V32m NULL.  Verifier side: V32m set, the others NULL. */
typedef struct bpg_flat_statement {
    uint64_t n, m, q, nnz;
    uint8_t* v32m;
    uint8_t* vbl32m;
    uint8_t* V32m;
    uint8_t* aL32n;
    uint8_t* aR32n;
    uint32_t* row_start;   /* q + 1 */
    uint32_t* term_var;    /* nnz variable tags */
    uint8_t* term_coef32;  /* nnz * 32 */
    char* com_names;       /* '\n'-terminated .coms names (C<w>-<limb>, D<line>-<sub>-<k>) in commit order */
} bpg_flat_statement;
int bpg_frontend_flatten_prover(const char* name, const char* instance, const char* witness, const char* gadgets,
                                const uint8_t* blinding_seed32, bpg_flat_statement** out);
int bpg_frontend_flatten_verifier(const char* name, const char* instance, const char* commitments, const char* gadgets,
                                  bpg_flat_statement** out);
void bpg_flat_statement_free(bpg_flat_statement* f);
/* mimc_hash(bytes) (big-endian 32-byte image as 32 LE scalar bytes) and the un-padded MiMC sponge over scalars
 * -- /root/reference/src/mimc_hash/mimc.rs:24-40,61-75.  Host only. */
int bpg_mimc_hash(const uint8_t* preimage, size_t len, uint8_t out32[32]);
int bpg_mimc_sponge(const uint8_t* scalars32n, size_t n, uint8_t out32[32]);

#ifdef __cplusplus
}
#endif
#endif /* BPG_H */
