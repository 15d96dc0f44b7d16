// Throughput entry points and the reference's own C ABI.
//
//  * c_prove / c_verify / free_proof + struct ProofArtifacts: exactly the symbols and the layout that
//    /root/reference/interfaces/ios/src/lib.rs:11-19,21,45,55 exports (the Android JNI shim,
//    /root/reference/interfaces/android/src/lib.rs, calls the same prove()/verify()), so an existing caller links
//    against libbpg.so unchanged.  They run on a process-global pool of contexts on one GPU (BPG_DEVICE, default 0).
//  * bpg_r1cs_prove_batch / bpg_r1cs_verify_batch: N independent flat statements over a caller-supplied set of contexts.
//    The library owns the host threads (one per context, each keeps one statement in flight on its stream), so the
//    sequential Merlin rng stream of one proof overlaps the MSMs of the others and the rng streams of the proofs in
//    flight are hashed eight at a time (merlin.cpp: RngBatcher) -- the orchestration that produces the headline number.
//  * bpg_prove_batch / bpg_verify_batch: the same for statements in the reference's text formats (front end included).
//
// Host-only translation unit (no kernels): everything below goes through the C ABI of r1cs.cu / frontend.cpp.
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/bpg.h"
#include "../../include/bulletproofs_gadgets.h"

void bpg_set_error(const char* fmt, ...);

namespace {

// ------------------------------------------------------------------------------------------
// worker threads: thread i owns ctxs[i] and pulls job indices until none are left
// ------------------------------------------------------------------------------------------
template <typename Fn>
int run_jobs(bpg_ctx* const* ctxs, size_t n_ctx, size_t n_jobs, Fn&& fn) {
    if (n_jobs == 0) return BPG_OK;
    if (!ctxs || n_ctx == 0) return BPG_E_ARG;
    for (size_t i = 0; i < n_ctx; i++)
        if (!ctxs[i]) return BPG_E_ARG;
    std::atomic<size_t> next{0};
    std::atomic<int> failed{0};
    auto work = [&](size_t w) {
        for (;;) {
            const size_t j = next.fetch_add(1, std::memory_order_relaxed);
            if (j >= n_jobs) return;
            if (fn(ctxs[w], j) != BPG_OK) failed.fetch_add(1, std::memory_order_relaxed);
        }
    };
    const size_t n_thr = n_ctx < n_jobs ? n_ctx : n_jobs;
    if (n_thr == 1) {
        work(0);
    } else {
        std::vector<std::thread> thr;
        thr.reserve(n_thr);
        for (size_t w = 0; w < n_thr; w++) thr.emplace_back(work, w);
        for (auto& t : thr) t.join();
    }
    return failed.load();  // number of jobs whose status is not BPG_OK
}

int verify_one(bpg_ctx* ctx, bpg_verify_job* j) {
    if (!j->label || (j->m && !j->V32m) || !j->proof) return j->status = BPG_E_ARG;
    bpg_transcript* t = bpg_transcript_new(j->label, j->label_len);
    bpg_verifier* v = nullptr;
    int rc = bpg_verifier_new(ctx, t, &v);
    if (!rc) rc = bpg_verifier_commit_batch(v, j->V32m, j->m, nullptr);
    if (!rc)
        rc = j->circuit ? bpg_verifier_attach(v, j->circuit)
                        : bpg_verifier_load_cs(v, j->n, j->row_start, j->term_var, j->term_coef32, j->q);
    if (!rc) rc = bpg_verifier_verify(v, j->proof, j->proof_len, j->rng_seed32);
    if (v) bpg_verifier_free(v);
    bpg_transcript_free(t);
    return j->status = rc;
}

int prove_one(bpg_ctx* ctx, bpg_prove_job* j) {
    j->proof_len = 0;
    if (!j->label || (j->m && (!j->v32m || !j->vbl32m || !j->V_out32m)) || !j->proof_out) return j->status = BPG_E_ARG;
    bpg_transcript* t = bpg_transcript_new(j->label, j->label_len);
    bpg_prover* p = nullptr;
    int rc = bpg_prover_new(ctx, t, &p);
    if (!rc) rc = bpg_prover_commit_batch(p, j->v32m, j->vbl32m, j->m, j->V_out32m, nullptr);
    if (!rc)
        rc = j->circuit ? bpg_prover_attach(p, j->circuit)
                        : bpg_prover_load_cs(p, j->aL32n, j->aR32n, j->n, j->row_start, j->term_var, j->term_coef32, j->q);
    if (!rc) rc = bpg_prover_prove(p, j->rng_seed32, j->proof_out, j->proof_cap, &j->proof_len);
    if (p) bpg_prover_free(p);
    bpg_transcript_free(t);
    if (!rc && (j->flags & BPG_JOB_VERIFY)) {  // Verifier::verify on the same context, right behind the proof
        bpg_verify_job v;
        memset(&v, 0, sizeof v);
        v.label = j->label, v.label_len = j->label_len;
        v.V32m = j->V_out32m, v.m = j->m, v.n = j->n;
        v.row_start = j->row_start, v.term_var = j->term_var, v.term_coef32 = j->term_coef32, v.q = j->q;
        v.circuit = j->circuit;
        v.proof = j->proof_out, v.proof_len = j->proof_len;
        v.rng_seed32 = j->verify_seed32;
        rc = verify_one(ctx, &v);
    }
    return j->status = rc;
}

// ------------------------------------------------------------------------------------------
// process-global contexts behind c_prove / c_verify
// ------------------------------------------------------------------------------------------
struct GlobalPool {
    std::mutex mu;
    std::condition_variable cv;
    std::vector<bpg_ctx*> idle;
    bpg_ctx* root = nullptr;
    size_t created = 0, limit = 8;
    bool failed = false;
    bpg_ctx* acquire() {
        std::unique_lock<std::mutex> lock(mu);
        for (;;) {
            if (!idle.empty()) {
                bpg_ctx* c = idle.back();
                idle.pop_back();
                return c;
            }
            if (failed) return nullptr;
            if (created < limit) {
                bpg_ctx* c = nullptr;
                int rc;
                if (!root) {
                    if (const char* e = getenv("BPG_GLOBAL_CONTEXTS")) {
                        const long v = atol(e);
                        if (v >= 1 && v <= 256) limit = (size_t)v;
                    }
                    const char* d = getenv("BPG_DEVICE");
                    rc = bpg_ctx_create(d ? atoi(d) : 0, &c);
                    if (!rc) root = c;
                } else {
                    rc = bpg_ctx_create_shared(root, &c);
                }
                if (rc) {
                    if (!root) failed = true;  // no usable GPU: every later call fails the same way (no CPU fallback)
                    return nullptr;
                }
                created++;
                return c;
            }
            cv.wait(lock);
        }
    }
    void release(bpg_ctx* c) {
        {
            std::lock_guard<std::mutex> lock(mu);
            idle.push_back(c);
        }
        cv.notify_one();
    }
};
GlobalPool& pool() {
    static GlobalPool* p = new GlobalPool();  // never destroyed: contexts must outlive static destructors of callers
    return *p;
}
struct PoolLease {
    bpg_ctx* c;
    PoolLease() : c(pool().acquire()) {}
    ~PoolLease() {
        if (c) pool().release(c);
    }
};

}  // namespace

extern "C" {

int bpg_r1cs_prove_batch(bpg_ctx* const* ctxs, size_t n_ctx, bpg_prove_job* jobs, size_t n_jobs) {
    if (n_jobs && !jobs) return BPG_E_ARG;
    return run_jobs(ctxs, n_ctx, n_jobs, [&](bpg_ctx* c, size_t j) { return prove_one(c, &jobs[j]); });
}

int bpg_r1cs_verify_batch(bpg_ctx* const* ctxs, size_t n_ctx, bpg_verify_job* jobs, size_t n_jobs) {
    if (n_jobs && !jobs) return BPG_E_ARG;
    return run_jobs(ctxs, n_ctx, n_jobs, [&](bpg_ctx* c, size_t j) { return verify_one(c, &jobs[j]); });
}

int bpg_prove_batch(bpg_ctx* const* ctxs, size_t n_ctx, bpg_text_job* jobs, size_t n_jobs) {
    if (n_jobs && !jobs) return BPG_E_ARG;
    return run_jobs(ctxs, n_ctx, n_jobs, [&](bpg_ctx* c, size_t j) {
        bpg_text_job& t = jobs[j];
        t.artifacts = nullptr;
        t.accepted = 0;
        if (!t.name || !t.instance || !t.witness || !t.gadgets) return t.status = BPG_E_ARG;
        int rc = bpg_prove(c, t.name, t.instance, t.witness, t.gadgets, t.blinding_seed32, t.rng_seed32, &t.artifacts);
        if (!rc && (t.flags & BPG_JOB_VERIFY))
            rc = bpg_verify(c, t.name, t.instance, t.gadgets, t.artifacts->commitments, t.artifacts->proof,
                            t.artifacts->proof_len, t.verify_seed32, &t.accepted);
        return t.status = rc;
    });
}

int bpg_verify_batch(bpg_ctx* const* ctxs, size_t n_ctx, bpg_text_job* jobs, size_t n_jobs) {
    if (n_jobs && !jobs) return BPG_E_ARG;
    return run_jobs(ctxs, n_ctx, n_jobs, [&](bpg_ctx* c, size_t j) {
        bpg_text_job& t = jobs[j];
        t.accepted = 0;
        if (!t.name || !t.instance || !t.gadgets || !t.commitments || !t.proof) return t.status = BPG_E_ARG;
        return t.status = bpg_verify(c, t.name, t.instance, t.gadgets, t.commitments, t.proof, t.proof_len, t.verify_seed32,
                                     &t.accepted);
    });
}

// ------------------------------------------------------------------------------------------
// the reference's C ABI (interfaces/ios/src/lib.rs)
// ------------------------------------------------------------------------------------------
struct ProofArtifacts* c_prove(const char* name, const char* instance, const char* witness, const char* gadgets) {
    if (!name || !instance || !witness || !gadgets) {
        bpg_set_error("c_prove: NULL argument");
        return nullptr;
    }
    PoolLease lease;
    if (!lease.c) return nullptr;  // bpg_last_error() holds the reason (no sm_100a device: there is no CPU fallback)
    bpg_proof_artifacts* a = nullptr;
    // blindings and the transcript-rng seed come from the OS, as the reference's thread_rng() draws do
    if (bpg_prove(lease.c, name, instance, witness, gadgets, nullptr, nullptr, &a) != BPG_OK || !a) return nullptr;
    struct ProofArtifacts* out = (struct ProofArtifacts*)calloc(1, sizeof *out);
    if (out) {
        out->commitments = a->commitments;  // ownership moves: both are malloc'ed, free_proof releases them
        out->proof = a->proof;
        out->proof_len = a->proof_len;
        out->proof_cap = a->proof_len;
        a->commitments = nullptr;
        a->proof = nullptr;
    }
    bpg_free_proof(a);
    return out;
}

bool c_verify(const char* name, const char* instance, const char* gadgets, const char* commitments, const uint8_t* proof,
              size_t proof_len) {
    if (!name || !instance || !gadgets || !commitments || !proof) {
        bpg_set_error("c_verify: NULL argument");
        return false;
    }
    PoolLease lease;
    if (!lease.c) return false;
    int accepted = 0;
    // the reference unwraps (aborts) on malformed input; here every error is `false` + bpg_last_error()
    if (bpg_verify(lease.c, name, instance, gadgets, commitments, proof, proof_len, nullptr, &accepted) != BPG_OK) return false;
    return accepted != 0;
}

void free_proof(struct ProofArtifacts* artifacts_pointer) {
    if (!artifacts_pointer) return;
    free((void*)artifacts_pointer->commitments);
    free((void*)artifacts_pointer->proof);
    free(artifacts_pointer);
}

}  // extern "C"
