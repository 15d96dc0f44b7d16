// R1CS constraint system + the Bulletproofs prover / verifier protocol drivers.
//
// Replaces, from the FairAds fork of dalek bulletproofs 2.1.0 (/root/reference/Cargo.lock:78-80, not
// vendored): r1cs::Prover::{new,commit,multiply,allocate_multiplier,constrain,prove},
// r1cs::Verifier::{new,commit,...,verify}, InnerProductProof::{create,verification_scalars},
// R1CSProof::{to_bytes,from_bytes} -- called from /root/reference/src/prove.rs:47,79-81,
// /root/reference/src/verify.rs:46,53,71, /root/reference/src/gadget.rs:32 and
// /root/reference/src/cs_buffer.rs:89-116,173-199.
//
// Division of labour: the Merlin transcript, the sequential TranscriptRng stream and a handful of
// per-proof challenge scalars stay on the host; every vector (a_L.., s_L.., w_L.., l, r, IPP state)
// lives in HBM and every group operation is a fixed-base MSM over [G | H | B | B_blinding].
// The IPP never folds points: round k's L and R are ONE two-bucket-set MSM over the original
// generators with scalars a_j * sG_i / b_j * sH_i, where sG/sH carry the folded challenge products.
// Identical group elements, hence identical proof bytes.
#include <stdlib.h>

#include <array>
#include <random>

#include "circuit.hpp"
#include "ctx.hpp"
#include "host_sc.hpp"
#include "kernels.hpp"
#include "merlin.hpp"

using bpg::Scalar;

enum { V_COMMITTED = 0, V_LEFT = 1, V_RIGHT = 2, V_OUT = 3, V_ONE = 4 };
static inline uint32_t var_kind(uint32_t v) { return v >> 29; }
static inline uint32_t var_idx(uint32_t v) { return v & ((1u << 29) - 1); }

// ------------------------------------------------------------------------------------------
// reusable device vectors
// ------------------------------------------------------------------------------------------
struct ProofWork {
    DevBuf<sc> sL, sR, w, ypow, yinv, zpow, l1, r0, r1, r3, lvec, rvec, sG, sH, mG, mH, partial, small;
    DevBuf<sc> vbl, dyn_s, ped_in;
    DevBuf<uint32_t> fail, fl_tickets, ipp_ticket;
    DevBuf<sc> fl_part;
    DevBuf<uint8_t> wide, dyn_enc;
    DevBuf<ge_ext> dyn_pts, dyn_blk;
    DevBuf<ge_niels> dyn_rows;  // variable-base Pippenger: one affine Niels row per caller point
    DevBuf<ge_ext> fold_pts;    // IPP: folded generators G'_j, H'_j once the vectors are short (prover_prove)
    uint8_t* h_pin = nullptr;  // pinned staging
    size_t h_pin_cap = 0;
    int pin(size_t n) {
        if (n <= h_pin_cap) return BPG_OK;
        if (h_pin) cudaFreeHost(h_pin);
        h_pin = nullptr;
        h_pin_cap = 0;
        const size_t want = n + n / 2 + 4096;  // (re-pinning costs tens of milliseconds of driver lock: grow generously)
        CUDA_TRY(cudaMallocHost((void**)&h_pin, want));
        h_pin_cap = want;
        return BPG_OK;
    }
    // Uploads of more than a few KB from pageable caller memory make cudaMemcpyAsync BLOCK until the stream reaches the
    // copy (measured with 48 statements in flight: 17-43 ms per call, tools/gpu_timeline.py); they go through this
    // second pinned buffer instead.  A region is written once per prove / verify / commit, and each of those ends with
    // a wait for the stream, so the next operation of the context finds it free.
    uint8_t* h_up = nullptr;
    size_t h_up_cap = 0;
    const void* stage_up(bpg_ctx* ctx, size_t offset, const void* src, size_t bytes) {
        if (bytes <= 16384) return src;  // copied inline by the driver
        if (offset + bytes > h_up_cap) {
            if (h_up) {
                if (ctx_sync(ctx) != cudaSuccess) return src;  // copies out of the old buffer may be in flight
                cudaFreeHost(h_up);
            }
            h_up = nullptr;
            h_up_cap = 0;
            const size_t want = (offset + bytes) * 2 + 4096;
            if (cudaMallocHost((void**)&h_up, want) != cudaSuccess) {
                cudaGetLastError();
                return src;  // (pageable copy: slower, still correct)
            }
            h_up_cap = want;
        }
        memcpy(h_up + offset, src, bytes);
        return h_up + offset;
    }
};
void r1cs_release_work(bpg_ctx* ctx) {
    ProofWork* p = ctx->pw;
    if (!p) return;
    DevBuf<sc>* bs[] = {&p->sL, &p->sR, &p->w, &p->ypow, &p->yinv, &p->zpow, &p->l1, &p->r0, &p->r1, &p->r3, &p->lvec,
                        &p->rvec, &p->sG, &p->sH, &p->mG, &p->mH, &p->partial, &p->small, &p->vbl, &p->dyn_s, &p->ped_in};
    for (auto* b : bs) b->release();
    p->fail.release();
    p->fl_tickets.release();
    p->ipp_ticket.release();
    p->fl_part.release();
    p->wide.release();
    p->dyn_enc.release();
    p->dyn_pts.release();
    p->dyn_blk.release();
    p->dyn_rows.release();
    p->fold_pts.release();
    if (p->h_pin) cudaFreeHost(p->h_pin);
    if (p->h_up) cudaFreeHost(p->h_up);
    delete p;
    ctx->pw = nullptr;
}
static ProofWork* work(bpg_ctx* ctx) {
    if (!ctx->pw) ctx->pw = new ProofWork();
    return ctx->pw;
}

// wide (64-byte) draws -> canonical scalars on the device
// wide[0 .. n) -> out0, wide[n .. 2n) -> out1 (s_L and s_R of the prover: one launch)
__global__ void __launch_bounds__(256) k_wide_reduce(const uint8_t* __restrict__ wide, sc* __restrict__ out0, sc* __restrict__ out1, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * n) return;
    sc* out = i < n ? out0 : out1 - n;
    const uint4* p = reinterpret_cast<const uint4*>(wide + 64 * (size_t)i);
    uint4 a = p[0], b = p[1], c = p[2], d = p[3];
    sc lo, hi;
    lo.v[0] = a.x, lo.v[1] = a.y, lo.v[2] = a.z, lo.v[3] = a.w, lo.v[4] = b.x, lo.v[5] = b.y, lo.v[6] = b.z, lo.v[7] = b.w;
    hi.v[0] = c.x, hi.v[1] = c.y, hi.v[2] = c.z, hi.v[3] = c.w, hi.v[4] = d.x, hi.v[5] = d.y, hi.v[6] = d.z, hi.v[7] = d.w;
    const uint32_t Rl[8] = SC_R_LIMBS;
    out[i] = sc_add(sc_reduce(lo), sc_mul(hi, sc_const(Rl)));
}

// raw 32-byte scalars (< 2^255) -> canonical, in place; *err |= 2 when bit 255 is set (not a valid Scalar)
__global__ void __launch_bounds__(256) k_sc_canon(sc* __restrict__ s, uint32_t n, uint32_t* __restrict__ err) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sc x = s[i];
    if (x.v[7] >> 31) atomicOr(err, 2u);
    s[i] = sc_reduce(x);
}

// ------------------------------------------------------------------------------------------
// constraint storage (host)
// ------------------------------------------------------------------------------------------
struct ConstraintStore {
    std::vector<uint32_t> term_var;
    std::vector<sc> term_coef;  // canonical
    std::vector<uint32_t> row_start{0};
    size_t num_constraints() const { return row_start.size() - 1; }
    void begin() {}
    void term(uint32_t var, const sc& coef) {
        term_var.push_back(var);
        term_coef.push_back(coef);
    }
    void end() { row_start.push_back((uint32_t)term_var.size()); }
    // all coefficients are checked BEFORE the first term is pushed: a failing call leaves the store untouched
    int add_lc(const uint32_t* vars, const uint8_t* coef32, size_t n) {
        if (n && (!vars || !coef32)) return BPG_E_ARG;
        for (size_t i = 0; i < n; i++)
            if (coef32[32 * i + 31] & 0x80) {
                bpg_set_error("coefficient %zu has bit 255 set", i);
                return BPG_E_ARG;
            }
        for (size_t i = 0; i < n; i++) term(vars[i], Scalar::from_bytes_mod_order(coef32 + 32 * i).s);
        return BPG_OK;
    }
    struct Mark {
        size_t terms, rows;
    };
    Mark mark() const { return {term_var.size(), row_start.size()}; }
    void rollback(const Mark& m) {
        term_var.resize(m.terms);
        term_coef.resize(m.terms);
        row_start.resize(m.rows);
    }
};

// ------------------------------------------------------------------------------------------
// prover / verifier objects
// ------------------------------------------------------------------------------------------
// `circ` is either a caller-owned resident circuit (bpg_prover_attach) or `owned`: the device copy that
// bpg_prover_load_cs builds directly from the caller's arrays (no host-side constraint store at all).
// Op-by-op calls (multiply / allocate_multiplier / constrain) fill the host store `cs` instead, which
// prove() uploads and transposes on the device.
struct bpg_prover {
    bpg_ctx* ctx;
    bpg::Transcript* T;
    const bpg_circuit* circ = nullptr;
    bpg_circuit* owned = nullptr;
    bpg::ProvingScope in_flight;                       // while Prover::prove runs
    ConstraintStore cs;
    std::vector<sc> aL, aR, aO, v, vbl;                // canonical
    std::vector<std::array<uint8_t, 32>> vbl_raw;     // as given (rekeys the transcript rng)
};
struct bpg_verifier {
    bpg_ctx* ctx;
    bpg::Transcript* T;
    const bpg_circuit* circ = nullptr;
    bpg_circuit* owned = nullptr;
    ConstraintStore cs;
    std::vector<std::array<uint8_t, 32>> V;
    uint64_t num_vars = 0;
};

// BPG_TRACE=1: per-phase wall times (stream synchronised at every mark) on stderr
struct Trace {
    bool on;
    cudaStream_t st;
    double t0, last;
    static double now() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    }
    Trace(cudaStream_t s) : on(getenv("BPG_TRACE") != nullptr), st(s) { t0 = last = on ? now() : 0; }
    void mark(const char* what) {
        if (!on) return;
        double a = now();
        cudaStreamSynchronize(st);  // trace mode only
        double b = now();
        fprintf(stderr, "[bpg trace] %-28s host %8.3f ms  +gpu-drain %8.3f ms  (t=%.3f)\n", what, a - last, b - a, b - t0);
        last = b;
    }
};

struct CscView {
    const uint32_t *col_start, *col_row, *long_targets;
    const sc* col_coef;
    uint32_t nt, n_long;
};
static CscView csc_view(const bpg_circuit* circ) {
    CscView v;
    v.col_start = circ->d_col_start;
    v.col_row = circ->d_col_row;
    v.col_coef = circ->d_col_coef;
    v.nt = circ->nt;
    v.long_targets = circ->d_long;
    v.n_long = circ->n_long;
    return v;
}
// per-proof circuit of the op-by-op path: uploads the host store and transposes it on the device
static int circuit_from_store(bpg_ctx* ctx, const ConstraintStore& cs, uint64_t n, uint64_t m, bpg_circuit** out, bool defer_check = false) {
    return circuit_build(ctx, n, m, cs.num_constraints(), cs.row_start.data(), cs.term_var.data(),
                         reinterpret_cast<const uint8_t*>(cs.term_coef.data()), true, out, defer_check);
}
struct CircuitGuard {  // frees a per-proof circuit on every exit path
    bpg_circuit* c = nullptr;
    ~CircuitGuard() { circuit_free(c); }
};

static void os_random(uint8_t out[32]) {
    std::random_device rd;
    for (int i = 0; i < 8; i++) {
        uint32_t x = rd();
        memcpy(out + 4 * i, &x, 4);
    }
}
static Scalar rng_scalar(bpg::TranscriptRng& rng) {
    uint8_t b[64];
    rng.fill_bytes(b, 64);
    return Scalar::from_bytes_wide(b);
}
static Scalar challenge_scalar(bpg::Transcript& T, const char* label) {
    uint8_t b[64];
    T.challenge_bytes(label, b, 64);
    return Scalar::from_bytes_wide(b);
}
static void append_scalar(bpg::Transcript& T, const char* label, const Scalar& s) {
    uint8_t b[32];
    s.to_bytes(b);
    T.append_message(label, b, 32);
}
static PowTable pow_table(const Scalar& base) {
    PowTable t;
    sc p = base.s;
    for (int k = 0; k < 32; k++) {
        t.p[k] = p;
        p = sc_mul(p, p);
    }
    return t;
}
static uint32_t next_pow2(uint64_t n) {
    uint32_t p = 1;
    while (p < n) p <<= 1;
    return p;
}
// k Pedersen commitments v_i*B + r_i*B_blinding -> compressed (host pointers, canonical scalars)
static int pedersen_batch(bpg_ctx* ctx, const sc* v, const sc* r, uint64_t k, uint8_t* out32k) {
    if (k == 0) return BPG_OK;
    int rc;
    // Keep the snapshot the caller is working with: prover_prove calls this between MSMs that index the generator table
    // by the capacity it read at its start, and another context of the same GPU may grow the shared store meanwhile
    // (superseded tables stay alive, and the B / B_blinding table does not depend on the capacity).
    if (!ctx->ped && (rc = gens_build(ctx, 1))) return rc;
    ProofWork* pw = work(ctx);
    cudaStream_t st = ctx->stream;
    if ((rc = pw->ped_in.ensure(2 * k)) || (rc = pw->dyn_pts.ensure(k)) || (rc = pw->dyn_enc.ensure(32 * k)))
        return rc;
    CUDA_TRY(cudaMemcpyAsync(pw->ped_in.p, pw->stage_up(ctx, 0, v, 32 * k), 32 * k, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(pw->ped_in.p + k, pw->stage_up(ctx, 32 * k, r, 32 * k), 32 * k, cudaMemcpyHostToDevice, st));
    pk_pedersen(st, ctx->ped, pw->ped_in.p, pw->ped_in.p + k, pw->dyn_pts.p, (uint32_t)k);
    ctx->launches++;
    if (k <= 8) {  // latency path: finish the serial inverse-square-root chain on the host
        ge_ext h[8];
        if ((rc = fetch_points(ctx, pw->dyn_pts.p, (uint32_t)k, h))) return rc;
        for (uint64_t i = 0; i < k; i++) host_ristretto_compress(out32k + 32 * i, h[i]);
    } else {
        pk_compress(st, pw->dyn_pts.p, pw->dyn_enc.p, (uint32_t)k);
        ctx->launches++;
        if ((rc = pw->pin(32 * k))) return rc;
        CUDA_TRY(cudaMemcpyAsync(pw->h_pin, pw->dyn_enc.p, 32 * k, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(ctx_sync(ctx));
        memcpy(out32k, pw->h_pin, 32 * k);
    }
    return BPG_OK;
}

static void seg_push(MsmSegments& segs, const sc* p, uint64_t base, uint64_t n, uint32_t set, uint32_t mode,
                     uint32_t period) {
    if (!p || n == 0) return;
    MsmSegment& s = segs.seg[segs.nseg++];
    s.scalars = reinterpret_cast<const uint32_t*>(p);
    s.point_base = (uint32_t)base;
    s.count = (uint32_t)n;
    s.set_id = set;
    s.mode = mode;
    s.period = period;
    segs.total += (uint32_t)n;
}

// ------------------------------------------------------------------------------------------
// Prover::prove
// ------------------------------------------------------------------------------------------
static int prover_prove(bpg_prover* P, const uint8_t* seed32, std::vector<uint8_t>* proof_out) {
    bpg_ctx* ctx = P->ctx;
    bpg::Transcript& T = *P->T;
    cudaStream_t st = ctx->stream;
    ProofWork* pw = work(ctx);
    int rc;
    const bpg_circuit* circ = P->circ;
    if (circ && (!circ->has_witness || circ->m != P->v.size() || !P->aL.empty() || P->cs.num_constraints())) {
        bpg_set_error("prove: attached circuit needs a witness, m = #commitments, and no other constraints");
        return BPG_E_ARG;
    }
    CircuitGuard per_proof;
    if (!circ) {  // op-by-op path: upload + transpose the host store now
        if ((rc = circuit_from_store(ctx, P->cs, P->aL.size(), P->v.size(), &per_proof.c, true))) return rc;
        if ((rc = circuit_set_witness(per_proof.c, reinterpret_cast<const uint8_t*>(P->aL.data()),
                                      reinterpret_cast<const uint8_t*>(P->aR.data()))))
            return rc;
        circ = per_proof.c;
    }
    const uint32_t n = circ->n, m = (uint32_t)P->v.size();
    const uint32_t npad = next_pow2(n ? n : 1);
    const uint32_t q = circ->q;
    uint32_t lg = 0;
    while ((1u << lg) < npad) lg++;

    Trace trace(st);
    PhaseClock phase(ctx->phase_wall_ns);
    P->in_flight.enter();
    struct Leave {
        bpg::ProvingScope& s;
        ~Leave() { s.leave(); }
    } leave_on_exit{P->in_flight};
    T.append_u64("m", m);
    uint8_t seed[32];
    if (seed32) memcpy(seed, seed32, 32); else os_random(seed);
    std::vector<const uint8_t*> wit;
    wit.reserve(m);
    for (auto& b : P->vbl_raw) wit.push_back(b.data());
    bpg::TranscriptRng rng = T.build_rng(wit, seed);

    if ((rc = gens_build(ctx, npad))) return rc;
    const uint64_t cap = ctx->table.capacity;
    const uint64_t iB = 2 * cap, iBb = 2 * cap + 1;

    // device vectors
    const size_t nn = npad;
    if ((rc = pw->sL.ensure(nn)) || (rc = pw->sR.ensure(nn)) || (rc = pw->w.ensure(3 * (size_t)n + m + 1)) || (rc = pw->ypow.ensure(nn)) ||
        (rc = pw->yinv.ensure(nn)) || (rc = pw->zpow.ensure(q + 1)) || (rc = pw->l1.ensure(nn)) ||
        (rc = pw->r0.ensure(nn)) || (rc = pw->r1.ensure(nn)) || (rc = pw->r3.ensure(nn)) ||
        (rc = pw->lvec.ensure(nn)) || (rc = pw->rvec.ensure(nn)) || (rc = pw->sG.ensure(nn)) ||
        (rc = pw->sH.ensure(nn)) || (rc = pw->mG.ensure(nn)) || (rc = pw->mH.ensure(nn)) ||
        (rc = pw->partial.ensure(SK_PARTIAL_SCALARS)) || (rc = pw->small.ensure(64)) ||
        (rc = pw->vbl.ensure(m + 1)) || (rc = pw->wide.ensure(128 * (size_t)n + 64)) ||
        (rc = ctx->d_points.ensure(64)) || (rc = pw->ipp_ticket.ensure(4)))
        return rc;
    if (pw->ipp_ticket.fresh) {  // ticket of the fused IPP round kernel: zero between launches
        CUDA_TRY(cudaMemsetAsync(pw->ipp_ticket.p, 0, pw->ipp_ticket.cap * 4, st));
        pw->ipp_ticket.fresh = false;
    }
    sc* small = pw->small.p;  // [0..2] blindings, [8..13] t1..t6, [16] t2_blinding, [20..21] cw, [24..25] a,b
    ge_ext* slots = ctx->d_points.p;

    const sc *d_aL = circ->d_aL, *d_aR = circ->d_aR, *d_aO = circ->d_aO;
    if (m) CUDA_TRY(cudaMemcpyAsync(pw->vbl.p, pw->stage_up(ctx, 0, P->vbl.data(), 32 * (size_t)m), 32 * (size_t)m, cudaMemcpyHostToDevice, st));

    // rng order: i, o, s blindings, then s_L[0..n), s_R[0..n)
    const Scalar i_bl = rng_scalar(rng), o_bl = rng_scalar(rng), s_bl = rng_scalar(rng);
    sc bl[3] = {i_bl.s, o_bl.s, s_bl.s};
    CUDA_TRY(cudaMemcpyAsync(small, bl, 96, cudaMemcpyHostToDevice, st));

    MsmSegments segs;
    memset(&segs, 0, sizeof segs);
    seg_push(segs, d_aL, 0, n, 0, 0, 1);
    seg_push(segs, d_aR, cap, n, 0, 0, 1);
    seg_push(segs, small + 0, iBb, 1, 0, 0, 1);
    if ((rc = msm_run(ctx, segs, 1, slots + 0))) return rc;  // A_I1
    memset(&segs, 0, sizeof segs);
    seg_push(segs, d_aO, 0, n, 0, 0, 1);
    seg_push(segs, small + 1, iBb, 1, 0, 0, 1);
    if ((rc = msm_run(ctx, segs, 1, slots + 1))) return rc;  // A_O1

    trace.mark("upload+A_I1+A_O1 launched");
    phase.lap(PH_SETUP);
    // the sequential STROBE stream runs on the host while the two MSMs above execute
    if (n) {
        if ((rc = pw->pin(128 * (size_t)n))) return rc;
        {
            CpuTimer cpu_rng(&ctx->cpu_rng_ns);
            rng.fill_many64(pw->h_pin, 2 * (size_t)n);  // 2n x fill_bytes(64); batched across proofs in flight
        }
        CUDA_TRY(cudaMemcpyAsync(pw->wide.p, pw->h_pin, 128 * (size_t)n, cudaMemcpyHostToDevice, st));
        k_wide_reduce<<<(2 * n + 255) / 256, 256, 0, st>>>(pw->wide.p, pw->sL.p, pw->sR.p, n);
        ctx->launches++;
    }
    trace.mark("rng s_L,s_R (host keccak)");
    phase.lap(PH_RNG);
    memset(&segs, 0, sizeof segs);
    seg_push(segs, pw->sL.p, 0, n, 0, 0, 1);
    seg_push(segs, pw->sR.p, cap, n, 0, 0, 1);
    seg_push(segs, small + 2, iBb, 1, 0, 0, 1);
    if ((rc = msm_run(ctx, segs, 1, slots + 2))) return rc;  // S1

    const CscView csc = csc_view(circ);

    trace.mark("S1 msm");
    ge_ext hp[8];
    if ((rc = fetch_points(ctx, slots, 3, hp))) return rc;
    uint8_t A_I1[32], A_O1[32], S1[32];
    host_ristretto_compress(A_I1, hp[0]);
    host_ristretto_compress(A_O1, hp[1]);
    host_ristretto_compress(S1, hp[2]);
    static const uint8_t ZERO32[32] = {0};
    T.append_message("A_I1", A_I1, 32);
    T.append_message("A_O1", A_O1, 32);
    T.append_message("S1", S1, 32);
    T.append_message("dom-sep", reinterpret_cast<const uint8_t*>("r1cs-1phase"), 11);
    T.append_message("A_I2", ZERO32, 32);
    T.append_message("A_O2", ZERO32, 32);
    T.append_message("S2", ZERO32, 32);
    const Scalar y = challenge_scalar(T, "y"), z = challenge_scalar(T, "z");
    trace.mark("fetch+compress A,S; y,z");
    phase.lap(PH_PHASE1);
    const Scalar y_inv = y.invert();

    {
        PowJobs pj;
        pj.count = 3;
        pj.out[0] = pw->ypow.p, pj.tbl[0] = pow_table(y), pj.n[0] = npad, pj.start[0] = 0;
        pj.out[1] = pw->yinv.p, pj.tbl[1] = pow_table(y_inv), pj.n[1] = npad, pj.start[1] = 0;
        pj.out[2] = pw->zpow.p, pj.tbl[2] = pow_table(z), pj.n[2] = q, pj.start[2] = 1;
        sk_powers_multi(st, pj);
    }
    sc* wL = pw->w.p;
    sc* wR = wL + n;
    sc* wO = wR + n;
    sc* wV = wO + n;
    if ((rc = pw->fl_part.ensure(64 * (size_t)csc.n_long + 1)) || (rc = pw->fl_tickets.ensure(csc.n_long + 1))) return rc;
    CUDA_TRY(cudaMemsetAsync(pw->fl_tickets.p, 0, 4 * (size_t)(csc.n_long + 1), st));
    sk_flatten(st, csc.col_start, csc.col_row, csc.col_coef, pw->zpow.p, pw->w.p, csc.nt - 1, 3 * n, csc.long_targets, csc.n_long,
               pw->fl_part.p, pw->fl_tickets.p);
    sk_lr_poly(st, d_aL, d_aR, d_aO, pw->sL.p, pw->sR.p, wL, wR, wO, pw->ypow.p, pw->yinv.p, pw->l1.p,
               pw->r0.p, pw->r1.p, pw->r3.p, pw->partial.p, small + 8, n);
    sk_dot(st, wV, pw->vbl.p, m, small + 16);
    ctx->launches += 5;
    trace.mark("powers+flatten+lr_poly");
    const sc* th = static_cast<const sc*>(d2h_stage(ctx, 0, small + 8, 9 * 32));
    if (!th) return BPG_E_CUDA;
    CUDA_TRY(ctx_sync(ctx));
    phase.lap(PH_POLY);
    const Scalar t1 = Scalar::from_sc(th[0]), t2 = Scalar::from_sc(th[1]), t3 = Scalar::from_sc(th[2]),
                 t4 = Scalar::from_sc(th[3]), t5 = Scalar::from_sc(th[4]), t6 = Scalar::from_sc(th[5]);
    const Scalar tb2 = Scalar::from_sc(th[8]);

    const Scalar tb1 = rng_scalar(rng), tb3 = rng_scalar(rng), tb4 = rng_scalar(rng), tb5 = rng_scalar(rng),
                 tb6 = rng_scalar(rng);
    sc tv[5] = {t1.s, t3.s, t4.s, t5.s, t6.s}, tr[5] = {tb1.s, tb3.s, tb4.s, tb5.s, tb6.s};
    uint8_t Tc[5][32];
    if ((rc = pedersen_batch(ctx, tv, tr, 5, &Tc[0][0]))) return rc;
    T.append_message("T_1", Tc[0], 32);
    T.append_message("T_3", Tc[1], 32);
    T.append_message("T_4", Tc[2], 32);
    T.append_message("T_5", Tc[3], 32);
    T.append_message("T_6", Tc[4], 32);
    trace.mark("T commitments");
    phase.lap(PH_TCOMMIT);
    const Scalar u = challenge_scalar(T, "u"), x = challenge_scalar(T, "x");

    auto poly6 = [&](const Scalar& c1, const Scalar& c2, const Scalar& c3, const Scalar& c4, const Scalar& c5,
                     const Scalar& c6) { return x * (c1 + x * (c2 + x * (c3 + x * (c4 + x * (c5 + x * c6))))); };
    const Scalar t_x = poly6(t1, t2, t3, t4, t5, t6);
    const Scalar t_x_blinding = poly6(tb1, tb2, tb3, tb4, tb5, tb6);
    const Scalar e_blinding = x * (i_bl + x * (o_bl + x * s_bl));

    sk_eval_lr(st, pw->l1.p, d_aO, pw->sL.p, pw->r0.p, pw->r1.p, pw->r3.p, pw->ypow.p, x.s, pw->lvec.p,
               pw->rvec.p, n, npad);
    append_scalar(T, "t_x", t_x);
    append_scalar(T, "t_x_blinding", t_x_blinding);
    append_scalar(T, "e_blinding", e_blinding);
    const Scalar w = challenge_scalar(T, "w");

    trace.mark("eval l,r; w");
    // ---- inner-product argument (InnerProductProof::create) ----
    T.append_message("dom-sep", reinterpret_cast<const uint8_t*>("ipp v1"), 6);
    T.append_u64("n", npad);
    sk_ipp_init(st, pw->sG.p, pw->sH.p, pw->yinv.p, u.s, n, npad);
    ctx->launches += 2;
    std::vector<uint8_t> LR(64 * (size_t)lg);
    uint32_t nk = npad;
    // Generator fold: the first rounds run over the ORIGINAL generators (one two-bucket-set fixed-base MSM
    // each, 2 n' points whatever the round).  Once the vectors are short, the folded generators
    //     G'_j = sum_{i = j mod nk} sG_i G_i ,  H'_j = sum_{i = j mod nk} sH_i H_i        (2 nk bucket sets, 8-bit windows)
    // are materialised by ONE MSM, and the remaining rounds are thread-per-point multiplications over those 2 nk points:
    // about as long in latency, next to nothing in GPU time -- the choice while several proofs share the GPU.
    int fold_n = ctx->ipp_fold_n;
    if (fold_n < 0) fold_n = bpg::proving_now() >= 4 ? 512 : 0;
    bool late = false;
    uint32_t base_n = npad;  // size of the generator basis the round works on
    // small statements are bound by the number of driver calls: their rounds run one fused scalar kernel (and one MSM kernel)
    const bool small_ipp = npad <= IPP_SMALL_MAX && ctx->ipp_fold_n <= 0 && !x_skip();
    Scalar prev_u(1), prev_uinv(1);
    for (uint32_t round = 0; round < lg; round++, nk >>= 1) {
        // (automatic mode folds large statements only: below 2^16 multipliers a round's MSM is cheaper than a thread-per-point round)
        if (!late && fold_n >= 2 && nk >= 2 && nk <= (uint32_t)fold_n && npad >= 16 * nk && (ctx->ipp_fold_n > 0 || npad >= (1u << 16))) {
            if ((rc = gens_build_fold_table(ctx))) return rc;
            if (ctx->fold_table.rows && ctx->fold_table.capacity == cap) {  // (a table of another capacity: keep the slow path)
                if ((rc = pw->fold_pts.ensure(2 * (size_t)nk + 2)) || (rc = pw->dyn_blk.ensure(2 * ((2 * (size_t)nk + 2) / 64 + 2)))) return rc;
                memset(&segs, 0, sizeof segs);
                seg_push(segs, pw->sG.p, 0, npad, 0, 3, nk);
                seg_push(segs, pw->sH.p, cap, npad, nk, 3, nk);
                if ((rc = msm_run_table(ctx, ctx->fold_table, segs, 2 * nk, pw->fold_pts.p))) return rc;
                sk_fill_one(st, pw->sG.p, nk);
                sk_fill_one(st, pw->sH.p, nk);
                ctx->launches += 2;
                late = true;
                base_n = nk;
                phase.lap(PH_IPP_EARLY);
            }
        }
        if (small_ipp) {  // previous round's fold + cross terms + MSM scalars: one launch
            sk_ipp_round_small(st, pw->lvec.p, pw->rvec.p, pw->sG.p, pw->sH.p, pw->mG.p, pw->mH.p, small + 20, w.s, prev_u.s, prev_uinv.s,
                               round > 0, base_n, nk);
            ctx->launches++;
        } else {
            if (!(x_skip() & 16)) sk_ipp_round_fused(st, pw->lvec.p, pw->rvec.p, pw->sG.p, pw->sH.p, pw->mG.p, pw->mH.p, pw->partial.p,
                                                     pw->ipp_ticket.p, small + 20, w.s, base_n, nk);
            ctx->launches++;
        }
        if (!late) {
            memset(&segs, 0, sizeof segs);
            seg_push(segs, pw->mG.p, 0, npad, 0, 1, nk);
            seg_push(segs, pw->mH.p, cap, npad, 0, 2, nk);
            seg_push(segs, small + 20, iB, 1, 0, 0, 1);  // c_L * w on B  (Q = w*B)
            seg_push(segs, small + 21, iB, 1, 1, 0, 1);  // c_R * w on B
            if ((rc = msm_run(ctx, segs, 2, slots + 4))) return rc;
        } else {
            if (!(x_skip() & 8)) pk_dyn_msm_lr(st, pw->fold_pts.p, ctx->gens_ext + iB, pw->mG.p, pw->mH.p, small + 20, base_n, nk, pw->dyn_blk.p, slots + 4);
            ctx->launches += 2;
        }
        if ((rc = fetch_points(ctx, slots + 4, 2, hp))) return rc;
        uint8_t* Lc = LR.data() + 64 * (size_t)round;
        host_ristretto_compress(Lc, hp[0]);
        host_ristretto_compress(Lc + 32, hp[1]);
        T.append_message("L", Lc, 32);
        T.append_message("R", Lc + 32, 32);
        const Scalar uk = challenge_scalar(T, "u");
        const Scalar uk_inv = uk.invert();
        if (small_ipp) {  // folded by the next round's kernel (after the last round: below)
            prev_u = uk;
            prev_uinv = uk_inv;
            if (round + 1 == lg) {
                sk_ipp_fold(st, pw->lvec.p, pw->rvec.p, pw->sG.p, pw->sH.p, uk.s, uk_inv.s, base_n, nk);
                ctx->launches++;
            }
        } else {
            if (!(x_skip() & 16)) sk_ipp_fold(st, pw->lvec.p, pw->rvec.p, pw->sG.p, pw->sH.p, uk.s, uk_inv.s, base_n, nk);
            ctx->launches++;
        }
    }
    trace.mark("ipp rounds");
    phase.lap(late ? PH_IPP_LATE : PH_IPP_EARLY);
    const sc* ab0 = static_cast<const sc*>(d2h_stage(ctx, 0, pw->lvec.p, 32));
    const sc* ab1 = static_cast<const sc*>(d2h_stage(ctx, 32, pw->rvec.p, 32));
    if (!ab0 || !ab1) return BPG_E_CUDA;
    CUDA_TRY(ctx_sync(ctx));
    const sc ab[2] = {*ab0, *ab1};
    CUDA_TRY(cudaGetLastError());

    trace.mark("final a,b");
    phase.lap(PH_FINAL);
    // ---- R1CSProof::to_bytes (1-phase) ----
    std::vector<uint8_t>& o = *proof_out;
    o.clear();
    o.push_back(0);
    auto put = [&](const uint8_t* b) { o.insert(o.end(), b, b + 32); };
    auto puts = [&](const Scalar& s) {
        uint8_t b[32];
        s.to_bytes(b);
        put(b);
    };
    put(A_I1);
    put(A_O1);
    put(S1);
    for (int k = 0; k < 5; k++) put(Tc[k]);
    puts(t_x);
    puts(t_x_blinding);
    puts(e_blinding);
    o.insert(o.end(), LR.begin(), LR.end());
    puts(Scalar::from_sc(ab[0]));
    puts(Scalar::from_sc(ab[1]));
    return BPG_OK;
}

// ------------------------------------------------------------------------------------------
// Verifier::verify
// ------------------------------------------------------------------------------------------
struct ParsedProof {
    uint8_t A_I1[32], A_O1[32], S1[32], A_I2[32], A_O2[32], S2[32], T1[32], T3[32], T4[32], T5[32], T6[32];
    Scalar t_x, t_x_blinding, e_blinding, a, b;
    std::vector<std::array<uint8_t, 32>> Lv, Rv;
};
static int parse_proof(const uint8_t* p, size_t len, ParsedProof* out) {
    if (len == 0) return BPG_E_FORMAT;
    const uint8_t version = p[0];
    const uint8_t* body = p + 1;
    size_t blen = len - 1;
    if (blen % 32 != 0) return BPG_E_FORMAT;
    size_t minlen;
    if (version == 0) minlen = 11 * 32;
    else if (version == 1) minlen = 14 * 32;
    else return BPG_E_FORMAT;
    if (blen < minlen) return BPG_E_FORMAT;
    size_t pos = 0;
    auto rd = [&](uint8_t* dst) {
        memcpy(dst, body + pos, 32);
        pos += 32;
    };
    auto rds = [&](Scalar* s) -> bool {
        *s = Scalar::from_bytes_raw(body + pos);
        pos += 32;
        return s->is_canonical();
    };
    rd(out->A_I1);
    rd(out->A_O1);
    rd(out->S1);
    if (version == 0) {
        memset(out->A_I2, 0, 32);
        memset(out->A_O2, 0, 32);
        memset(out->S2, 0, 32);
    } else {
        rd(out->A_I2);
        rd(out->A_O2);
        rd(out->S2);
    }
    rd(out->T1);
    rd(out->T3);
    rd(out->T4);
    rd(out->T5);
    rd(out->T6);
    if (!rds(&out->t_x) || !rds(&out->t_x_blinding) || !rds(&out->e_blinding)) return BPG_E_FORMAT;
    const size_t ne = (blen - pos) / 32;
    if (ne < 2 || (ne - 2) % 2 != 0) return BPG_E_FORMAT;
    const size_t lg = (ne - 2) / 2;
    if (lg >= 32) return BPG_E_FORMAT;
    out->Lv.resize(lg);
    out->Rv.resize(lg);
    for (size_t k = 0; k < lg; k++) {
        rd(out->Lv[k].data());
        rd(out->Rv[k].data());
    }
    if (!rds(&out->a) || !rds(&out->b)) return BPG_E_FORMAT;
    return BPG_OK;
}
static bool is_zero32(const uint8_t* b) {
    uint8_t o = 0;
    for (int i = 0; i < 32; i++) o |= b[i];
    return o == 0;
}

static int verifier_verify(bpg_verifier* Vf, const uint8_t* proof, size_t proof_len, const uint8_t* seed32) {
    bpg_ctx* ctx = Vf->ctx;
    bpg::Transcript& T = *Vf->T;
    cudaStream_t st = ctx->stream;
    ProofWork* pw = work(ctx);
    int rc;
    ParsedProof pr;
    if ((rc = parse_proof(proof, proof_len, &pr))) {
        bpg_set_error("malformed proof bytes");
        return rc;
    }
    const bpg_circuit* circ = Vf->circ;
    if (circ && (circ->m != Vf->V.size() || Vf->num_vars || Vf->cs.num_constraints())) {
        bpg_set_error("verify: attached circuit needs m = #commitments and no other constraints");
        return BPG_E_ARG;
    }
    CircuitGuard per_proof;
    if (!circ) {
        if ((rc = circuit_from_store(ctx, Vf->cs, Vf->num_vars, Vf->V.size(), &per_proof.c))) return rc;
        circ = per_proof.c;
    }
    const uint32_t n = circ->n, m = (uint32_t)Vf->V.size();
    const uint32_t npad = next_pow2(n ? n : 1);
    const uint32_t q = circ->q;
    uint32_t lg = 0;
    while ((1u << lg) < npad) lg++;

    Trace trace(st);
    PhaseClock phase(ctx->phase_wall_ns);
    T.append_u64("m", m);
#define VALIDATE_APPEND(label, pt)                     \
    do {                                               \
        if (is_zero32(pt)) return BPG_E_VERIFY;        \
        T.append_message(label, pt, 32);               \
    } while (0)
    VALIDATE_APPEND("A_I1", pr.A_I1);
    VALIDATE_APPEND("A_O1", pr.A_O1);
    VALIDATE_APPEND("S1", pr.S1);
    T.append_message("dom-sep", reinterpret_cast<const uint8_t*>("r1cs-1phase"), 11);
    T.append_message("A_I2", pr.A_I2, 32);
    T.append_message("A_O2", pr.A_O2, 32);
    T.append_message("S2", pr.S2, 32);
    const Scalar y = challenge_scalar(T, "y"), z = challenge_scalar(T, "z");
    VALIDATE_APPEND("T_1", pr.T1);
    VALIDATE_APPEND("T_3", pr.T3);
    VALIDATE_APPEND("T_4", pr.T4);
    VALIDATE_APPEND("T_5", pr.T5);
    VALIDATE_APPEND("T_6", pr.T6);
    const Scalar u = challenge_scalar(T, "u"), x = challenge_scalar(T, "x");
    append_scalar(T, "t_x", pr.t_x);
    append_scalar(T, "t_x_blinding", pr.t_x_blinding);
    append_scalar(T, "e_blinding", pr.e_blinding);
    const Scalar w = challenge_scalar(T, "w");

    // InnerProductProof::verification_scalars
    if (pr.Lv.size() != lg) return BPG_E_VERIFY;  // n != 1 << lg_n
    T.append_message("dom-sep", reinterpret_cast<const uint8_t*>("ipp v1"), 6);
    T.append_u64("n", npad);
    std::vector<Scalar> uk(lg);
    for (uint32_t k = 0; k < lg; k++) {
        VALIDATE_APPEND("L", pr.Lv[k].data());
        VALIDATE_APPEND("R", pr.Rv[k].data());
        uk[k] = challenge_scalar(T, "u");
    }
    // batch inversion of [y, u_0 .. u_{lg-1}] (Montgomery's trick)
    std::vector<Scalar> inv_in(lg + 1), pref(lg + 2), inv_out(lg + 1);
    inv_in[0] = y;
    for (uint32_t k = 0; k < lg; k++) inv_in[k + 1] = uk[k];
    pref[0] = Scalar(1);
    for (uint32_t k = 0; k <= lg; k++) pref[k + 1] = pref[k] * inv_in[k];
    Scalar run = pref[lg + 1].invert();
    for (int k = (int)lg; k >= 0; k--) {
        inv_out[k] = run * pref[k];
        run = run * inv_in[k];
    }
    const Scalar y_inv = inv_out[0];

    uint8_t seed[32];
    if (seed32) memcpy(seed, seed32, 32); else os_random(seed);
    bpg::TranscriptRng rng = T.build_rng({}, seed);
    const Scalar r = rng_scalar(rng);
    const Scalar xx = x * x, rxx = r * xx, xxx = x * xx;

    trace.mark("v: transcript+challenges");
    if ((rc = gens_build(ctx, npad))) return rc;
    const uint64_t cap = ctx->table.capacity;
    const uint64_t iB = 2 * cap, iBb = 2 * cap + 1;

    const CscView csc = csc_view(circ);
    const uint32_t ndyn = 6 + m + 5 + 2 * lg;
    const size_t nn = npad;
    if ((rc = pw->w.ensure(3 * (size_t)n + m + 1)) || (rc = pw->yinv.ensure(nn)) || (rc = pw->zpow.ensure(q + 1)) ||
        (rc = pw->mG.ensure(nn)) || (rc = pw->mH.ensure(nn)) || (rc = pw->partial.ensure(SK_PARTIAL_SCALARS)) ||
        (rc = pw->small.ensure(64)) || (rc = pw->dyn_s.ensure(ndyn)) || (rc = pw->dyn_enc.ensure(32 * (size_t)ndyn)) ||
        (rc = pw->dyn_pts.ensure(ndyn)) || (rc = pw->dyn_blk.ensure(ndyn / 64 + 2)) || (rc = pw->fail.ensure(1)) ||
        (rc = ctx->d_points.ensure(64)))
        return rc;
    sc* small = pw->small.p;  // [0] delta, [1] sB, [2] sBb
    {
        PowJobs pj;
        pj.count = 2;
        pj.out[0] = pw->yinv.p, pj.tbl[0] = pow_table(y_inv), pj.n[0] = npad, pj.start[0] = 0;
        pj.out[1] = pw->zpow.p, pj.tbl[1] = pow_table(z), pj.n[1] = q, pj.start[1] = 1;
        sk_powers_multi(st, pj);
    }
    sc* wL = pw->w.p;
    sc* wR = wL + n;
    sc* wO = wR + n;
    sc* wV = wO + n;
    sc* wc = wV + m;
    if ((rc = pw->fl_part.ensure(64 * (size_t)csc.n_long + 1)) || (rc = pw->fl_tickets.ensure(csc.n_long + 1))) return rc;
    CUDA_TRY(cudaMemsetAsync(pw->fl_tickets.p, 0, 4 * (size_t)(csc.n_long + 1), st));
    sk_flatten(st, csc.col_start, csc.col_row, csc.col_coef, pw->zpow.p, pw->w.p, csc.nt, 3 * n, csc.long_targets, csc.n_long,
               pw->fl_part.p, pw->fl_tickets.p);
    trace.mark("v: csc+powers+flatten");
    VerChallenges ch;
    memset(&ch, 0, sizeof ch);
    for (uint32_t k = 0; k < lg; k++) {
        ch.u[k] = uk[k].s;
        ch.uinv[k] = inv_out[k + 1].s;
    }
    ch.x = x.s;
    ch.a = pr.a.s;
    ch.b = pr.b.s;
    ch.u_pad = u.s;
    sk_ver_scalars(st, ch, wL, wR, wO, pw->yinv.p, pw->mG.p, pw->mH.p, pw->partial.p, small + 0, n, npad, lg);
    trace.mark("v: ver_scalars");
    const Scalar w_tab = w * (pr.t_x - pr.a * pr.b);
    const Scalar sBb = -pr.e_blinding - r * pr.t_x_blinding;
    // dynamic scalars, in dalek's point order: A_I1 A_O1 S1 A_I2 A_O2 S2 | V | T_1 T_3 T_4 T_5 T_6 | L.. | R..
    std::vector<sc> hs(ndyn);
    std::vector<uint8_t> henc(32 * (size_t)ndyn);
    {
        const Scalar head[6] = {x, xx, xxx, u * x, u * xx, u * xxx};
        const uint8_t* hp6[6] = {pr.A_I1, pr.A_O1, pr.S1, pr.A_I2, pr.A_O2, pr.S2};
        for (int k = 0; k < 6; k++) {
            hs[k] = head[k].s;
            memcpy(&henc[32 * k], hp6[k], 32);
        }
        for (uint32_t j = 0; j < m; j++) {
            hs[6 + j] = sc_zero();  // filled on the device (wV_j * r x^2)
            memcpy(&henc[32 * (6 + j)], Vf->V[j].data(), 32);
        }
        const Scalar Ts[5] = {r * x, rxx * x, rxx * xx, rxx * xxx, rxx * xx * xx};
        const uint8_t* Tp[5] = {pr.T1, pr.T3, pr.T4, pr.T5, pr.T6};
        for (int k = 0; k < 5; k++) {
            hs[6 + m + k] = Ts[k].s;
            memcpy(&henc[32 * (6 + m + k)], Tp[k], 32);
        }
        for (uint32_t k = 0; k < lg; k++) {
            hs[11 + m + k] = (uk[k] * uk[k]).s;
            memcpy(&henc[32 * (11 + m + k)], pr.Lv[k].data(), 32);
            hs[11 + m + lg + k] = (inv_out[k + 1] * inv_out[k + 1]).s;
            memcpy(&henc[32 * (11 + m + lg + k)], pr.Rv[k].data(), 32);
        }
    }
    CUDA_TRY(cudaMemcpyAsync(pw->dyn_s.p, pw->stage_up(ctx, 0, hs.data(), 32 * (size_t)ndyn), 32 * (size_t)ndyn, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(pw->dyn_enc.p, pw->stage_up(ctx, 32 * (size_t)ndyn, henc.data(), 32 * (size_t)ndyn), 32 * (size_t)ndyn,
                             cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(small + 2, &sBb.s, 32, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(pw->fail.p, 0, 4, st));
    sk_ver_head(st, wV, wc, small + 0, rxx.s, r.s, xx.s, w_tab.s, pr.t_x.s, pw->dyn_s.p + 6, small + 1, m);
    pk_decompress(st, pw->dyn_enc.p, pw->dyn_pts.p, ndyn, pw->fail.p);
    trace.mark("v: head+decompress");
    ge_ext* slots = ctx->d_points.p;
    pk_dyn_msm(st, pw->dyn_pts.p, pw->dyn_s.p, ndyn, pw->dyn_blk.p, slots + 8);
    ctx->launches += 9;
    trace.mark("v: dyn msm");
    MsmSegments segs;
    memset(&segs, 0, sizeof segs);
    seg_push(segs, pw->mG.p, 0, npad, 0, 0, 1);
    seg_push(segs, pw->mH.p, cap, npad, 0, 0, 1);
    seg_push(segs, small + 1, iB, 1, 0, 0, 1);
    seg_push(segs, small + 2, iBb, 1, 0, 0, 1);
    if ((rc = msm_run(ctx, segs, 1, slots + 9))) return rc;
    pk_add2(st, slots + 8, slots + 9, slots + 10);
    ctx->launches++;
    trace.mark("v: fixed msm");
    phase.lap(PH_V_HEAD);  // host work + launches of the whole verification
    const uint32_t* failp = static_cast<const uint32_t*>(d2h_stage(ctx, 0, pw->fail.p, 4));
    ge_ext res;
    if (!failp) return BPG_E_CUDA;
    if ((rc = fetch_points(ctx, slots + 10, 1, &res))) return rc;
    phase.lap(PH_V_MSM);  // the wait for its kernels
    CUDA_TRY(cudaGetLastError());
    if (*failp) return BPG_E_VERIFY;  // a point did not decode
    return host_is_ristretto_identity(res) ? BPG_OK : BPG_E_VERIFY;
}

// ------------------------------------------------------------------------------------------
// extern "C"
// ------------------------------------------------------------------------------------------
static sc eval_lc(const bpg_prover* p, const uint32_t* vars, const sc* coefs, size_t n, bool* ok) {
    sc acc = sc_zero();
    for (size_t i = 0; i < n; i++) {
        const uint32_t k = var_kind(vars[i]), idx = var_idx(vars[i]);
        sc val;
        switch (k) {
            case V_LEFT: if (idx >= p->aL.size()) { *ok = false; return acc; } val = p->aL[idx]; break;
            case V_RIGHT: if (idx >= p->aR.size()) { *ok = false; return acc; } val = p->aR[idx]; break;
            case V_OUT: if (idx >= p->aO.size()) { *ok = false; return acc; } val = p->aO[idx]; break;
            case V_COMMITTED: if (idx >= p->v.size()) { *ok = false; return acc; } val = p->v[idx]; break;
            case V_ONE: val = sc_one(); break;
            default: *ok = false; return acc;
        }
        acc = sc_add(acc, sc_mul(coefs[i], val));
    }
    return acc;
}

extern "C" {

int bpg_pedersen_commit_batch(bpg_ctx* ctx, const uint8_t* v32k, const uint8_t* r32k, uint64_t k, uint8_t* out32k) {
    if (!ctx || (k && (!v32k || !r32k || !out32k))) return BPG_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    std::vector<sc> v(k), r(k);
    for (uint64_t i = 0; i < k; i++) {
        if ((v32k[32 * i + 31] | r32k[32 * i + 31]) & 0x80) return BPG_E_ARG;
        v[i] = Scalar::from_bytes_mod_order(v32k + 32 * i).s;
        r[i] = Scalar::from_bytes_mod_order(r32k + 32 * i).s;
    }
    return pedersen_batch(ctx, v.data(), r.data(), k, out32k);
}

int bpg_msm(bpg_ctx* ctx, const uint8_t* scalars32n, const uint8_t* points32n, uint64_t n, uint8_t out32[32]) {
    if (!ctx || !out32 || (n && (!scalars32n || !points32n))) return BPG_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    ProofWork* pw = work(ctx);
    cudaStream_t st = ctx->stream;
    int rc;
    if (n >= (1ull << 27)) {
        bpg_set_error("msm: more than 2^27 points");
        return BPG_E_ARG;
    }
    static const uint64_t pip_min = [] {  // BPG_VARBASE_MIN: smallest n that takes the Pippenger path (A/B measurements)
        const char* e = getenv("BPG_VARBASE_MIN");
        return e ? (uint64_t)atoll(e) : (uint64_t)512;
    }();
    const bool pippenger = n >= pip_min;
    if ((rc = pw->dyn_s.ensure(n + 1)) || (rc = pw->dyn_enc.ensure(32 * n + 32)) || (rc = pw->fail.ensure(2)) ||
        (rc = ctx->d_points.ensure(64)))
        return rc;
    CUDA_TRY(cudaMemsetAsync(pw->fail.p, 0, 8, st));  // [0] undecodable points, [1] invalid scalars
    if (n) {  // scalars go up as given and are reduced mod l on the device
        CUDA_TRY(cudaMemcpyAsync(pw->dyn_s.p, scalars32n, 32 * n, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(pw->dyn_enc.p, points32n, 32 * n, cudaMemcpyHostToDevice, st));
        k_sc_canon<<<(uint32_t)((n + 255) / 256), 256, 0, st>>>(pw->dyn_s.p, (uint32_t)n, pw->fail.p + 1);
        ctx->launches++;
    }
    if (pippenger) {
        // Variable-base Pippenger (dalek: vartime_multiscalar_mul above 190 points, reached from
        // /root/reference/src/verify.rs:71): the caller's points become a one-row-per-point table, window w of every
        // scalar goes to bucket set w through the same sort / accumulate / reduce kernels as the fixed-base MSM, and the
        // K window sums are combined with c doublings each.  Window width by size: ~2^(c-1) buckets <= points / 4.
        if ((rc = pw->dyn_rows.ensure(n))) return rc;
        int c = 8;
        while (c < 16 && (1ull << (c + 2)) <= n) c++;
        FixedTable tb;
        tb.rows = pw->dyn_rows.p;
        tb.n_points = (uint32_t)n;
        tb.c = c;
        tb.K = (254 + c - 1) / c;  // one spare bit: the signed recoding of a 253-bit scalar carries out of bit 252
        tb.capacity = n;
        pk_decompress_niels(st, pw->dyn_enc.p, pw->dyn_rows.p, (uint32_t)n, pw->fail.p);
        MsmSegments segs;
        memset(&segs, 0, sizeof segs);
        seg_push(segs, pw->dyn_s.p, 0, n, 0, 0, 1);
        segs.var_base = 1;
        if ((rc = msm_run_table(ctx, tb, segs, (uint32_t)tb.K, ctx->d_points.p + 16))) return rc;
        pk_window_combine(st, ctx->d_points.p + 16, tb.K, c, ctx->d_points.p + 12);
        ctx->launches += 2;
    } else {
        if ((rc = pw->dyn_pts.ensure(n + 1)) || (rc = pw->dyn_blk.ensure(n / 64 + 2))) return rc;
        pk_decompress(st, pw->dyn_enc.p, pw->dyn_pts.p, (uint32_t)n, pw->fail.p);
        pk_dyn_msm(st, pw->dyn_pts.p, pw->dyn_s.p, (uint32_t)n, pw->dyn_blk.p, ctx->d_points.p + 12);
        ctx->launches += 3;
    }
    const uint32_t* failp = static_cast<const uint32_t*>(d2h_stage(ctx, 0, pw->fail.p, 8));
    ge_ext res;
    if (!failp) return BPG_E_CUDA;
    if ((rc = fetch_points(ctx, ctx->d_points.p + 12, 1, &res))) return rc;
    const uint32_t fail = failp[0];
    if (failp[1]) {
        bpg_set_error("msm: scalar with bit 255 set (not a valid Scalar)");
        return BPG_E_ARG;
    }
    if (fail) {
        bpg_set_error("msm: %u point(s) failed to decompress", fail);
        return BPG_E_VERIFY;
    }
    host_ristretto_compress(out32, res);
    return BPG_OK;
}

int bpg_prover_new(bpg_ctx* ctx, bpg_transcript* t, bpg_prover** out) {
    if (!ctx || !t || !out) return BPG_E_ARG;
    bpg_prover* p = new bpg_prover();
    p->ctx = ctx;
    p->T = &t->t;
    p->T->append_message("dom-sep", reinterpret_cast<const uint8_t*>("r1cs v1"), 7);
    // (the rng batcher sizes its batches by the provers INSIDE prove(): a prover that is merely alive -- created and kept,
    // or abandoned -- must not make others wait for a partner that never comes)
    *out = p;
    return BPG_OK;
}
void bpg_prover_free(bpg_prover* p) {
    if (!p) return;
    circuit_free(p->owned);
    delete p;
}

int bpg_prover_commit_batch(bpg_prover* p, const uint8_t* v32k, const uint8_t* vb32k, uint64_t k, uint8_t* V_out32k,
                            uint32_t* first_var_out) {
    if (!p || (k && (!v32k || !vb32k || !V_out32k))) return BPG_E_ARG;
    CpuTimer cpu(&p->ctx->cpu_commit_ns);
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    std::vector<sc> v(k), r(k);
    for (uint64_t i = 0; i < k; i++) {
        if ((v32k[32 * i + 31] | vb32k[32 * i + 31]) & 0x80) {
            bpg_set_error("commit: scalar with bit 255 set");
            return BPG_E_ARG;
        }
        v[i] = Scalar::from_bytes_mod_order(v32k + 32 * i).s;
        r[i] = Scalar::from_bytes_mod_order(vb32k + 32 * i).s;
    }
    int rc = pedersen_batch(p->ctx, v.data(), r.data(), k, V_out32k);
    if (rc) return rc;
    if (first_var_out) *first_var_out = (uint32_t)p->v.size();
    for (uint64_t i = 0; i < k; i++) {
        p->v.push_back(v[i]);
        p->vbl.push_back(r[i]);
        std::array<uint8_t, 32> raw;
        memcpy(raw.data(), vb32k + 32 * i, 32);
        p->vbl_raw.push_back(raw);
        p->T->append_message("V", V_out32k + 32 * i, 32);
    }
    return BPG_OK;
}
int bpg_prover_commit(bpg_prover* p, const uint8_t v[32], const uint8_t v_blinding[32], uint8_t V_out[32],
                      uint32_t* var_out) {
    return bpg_prover_commit_batch(p, v, v_blinding, 1, V_out, var_out);
}

int bpg_prover_allocate_multiplier(bpg_prover* p, const uint8_t l[32], const uint8_t r[32], uint32_t vars_out[3]) {
    if (!p || !vars_out) return BPG_E_ARG;
    if (!l || !r) return BPG_E_MISSING_ASSIGNMENT;
    if ((l[31] | r[31]) & 0x80) return BPG_E_ARG;
    const sc ls = Scalar::from_bytes_mod_order(l).s, rs = Scalar::from_bytes_mod_order(r).s;
    const uint32_t i = (uint32_t)p->aL.size();
    p->aL.push_back(ls);
    p->aR.push_back(rs);
    p->aO.push_back(sc_mul(ls, rs));
    vars_out[0] = BPG_VAR_LEFT(i);
    vars_out[1] = BPG_VAR_RIGHT(i);
    vars_out[2] = BPG_VAR_OUT(i);
    return BPG_OK;
}

int bpg_prover_multiply(bpg_prover* p, const uint32_t* lvars, const uint8_t* lcoef32, size_t ln, const uint32_t* rvars,
                        const uint8_t* rcoef32, size_t rn, uint32_t vars_out[3]) {
    if (!p || !vars_out) return BPG_E_ARG;
    const uint32_t i = (uint32_t)p->aL.size();
    ConstraintStore& cs = p->cs;
    const ConstraintStore::Mark entry = cs.mark();  // every failure below restores the store to this state
    const size_t mark_v = cs.term_var.size();
    int rc;
    // constraint "left - L_i": terms of left, then (-1) L_i ; same for right
    if ((rc = cs.add_lc(lvars, lcoef32, ln))) return rc;
    bool ok = true;
    const sc lval = eval_lc(p, lvars, cs.term_coef.data() + mark_v, ln, &ok);
    const sc minus_one = sc_neg(sc_one());
    cs.term(BPG_VAR_LEFT(i), minus_one);
    cs.end();
    const size_t mark_r = cs.term_var.size();
    if (ok && (rc = cs.add_lc(rvars, rcoef32, rn))) {
        cs.rollback(entry);
        return rc;
    }
    const sc rval = ok ? eval_lc(p, rvars, cs.term_coef.data() + mark_r, rn, &ok) : sc_zero();
    if (!ok) {
        bpg_set_error("multiply: linear combination references an unallocated variable");
        cs.rollback(entry);
        return BPG_E_ARG;
    }
    cs.term(BPG_VAR_RIGHT(i), minus_one);
    cs.end();
    p->aL.push_back(lval);
    p->aR.push_back(rval);
    p->aO.push_back(sc_mul(lval, rval));
    vars_out[0] = BPG_VAR_LEFT(i);
    vars_out[1] = BPG_VAR_RIGHT(i);
    vars_out[2] = BPG_VAR_OUT(i);
    return BPG_OK;
}

int bpg_prover_constrain(bpg_prover* p, const uint32_t* vars, const uint8_t* coef32, size_t n) {
    if (!p) return BPG_E_ARG;
    int rc = p->cs.add_lc(vars, coef32, n);
    if (rc) return rc;
    p->cs.end();
    return BPG_OK;
}
int bpg_prover_load_cs(bpg_prover* p, const uint8_t* aL32n, const uint8_t* aR32n, uint64_t n, const uint32_t* row_start,
                       const uint32_t* term_var, const uint8_t* term_coef32, uint64_t q) {
    if (!p || (n && (!aL32n || !aR32n))) return BPG_E_ARG;
    if (p->circ || !p->aL.empty() || p->cs.num_constraints()) {
        bpg_set_error("load_cs: the bulk loader must be the only constraint-system call on a prover");
        return BPG_E_ARG;
    }
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    // straight from the caller's arrays to HBM: upload, reduce mod l, a_O = a_L*a_R and the transposition all
    // run on the device (no host-side constraint store)
    int rc = circuit_build(p->ctx, n, p->v.size(), q, row_start, term_var, term_coef32, true, &p->owned, true);
    if (rc) return rc;
    if ((rc = circuit_set_witness(p->owned, aL32n, aR32n))) {
        circuit_free(p->owned);
        p->owned = nullptr;
        return rc;
    }
    p->circ = p->owned;
    return BPG_OK;
}
int bpg_prover_load_cs_bits(bpg_prover* p, uint64_t n, const bpg_bit_run* runs, uint64_t n_runs, const uint8_t* aL32h,
                            const uint8_t* aR32h, const uint32_t* host_index, uint64_t h, const uint32_t* row_start,
                            const uint32_t* term_var, const uint8_t* term_coef32, uint64_t q) {
    if (!p || (n_runs && !runs) || (h && (!aL32h || !aR32h || !host_index))) return BPG_E_ARG;
    if (p->circ || !p->aL.empty() || p->cs.num_constraints()) {
        bpg_set_error("load_cs: the bulk loader must be the only constraint-system call on a prover");
        return BPG_E_ARG;
    }
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    int rc = circuit_build(p->ctx, n, p->v.size(), q, row_start, term_var, term_coef32, true, &p->owned, true);
    if (rc) return rc;
    if ((rc = circuit_set_witness_bits(p->owned, runs, n_runs, aL32h, aR32h, host_index, h))) {
        circuit_free(p->owned);
        p->owned = nullptr;
        return rc;
    }
    p->circ = p->owned;
    return BPG_OK;
}
uint64_t bpg_prover_num_constraints(const bpg_prover* p) {
    return !p ? 0 : p->circ ? p->circ->q : p->cs.num_constraints();
}
uint64_t bpg_prover_num_multipliers(const bpg_prover* p) { return !p ? 0 : p->circ ? p->circ->n : p->aL.size(); }

int bpg_prover_prove(bpg_prover* p, const uint8_t* rng_seed32, uint8_t* proof_out, size_t proof_cap, size_t* proof_len) {
    if (!p || !proof_out || !proof_len) return BPG_E_ARG;
    CpuTimer cpu(&p->ctx->cpu_prove_ns);
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    std::vector<uint8_t> proof;
    int rc = prover_prove(p, rng_seed32, &proof);
    if (rc) return rc;
    if (proof.size() > proof_cap) {
        bpg_set_error("proof buffer too small: need %zu", proof.size());
        *proof_len = proof.size();
        return BPG_E_ARG;
    }
    memcpy(proof_out, proof.data(), proof.size());
    *proof_len = proof.size();
    return BPG_OK;
}

int bpg_prover_attach(bpg_prover* p, const bpg_circuit* c) {
    if (!p || !c || c->device != p->ctx->device || p->circ) return BPG_E_ARG;  // same GPU, nothing attached yet
    p->circ = c;
    return BPG_OK;
}
int bpg_verifier_attach(bpg_verifier* v, const bpg_circuit* c) {
    if (!v || !c || c->device != v->ctx->device || v->circ) return BPG_E_ARG;
    v->circ = c;
    return BPG_OK;
}

int bpg_verifier_new(bpg_ctx* ctx, bpg_transcript* t, bpg_verifier** out) {
    if (!ctx || !t || !out) return BPG_E_ARG;
    bpg_verifier* v = new bpg_verifier();
    v->ctx = ctx;
    v->T = &t->t;
    v->T->append_message("dom-sep", reinterpret_cast<const uint8_t*>("r1cs v1"), 7);
    *out = v;
    return BPG_OK;
}
void bpg_verifier_free(bpg_verifier* v) {
    if (!v) return;
    circuit_free(v->owned);
    delete v;
}
int bpg_verifier_commit(bpg_verifier* v, const uint8_t V[32], uint32_t* var_out) {
    if (!v || !V) return BPG_E_ARG;
    std::array<uint8_t, 32> a;
    memcpy(a.data(), V, 32);
    if (var_out) *var_out = BPG_VAR_COMMITTED((uint32_t)v->V.size());
    v->V.push_back(a);
    v->T->append_message("V", V, 32);
    return BPG_OK;
}
int bpg_verifier_commit_batch(bpg_verifier* v, const uint8_t* V32k, uint64_t k, uint32_t* first_var_out) {
    if (!v || (k && !V32k)) return BPG_E_ARG;
    if (first_var_out) *first_var_out = BPG_VAR_COMMITTED((uint32_t)v->V.size());
    for (uint64_t i = 0; i < k; i++) {
        int rc = bpg_verifier_commit(v, V32k + 32 * i, nullptr);
        if (rc) return rc;
    }
    return BPG_OK;
}
int bpg_verifier_allocate_multiplier(bpg_verifier* v, uint32_t vars_out[3]) {
    if (!v || !vars_out) return BPG_E_ARG;
    const uint32_t i = (uint32_t)v->num_vars++;
    vars_out[0] = BPG_VAR_LEFT(i);
    vars_out[1] = BPG_VAR_RIGHT(i);
    vars_out[2] = BPG_VAR_OUT(i);
    return BPG_OK;
}
int bpg_verifier_multiply(bpg_verifier* v, const uint32_t* lvars, const uint8_t* lcoef32, size_t ln,
                          const uint32_t* rvars, const uint8_t* rcoef32, size_t rn, uint32_t vars_out[3]) {
    if (!v || !vars_out) return BPG_E_ARG;
    const uint32_t i = (uint32_t)v->num_vars;
    int rc;
    const sc minus_one = sc_neg(sc_one());
    const ConstraintStore::Mark entry = v->cs.mark();
    if ((rc = v->cs.add_lc(lvars, lcoef32, ln))) return rc;
    v->cs.term(BPG_VAR_LEFT(i), minus_one);
    v->cs.end();
    if ((rc = v->cs.add_lc(rvars, rcoef32, rn))) {
        v->cs.rollback(entry);  // the left row must not stay behind without its multiplier
        return rc;
    }
    v->cs.term(BPG_VAR_RIGHT(i), minus_one);
    v->cs.end();
    v->num_vars++;
    vars_out[0] = BPG_VAR_LEFT(i);
    vars_out[1] = BPG_VAR_RIGHT(i);
    vars_out[2] = BPG_VAR_OUT(i);
    return BPG_OK;
}
int bpg_verifier_constrain(bpg_verifier* v, const uint32_t* vars, const uint8_t* coef32, size_t n) {
    if (!v) return BPG_E_ARG;
    int rc = v->cs.add_lc(vars, coef32, n);
    if (rc) return rc;
    v->cs.end();
    return BPG_OK;
}
int bpg_verifier_load_cs(bpg_verifier* v, uint64_t n, const uint32_t* row_start, const uint32_t* term_var,
                         const uint8_t* term_coef32, uint64_t q) {
    if (!v) return BPG_E_ARG;
    if (v->circ || v->num_vars || v->cs.num_constraints()) {
        bpg_set_error("load_cs: the bulk loader must be the only constraint-system call on a verifier");
        return BPG_E_ARG;
    }
    CUDA_TRY(cudaSetDevice(v->ctx->device));
    int rc = circuit_build(v->ctx, n, v->V.size(), q, row_start, term_var, term_coef32, true, &v->owned);
    if (rc) return rc;
    v->circ = v->owned;
    return BPG_OK;
}
uint64_t bpg_verifier_num_vars(const bpg_verifier* v) { return !v ? 0 : v->circ ? v->circ->n : v->num_vars; }
int bpg_verifier_verify(bpg_verifier* v, const uint8_t* proof, size_t proof_len, const uint8_t* rng_seed32) {
    if (!v || !proof) return BPG_E_ARG;
    CpuTimer cpu(&v->ctx->cpu_verify_ns);
    CUDA_TRY(cudaSetDevice(v->ctx->device));
    return verifier_verify(v, proof, proof_len, rng_seed32);
}

}  // extern "C"
