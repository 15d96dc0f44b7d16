// Per-GPU context: streams, generator tables, MSM work buffers.  One context per GPU; contexts
// are independent (thread-per-GPU safe); a single context is not re-entrant.
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <time.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/bpg.h"
#include "ge25519.cuh"
#include "sc25519.cuh"

void bpg_set_error(const char* fmt, ...);

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            bpg_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));    \
            return BPG_E_CUDA;                                                                      \
        }                                                                                           \
    } while (0)

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
    bool fresh = false;  // (re)allocated since the owner last cleared the flag: contents undefined
    size_t cap_hint = 0;  // capacity before the last reallocation
    int ensure(size_t n) {
        if (n <= cap) return BPG_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap_hint = cap;
        cap = 0;
        // cudaFree / cudaMalloc synchronise the whole device and hold the driver lock for milliseconds while other
        // statements are in flight (tools/gpu_timeline.py): grow by at least half so that statements of slightly
        // different sizes settle after a few reallocations instead of one per new maximum
        size_t want = n + n / 8 + 64;
        if (want < cap_hint + cap_hint / 2) want = cap_hint + cap_hint / 2;
        CUDA_TRY(cudaMalloc((void**)&p, want * sizeof(T)));
        cap_hint = want;
        cap = want;
        fresh = true;
        return BPG_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// One segment of an MSM: `count` scalars (32 B LE each, canonical) multiplying table points
// point_base .. point_base+count-1.  The bucket set an element goes to is
//   mode 0: set_id
//   mode 1: set_id + ((i mod period) >= period/2 ? 0 : 1)      (IPP G rule: G_R -> L, G_L -> R)
//   mode 2: set_id + ((i mod period) <  period/2 ? 0 : 1)      (IPP H rule: H_L -> L, H_R -> R)
//   mode 3: set_id + (i mod period)                            (IPP generator fold: one output point per residue)
struct MsmSegment {
    const uint32_t* scalars;
    uint32_t point_base;
    uint32_t count;
    uint32_t set_id;
    uint32_t mode;
    uint32_t period;
};
#define MSM_MAX_SEGMENTS 8
struct MsmSegments {
    MsmSegment seg[MSM_MAX_SEGMENTS];
    uint32_t nseg;
    uint32_t total;     // sum of counts
    uint32_t var_base;  // variable-base form: ONE table row per point (no per-window multiples), window w accumulates
                        // into bucket set set_id + w; the caller combines the K window sums with doublings
};

// Fixed-base table: rows[w][i] = affine Niels form of 2^(c*w) * P_i, i < n_points (row stride = n_points)
struct FixedTable {
    ge_niels* rows = nullptr;
    uint32_t n_points = 0;
    int c = 0, K = 0;
    uint64_t capacity = 0;  // generator capacity: points are [G_0..G_cap-1, H_0..H_cap-1, B, B_blinding]
};

// Generator tables of one GPU, shared by every context created on it with bpg_ctx_create_shared
// (contexts = independent streams + work buffers; the 400 MB fixed-base table exists once).
// Superseded tables are kept until the store dies, so a context may keep using its snapshot while
// another one grows the capacity.
struct GensStore {
    std::mutex mu;
    FixedTable table;
    FixedTable small;  // 8-bit windows over the first `small.capacity` generators: MSMs of a few thousand points
    FixedTable fold;   // 8-bit windows over ALL generators (built on first use): the IPP generator fold, whose 2 n_r bucket
                       // sets can only afford 128 buckets each
    ge_ext* gens_ext = nullptr;
    ge_niels* ped = nullptr;
    int window_bits = 0;  // 0 = auto
    std::vector<void*> garbage;
    int refs = 1;
};

// Geometry of one MSM's bucket accumulation, derived ON THE DEVICE from the entry count (scan.cu: k_scan_meta)
struct MsmMeta {
    uint32_t E;        // entries (non-zero digits)
    uint32_t CL;       // entries per chunk (= per thread of k_accumulate)
    uint32_t nchunks;  // ceil(E / CL)
    uint32_t pad;
};
#define MSM_STAGES 7  // digits+histogram, scan, digits+scatter, accumulate, bucket reduce, final, whole
#define REDUCE_MAXV 16  // values per segment of the bucket-reduction butterfly (R, M_0 .. M_14)
struct MsmWork {
    DevBuf<uint32_t> hist;          // [nsets*nb] counts (zero between MSMs: the scan re-zeroes it)
    bool hist_dirty = true;         // needs a memset before the next histogram pass
    DevBuf<uint32_t> bucket_off;    // [nsets*nb + 1] exclusive scan of the counts
    DevBuf<MsmMeta> meta;
    DevBuf<uint32_t> entries;       // [K*N] (row | sign << 31), sorted by bucket
    DevBuf<ge_ext> partials;        // slot (chunk t, bucket b) = t + b
    DevBuf<ge_ext> blockres;        // CTA / group / set states of the reduction butterfly (msm.cu: ReduceScratch)
    DevBuf<unsigned long long> reduce_dbg;  // diagnostic phase stamps of k_bucket_reduce (BPG_REDUCE_TRACE)
    DevBuf<uint32_t> reduce_cnt;    // arrival counters of the reduction's group and set stages (zero between launches)
    DevBuf<uint32_t> scan_tmp;      // tile sums of the bucket scan (multi-CTA fallback)
    DevBuf<uint32_t> sort_cnt, sort_base;  // shared-memory sort: [bins][CTAs] counts and their exclusive scan
    DevBuf<uint32_t> sort_small;           // arrival counter, bin totals, bin starts
    DevBuf<uint32_t> sort_val;             // row|sign of every entry, grouped by coarse bin
    DevBuf<uint8_t> sort_fine;             // its fine bucket (bucket & 255)
    DevBuf<uint32_t> small_cnt;     // one-kernel small MSM: arrival counter per bucket set (zero between launches)
    DevBuf<uint32_t> tickets;       // [16][points] rank of each entry inside its bucket (histogram pass -> scatter pass)
};

struct bpg_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaMemPool_t pool = nullptr;  // private stream-ordered pool: per-proof circuits never create cross-stream dependencies
    GensStore* store = nullptr;  // shared per GPU
    // snapshots of the store taken by gens_build() at the start of every operation
    FixedTable table;
    FixedTable small_table;
    FixedTable fold_table;
    ge_ext* gens_ext = nullptr;  // untabulated generators (extended), same order as the table
    MsmWork work;
    ge_ext* h_result = nullptr;  // pinned
    DevBuf<uint32_t> d_scalars;  // staging for host-provided scalars
    DevBuf<ge_ext> d_points;     // result slots of asynchronous MSMs
    ge_niels* ped = nullptr;     // radix-16 tables of B and B_blinding (points.cu); snapshot
    struct ProofWork* pw = nullptr;  // reusable device vectors of the R1CS driver (r1cs.cu)
    int task_len = 0;         // entries per k_accumulate thread; 0 = derived on the device: one full wave of equal chunks
    int target_chunks = 0;    // chunks aimed at when task_len == 0 (0 = SMs x resident CTAs x threads), at least cl_min entries each
    int cl_min = 8;
    int ipp_fold_n = -1;      // IPP: fold the generators once the vectors are this short (0 = never, -1 = automatic: 512 while
                              // several proofs are in flight, never for a lone proof -- it trades latency for GPU time)
    int acc_variant = 0;      // k_accumulate variant (msm.cu): 0 = 4 CTAs/SM, 1 = next row prefetched (registers), 2 = 5 CTAs/SM, 3 = next row by cp.async
    int use_tickets = 1;  // scatter pass without atomics (msm.cu k_digits)
    bool sort_attr_set = false;
    bool small_attr_set = false;
    bool acc_attr_set = false;
    int acc_smem_pad = 0;      // bytes of unused dynamic shared memory per k_accumulate CTA (occupancy cap, see msm.cu)
    int use_small_kernel = 1;  // MSMs of at most SMALL_KERNEL_MAX_POINTS points on the 8-bit table: ONE launch (msm.cu k_msm_small)
    int use_smem_sort = 1;  // two-level shared-memory counting sort where it applies (msm.cu k_sort_*), else k_digits
    int sm_count = 148;
    // counters for bench.py ("gpu_launches")
    uint64_t launches = 0;
    // timing of the dominant kernel (accumulate), CUDA events on ctx stream
    cudaEvent_t ev_stage[MSM_STAGES + 1] = {nullptr};  // time_accum mode: one event after every MSM stage
    double sum_stage_ms[MSM_STAGES] = {0};             // accumulated since the last reset
    uint64_t timed_msms = 0;
    uint32_t last_chunk_len = 0;
    cudaEvent_t ev_sync = nullptr;  // blocking-sync event behind ctx_sync()
    float last_accum_ms = 0.f;
    uint64_t last_entries = 0;
    double sum_scatter_ms = 0;  // digits + scatter kernel (the sort stage), same mode
    uint64_t sum_points = 0;    // scalars decomposed
    double sum_accum_ms = 0;  // accumulated over all MSMs since the last reset (time_accum mode)
    uint64_t sum_entries = 0;
    bool time_accum = false;
    // host CPU accounting (thread CPU time, ns): inside ctx_sync, and inside the C-ABI calls commit / prove / verify / load
    uint64_t cpu_sync_ns = 0, cpu_commit_ns = 0, cpu_prove_ns = 0, cpu_verify_ns = 0, cpu_load_ns = 0, cpu_rng_ns = 0;
    // WALL time of the calling thread between the phase boundaries of prove / verify (no added synchronisation: a phase
    // ends where the driver waits for the stream anyway), summed over the statements of this context; see PHASE_NAMES
    uint64_t phase_wall_ns[16] = {0};
};
// phases of phase_wall_ns (bpg_ctx_get "phase_ns_<i>")
enum {
    PH_SETUP = 0, PH_RNG, PH_PHASE1, PH_POLY, PH_TCOMMIT, PH_IPP_EARLY, PH_IPP_FOLD, PH_IPP_LATE, PH_FINAL,
    PH_V_HEAD, PH_V_SCALARS, PH_V_MSM, PH_COUNT
};
// BPG_X_SKIP (measurement only, results are WRONG): bit mask of kernels NOT launched, to read the marginal cost of a stage
// in the concurrent regime (tools/gpu_timeline.py): 1 bucket reduce, 2 accumulate, 4 sort stage, 8 late IPP rounds'
// point kernels, 16 IPP scalar kernels
inline int x_skip() {
    static const int v = [] { const char* e = getenv("BPG_X_SKIP"); return e ? atoi(e) : 0; }();
    return v;
}
struct PhaseClock {  // lap(i): everything since the previous lap goes to phase i
    uint64_t* acc;
    uint64_t last;
    static uint64_t now() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
    }
    explicit PhaseClock(uint64_t* a) : acc(a), last(now()) {}
    void lap(int i) {
        const uint64_t t = now();
        acc[i] += t - last;
        last = t;
    }
};
struct CpuTimer {  // adds the calling thread's CPU time between construction and destruction to *acc
    uint64_t* acc;
    uint64_t t0;
    static uint64_t now() {
        timespec ts;
        clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts);
        return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
    }
    explicit CpuTimer(uint64_t* a) : acc(a), t0(now()) {}
    ~CpuTimer() { *acc += now() - t0; }
};

// Waits for ctx->stream: a short poll (single proofs keep their latency) and then a BLOCKING wait, so that the many
// host threads of a throughput run sleep instead of spinning on the few cores of a multi-GPU host.
cudaError_t ctx_sync(bpg_ctx* ctx);

// msm.cu
// asynchronous on ctx->stream; writes nsets extended points to the DEVICE array d_out
int msm_run(bpg_ctx* ctx, const MsmSegments& segs, uint32_t nsets, ge_ext* d_out);
// the same over an explicit table (variable-base MSM: rows = the affine Niels forms of the caller's points)
int msm_run_table(bpg_ctx* ctx, const FixedTable& tb, const MsmSegments& segs, uint32_t nsets, ge_ext* d_out);
// copies n <= 32 extended points from device to host (through pinned staging) and waits for the stream
int fetch_points(bpg_ctx* ctx, const ge_ext* d_pts, uint32_t n, ge_ext* h_out);
// queues a small device->host copy into the pinned staging page (offset + bytes <= 4096); valid after ctx_sync()
void* d2h_stage(bpg_ctx* ctx, size_t offset, const void* d_src, size_t bytes);
// r1cs.cu
void r1cs_release_work(bpg_ctx* ctx);
// gens.cu
int gens_build(bpg_ctx* ctx, uint64_t capacity);  // ensures capacity in the shared store and refreshes ctx snapshots
int gens_build_fold_table(bpg_ctx* ctx);          // 8-bit-window table over all generators of the current capacity
void gens_store_release(GensStore* g);
int gens_compress_range(bpg_ctx* ctx, int which, uint64_t start, uint64_t count, uint8_t* out);
// host_fe.cpp
void host_ristretto_compress(uint8_t out[32], const ge_ext& p);
bool host_is_ristretto_identity(const ge_ext& p);
