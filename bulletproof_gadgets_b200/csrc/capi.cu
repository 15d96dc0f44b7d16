// extern "C" surface of libbpg.so: context, generators, MSM and transcript entry points.
// The R1CS prover/verifier entry points live in r1cs.cu, the statement front end in frontend.cpp.
#include <sched.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "ctx.hpp"
#include "host_sc.hpp"
#include "merlin.hpp"

static thread_local char g_err[512] = "";

void bpg_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

void host_ristretto_compress(uint8_t out[32], const ge_ext& p) { ge_ristretto_compress(out, p); }
// Few proofs in flight in this process: spin on the event (lowest latency).  Many: sleep in the driver
// (cudaEventBlockingSync), because dozens of spinning host threads starve the few cores of a multi-GPU host.
static std::atomic<int> g_waiters{0};
static int sync_mode() {  // BPG_SYNC = spin | yield | block | auto (default)
    static const int mode = [] {
        const char* e = getenv("BPG_SYNC");
        if (!e) return 0;
        return !strcmp(e, "spin") ? 1 : !strcmp(e, "yield") ? 2 : !strcmp(e, "block") ? 3 : 0;
    }();
    return mode;
}
cudaError_t ctx_sync(bpg_ctx* ctx) {
    CpuTimer cpu(&ctx->cpu_sync_ns);
    cudaError_t e = cudaEventRecord(ctx->ev_sync, ctx->stream);
    if (e != cudaSuccess) return e;
    const int mode = sync_mode();
    const int w = ++g_waiters;
    int polls = 0;
    if (mode == 0 && w > 4) {  // many statements in flight: every driver call costs a turn at a lock they all share -- no polling at all
        e = cudaEventSynchronize(ctx->ev_sync);
        --g_waiters;
        return e;
    }
    for (;;) {
        e = cudaEventQuery(ctx->ev_sync);
        if (e != cudaErrorNotReady) break;
        ++polls;
        if (mode == 1) continue;
        if (mode == 2) {
            if (polls >= 8) sched_yield();
            continue;
        }
        // under load every driver call contends with the other threads of the process (a call costs ~3 us of a lock all
        // threads share, and small statements are bound by exactly that): go to sleep at once
        if (polls >= ((w > 4 || mode == 3) ? 1 : 32) && (mode == 3 || w > 2 || g_waiters.load(std::memory_order_relaxed) > 2)) {
            e = cudaEventSynchronize(ctx->ev_sync);
            break;
        }
    }
    --g_waiters;
    return e;
}
// Small device->host reads go through the context's pinned staging page: a copy into pageable memory would make the
// driver spin-wait inside cudaMemcpyAsync, and spinning host threads are what starves a multi-GPU host of cores.
// Layout of the 8 KiB page: [0, 4096) points of fetch_points, [4096, 8192) staged words (d2h_stage).
int fetch_points(bpg_ctx* ctx, const ge_ext* d_pts, uint32_t n, ge_ext* h_out) {
    if (n > 32) return BPG_E_ARG;
    CUDA_TRY(cudaMemcpyAsync(ctx->h_result, d_pts, sizeof(ge_ext) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx_sync(ctx));
    if (h_out != ctx->h_result) memcpy(h_out, ctx->h_result, sizeof(ge_ext) * n);
    return BPG_OK;
}
void* d2h_stage(bpg_ctx* ctx, size_t offset, const void* d_src, size_t bytes) {
    uint8_t* dst = reinterpret_cast<uint8_t*>(ctx->h_result) + 4096 + offset;
    if (offset + bytes > 4096 || cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
        return nullptr;
    return dst;
}
bool host_is_ristretto_identity(const ge_ext& p) { return ge_is_ristretto_identity(p); }

// Device self-test of the field layer: the dedicated squaring against the general product, on pseudo-random 256-bit
// values whose limbs are biased towards 0 and 2^32-1 (carry chains) -- see bpg_selftest_field.
__global__ void __launch_bounds__(128) k_selftest_field(uint64_t n, uint64_t seed, uint32_t* __restrict__ bad) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t x = seed + 0x9e3779b97f4a7c15ull * (i + 1);
    auto next = [&]() {
        x += 0x9e3779b97f4a7c15ull;
        uint64_t z = x;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    };
    fe a;
    const uint64_t shape = next();
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint32_t sel = (uint32_t)(shape >> (4 * k)) & 15u;
        const uint32_t r = (uint32_t)next();
        a.v[k] = sel == 0 ? 0u : sel == 1 ? 0xffffffffu : sel == 2 ? 0xfffffffeu : sel == 3 ? 1u : r;
    }
    const fe s1 = fe_canon(fe_sqr(a)), s2 = fe_canon(fe_mul(a, a));
    // (a + 1)^2 - a^2 - 2a - 1 == 0 ties the squaring to the adder as well
    const fe a1 = fe_add(a, fe_one());
    const fe lhs = fe_sub(fe_sub(fe_sub(fe_sqr(a1), fe_sqr(a)), fe_add(a, a)), fe_one());
    uint32_t diff = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) diff |= s1.v[k] ^ s2.v[k];
    if (diff || !fe_is_zero(lhs)) atomicAdd(bad, 1u);
}

extern "C" {

const char* bpg_last_error(void) { return g_err; }

int bpg_selftest_field(bpg_ctx* ctx, uint64_t n, uint64_t seed, uint64_t* mismatches) {
    if (!ctx || !mismatches) return BPG_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    uint32_t* d = nullptr;
    CUDA_TRY(cudaMalloc((void**)&d, 4));
    CUDA_TRY(cudaMemsetAsync(d, 0, 4, ctx->stream));
    if (n) k_selftest_field<<<(uint32_t)((n + 127) / 128), 128, 0, ctx->stream>>>(n, seed, d);
    ctx->launches++;
    uint32_t h = 0;
    cudaError_t e = cudaMemcpyAsync(&h, d, 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = ctx_sync(ctx);
    cudaFree(d);
    CUDA_TRY(e);
    *mismatches = h;
    return BPG_OK;
}

static int ctx_init(bpg_ctx* ctx, int device);
// takes over one reference of `store`: released again when the context cannot be built
static int ctx_new(int device, GensStore* store, bpg_ctx** out) {
    bpg_ctx* ctx = new bpg_ctx();
    ctx->device = device;
    ctx->store = store;
    const int rc = ctx_init(ctx, device);
    if (rc != BPG_OK) {
        bpg_ctx_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return BPG_OK;
}
static int ctx_init(bpg_ctx* ctx, int device) {
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    {
        // Per-proof device buffers (bulk-loaded circuits) come from a pool private to this context: with the device's
        // default pool, memory freed on one stream and reused on another makes the driver chain the two streams, which
        // serialises proofs that are otherwise independent.  Freed memory stays cached (release threshold = max).
        cudaMemPoolProps props;
        memset(&props, 0, sizeof props);
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        CUDA_TRY(cudaMemPoolCreate(&ctx->pool, &props));
        uint64_t keep = ~0ull;
        CUDA_TRY(cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    for (int k = 0; k <= MSM_STAGES; k++) CUDA_TRY(cudaEventCreate(&ctx->ev_stage[k]));
    {
        int sms = 0;
        CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        if (sms > 0) ctx->sm_count = sms;
    }
    CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_sync, cudaEventBlockingSync | cudaEventDisableTiming));
    CUDA_TRY(cudaMallocHost((void**)&ctx->h_result, 64 * sizeof(ge_ext)));
    if (const char* e = getenv("BPG_TASK_LEN")) ctx->task_len = atoi(e) > 0 && atoi(e) < (1 << 20) ? atoi(e) : 0;
    if (const char* e = getenv("BPG_TARGET_CHUNKS")) ctx->target_chunks = atoi(e) > 0 ? atoi(e) : 0;
    if (const char* e = getenv("BPG_ACC_SMEM_PAD")) ctx->acc_smem_pad = atoi(e) > 0 && atoi(e) <= 200 * 1024 ? atoi(e) : 0;
    if (const char* e = getenv("BPG_ACC_VARIANT")) ctx->acc_variant = atoi(e) >= 0 && atoi(e) <= 3 ? atoi(e) : 0;
    if (const char* e = getenv("BPG_TICKETS")) ctx->use_tickets = atoi(e) != 0;
    if (const char* e = getenv("BPG_SMALL_KERNEL")) ctx->use_small_kernel = atoi(e) != 0;
    if (const char* e = getenv("BPG_SMEM_SORT")) ctx->use_smem_sort = atoi(e) != 0;
    if (const char* e = getenv("BPG_IPP_FOLD_N")) ctx->ipp_fold_n = atoi(e);
    return BPG_OK;
}

int bpg_ctx_create(int device, bpg_ctx** out) {
    if (!out) return BPG_E_ARG;
    *out = nullptr;
    // Throughput comes from dozens of contexts (streams) per GPU.  With the driver's default of 8 hardware work queues, every
    // stream waits behind the unfinished kernel chains of the streams sharing its queue: 32 queues and 96 statements in flight
    // measured 6.9 instead of 7.4 ms per config-2 statement (profiles/r02_ab5.jsonl).  Only effective when this is the first CUDA
    // call of the process; never overrides the caller's setting.
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        bpg_set_error("no CUDA device available (%s); this library has no CPU fallback",
                      e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return BPG_E_CUDA;
    }
    if (device < 0 || device >= ndev) {
        bpg_set_error("device %d out of range (have %d)", device, ndev);
        return BPG_E_ARG;
    }
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        bpg_set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return BPG_E_CUDA;
    }
    return ctx_new(device, new GensStore(), out);
}

int bpg_ctx_create_shared(bpg_ctx* parent, bpg_ctx** out) {
    if (!parent || !out) return BPG_E_ARG;
    *out = nullptr;
    CUDA_TRY(cudaSetDevice(parent->device));
    {
        std::lock_guard<std::mutex> lock(parent->store->mu);
        parent->store->refs++;
    }
    int rc = ctx_new(parent->device, parent->store, out);
    if (rc == BPG_OK) {
        (*out)->task_len = parent->task_len;
        (*out)->target_chunks = parent->target_chunks;
        (*out)->cl_min = parent->cl_min;
        (*out)->acc_variant = parent->acc_variant;
        (*out)->acc_smem_pad = parent->acc_smem_pad;
        (*out)->ipp_fold_n = parent->ipp_fold_n;
        (*out)->use_tickets = parent->use_tickets;
        (*out)->use_smem_sort = parent->use_smem_sort;
    }
    return rc;
}

void bpg_ctx_destroy(bpg_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream && ctx->ev_sync) ctx_sync(ctx);  // (a context whose construction failed half-way has neither)
    MsmWork& w = ctx->work;
    w.hist.release();
    w.bucket_off.release();
    w.meta.release();
    w.entries.release();
    w.partials.release();
    w.blockres.release();
    w.reduce_cnt.release();
    w.reduce_dbg.release();
    w.scan_tmp.release();
    w.tickets.release();
    w.sort_cnt.release();
    w.sort_base.release();
    w.sort_val.release();
    w.sort_fine.release();
    w.sort_small.release();
    ctx->d_scalars.release();
    ctx->d_points.release();
    r1cs_release_work(ctx);
    if (ctx->h_result) cudaFreeHost(ctx->h_result);
    for (int k = 0; k <= MSM_STAGES; k++)
        if (ctx->ev_stage[k]) cudaEventDestroy(ctx->ev_stage[k]);
    if (ctx->ev_sync) cudaEventDestroy(ctx->ev_sync);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
    if (ctx->store) gens_store_release(ctx->store);
    delete ctx;
}

void* bpg_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        bpg_set_error("bpg_host_alloc(%zu): pinned allocation failed", bytes);
        return nullptr;
    }
    return p;
}
void bpg_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int bpg_ctx_set(bpg_ctx* ctx, const char* key, int64_t value) {
    if (!ctx || !key) return BPG_E_ARG;
    std::string k(key);
    if (k == "task_len") {
        if (value < 0 || value >= (1 << 20)) return BPG_E_ARG;  // 0 = derived on the device from the entry count
        ctx->task_len = (int)value;
    } else if (k == "target_chunks") {
        if (value < 0 || value > (1 << 24)) return BPG_E_ARG;
        ctx->target_chunks = (int)value;
    } else if (k == "ipp_fold_n") {
        if (value < -1 || value > 4096 || (value > 0 && (value & (value - 1)))) return BPG_E_ARG;  // -1 auto, 0 never, or 2^k
        ctx->ipp_fold_n = (int)value;
    } else if (k == "acc_variant") {
        if (value < 0 || value > 3) return BPG_E_ARG;
        ctx->acc_variant = (int)value;
    } else if (k == "cl_min") {
        if (value < 1 || value > 4096) return BPG_E_ARG;
        ctx->cl_min = (int)value;
    } else if (k == "smem_sort") {
        ctx->use_smem_sort = value != 0;
    } else if (k == "tickets") {
        ctx->use_tickets = value != 0;
    } else if (k == "window_bits") {
        if (value != 0 && (value < 4 || value > 16)) return BPG_E_ARG;
        GensStore* g = ctx->store;
        std::lock_guard<std::mutex> lock(g->mu);
        if (g->table.rows && value != g->window_bits) {  // force a rebuild on next ensure
            g->garbage.push_back(g->table.rows);
            g->table = FixedTable();
            ctx->table = FixedTable();
        }
        g->window_bits = (int)value;
    } else if (k == "time_accum") {
        ctx->time_accum = value != 0;
        ctx->sum_accum_ms = 0;
        ctx->sum_entries = 0;
        ctx->sum_scatter_ms = 0;
        ctx->sum_points = 0;
        ctx->timed_msms = 0;
        for (int i = 0; i < MSM_STAGES; i++) ctx->sum_stage_ms[i] = 0;
    } else {
        return BPG_E_ARG;
    }
    return BPG_OK;
}

int64_t bpg_ctx_get(bpg_ctx* ctx, const char* key) {
    if (!ctx || !key) return -1;
    std::string k(key);
    if (k == "launches") return (int64_t)ctx->launches;
    if (k == "accum_us") return (int64_t)(ctx->last_accum_ms * 1000.0f);
    if (k == "accum_ns") return (int64_t)((double)ctx->last_accum_ms * 1e6);
    if (k == "accum_entries") return (int64_t)ctx->last_entries;
    if (k == "sum_accum_ns") return (int64_t)(ctx->sum_accum_ms * 1e6);
    if (k == "sum_entries") return (int64_t)ctx->sum_entries;
    if (k == "sum_scatter_ns") return (int64_t)(ctx->sum_scatter_ms * 1e6);
    if (k == "sum_points") return (int64_t)ctx->sum_points;
    if (k == "timed_msms") return (int64_t)ctx->timed_msms;
    if (k.rfind("stage_ns_", 0) == 0) {  // stage_ns_0 .. stage_ns_6: see MSM_STAGES (ctx.hpp)
        const int i = atoi(k.c_str() + 9);
        return i >= 0 && i < MSM_STAGES ? (int64_t)(ctx->sum_stage_ms[i] * 1e6) : -1;
    }
    if (k.rfind("phase_ns_", 0) == 0) {  // wall time per prove / verify phase (ctx.hpp PH_*)
        const int i = atoi(k.c_str() + 9);
        return i >= 0 && i < PH_COUNT ? (int64_t)ctx->phase_wall_ns[i] : -1;
    }
    if (k == "task_len") return ctx->task_len;
    if (k == "last_chunk_len") return (int64_t)ctx->last_chunk_len;
    if (k == "cpu_sync_ns") return (int64_t)ctx->cpu_sync_ns;
    if (k == "cpu_commit_ns") return (int64_t)ctx->cpu_commit_ns;
    if (k == "cpu_prove_ns") return (int64_t)ctx->cpu_prove_ns;
    if (k == "cpu_verify_ns") return (int64_t)ctx->cpu_verify_ns;
    if (k == "cpu_load_ns") return (int64_t)ctx->cpu_load_ns;
    if (k == "cpu_rng_ns") return (int64_t)ctx->cpu_rng_ns;
    if (k == "capacity") return (int64_t)ctx->table.capacity;
    if (k == "window_bits") return ctx->table.c;
    if (k == "windows") return ctx->table.K;
    if (k == "stream") return (int64_t)(intptr_t)ctx->stream;
    return -1;
}

int bpg_gens_ensure(bpg_ctx* ctx, uint64_t capacity) {
    if (!ctx) return BPG_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    return gens_build(ctx, capacity);
}

int bpg_gens_compressed(bpg_ctx* ctx, int which, uint64_t start, uint64_t count, uint8_t* out) {
    if (!ctx || !out) return BPG_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    return gens_compress_range(ctx, which, start, count, out);
}

static int msm_gens_common(bpg_ctx* ctx, const uint32_t* d_sG, uint64_t nG, const uint32_t* d_sH, uint64_t nH,
                           const uint32_t* d_sB, const uint32_t* d_sBb, uint8_t out32[32], uint64_t g_start = 0,
                           uint64_t h_start = 0) {
    const uint64_t cap = ctx->table.capacity;
    if (nG > cap || nH > cap || g_start > cap - nG || h_start > cap - nH) {  // no wrap-around in the sums
        bpg_set_error("msm: %llu/%llu scalars exceed generator capacity %llu", (unsigned long long)nG,
                      (unsigned long long)nH, (unsigned long long)cap);
        return BPG_E_GENS_LEN;
    }
    MsmSegments segs;
    memset(&segs, 0, sizeof segs);
    auto push = [&](const uint32_t* p, uint64_t base, uint64_t n) {
        if (!p || n == 0) return;
        MsmSegment& s = segs.seg[segs.nseg++];
        s.scalars = p;
        s.point_base = (uint32_t)base;
        s.count = (uint32_t)n;
        s.set_id = 0;
        s.mode = 0;
        s.period = 1;
        segs.total += (uint32_t)n;
    };
    push(d_sG, g_start, nG);
    push(d_sH, cap + h_start, nH);
    push(d_sB, 2 * cap, 1);
    push(d_sBb, 2 * cap + 1, 1);
    int rc;
    if ((rc = ctx->d_points.ensure(64))) return rc;
    if ((rc = msm_run(ctx, segs, 1, ctx->d_points.p))) return rc;
    if ((rc = fetch_points(ctx, ctx->d_points.p, 1, ctx->h_result))) return rc;
    host_ristretto_compress(out32, ctx->h_result[0]);
    return BPG_OK;
}

static int check_scalars(const uint8_t* s, uint64_t n) {
    for (uint64_t i = 0; i < n; i++)
        if (s[32 * i + 31] & 0x80) {
            bpg_set_error("scalar %llu has bit 255 set (not a valid Scalar)", (unsigned long long)i);
            return BPG_E_ARG;
        }
    return BPG_OK;
}

int bpg_msm_gens_range(bpg_ctx* ctx, const uint8_t* sG, uint64_t g_start, uint64_t nG, const uint8_t* sH, uint64_t h_start,
                       uint64_t nH, const uint8_t* sB, const uint8_t* sBb, uint8_t out32[32]) {
    if (!ctx || !out32) return BPG_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (!sG) nG = 0;
    if (!sH) nH = 0;
    if (nG > (1ull << 31) || nH > (1ull << 31) || g_start > (1ull << 31) || h_start > (1ull << 31)) {
        bpg_set_error("msm: range exceeds 2^31 points");
        return BPG_E_GENS_LEN;
    }
    int rc;
    if ((rc = check_scalars(sG, nG)) || (rc = check_scalars(sH, nH)) || (sB && (rc = check_scalars(sB, 1))) ||
        (sBb && (rc = check_scalars(sBb, 1))))
        return rc;
    uint64_t need = g_start + nG > h_start + nH ? g_start + nG : h_start + nH;
    if ((rc = gens_build(ctx, need ? need : 1))) return rc;
    const uint64_t total = nG + nH + 2;
    if ((rc = ctx->d_scalars.ensure(total * 8))) return rc;
    uint32_t* d = ctx->d_scalars.p;
    cudaStream_t st = ctx->stream;
    if (nG) CUDA_TRY(cudaMemcpyAsync(d, sG, 32 * nG, cudaMemcpyHostToDevice, st));
    if (nH) CUDA_TRY(cudaMemcpyAsync(d + 8 * nG, sH, 32 * nH, cudaMemcpyHostToDevice, st));
    if (sB) CUDA_TRY(cudaMemcpyAsync(d + 8 * (nG + nH), sB, 32, cudaMemcpyHostToDevice, st));
    if (sBb) CUDA_TRY(cudaMemcpyAsync(d + 8 * (nG + nH + 1), sBb, 32, cudaMemcpyHostToDevice, st));
    return msm_gens_common(ctx, nG ? d : nullptr, nG, nH ? d + 8 * nG : nullptr, nH, sB ? d + 8 * (nG + nH) : nullptr,
                           sBb ? d + 8 * (nG + nH + 1) : nullptr, out32, g_start, h_start);
}

int bpg_msm_gens(bpg_ctx* ctx, const uint8_t* sG, uint64_t nG, const uint8_t* sH, uint64_t nH, const uint8_t* sB,
                 const uint8_t* sBb, uint8_t out32[32]) {
    return bpg_msm_gens_range(ctx, sG, 0, nG, sH, 0, nH, sB, sBb, out32);
}

// host-only: sum of n compressed ristretto points (partial results of a point-range sharded MSM)
int bpg_point_sum(const uint8_t* points32n, uint64_t n, uint8_t out32[32]) {
    if (!out32 || (n && !points32n)) return BPG_E_ARG;
    ge_ext acc = ge_identity();
    for (uint64_t i = 0; i < n; i++) {
        ge_ext p;
        if (!ge_ristretto_decompress(&p, points32n + 32 * i)) {
            bpg_set_error("point_sum: point %llu does not decode", (unsigned long long)i);
            return BPG_E_VERIFY;
        }
        acc = ge_add(acc, p);
    }
    ge_ristretto_compress(out32, acc);
    return BPG_OK;
}

int bpg_msm_gens_range_dev(bpg_ctx* ctx, const void* d_sG, uint64_t g_start, uint64_t nG, const void* d_sH, uint64_t h_start,
                           uint64_t nH, const void* d_sB, const void* d_sBb, uint8_t out32[32]) {
    if (!ctx || !out32) return BPG_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (!d_sG) nG = 0;
    if (!d_sH) nH = 0;
    if (nG > (1ull << 31) || nH > (1ull << 31) || g_start > (1ull << 31) || h_start > (1ull << 31)) return BPG_E_GENS_LEN;
    const uint64_t need = g_start + nG > h_start + nH ? g_start + nG : h_start + nH;
    int rc;
    if ((rc = gens_build(ctx, need ? need : 1))) return rc;
    return msm_gens_common(ctx, (const uint32_t*)d_sG, nG, (const uint32_t*)d_sH, nH, (const uint32_t*)d_sB,
                           (const uint32_t*)d_sBb, out32, g_start, h_start);
}

int bpg_msm_gens_dev(bpg_ctx* ctx, const void* d_sG, uint64_t nG, const void* d_sH, uint64_t nH, const void* d_sB,
                     const void* d_sBb, uint8_t out32[32]) {
    return bpg_msm_gens_range_dev(ctx, d_sG, 0, nG, d_sH, 0, nH, d_sB, d_sBb, out32);
}

// ---------------------------------------------------------------------------- transcript
bpg_transcript* bpg_transcript_new(const uint8_t* label, size_t len) { return new bpg_transcript(label, len); }
bpg_transcript* bpg_transcript_clone(const bpg_transcript* t) { return t ? new bpg_transcript(*t) : nullptr; }
void bpg_transcript_free(bpg_transcript* t) { delete t; }
void bpg_transcript_append_message(bpg_transcript* t, const uint8_t* label, size_t label_len, const uint8_t* msg,
                                   size_t msg_len) {
    t->t.append_message(label, label_len, msg, msg_len);
}
void bpg_transcript_challenge_bytes(bpg_transcript* t, const uint8_t* label, size_t label_len, uint8_t* out,
                                    size_t out_len) {
    t->t.challenge_bytes(label, label_len, out, out_len);
}

// TranscriptRngBuilder::{rekey_with_witness_bytes("v_blinding", w)}*.finalize(seed) followed by n x fill_bytes(64):
// the s_L / s_R stream of Prover::prove (host only; concurrent callers are batched eight streams at a time).
int bpg_transcript_rng_fill64(const bpg_transcript* t, const uint8_t* witness32k, size_t k, const uint8_t seed32[32],
                              size_t warm, uint8_t* out64n, size_t n) {
    if (!t || !seed32 || (k && !witness32k) || (n && !out64n)) return BPG_E_ARG;
    std::vector<const uint8_t*> wit;
    for (size_t i = 0; i < k; i++) wit.push_back(witness32k + 32 * i);
    bpg::TranscriptRng rng = t->t.build_rng(wit, seed32);
    uint8_t tmp[64];
    bpg::ProvingScope in_flight;
    in_flight.enter();
    for (size_t i = 0; i < warm; i++) rng.fill_bytes(tmp, 64);  // e.g. the three blinding draws that come first
    rng.fill_many64(out64n, n);
    return BPG_OK;
}
int64_t bpg_rng_batcher_stat(int which) {
    uint64_t s[3];
    bpg::rng_batcher_stats(s);
    return which >= 0 && which < 3 ? (int64_t)s[which] : -1;
}

}  // extern "C"
