// Generator derivation and the device-resident fixed-base window tables.
//
// Replaces bulletproofs `BulletproofGens::new(capacity, 1)` / `GeneratorsChain` / `PedersenGens::default`
// (generators.rs of the FairAds fork, /root/reference/Cargo.lock:78-80, not vendored) as called at
// /root/reference/src/prove.rs:46,78 and /root/reference/src/verify.rs:45,70 (SURVEY.md rows a2, K8,
// K9, f1).  The SHAKE256 stream is sequential and stays on the host; the 4*capacity Elligator maps,
// the window multiples 2^(c*w)*P and the affine-Niels normalisation run on the GPU.  The reference
// rebuilds the generators on every run; here they are cached per context.
#include "ctx.hpp"
#include "kernels.hpp"
#include "merlin.hpp"

#define SMALL_TABLE_CAP 4096u

__global__ void __launch_bounds__(128) k_uniform_to_ext(const uint8_t* __restrict__ uniform, ge_ext* __restrict__ out,
                                                        uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    __align__(16) uint8_t b[64];
    const uint4* src = reinterpret_cast<const uint4*>(uniform + 64 * (size_t)i);
    uint4* dst = reinterpret_cast<uint4*>(b);
#pragma unroll
    for (int k = 0; k < 4; k++) dst[k] = src[k];
    out[i] = ge_from_uniform_bytes(b);
}

__global__ void __launch_bounds__(128) k_decompress(const uint8_t* __restrict__ in, ge_ext* __restrict__ out,
                                                    uint32_t n, uint32_t* __restrict__ fail) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t b[32];
    for (int k = 0; k < 32; k++) b[k] = in[32 * (size_t)i + k];
    ge_ext p;
    if (!ge_ristretto_decompress(&p, b)) {
        atomicAdd(fail, 1u);
        p = ge_identity();
    }
    out[i] = p;
}

__global__ void __launch_bounds__(128) k_compress(const ge_ext* __restrict__ in, uint8_t* __restrict__ out,
                                                  uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t b[32];
    ge_ristretto_compress(b, in[i]);
    for (int k = 0; k < 32; k++) out[32 * (size_t)i + k] = b[k];
}

// tmp[w*n + i] = 2^(c*w) * P_i
__global__ void __launch_bounds__(128) k_window_multiples(const ge_ext* __restrict__ pts, ge_ext* __restrict__ tmp,
                                                          uint32_t n, int c, int K) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ge_ext p = pts[i];
    for (int w = 0; w < K; w++) {
        tmp[(size_t)w * n + i] = p;
        if (w + 1 < K) {
#pragma unroll 1
            for (int k = 0; k < c; k++) p = ge_dbl(p);
        }
    }
}

// rows[e] = affine Niels of tmp[e]
__global__ void __launch_bounds__(128) k_to_niels(const ge_ext* __restrict__ tmp, ge_niels* __restrict__ rows,
                                                  uint64_t total) {
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    ge_ext p = tmp[e];
    rows[e] = ge_to_niels(p, fe_invert(p.Z));
}

static const uint8_t BASEPOINT_COMPRESSED[32] = {0xe2, 0xf2, 0xae, 0x0a, 0x6a, 0xbc, 0x4e, 0x71, 0xa8, 0x84, 0xa9,
                                                 0x61, 0xc5, 0x00, 0x51, 0x5f, 0x58, 0xe3, 0x0b, 0x6a, 0xa5, 0x82,
                                                 0xdd, 0x8d, 0xb6, 0xa6, 0x59, 0x45, 0xe0, 0x8d, 0x2d, 0x76};

static void snapshot(bpg_ctx* ctx) {
    GensStore* g = ctx->store;
    ctx->table = g->table;
    ctx->small_table = g->small;
    ctx->fold_table = g->fold;
    ctx->gens_ext = g->gens_ext;
    ctx->ped = g->ped;
}

int gens_build(bpg_ctx* ctx, uint64_t capacity) {
    if (capacity == 0) capacity = 1;
    GensStore* g = ctx->store;
    std::lock_guard<std::mutex> lock(g->mu);
    if (g->table.rows && g->table.capacity >= capacity) {
        snapshot(ctx);
        return BPG_OK;
    }
    // grow geometrically so repeated small requests do not rebuild
    uint64_t cap = g->table.capacity ? g->table.capacity : 1;
    while (cap < capacity) cap *= 2;
    const int c = g->window_bits ? g->window_bits : 16;
    const int K = (256 + c - 1) / c;
    const uint64_t n = 2 * cap + 2;
    if ((uint64_t)K * n >= (1ull << 31)) {
        bpg_set_error("gens_build: capacity %llu too large", (unsigned long long)cap);
        return BPG_E_ARG;
    }
    cudaStream_t st = ctx->stream;

    // host: 64 uniform bytes per point.  [G chain | H chain | unused(B) | B_blinding]
    std::vector<uint8_t> uni(64 * n, 0);
    for (int which = 0; which < 2; which++) {
        bpg::Sponge sh = bpg::shake256();
        const uint8_t label[5] = {(uint8_t)(which ? 'H' : 'G'), 0, 0, 0, 0};  // party index 0, LE32
        sh.absorb(reinterpret_cast<const uint8_t*>("GeneratorsChain"), 15);
        sh.absorb(label, 5);
        sh.squeeze(uni.data() + 64 * cap * which, 64 * cap);
    }
    bpg::sha3_512(BASEPOINT_COMPRESSED, 32, uni.data() + 64 * (2 * cap + 1));

    uint8_t* d_uni = nullptr;
    uint8_t* d_bp = nullptr;
    uint32_t* d_fail = nullptr;
    ge_ext* d_ext = nullptr;
    ge_ext* d_tmp = nullptr;
    ge_niels* d_rows = nullptr;
    ge_niels* d_ped = nullptr;
    CUDA_TRY(cudaMalloc((void**)&d_uni, uni.size()));
    CUDA_TRY(cudaMalloc((void**)&d_bp, 32));
    CUDA_TRY(cudaMalloc((void**)&d_fail, 4));
    CUDA_TRY(cudaMalloc((void**)&d_ext, n * sizeof(ge_ext)));
    CUDA_TRY(cudaMalloc((void**)&d_tmp, (size_t)K * n * sizeof(ge_ext)));
    CUDA_TRY(cudaMalloc((void**)&d_rows, (size_t)K * n * sizeof(ge_niels)));
    CUDA_TRY(cudaMalloc((void**)&d_ped, 1024 * sizeof(ge_niels)));
    CUDA_TRY(cudaMemcpyAsync(d_uni, uni.data(), uni.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_bp, BASEPOINT_COMPRESSED, 32, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(d_fail, 0, 4, st));
    const uint32_t n32 = (uint32_t)n;
    k_uniform_to_ext<<<(n32 + 127) / 128, 128, 0, st>>>(d_uni, d_ext, n32);
    k_decompress<<<1, 128, 0, st>>>(d_bp, d_ext + 2 * cap, 1, d_fail);  // B overwrites its placeholder
    k_window_multiples<<<(n32 + 127) / 128, 128, 0, st>>>(d_ext, d_tmp, n32, c, K);
    const uint64_t total = (uint64_t)K * n;
    k_to_niels<<<(uint32_t)((total + 127) / 128), 128, 0, st>>>(d_tmp, d_rows, total);
    pk_pedersen_table(st, d_ext, (uint32_t)(2 * cap), d_ped);
    ctx->launches += 5;
    // Small-MSM table: 8-bit windows (32 rows per point, 128 signed buckets) over [G_0..G_sc-1 | H_0..H_sc-1 | B | B~].
    // With 2^15 buckets a 1 000-point MSM spends its time walking empty buckets; with 128 it does not.
    const uint64_t sc_cap = cap < SMALL_TABLE_CAP ? cap : SMALL_TABLE_CAP;
    const uint32_t ns = (uint32_t)(2 * sc_cap + 2);
    const int cs = 8, Ks = 32;
    ge_ext *d_sub = nullptr, *d_stmp = nullptr;
    ge_niels* d_srows = nullptr;
    CUDA_TRY(cudaMalloc((void**)&d_sub, ns * sizeof(ge_ext)));
    CUDA_TRY(cudaMalloc((void**)&d_stmp, (size_t)Ks * ns * sizeof(ge_ext)));
    CUDA_TRY(cudaMalloc((void**)&d_srows, (size_t)Ks * ns * sizeof(ge_niels)));
    CUDA_TRY(cudaMemcpyAsync(d_sub, d_ext, sc_cap * sizeof(ge_ext), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_sub + sc_cap, d_ext + cap, sc_cap * sizeof(ge_ext), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_sub + 2 * sc_cap, d_ext + 2 * cap, 2 * sizeof(ge_ext), cudaMemcpyDeviceToDevice, st));
    k_window_multiples<<<(ns + 127) / 128, 128, 0, st>>>(d_sub, d_stmp, ns, cs, Ks);
    k_to_niels<<<(uint32_t)(((uint64_t)Ks * ns + 127) / 128), 128, 0, st>>>(d_stmp, d_srows, (uint64_t)Ks * ns);
    ctx->launches += 2;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(ctx_sync(ctx));
    cudaFree(d_uni);
    cudaFree(d_bp);
    cudaFree(d_fail);
    cudaFree(d_tmp);
    cudaFree(d_sub);
    cudaFree(d_stmp);
    if (g->small.rows) g->garbage.push_back(g->small.rows);
    g->small.rows = d_srows;
    g->small.n_points = ns;
    g->small.c = cs;
    g->small.K = Ks;
    g->small.capacity = sc_cap;
    // other contexts of this GPU may still be reading the superseded tables: keep them until the store dies
    if (g->fold.rows) g->garbage.push_back(g->fold.rows);
    g->fold = FixedTable();  // rebuilt on demand for the new capacity
    if (g->table.rows) g->garbage.push_back(g->table.rows);
    if (g->gens_ext) g->garbage.push_back(g->gens_ext);
    if (g->ped) g->garbage.push_back(g->ped);
    g->gens_ext = d_ext;
    g->ped = d_ped;
    g->table.rows = d_rows;
    g->table.n_points = n32;
    g->table.c = c;
    g->table.K = K;
    g->table.capacity = cap;
    snapshot(ctx);
    return BPG_OK;
}

// 8-bit windows (32 rows per point) over [G | H | B | B~] in the index space of the big table: 805 MB at capacity 2^17.
int gens_build_fold_table(bpg_ctx* ctx) {
    GensStore* g = ctx->store;
    std::lock_guard<std::mutex> lock(g->mu);
    if (g->fold.rows && g->fold.capacity == g->table.capacity) {
        ctx->fold_table = g->fold;  // (only this: the caller keeps the snapshot of the other tables it is working with)
        return BPG_OK;
    }
    if (!g->gens_ext || !g->table.rows) return BPG_E_ARG;
    const uint32_t n = g->table.n_points;
    const int c = 8, K = 32;
    cudaStream_t st = ctx->stream;
    ge_ext* d_tmp = nullptr;
    ge_niels* d_rows = nullptr;
    CUDA_TRY(cudaMalloc((void**)&d_tmp, (size_t)K * n * sizeof(ge_ext)));
    CUDA_TRY(cudaMalloc((void**)&d_rows, (size_t)K * n * sizeof(ge_niels)));
    k_window_multiples<<<(n + 127) / 128, 128, 0, st>>>(g->gens_ext, d_tmp, n, c, K);
    k_to_niels<<<(uint32_t)(((uint64_t)K * n + 127) / 128), 128, 0, st>>>(d_tmp, d_rows, (uint64_t)K * n);
    ctx->launches += 2;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(ctx_sync(ctx));
    cudaFree(d_tmp);
    if (g->fold.rows) g->garbage.push_back(g->fold.rows);
    g->fold.rows = d_rows;
    g->fold.n_points = n;
    g->fold.c = c;
    g->fold.K = K;
    g->fold.capacity = g->table.capacity;
    ctx->fold_table = g->fold;
    return BPG_OK;
}

void gens_store_release(GensStore* g) {
    bool last;
    {
        std::lock_guard<std::mutex> lock(g->mu);
        last = --g->refs == 0;
    }
    if (!last) return;
    if (g->table.rows) cudaFree(g->table.rows);
    if (g->small.rows) cudaFree(g->small.rows);
    if (g->fold.rows) cudaFree(g->fold.rows);
    if (g->gens_ext) cudaFree(g->gens_ext);
    if (g->ped) cudaFree(g->ped);
    for (void* p : g->garbage) cudaFree(p);
    delete g;
}

int gens_compress_range(bpg_ctx* ctx, int which, uint64_t start, uint64_t count, uint8_t* out) {
    const uint64_t cap = ctx->table.capacity;
    if (!ctx->gens_ext) return BPG_E_ARG;
    uint64_t base;
    if (which == 0 || which == 1) {
        if (start + count > cap) return BPG_E_GENS_LEN;
        base = (which ? cap : 0) + start;
    } else if (which == 2 || which == 3) {
        if (start != 0 || count != 1) return BPG_E_ARG;
        base = 2 * cap + (which - 2);
    } else {
        return BPG_E_ARG;
    }
    if (count == 0) return BPG_OK;
    uint8_t* d_out = nullptr;
    CUDA_TRY(cudaMalloc((void**)&d_out, 32 * count));
    k_compress<<<(uint32_t)((count + 127) / 128), 128, 0, ctx->stream>>>(ctx->gens_ext + base, d_out, (uint32_t)count);
    ctx->launches++;
    CUDA_TRY(cudaMemcpyAsync(out, d_out, 32 * count, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx_sync(ctx));
    cudaFree(d_out);
    return BPG_OK;
}
