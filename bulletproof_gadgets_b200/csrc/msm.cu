// Fixed-base Pippenger multiscalar multiplication over Edwards25519 for sm_100a.
//
// Replaces curve25519-dalek 3.2.0 Straus `multiscalar_mul` and Pippenger/Straus
// `vartime_multiscalar_mul` / `optional_multiscalar_mul` (backend/serial/scalar_mul/{straus,
// pippenger}.rs; /root/reference/Cargo.lock:155-157, not vendored) as they are used by
// bulletproofs for A_I1/A_O1/S1, the IPP L/R points and the verifier's mega-MSM -- reached from
// /root/reference/src/prove.rs:79 and /root/reference/src/verify.rs:71.  SURVEY.md rows K3/K4.
// Any correct MSM yields the same group element, hence the same 32 ristretto bytes.
//
// Every base point of the hot path is a fixed generator (G_i, H_i, B, B_blinding), so the table
// holds, per point, the affine Niels form of 2^(c*w) * P for every window w.  All windows then
// share ONE bucket set: no per-window doubling chain (a serial latency tail on a GPU), 16x
// fuller buckets (better balance), and the bucket reduction is paid once.
//
// Stages (all on ctx->stream, no host round trip until the final 128-byte result):
//   1 digits+histogram   signed c-bit digits, one global-atomic histogram per bucket set  [HBM/L2]
//   2 scan               bucket offsets + balanced task split (<= task_len entries each)
//   3 digits+scatter     counting-sort of (row | sign) entries by bucket                  [HBM/L2]
//   4 tasks              task descriptors
//   5 accumulate         one thread per task: mixed adds of gathered Niels rows           [IMAD]
//   6 reduce             weighted running sums over task partials, block tree, final tree
#include "ctx.hpp"

#define REDUCE_BLOCKS 128
#define REDUCE_THREADS 64
#define ACC_THREADS 128

// ------------------------------------------------------------------------------------------
// stage 1 / 3: signed digit decomposition
// ------------------------------------------------------------------------------------------
template <bool SCATTER>
__global__ void __launch_bounds__(256) k_digits(MsmSegments segs, int c, int K, uint32_t nb, uint32_t n_points,
                                                uint32_t* __restrict__ hist, const uint32_t* __restrict__ bucket_off,
                                                uint32_t* __restrict__ entries) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= segs.total) return;
    // locate the segment
    uint32_t si = 0, base = 0, end = 0;
#pragma unroll
    for (int k = 0; k < MSM_MAX_SEGMENTS; k++) {
        if (k < (int)segs.nseg) {
            end += segs.seg[k].count;
            if (g >= end) {  // ends are non-decreasing, so this is true for a prefix of k only
                base = end;
                si = k + 1;
            }
        }
    }
    const MsmSegment sg = segs.seg[si];
    const uint32_t i = g - base;
    const uint4* sp = reinterpret_cast<const uint4*>(sg.scalars) + 2 * (size_t)i;
    uint4 lo = __ldg(sp), hi = __ldg(sp + 1);
    uint32_t s[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    if ((s[0] | s[1] | s[2] | s[3] | s[4] | s[5] | s[6] | s[7]) == 0) return;
    uint32_t set = sg.set_id;
    if (sg.mode == 1) set += ((i % sg.period) >= (sg.period >> 1)) ? 0u : 1u;
    if (sg.mode == 2) set += ((i % sg.period) < (sg.period >> 1)) ? 0u : 1u;
    const uint32_t point = sg.point_base + i;
    const uint32_t mask = (1u << c) - 1u, half = 1u << (c - 1);
    uint32_t carry = 0;
    for (int w = 0; w < K; w++) {
        uint32_t raw = (s[0] & mask) + carry;
#pragma unroll
        for (int k = 0; k < 7; k++) s[k] = __funnelshift_r(s[k], s[k + 1], c);
        s[7] >>= c;
        uint32_t neg = raw > half;
        uint32_t mag = neg ? ((1u << c) - raw) : raw;
        carry = neg;
        if (mag != 0) {
            uint32_t gb = set * nb + (mag - 1);
            if (!SCATTER) {
                atomicAdd(&hist[gb], 1u);
            } else {
                uint32_t pos = atomicAdd(&hist[gb], 1u);
                entries[bucket_off[gb] + pos] = ((uint32_t)w * n_points + point) | (neg << 31);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// stage 2: exclusive scans (entries, tasks) over all buckets of all sets.  The two sums ride one
// 64-bit word (entries low, tasks high; both < 2^32), tiles of 2048 buckets, three small launches.
// ------------------------------------------------------------------------------------------
#define BS_THREADS 256
#define BS_ITEMS 8
#define BS_TILE (BS_THREADS * BS_ITEMS)
__device__ __forceinline__ uint64_t bs_block_exclusive(uint64_t v, uint64_t* total, uint64_t* sh /*[32]*/) {
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint64_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += t;
    }
    if (lane == 31) sh[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint64_t w = lane < nw ? sh[lane] : 0ull;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= (uint32_t)o) w += t;
        }
        sh[lane] = w;
    }
    __syncthreads();
    const uint64_t base = wid ? sh[wid - 1] : 0ull;
    *total = sh[nw - 1];
    __syncthreads();
    return base + inc - v;
}
__global__ void __launch_bounds__(BS_THREADS) k_bscan_tiles(const uint32_t* __restrict__ hist, uint32_t* __restrict__ bucket_off,
                                                            uint32_t* __restrict__ task_off, uint64_t* __restrict__ tile_sum,
                                                            uint32_t G, uint32_t task_len) {
    __shared__ uint64_t sh[32];
    const uint32_t base = blockIdx.x * BS_TILE + threadIdx.x * BS_ITEMS;
    uint64_t v[BS_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < BS_ITEMS; k++) {
        const uint32_t cnt = base + k < G ? hist[base + k] : 0u;
        v[k] = (uint64_t)cnt | ((uint64_t)((cnt + task_len - 1) / task_len) << 32);
        s += v[k];
    }
    uint64_t total;
    uint64_t pre = bs_block_exclusive(s, &total, sh);
#pragma unroll
    for (int k = 0; k < BS_ITEMS; k++) {
        if (base + k < G) {
            bucket_off[base + k] = (uint32_t)pre;
            task_off[base + k] = (uint32_t)(pre >> 32);
        }
        pre += v[k];
    }
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}
__global__ void __launch_bounds__(1024) k_bscan_top(uint64_t* __restrict__ tile_sum, uint32_t ntiles) {
    __shared__ uint64_t sh[32];
    uint64_t carry = 0;
    for (uint32_t b0 = 0; b0 < ntiles; b0 += 1024) {
        const uint32_t i = b0 + threadIdx.x;
        const uint64_t v = i < ntiles ? tile_sum[i] : 0ull;
        uint64_t total;
        const uint64_t pre = bs_block_exclusive(v, &total, sh);
        if (i < ntiles) tile_sum[i] = carry + pre;
        carry += total;
    }
    if (threadIdx.x == 0) tile_sum[ntiles] = carry;
}
// adds the tile offsets, clears the histogram (it becomes the scatter cursor) and writes
// meta[0] = #tasks, meta[1+s] = first task of set s, meta[1+nsets] = #tasks
__global__ void __launch_bounds__(BS_THREADS) k_bscan_finish(uint32_t* __restrict__ hist, uint32_t* __restrict__ bucket_off,
                                                             uint32_t* __restrict__ task_off, const uint64_t* __restrict__ tile_sum,
                                                             uint32_t* __restrict__ meta, uint32_t G, uint32_t nb, uint32_t nsets,
                                                             uint32_t ntiles) {
    const uint64_t add = tile_sum[blockIdx.x];
    const uint32_t base = blockIdx.x * BS_TILE + threadIdx.x * BS_ITEMS;
#pragma unroll
    for (int k = 0; k < BS_ITEMS; k++) {
        const uint32_t b = base + k;
        if (b < G) {
            const uint32_t to = task_off[b] + (uint32_t)(add >> 32);
            bucket_off[b] += (uint32_t)add;
            task_off[b] = to;
            hist[b] = 0;
            if (b % nb == 0) meta[1 + b / nb] = to;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const uint64_t tot = tile_sum[ntiles];
        bucket_off[G] = (uint32_t)tot;
        task_off[G] = (uint32_t)(tot >> 32);
        meta[0] = (uint32_t)(tot >> 32);
        meta[1 + nsets] = (uint32_t)(tot >> 32);
    }
}

// ------------------------------------------------------------------------------------------
// stage 4: task descriptors (balanced split of each bucket)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_tasks(const uint32_t* __restrict__ bucket_off,
                                               const uint32_t* __restrict__ task_off, uint2* __restrict__ tasks,
                                               uint32_t G) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= G) return;
    const uint32_t off = bucket_off[b], cnt = bucket_off[b + 1] - off;
    const uint32_t t0 = task_off[b], nt = task_off[b + 1] - t0;
    for (uint32_t k = 0; k < nt; k++) {
        uint32_t a = (uint32_t)(((uint64_t)k * cnt) / nt), e = (uint32_t)(((uint64_t)(k + 1) * cnt) / nt);
        tasks[t0 + k] = make_uint2(off + a, (b << 8) | (e - a));
    }
}

// ------------------------------------------------------------------------------------------
// stage 5: bucket accumulation -- the IMAD-bound kernel
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ ge_niels load_niels(const ge_niels* __restrict__ rows, uint32_t row) {
    const uint4* p = reinterpret_cast<const uint4*>(rows + row);
    uint4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3), e = __ldg(p + 4), f = __ldg(p + 5);
    ge_niels q;
    q.yp.v[0] = a.x, q.yp.v[1] = a.y, q.yp.v[2] = a.z, q.yp.v[3] = a.w;
    q.yp.v[4] = b.x, q.yp.v[5] = b.y, q.yp.v[6] = b.z, q.yp.v[7] = b.w;
    q.ym.v[0] = c.x, q.ym.v[1] = c.y, q.ym.v[2] = c.z, q.ym.v[3] = c.w;
    q.ym.v[4] = d.x, q.ym.v[5] = d.y, q.ym.v[6] = d.z, q.ym.v[7] = d.w;
    q.t2d.v[0] = e.x, q.t2d.v[1] = e.y, q.t2d.v[2] = e.z, q.t2d.v[3] = e.w;
    q.t2d.v[4] = f.x, q.t2d.v[5] = f.y, q.t2d.v[6] = f.z, q.t2d.v[7] = f.w;
    return q;
}
__device__ __forceinline__ void store_ext(ge_ext* dst, const ge_ext& p) {
    uint4* o = reinterpret_cast<uint4*>(dst);
    const fe* f[4] = {&p.X, &p.Y, &p.Z, &p.T};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        o[2 * k] = make_uint4(f[k]->v[0], f[k]->v[1], f[k]->v[2], f[k]->v[3]);
        o[2 * k + 1] = make_uint4(f[k]->v[4], f[k]->v[5], f[k]->v[6], f[k]->v[7]);
    }
}
__device__ __forceinline__ ge_ext load_ext(const ge_ext* src) {
    const uint4* o = reinterpret_cast<const uint4*>(src);
    ge_ext p;
    fe* f[4] = {&p.X, &p.Y, &p.Z, &p.T};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint4 a = o[2 * k], b = o[2 * k + 1];
        f[k]->v[0] = a.x, f[k]->v[1] = a.y, f[k]->v[2] = a.z, f[k]->v[3] = a.w;
        f[k]->v[4] = b.x, f[k]->v[5] = b.y, f[k]->v[6] = b.z, f[k]->v[7] = b.w;
    }
    return p;
}

__global__ void __launch_bounds__(ACC_THREADS, 4)
    k_accumulate(const ge_niels* __restrict__ rows, const uint32_t* __restrict__ entries,
                 const uint2* __restrict__ tasks, const uint32_t* __restrict__ meta, ge_ext* __restrict__ partials) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= meta[0]) return;
    const uint2 td = tasks[t];
    const uint32_t cnt = td.y & 0xffu;
    const uint32_t* ep = entries + td.x;
    ge_ext acc = ge_identity();
    uint32_t ent = __ldg(ep);
#pragma unroll 1
    for (uint32_t e = 0; e < cnt; e++) {
        ge_niels q = load_niels(rows, ent & 0x7fffffffu);
        const bool neg = ent >> 31;
        if (e + 1 < cnt) ent = __ldg(ep + e + 1);
        acc = ge_madd(acc, q, neg);
    }
    store_ext(partials + t, acc);
}

// ------------------------------------------------------------------------------------------
// stage 6: sum_t weight(t) * partial(t), weight = bucket index + 1
// ------------------------------------------------------------------------------------------
__device__ ge_ext ge_mul_small(const ge_ext& p, uint32_t k) {
    if (k == 0) return ge_identity();
    if (k == 1) return p;
    int top = 31 - __clz(k);
    ge_ext acc = p;
#pragma unroll 1
    for (int b = top - 1; b >= 0; b--) {
        acc = ge_dbl(acc);
        if ((k >> b) & 1u) acc = ge_add(acc, p);
    }
    return acc;
}

__device__ __forceinline__ void block_tree_reduce(ge_ext* sh, ge_ext& mine, uint32_t tid, uint32_t n) {
    store_ext(sh + tid, mine);
    __syncthreads();
    for (uint32_t s = n >> 1; s > 0; s >>= 1) {
        if (tid < s) {
            ge_ext a = load_ext(sh + tid), b = load_ext(sh + tid + s);
            a = ge_add(a, b);
            store_ext(sh + tid, a);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(REDUCE_THREADS)
    k_reduce_chunks(const ge_ext* __restrict__ partials, const uint2* __restrict__ tasks,
                    const uint32_t* __restrict__ meta, uint32_t nb, ge_ext* __restrict__ blockres) {
    __shared__ ge_ext sh[REDUCE_THREADS];
    const uint32_t s = blockIdx.y;
    const uint32_t tb = meta[1 + s], te = meta[2 + s];
    const uint32_t nt = te - tb, nchunks = REDUCE_BLOCKS * REDUCE_THREADS;
    const uint32_t per = (nt + nchunks - 1) / nchunks;
    const uint32_t cidx = blockIdx.x * REDUCE_THREADS + threadIdx.x;
    const uint64_t a64 = (uint64_t)tb + (uint64_t)cidx * per;
    ge_ext total = ge_identity();
    if (per > 0 && a64 < te) {
        const uint32_t a = (uint32_t)a64, b = min(a + per, te);
        const uint32_t wbase = s * nb;
        ge_ext R = ge_identity(), S = ge_identity();
        uint32_t wprev = (tasks[b - 1].y >> 8) - wbase + 1;
#pragma unroll 1
        for (uint32_t t = b; t-- > a;) {
            uint32_t w = (tasks[t].y >> 8) - wbase + 1;
            if (w != wprev) {
                uint32_t gap = wprev - w;
                S = ge_add(S, gap == 1 ? R : ge_mul_small(R, gap));
                wprev = w;
            }
            R = ge_add(R, load_ext(partials + t));
        }
        total = ge_add(S, ge_mul_small(R, wprev));
    }
    block_tree_reduce(sh, total, threadIdx.x, REDUCE_THREADS);
    if (threadIdx.x == 0) store_ext(blockres + s * REDUCE_BLOCKS + blockIdx.x, load_ext(sh));
}

__global__ void __launch_bounds__(REDUCE_BLOCKS) k_reduce_final(const ge_ext* __restrict__ blockres,
                                                                  ge_ext* __restrict__ result) {
    __shared__ ge_ext sh[REDUCE_BLOCKS];
    const uint32_t s = blockIdx.x;
    ge_ext mine = load_ext(blockres + s * REDUCE_BLOCKS + threadIdx.x);
    block_tree_reduce(sh, mine, threadIdx.x, REDUCE_BLOCKS);
    if (threadIdx.x == 0) store_ext(result + s, load_ext(sh));
}

// ------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------
int msm_run(bpg_ctx* ctx, const MsmSegments& segs, uint32_t nsets, ge_ext* d_out) {
    const FixedTable& tb = ctx->table;
    if (!tb.rows) {
        bpg_set_error("msm_run: generator table not built");
        return BPG_E_ARG;
    }
    if (nsets == 0 || segs.nseg > MSM_MAX_SEGMENTS) return BPG_E_ARG;
    cudaStream_t st = ctx->stream;
    const uint32_t nb = 1u << (tb.c - 1);
    const uint32_t G = nsets * nb;
    if ((uint64_t)G >= (1ull << 24)) {
        bpg_set_error("msm_run: too many buckets");
        return BPG_E_ARG;
    }
    const uint64_t total = segs.total;
    const uint64_t max_entries = (uint64_t)tb.K * total;
    const uint32_t T = (uint32_t)ctx->task_len;
    const uint64_t max_tasks = max_entries / T + G + 1;
    if (max_entries >= (1ull << 32) || (uint64_t)tb.K * tb.n_points >= (1ull << 31)) {
        bpg_set_error("msm_run: problem too large for 32-bit entry indices");
        return BPG_E_ARG;
    }
    MsmWork& w = ctx->work;
    int rc;
    if ((rc = w.hist.ensure(G)) || (rc = w.bucket_off.ensure(G + 1)) || (rc = w.task_off.ensure(G + 1)) ||
        (rc = w.entries.ensure(max_entries + 1)) || (rc = w.tasks.ensure(max_tasks)) ||
        (rc = w.partials.ensure(max_tasks)) || (rc = w.blockres.ensure((size_t)nsets * REDUCE_BLOCKS)) ||
        (rc = w.meta.ensure(nsets + 3)) || (rc = w.tile_sum.ensure(G / BS_TILE + 3)))
        return rc;

    CUDA_TRY(cudaMemsetAsync(w.hist.p, 0, (size_t)G * 4, st));
    if (total > 0) {
        const uint32_t blocks = (uint32_t)((total + 255) / 256);
        k_digits<false><<<blocks, 256, 0, st>>>(segs, tb.c, tb.K, nb, tb.n_points, w.hist.p, nullptr, nullptr);
        ctx->launches++;
    }
    {
        const uint32_t ntiles = (G + BS_TILE - 1) / BS_TILE;
        k_bscan_tiles<<<ntiles, BS_THREADS, 0, st>>>(w.hist.p, w.bucket_off.p, w.task_off.p, w.tile_sum.p, G, T);
        k_bscan_top<<<1, 1024, 0, st>>>(w.tile_sum.p, ntiles);
        k_bscan_finish<<<ntiles, BS_THREADS, 0, st>>>(w.hist.p, w.bucket_off.p, w.task_off.p, w.tile_sum.p, w.meta.p, G, nb,
                                                      nsets, ntiles);
        ctx->launches += 3;
    }
    if (total > 0) {
        const uint32_t blocks = (uint32_t)((total + 255) / 256);
        k_digits<true><<<blocks, 256, 0, st>>>(segs, tb.c, tb.K, nb, tb.n_points, w.hist.p, w.bucket_off.p,
                                              w.entries.p);
        ctx->launches++;
    }
    k_tasks<<<(G + 255) / 256, 256, 0, st>>>(w.bucket_off.p, w.task_off.p, w.tasks.p, G);
    ctx->launches++;
    {
        const uint32_t blocks = (uint32_t)((max_tasks + ACC_THREADS - 1) / ACC_THREADS);
        if (ctx->time_accum) CUDA_TRY(cudaEventRecord(ctx->ev_a, st));
        k_accumulate<<<blocks, ACC_THREADS, 0, st>>>(tb.rows, w.entries.p, w.tasks.p, w.meta.p, w.partials.p);
        if (ctx->time_accum) CUDA_TRY(cudaEventRecord(ctx->ev_b, st));
        ctx->launches++;
    }
    k_reduce_chunks<<<dim3(REDUCE_BLOCKS, nsets), REDUCE_THREADS, 0, st>>>(w.partials.p, w.tasks.p, w.meta.p, nb,
                                                                          w.blockres.p);
    k_reduce_final<<<nsets, REDUCE_BLOCKS, 0, st>>>(w.blockres.p, d_out);
    ctx->launches += 2;
    CUDA_TRY(cudaGetLastError());
    if (ctx->time_accum) {  // diagnostic mode: synchronous, reads back the entry count
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaEventElapsedTime(&ctx->last_accum_ms, ctx->ev_a, ctx->ev_b));
        uint32_t ne = 0;
        CUDA_TRY(cudaMemcpy(&ne, w.bucket_off.p + G, 4, cudaMemcpyDeviceToHost));
        ctx->last_entries = ne;
        ctx->sum_accum_ms += ctx->last_accum_ms;
        ctx->sum_entries += ne;
    }
    return BPG_OK;
}
