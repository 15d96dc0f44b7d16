// Multi-CTA exclusive prefix sum over uint32 (bucket offsets of the MSM sort, column offsets of the
// transposed constraint matrix).  Three small launches: per-tile scan, scan of tile totals, add-back.
#include "circuit.hpp"

#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total, uint32_t* sh /*[32]*/) {
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += t;
    }
    if (lane == 31) sh[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < (blockDim.x >> 5) ? sh[lane] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= (uint32_t)o) w += t;
        }
        sh[lane] = w;
    }
    __syncthreads();
    const uint32_t base = wid ? sh[wid - 1] : 0u;
    *total = sh[(blockDim.x >> 5) - 1];
    __syncthreads();
    return base + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                             uint32_t n, uint32_t* __restrict__ tile_sum) {
    __shared__ uint32_t sh[32];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = base + k < n ? in[base + k] : 0u;
        s += v[k];
    }
    uint32_t total;
    uint32_t pre = block_exclusive_scan(s, &total, sh);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) out[base + k] = pre;
        pre += v[k];
    }
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

// exclusive scan of the tile totals in place; tile_sum[ntiles] = grand total
__global__ void __launch_bounds__(1024) k_scan_top(uint32_t* __restrict__ tile_sum, uint32_t ntiles) {
    __shared__ uint32_t sh[32];
    uint32_t carry = 0;
    for (uint32_t b0 = 0; b0 < ntiles; b0 += 1024) {
        const uint32_t i = b0 + threadIdx.x;
        const uint32_t v = i < ntiles ? tile_sum[i] : 0u;
        uint32_t total;
        const uint32_t pre = block_exclusive_scan(v, &total, sh);
        if (i < ntiles) tile_sum[i] = carry + pre;
        carry += total;
    }
    if (threadIdx.x == 0) tile_sum[ntiles] = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(uint32_t* __restrict__ out, uint32_t n,
                                                           const uint32_t* __restrict__ tile_sum, uint32_t ntiles) {
    const uint32_t add = tile_sum[blockIdx.x];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
        if (base + k < n) out[base + k] += add;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = tile_sum[ntiles];
}

// ------------------------------------------------------------------------------------------
// MSM bucket offsets in ONE launch: a single 1024-thread CTA scans the histogram (G <= 65 536 in the hot path: two bucket
// sets of 2^15), writes the G+1 offsets, re-zeroes the histogram for the next MSM and derives the chunk geometry of the
// bucket-accumulation kernel from the entry count E (msm.cu).
// ------------------------------------------------------------------------------------------
// Warp w owns the contiguous range [w * span, (w + 1) * span), span = G / 32 rounded up to a multiple of 128, and walks it
// in steps of 128 counters: one coalesced uint4 load per lane and step, all steps in flight at once, a shuffle scan per
// step, then the 32 warp totals are scanned through shared memory.  G <= 32 768 is one pass (8 steps); larger G loops (two bucket sets: two passes).
#define SM_STEPS 8
__global__ void __launch_bounds__(1024) k_scan_meta(uint32_t* __restrict__ hist, uint32_t* __restrict__ off, uint32_t G,
                                                    uint32_t target_chunks, uint32_t cl_min, uint32_t cl_fixed,
                                                    MsmMeta* __restrict__ meta) {
    __shared__ uint32_t sh_tot[32];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool vec = (G % 4) == 0;
    uint32_t carry = 0;
    for (uint32_t tile = 0; tile < G; tile += 32u * 128u * SM_STEPS) {
        const uint32_t tile_n = min(G - tile, 32u * 128u * SM_STEPS);
        const uint32_t span = ((tile_n + 31) / 32 + 127) / 128 * 128;
        const uint32_t steps = span / 128;
        const uint32_t w0 = tile + wid * span;
        uint32_t v[SM_STEPS][4];
#pragma unroll
        for (int k = 0; k < SM_STEPS; k++) {
            const uint32_t idx = w0 + k * 128 + lane * 4;
            v[k][0] = v[k][1] = v[k][2] = v[k][3] = 0;
            if ((uint32_t)k < steps && idx < tile + tile_n) {
                if (vec) {
                    const uint4 q = *reinterpret_cast<const uint4*>(hist + idx);
                    v[k][0] = q.x, v[k][1] = q.y, v[k][2] = q.z, v[k][3] = q.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (idx + j < tile + tile_n) v[k][j] = hist[idx + j];
                }
            }
        }
        // exclusive prefix of every lane's group of four inside the warp's range
        uint32_t pre[SM_STEPS], run = 0;
#pragma unroll
        for (int k = 0; k < SM_STEPS; k++) {
            const uint32_t s = v[k][0] + v[k][1] + v[k][2] + v[k][3];
            uint32_t inc = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= (uint32_t)o) inc += t;
            }
            pre[k] = run + inc - s;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) sh_tot[wid] = run;
        __syncthreads();
        uint32_t wtot = sh_tot[lane], winc = wtot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= (uint32_t)o) winc += t;
        }
        const uint32_t wbase = carry + __shfl_sync(0xffffffffu, winc - wtot, wid);
        const uint32_t tile_total = __shfl_sync(0xffffffffu, winc, 31);
#pragma unroll
        for (int k = 0; k < SM_STEPS; k++) {
            const uint32_t idx = w0 + k * 128 + lane * 4;
            if ((uint32_t)k < steps && idx < tile + tile_n) {
                uint32_t p = wbase + pre[k];
                if (vec) {
                    uint4 q;
                    q.x = p, p += v[k][0];
                    q.y = p, p += v[k][1];
                    q.z = p, p += v[k][2];
                    q.w = p;
                    *reinterpret_cast<uint4*>(off + idx) = q;
                    *reinterpret_cast<uint4*>(hist + idx) = make_uint4(0u, 0u, 0u, 0u);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if (idx + j < tile + tile_n) {
                            off[idx + j] = p;
                            hist[idx + j] = 0u;
                        }
                        p += v[k][j];
                    }
                }
            }
        }
        carry += tile_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const uint32_t E = carry;
        off[G] = E;
        uint32_t cl = cl_fixed;
        if (cl == 0) {
            cl = (uint32_t)(((uint64_t)E + target_chunks - 1) / target_chunks);
            if (cl < cl_min) cl = cl_min;
        }
        meta->E = E;
        meta->CL = cl;
        meta->nchunks = (uint32_t)(((uint64_t)E + cl - 1) / cl);
        meta->pad = 0;
    }
}
void dev_scan_meta(cudaStream_t st, uint32_t* hist, uint32_t* off, uint32_t G, uint32_t target_chunks, uint32_t cl_min,
                   uint32_t cl_fixed, MsmMeta* meta) {
    k_scan_meta<<<1, 1024, 0, st>>>(hist, off, G, target_chunks, cl_min, cl_fixed, meta);
}

void dev_exclusive_scan_u32(cudaStream_t st, const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* scratch) {
    const uint32_t ntiles = n ? (n + SCAN_TILE - 1) / SCAN_TILE : 1;
    k_scan_tiles<<<ntiles, SCAN_THREADS, 0, st>>>(in, out, n, scratch);
    k_scan_top<<<1, 1024, 0, st>>>(scratch, ntiles);
    k_scan_add<<<ntiles, SCAN_THREADS, 0, st>>>(out, n, scratch, ntiles);
}
