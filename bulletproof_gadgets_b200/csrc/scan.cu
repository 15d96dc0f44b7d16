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
// Each thread owns PER consecutive counters (PER = G / 1024 rounded up to a multiple of 4, <= 64), so G <= 65 536 is ONE
// pass: all loads in flight at once, one block scan, stores.  Larger G loops over tiles of 65 536.
#define SM_MAXPER 64
template <int PER>
__device__ __forceinline__ uint32_t scan_tile(uint32_t* __restrict__ hist, uint32_t* __restrict__ off, uint32_t base, uint32_t G,
                                              uint32_t carry, uint32_t* sh) {
    const uint32_t idx = base + threadIdx.x * PER;
    uint32_t v[PER], s = 0;
    const bool vec = (G % 4) == 0;  // base and idx are multiples of 4: whole uint4 groups are inside or outside
    if (vec) {
#pragma unroll
        for (int k = 0; k < PER / 4; k++) {
            uint4 q = make_uint4(0u, 0u, 0u, 0u);
            if (idx + 4 * k < G) q = *reinterpret_cast<const uint4*>(hist + idx + 4 * k);
            v[4 * k] = q.x, v[4 * k + 1] = q.y, v[4 * k + 2] = q.z, v[4 * k + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < PER; k++) v[k] = idx + k < G ? hist[idx + k] : 0u;
    }
#pragma unroll
    for (int k = 0; k < PER; k++) s += v[k];
    uint32_t total;
    uint32_t pre = carry + block_exclusive_scan(s, &total, sh);
    if (vec) {
#pragma unroll
        for (int k = 0; k < PER / 4; k++) {
            uint4 q;
            q.x = pre, pre += v[4 * k];
            q.y = pre, pre += v[4 * k + 1];
            q.z = pre, pre += v[4 * k + 2];
            q.w = pre, pre += v[4 * k + 3];
            if (idx + 4 * k < G) {
                *reinterpret_cast<uint4*>(off + idx + 4 * k) = q;
                *reinterpret_cast<uint4*>(hist + idx + 4 * k) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < PER; k++) {
            if (idx + k < G) {
                off[idx + k] = pre;
                hist[idx + k] = 0u;
            }
            pre += v[k];
        }
    }
    return total;
}
__global__ void __launch_bounds__(1024) k_scan_meta(uint32_t* __restrict__ hist, uint32_t* __restrict__ off, uint32_t G,
                                                    uint32_t target_chunks, uint32_t cl_min, uint32_t cl_fixed,
                                                    MsmMeta* __restrict__ meta) {
    __shared__ uint32_t sh[32];
    uint32_t carry = 0;
    if (G <= 1024 * 4) carry = scan_tile<4>(hist, off, 0, G, 0, sh);
    else if (G <= 1024 * 16) carry = scan_tile<16>(hist, off, 0, G, 0, sh);
    else if (G <= 1024 * 32) carry = scan_tile<32>(hist, off, 0, G, 0, sh);
    else
        for (uint32_t base = 0; base < G; base += 1024 * SM_MAXPER) carry += scan_tile<SM_MAXPER>(hist, off, base, G, carry, sh);
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
