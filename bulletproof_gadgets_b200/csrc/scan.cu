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

void dev_exclusive_scan_u32(cudaStream_t st, const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* scratch) {
    const uint32_t ntiles = n ? (n + SCAN_TILE - 1) / SCAN_TILE : 1;
    k_scan_tiles<<<ntiles, SCAN_THREADS, 0, st>>>(in, out, n, scratch);
    k_scan_top<<<1, 1024, 0, st>>>(scratch, ntiles);
    k_scan_add<<<ntiles, SCAN_THREADS, 0, st>>>(out, n, scratch, ntiles);
}
