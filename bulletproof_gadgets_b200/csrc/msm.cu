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
//   2 scan               bucket offsets + chunk geometry from the entry count, one CTA (scan.cu)
//   3 digits+scatter     counting-sort of (row | sign) entries by bucket                  [HBM/L2]
//   4 accumulate         one thread per chunk of exactly CL sorted entries: mixed adds of gathered Niels rows; a
//                        chunk that crosses bucket boundaries emits one partial per bucket, so every lane runs
//                        the same number of additions (no divergence).  CL is chosen on the device so that the
//                        chunks fill exactly one wave of resident CTAs (no tail)           [IMAD]
//   5 bucket reduce      thread per bucket: sum of its partial slots; then a butterfly over the bucket INDEX bits
//                        keeps, per segment, the plain sum R and the bit marginals M_j = sum of the buckets whose
//                        index has bit j set -- one point addition per lane and level, no scalar multiplications
//   6 final              remaining butterfly levels across CTAs, then sum (b+1) S_b = R + sum_j 2^j M_j
#include <stdlib.h>

#include <cooperative_groups.h>

#include <algorithm>
#include <vector>

#include "circuit.hpp"
#include "ctx.hpp"

#define BR_THREADS 128  // buckets per CTA of k_bucket_reduce
#define BR_CAP 8       // partial slots a lane sums alone before the warp shares the rest of a heavy bucket
#define SMALL_MSM_MAX_POINTS 4096u  // MSMs up to 2 x this many points use the 8-bit-window table
#define ACC_THREADS 128

// ------------------------------------------------------------------------------------------
// stage 1 / 3: signed digit decomposition
// ------------------------------------------------------------------------------------------
#define DG_NONE 0xffffffffu
template <bool SCATTER>
__global__ void __launch_bounds__(256) k_digits(MsmSegments segs, int c, int K, uint32_t nb, uint32_t n_points,
                                                uint32_t* __restrict__ hist, const uint32_t* __restrict__ bucket_off,
                                                uint32_t* __restrict__ entries, uint32_t* __restrict__ tickets) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = g < segs.total;  // out-of-range lanes stay: the histogram pass uses warp-wide votes
    const uint32_t gi = in ? g : 0u;
    // locate the segment
    uint32_t si = 0, base = 0, end = 0;
#pragma unroll
    for (int k = 0; k < MSM_MAX_SEGMENTS; k++) {
        if (k < (int)segs.nseg) {
            end += segs.seg[k].count;
            if (gi >= end) {  // ends are non-decreasing, so this is true for a prefix of k only
                base = end;
                si = k + 1;
            }
        }
    }
    const MsmSegment sg = segs.seg[si];
    const uint32_t i = gi - base;
    uint32_t s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (in) {
        const uint4* sp = reinterpret_cast<const uint4*>(sg.scalars) + 2 * (size_t)i;
        const uint4 lo = __ldg(sp), hi = __ldg(sp + 1);
        s[0] = lo.x, s[1] = lo.y, s[2] = lo.z, s[3] = lo.w, s[4] = hi.x, s[5] = hi.y, s[6] = hi.z, s[7] = hi.w;
    }
    const bool nz = (s[0] | s[1] | s[2] | s[3] | s[4] | s[5] | s[6] | s[7]) != 0;  // zero scalars contribute nothing
    uint32_t set = sg.set_id;
    if (sg.mode == 1) set += ((i % sg.period) >= (sg.period >> 1)) ? 0u : 1u;
    if (sg.mode == 2) set += ((i % sg.period) < (sg.period >> 1)) ? 0u : 1u;
    if (sg.mode == 3) set += i % sg.period;
    const uint32_t point = sg.point_base + i;
    const uint32_t row_stride = segs.var_base ? 0u : n_points, set_stride = segs.var_base ? 1u : 0u;
    const uint32_t mask = (1u << c) - 1u, half = 1u << (c - 1);
    uint32_t carry = 0;
    if (K <= 16) {
        // all digits first, then all atomics back to back (16 independent L2 round trips in flight instead of a
        // dependent chain), then the scattered stores
        uint32_t gbv[16], entv[16];
#pragma unroll
        for (int w = 0; w < 16; w++) {
            gbv[w] = DG_NONE;
            entv[w] = 0;
            if (w < K) {
                uint32_t raw = (s[0] & mask) + carry;
#pragma unroll
                for (int k = 0; k < 7; k++) s[k] = __funnelshift_r(s[k], s[k + 1], c);
                s[7] >>= c;
                const uint32_t neg = raw > half;
                const uint32_t mag = neg ? ((1u << c) - raw) : raw;
                carry = neg;
                if (mag != 0) {  // implies nz
                    gbv[w] = (set + (uint32_t)w * set_stride) * nb + (mag - 1);
                    entv[w] = ((uint32_t)w * row_stride + point) | (neg << 31);
                }
            }
        }
        // With `tickets` the histogram pass keeps what its atomicAdd returns -- the rank of the entry inside its bucket
        // -- at [w][g] (coalesced), and the scatter pass is bucket_off + rank: no second round of atomics.
        if (!SCATTER) {
            if (tickets) {
                const uint32_t lane = threadIdx.x & 31, lt = (1u << lane) - 1u;
                uint32_t tk[16];
                // Structured scalars (a_L in {0,1}, a_R in {0,-1}: every entry of a window falls into ONE bucket) would
                // serialise on a single L2 address.  Windows in which neighbouring lanes collide are found first (one vote
                // per window); the common case -- none -- keeps the 16 independent atomics of a thread in flight together.
                uint32_t dup = 0;
#pragma unroll
                for (int w = 0; w < 16; w++) {
                    const uint32_t nxt = __shfl_down_sync(0xffffffffu, gbv[w], 1);
                    if (__any_sync(0xffffffffu, gbv[w] != DG_NONE && gbv[w] == nxt && lane < 31)) dup |= 1u << w;
                }
                if (dup == 0) {
#pragma unroll
                    for (int w = 0; w < 16; w++)
                        if (gbv[w] != DG_NONE) tk[w] = atomicAdd(&hist[gbv[w]], 1u);
                } else {
                    // the lanes of a warp that share a bucket send one atomic for all of them and split the returned
                    // range by lane order
#pragma unroll
                    for (int w = 0; w < 16; w++) {
                        const uint32_t key = gbv[w];
                        if ((dup >> w) & 1u) {
                            const uint32_t peers = __match_any_sync(0xffffffffu, key);
                            const int leader = __ffs(peers) - 1;
                            uint32_t b0 = 0;
                            if (key != DG_NONE && (int)lane == leader) b0 = atomicAdd(&hist[key], (uint32_t)__popc(peers));
                            b0 = __shfl_sync(0xffffffffu, b0, leader);
                            tk[w] = b0 + __popc(peers & lt);
                        } else if (key != DG_NONE) {
                            tk[w] = atomicAdd(&hist[key], 1u);
                        }
                    }
                }
                if (in) {
#pragma unroll
                    for (int w = 0; w < 16; w++)
                        if (gbv[w] != DG_NONE) tickets[(size_t)w * segs.total + g] = tk[w];
                }
            } else {
#pragma unroll
                for (int w = 0; w < 16; w++)
                    if (gbv[w] != DG_NONE) atomicAdd(&hist[gbv[w]], 1u);
            }
        } else {
            uint32_t pos[16];
            if (tickets) {
#pragma unroll
                for (int w = 0; w < 16; w++)
                    if (gbv[w] != DG_NONE) pos[w] = bucket_off[gbv[w]] + tickets[(size_t)w * segs.total + g];
            } else {
#pragma unroll
                for (int w = 0; w < 16; w++)
                    if (gbv[w] != DG_NONE) pos[w] = bucket_off[gbv[w]] + atomicAdd(&hist[gbv[w]], 1u);
            }
#pragma unroll
            for (int w = 0; w < 16; w++)
                if (gbv[w] != DG_NONE) entries[pos[w]] = entv[w];
        }
        return;
    }
    if (!nz) return;
    for (int w = 0; w < K; w++) {
        uint32_t raw = (s[0] & mask) + carry;
#pragma unroll
        for (int k = 0; k < 7; k++) s[k] = __funnelshift_r(s[k], s[k + 1], c);
        s[7] >>= c;
        uint32_t neg = raw > half;
        uint32_t mag = neg ? ((1u << c) - raw) : raw;
        carry = neg;
        if (mag != 0) {
            uint32_t gb = (set + (uint32_t)w * set_stride) * nb + (mag - 1);
            if (!SCATTER) {
                atomicAdd(&hist[gb], 1u);
            } else {
                uint32_t pos = atomicAdd(&hist[gb], 1u);
                entries[bucket_off[gb] + pos] = ((uint32_t)w * row_stride + point) | (neg << 31);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// stages 1-3, shared-memory form (two-level counting sort; used when a bucket set has >= 256 buckets and at most 65 536
// buckets exist in all): no global atomics and no scattered 4-byte stores.
//   k_sort_count    CTA = 512 scalars: digits, histogram of the COARSE bin (bucket >> 8) in shared memory,
//                   one column of the [bin][CTA] count matrix per CTA
//   (scan)          exclusive scan of the matrix in bin-major order = where each CTA's share of each bin starts;
//                   the same launch derives the chunk geometry from the entry count (scan.cu: k_scan_meta)
//   k_sort_scatter  digits again; the CTA's <= 8192 entries are counting-sorted by bin in shared memory and copied out
//                   as (bucket, row|sign) pairs: every bin run of a CTA is one contiguous, coalesced store
//   k_sort_bins     one thread-block CLUSTER per bin (its entries are contiguous now): the CTAs count the 256 fine buckets of
//                   their slices in shared memory, exchange the counts through distributed shared memory, write the
//                   bucket offsets (= bucket_off, no global scan) and store their slices sorted by fine bucket
// ------------------------------------------------------------------------------------------
#define SORT_PTS 512          // scalars per CTA of k_sort_count / k_sort_scatter (2 per thread)

// digits of scalar g -> up to 16 (bucket, row|sign) pairs; bucket DG_NONE = no entry.  K <= 16.
__device__ __forceinline__ void decode_scalar(const MsmSegments& segs, uint32_t g, int c, int K, uint32_t nb, uint32_t n_points,
                                              uint32_t gbv[16], uint32_t entv[16]) {
#pragma unroll
    for (int w = 0; w < 16; w++) gbv[w] = DG_NONE, entv[w] = 0;
    if (g >= segs.total) return;
    uint32_t si = 0, base = 0, end = 0;
#pragma unroll
    for (int k = 0; k < MSM_MAX_SEGMENTS; k++) {
        if (k < (int)segs.nseg) {
            end += segs.seg[k].count;
            if (g >= end) {
                base = end;
                si = k + 1;
            }
        }
    }
    const MsmSegment sg = segs.seg[si];
    const uint32_t i = g - base;
    const uint4* sp = reinterpret_cast<const uint4*>(sg.scalars) + 2 * (size_t)i;
    const uint4 lo = __ldg(sp), hi = __ldg(sp + 1);
    uint32_t s[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    uint32_t set = sg.set_id;
    if (sg.mode == 1) set += ((i % sg.period) >= (sg.period >> 1)) ? 0u : 1u;
    if (sg.mode == 2) set += ((i % sg.period) < (sg.period >> 1)) ? 0u : 1u;
    if (sg.mode == 3) set += i % sg.period;
    const uint32_t point = sg.point_base + i;
    const uint32_t row_stride = segs.var_base ? 0u : n_points, set_stride = segs.var_base ? 1u : 0u;
    const uint32_t mask = (1u << c) - 1u, half = 1u << (c - 1);
    uint32_t carry = 0;
    if (c == 16) {  // the big table's window: digit w is a half word of limb w / 2 -- no shifting of the whole scalar per window
#pragma unroll
        for (int w = 0; w < 16; w++) {
            if (w < K) {
                const uint32_t raw = ((w & 1) ? (s[w >> 1] >> 16) : (s[w >> 1] & 0xffffu)) + carry;
                const uint32_t neg = raw > 0x8000u;
                const uint32_t mag = neg ? (0x10000u - raw) : raw;
                carry = neg;
                if (mag != 0) {
                    gbv[w] = (set + (uint32_t)w * set_stride) * nb + (mag - 1);
                    entv[w] = ((uint32_t)w * row_stride + point) | (neg << 31);
                }
            }
        }
        return;
    }
#pragma unroll
    for (int w = 0; w < 16; w++) {
        if (w < K) {
            const uint32_t raw = (s[0] & mask) + carry;
#pragma unroll
            for (int k = 0; k < 7; k++) s[k] = __funnelshift_r(s[k], s[k + 1], c);
            s[7] >>= c;
            const uint32_t neg = raw > half;
            const uint32_t mag = neg ? ((1u << c) - raw) : raw;
            carry = neg;
            if (mag != 0) {
                gbv[w] = (set + (uint32_t)w * set_stride) * nb + (mag - 1);
                entv[w] = ((uint32_t)w * row_stride + point) | (neg << 31);
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_sort_count(MsmSegments segs, int c, int K, uint32_t nb, uint32_t n_points,
                                                    uint32_t nbins, uint32_t* __restrict__ cnt /* [nbins][gridDim.x] */) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll 1
    for (int r = 0; r < SORT_PTS / 256; r++) {
        uint32_t gbv[16], entv[16];
        decode_scalar(segs, blockIdx.x * SORT_PTS + r * 256 + threadIdx.x, c, K, nb, n_points, gbv, entv);
#pragma unroll
        for (int w = 0; w < 16; w++)
            if (gbv[w] != DG_NONE) atomicAdd(&h[gbv[w] >> 8], 1u);
    }
    __syncthreads();
    if (threadIdx.x < nbins) cnt[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];
}

__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* wt /* [8] */) {  // threads 0..255, 256 values
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += t;
    }
    if (lane == 31) wt[wid] = inc;
    __syncthreads();
    uint32_t add = 0;
    for (uint32_t k = 0; k < wid; k++) add += wt[k];
    __syncthreads();
    return add + inc - v;
}
// Offsets of the [bin][CTA] count matrix without a global scan: CTA b scans row b (where each scatter CTA's share of bin b
// starts INSIDE the bin) and publishes the bin total; the CTA that arrives last scans the <= 256 totals into bin_start[]
// and derives the chunk geometry from the entry count.  Consumers add bin_start[bin] + pre[bin][cta].
__global__ void __launch_bounds__(256) k_sort_scan(const uint32_t* __restrict__ cnt, uint32_t nctas, uint32_t nbins,
                                                   uint32_t* __restrict__ pre, uint32_t* __restrict__ tot /* [nbins] */,
                                                   uint32_t* __restrict__ bin_start /* [nbins + 1] */, uint32_t* __restrict__ arrive,
                                                   uint32_t target_chunks, uint32_t cl_min, uint32_t cl_fixed, MsmMeta* __restrict__ meta) {
    __shared__ uint32_t wt[8];
    __shared__ uint32_t sh_last;
    const uint32_t tid = threadIdx.x, b = blockIdx.x;
    uint32_t carry = 0;
    for (uint32_t t0 = 0; t0 < nctas; t0 += 256) {
        const uint32_t i = t0 + tid, v = i < nctas ? cnt[(size_t)b * nctas + i] : 0u;
        const uint32_t ex = block_excl_scan_256(v, wt);
        if (i < nctas) pre[(size_t)b * nctas + i] = carry + ex;
        __shared__ uint32_t tile_total;
        if (tid == 255) tile_total = ex + v;
        __syncthreads();
        carry += tile_total;
        __syncthreads();
    }
    if (tid == 0) {
        tot[b] = carry;
        __threadfence();
        sh_last = atomicAdd(arrive, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!sh_last) return;
    __threadfence();
    const uint32_t v = tid < nbins ? __ldcg(tot + tid) : 0u;
    const uint32_t ex = block_excl_scan_256(v, wt);
    if (tid < nbins) bin_start[tid] = ex;
    if (tid == 255) {
        const uint32_t E = ex + v;
        bin_start[nbins] = E;
        uint32_t cl = cl_fixed;
        if (cl == 0) {
            cl = (uint32_t)(((uint64_t)E + target_chunks - 1) / target_chunks);
            if (cl < cl_min) cl = cl_min;
        }
        meta->E = E;
        meta->CL = cl;
        meta->nchunks = (uint32_t)(((uint64_t)E + cl - 1) / cl);
        meta->pad = 0;
        *arrive = 0;
    }
}

__global__ void __launch_bounds__(256) k_sort_scatter(MsmSegments segs, int c, int K, uint32_t nb, uint32_t n_points,
                                                      uint32_t nbins, const uint32_t* __restrict__ pre, const uint32_t* __restrict__ bin_start,
                                                      uint32_t* __restrict__ inter_val, uint8_t* __restrict__ inter_fine) {
    // staging: row|sign (4 B) and the 16-bit bucket (2 B) of the CTA's <= SORT_PTS * 16 entries, sorted by coarse bin
    extern __shared__ uint32_t st_dyn[];  // SORT_PTS * 16 * 6 bytes
    uint32_t* st_val = st_dyn;
    uint16_t* st_key = reinterpret_cast<uint16_t*>(st_dyn + SORT_PTS * 16);
    __shared__ uint32_t h[256], loc[256], cur[256], gbase[256], wt[8];
    const uint32_t tid = threadIdx.x;
    h[tid] = 0;
    __syncthreads();
#pragma unroll 1
    for (int r = 0; r < SORT_PTS / 256; r++) {
        uint32_t gbv[16], entv[16];
        decode_scalar(segs, blockIdx.x * SORT_PTS + r * 256 + tid, c, K, nb, n_points, gbv, entv);
#pragma unroll
        for (int w = 0; w < 16; w++)
            if (gbv[w] != DG_NONE) atomicAdd(&h[gbv[w] >> 8], 1u);
    }
    __syncthreads();
    {
        const uint32_t ex = block_excl_scan_256(h[tid], wt);
        loc[tid] = ex;
        cur[tid] = ex;
        gbase[tid] = tid < nbins ? bin_start[tid] + pre[(size_t)tid * gridDim.x + blockIdx.x] : 0u;
    }
    __syncthreads();
#pragma unroll 1
    for (int r = 0; r < SORT_PTS / 256; r++) {
        uint32_t gbv[16], entv[16];
        decode_scalar(segs, blockIdx.x * SORT_PTS + r * 256 + tid, c, K, nb, n_points, gbv, entv);
#pragma unroll
        for (int w = 0; w < 16; w++)
            if (gbv[w] != DG_NONE) {
                const uint32_t i = atomicAdd(&cur[gbv[w] >> 8], 1u);
                st_val[i] = entv[w];
                st_key[i] = (uint16_t)gbv[w];
            }
    }
    __syncthreads();
    const uint32_t n_local = loc[255] + h[255];
    for (uint32_t j = tid; j < n_local; j += 256) {
        const uint32_t key = st_key[j], bin = key >> 8, dst = gbase[bin] + (j - loc[bin]);
        inter_val[dst] = st_val[j];
        inter_fine[dst] = (uint8_t)key;
    }
}

// bin b = entries [bin_start[b], bin_start[b + 1]) of `inter`, all of buckets [256 b, 256 b + 256).
// One thread-block CLUSTER of SORT_CLUSTER CTAs per bin: every CTA counts the fine buckets of its slice in its own shared
// memory, reads the other CTAs' counts through distributed shared memory (no global histogram, no second kernel), and
// places its slice -- through a shared-memory staging buffer, so that every (CTA, bucket) run leaves as one contiguous store.
// A bin of any size is spread over the cluster, which keeps structured scalars (one bucket holding most entries) parallel.
#define SORT_CLUSTER 8
#define SORT_SLICE_SMEM 4096u  // slice entries staged in shared memory (20 KB); larger slices store directly
#define SORT_REG 16  // slice entries a thread keeps in registers between counting and placing (slices up to 4096)
__global__ void __cluster_dims__(SORT_CLUSTER, 1, 1) __launch_bounds__(256)
    k_sort_bins(const uint32_t* __restrict__ inter_val, const uint8_t* __restrict__ inter_fine, const uint32_t* __restrict__ bin_start,
                uint32_t nbins, uint32_t* __restrict__ bucket_off, uint32_t* __restrict__ entries) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ uint32_t h[256], lofs[256], gofs[256], cur[256], wt[8];
    __shared__ uint32_t out_s[SORT_SLICE_SMEM];
    __shared__ uint8_t out_k[SORT_SLICE_SMEM];
    const uint32_t tid = threadIdx.x, rank = cluster.block_rank(), bin = blockIdx.x / SORT_CLUSTER;
    const uint32_t b0 = bin_start[bin], b1 = bin_start[bin + 1];  // bin_start[nbins] = E
    const uint32_t n = b1 - b0, slice = (n + SORT_CLUSTER - 1) / SORT_CLUSTER;
    const uint32_t s0 = min(b0 + rank * slice, b1), s1 = min(s0 + slice, b1), ns = s1 - s0;
    const bool in_regs = ns <= SORT_REG * 256;
    h[tid] = 0;
    __syncthreads();
    // count; a slice of up to 4096 entries stays in registers for the placing pass (one read of the intermediate array)
    uint32_t rv[SORT_REG], rk[SORT_REG];
    if (in_regs) {
#pragma unroll
        for (int u = 0; u < SORT_REG; u++) {
            const uint32_t j = s0 + u * 256 + tid;
            rk[u] = j < s1 ? (uint32_t)inter_fine[j] : DG_NONE;
            rv[u] = j < s1 ? inter_val[j] : 0u;
        }
#pragma unroll
        for (int u = 0; u < SORT_REG; u++)
            if (rk[u] != DG_NONE) atomicAdd(&h[rk[u]], 1u);
    } else {
        for (uint32_t j0 = s0 + tid; j0 < s1; j0 += 8 * 256) {
            uint32_t k[8];
#pragma unroll
            for (int u = 0; u < 8; u++) k[u] = j0 + u * 256 < s1 ? (uint32_t)inter_fine[j0 + u * 256] : DG_NONE;
#pragma unroll
            for (int u = 0; u < 8; u++)
                if (k[u] != DG_NONE) atomicAdd(&h[k[u]], 1u);
        }
    }
    cluster.sync();
    // fine bucket `tid`: entries of the lower-ranked CTAs, and of the whole bin
    uint32_t before = 0, total = 0;
#pragma unroll
    for (uint32_t q = 0; q < SORT_CLUSTER; q++) {
        const uint32_t c = cluster.map_shared_rank(h, q)[tid];
        before += q < rank ? c : 0u;
        total += c;
    }
    const uint32_t off = block_excl_scan_256(total, wt);   // start of the bucket inside the bin
    const uint32_t lo = block_excl_scan_256(h[tid], wt);   // start of the bucket inside this CTA's slice
    gofs[tid] = b0 + off + before;                         // where this CTA's run of the bucket goes
    lofs[tid] = lo;
    cur[tid] = lo;
    if (rank == 0) {
        bucket_off[(size_t)bin * 256 + tid] = b0 + off;
        if (bin == nbins - 1 && tid == 255) bucket_off[(size_t)nbins * 256] = b1;
    }
    cluster.sync();  // nobody leaves (or reuses h) while its counts are still being read
    if (in_regs) {
#pragma unroll
        for (int u = 0; u < SORT_REG; u++)
            if (rk[u] != DG_NONE) {
                const uint32_t i = atomicAdd(&cur[rk[u]], 1u);
                out_s[i] = rv[u];
                out_k[i] = (uint8_t)rk[u];
            }
        __syncthreads();
        for (uint32_t i = tid; i < ns; i += 256) {
            const uint32_t k = out_k[i];
            entries[gofs[k] + (i - lofs[k])] = out_s[i];
        }
    } else {  // an oversized slice (structured scalars: a few buckets hold everything, consecutive ranks = consecutive words)
        for (uint32_t j0 = s0 + tid; j0 < s1; j0 += 8 * 256) {
            uint32_t k[8], v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const bool ok = j0 + u * 256 < s1;
                k[u] = ok ? (uint32_t)inter_fine[j0 + u * 256] : DG_NONE;
                v[u] = ok ? inter_val[j0 + u * 256] : 0u;
            }
#pragma unroll
            for (int u = 0; u < 8; u++)
                if (k[u] != DG_NONE) entries[gofs[k[u]] + (atomicAdd(&cur[k[u]], 1u) - lofs[k[u]])] = v[u];
        }
    }
}

// ------------------------------------------------------------------------------------------
// stage 4: bucket accumulation -- the IMAD-bound kernel
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
// the (y+x, y-x) pair comes back as (qa, qb) = neg ? (y+x, y-x) : (y-x, y+x): the sign of the digit swaps two addresses
struct niels_swapped {
    fe qa, qb, t2d;
};
__device__ __forceinline__ niels_swapped load_niels_swapped(const ge_niels* __restrict__ rows, uint32_t row, bool neg) {
    const uint4* p = reinterpret_cast<const uint4*>(rows + row);
    const int oa = neg ? 0 : 2, ob = 2 - oa;
    uint4 a = __ldg(p + oa), b = __ldg(p + oa + 1), c = __ldg(p + ob), d = __ldg(p + ob + 1), e = __ldg(p + 4), f = __ldg(p + 5);
    niels_swapped q;
    q.qa.v[0] = a.x, q.qa.v[1] = a.y, q.qa.v[2] = a.z, q.qa.v[3] = a.w;
    q.qa.v[4] = b.x, q.qa.v[5] = b.y, q.qa.v[6] = b.z, q.qa.v[7] = b.w;
    q.qb.v[0] = c.x, q.qb.v[1] = c.y, q.qb.v[2] = c.z, q.qb.v[3] = c.w;
    q.qb.v[4] = d.x, q.qb.v[5] = d.y, q.qb.v[6] = d.z, q.qb.v[7] = d.w;
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

__device__ __forceinline__ ge_ext shfl_ext(const ge_ext& p, int src) {
    ge_ext r;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r.X.v[i] = __shfl_sync(0xffffffffu, p.X.v[i], src);
        r.Y.v[i] = __shfl_sync(0xffffffffu, p.Y.v[i], src);
        r.Z.v[i] = __shfl_sync(0xffffffffu, p.Z.v[i], src);
        r.T.v[i] = __shfl_sync(0xffffffffu, p.T.v[i], src);
    }
    return r;
}

// smallest j in (lo, hi] with bucket_off[j] > e; the entry e lives in bucket j - 1
__device__ __forceinline__ uint32_t bucket_upper(const uint32_t* __restrict__ bucket_off, uint32_t lo, uint32_t hi, uint32_t e) {
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (bucket_off[mid] > e) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// Chunk t = sorted entries [t*CL, (t+1)*CL).  Partial slot of (chunk t, bucket b) = t + b: buckets are sorted along the
// entries, so the sum is unique, and the slots of one bucket are contiguous: chunks off[b]/CL .. (off[b+1]-1)/CL, every one
// of them used.  A warp whose 32 chunks all lie inside ONE bucket (0/1-valued a_L vectors, the 16 digit buckets of
// a_R = -1: a single bucket holds 2^16 .. 2^21 entries; small MSMs on the 128-bucket table) adds its lanes' sums with
// shuffles and emits one partial at the slot of its first chunk, and a CTA whose warps all did so emits one partial for
// the CTA; k_bucket_reduce derives the same predicates from bucket_off and skips the other slots.
// VARIANT 0: 4 CTAs per SM (<= 128 registers); 1: the next Niels row in flight during the addition, 3 CTAs per SM;
// 2: 5 CTAs per SM (<= 96 registers, a few spilled words outside the inner loop)
// (MEASURED, not the default: 320.7 vs 322.8 us at 2^18 points, 1274 vs 1297 at 2^20 with cp.async.ca -- the kernel is bound by
// the multiplier pipe, not by the gather; with cp.async.cg, past L1, it LOSES 27 %: the six granules of a row share 128-byte
// lines with each other and with the neighbouring rows, and L1 serves half of them.)
// VARIANT 3: the next Niels row travels global -> shared memory with cp.async (LDGSTS: no registers while in flight, so the
// kernel keeps its 4 CTAs per SM) while the current addition runs; every thread owns two 96-byte slots, laid out granule-major
// ([buffer][16-byte granule][thread]) so that neither the asynchronous stores nor the LDS.128 reads conflict.
__device__ __forceinline__ void acc_prefetch_row(uint4* slot /* &pre[buf][0][tid] */, const ge_niels* __restrict__ rows, uint32_t row) {
    const uint4* g = reinterpret_cast<const uint4*>(rows + row);
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const uint32_t sa = (uint32_t)__cvta_generic_to_shared(slot + k * ACC_THREADS);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(g + k) : "memory");
    }
}
__device__ __forceinline__ fe acc_lds_fe(const uint4* slot, int g0) {
    const uint4 a = slot[g0 * ACC_THREADS], b = slot[(g0 + 1) * ACC_THREADS];
    fe r;
    r.v[0] = a.x, r.v[1] = a.y, r.v[2] = a.z, r.v[3] = a.w, r.v[4] = b.x, r.v[5] = b.y, r.v[6] = b.z, r.v[7] = b.w;
    return r;
}
template <int VARIANT>
__global__ void __launch_bounds__(ACC_THREADS, VARIANT == 1 ? 3 : VARIANT == 2 ? 5 : 4)
    k_accumulate(const ge_niels* __restrict__ rows, const uint32_t* __restrict__ entries,
                 const uint32_t* __restrict__ bucket_off, const MsmMeta* __restrict__ meta, uint32_t G,
                 ge_ext* __restrict__ partials) {
    __shared__ ge_ext sh[ACC_THREADS / 32];
    __shared__ uint32_t sh_b0;
    __shared__ uint4 pre[VARIANT == 3 ? 2 * 6 * ACC_THREADS : 1];
    const uint32_t E = meta->E, CL = meta->CL;
    const uint64_t blk_e0 = (uint64_t)blockIdx.x * ACC_THREADS * CL;
    if (blk_e0 >= E) return;
    const uint32_t t = blockIdx.x * ACC_THREADS + threadIdx.x;
    const uint64_t e0w = (uint64_t)t * CL;
    const bool active = e0w < E;
    const uint32_t e0 = active ? (uint32_t)e0w : E;
    const uint32_t e1 = (uint32_t)min((uint64_t)e0 + CL, (uint64_t)E);
    uint32_t b = 0, next = 0;
    if (active) {
        const uint32_t j = bucket_upper(bucket_off, 0, G, e0);
        b = j - 1;
        next = bucket_off[j];
    }
    if (threadIdx.x == 0) sh_b0 = b;
    __syncthreads();
    const uint32_t b0 = sh_b0;
    const uint64_t blk_e1 = min(blk_e0 + (uint64_t)ACC_THREADS * CL, (uint64_t)E);
    const bool uniform = blk_e1 <= bucket_off[b0 + 1];
    // the same test per warp (32 consecutive chunks inside one bucket)
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t bw0 = __shfl_sync(0xffffffffu, b, 0);
    const uint64_t w_e0 = (uint64_t)(t - lane) * CL;
    const uint64_t w_e1 = min(w_e0 + (uint64_t)32 * CL, (uint64_t)E);
    const bool w_uniform = w_e0 < E && w_e1 <= bucket_off[bw0 + 1];
    ge_ext acc = ge_identity();
    if (active && VARIANT == 1) {  // variant: the Niels row of entry e+1 is in flight while entry e is added
        uint32_t ent = __ldg(entries + e0);
        ge_niels q = load_niels(rows, ent & 0x7fffffffu);
#pragma unroll 1
        for (uint32_t e = e0; e < e1; e++) {
            if (e == next) {
                store_ext(partials + t + b, acc);
                acc = ge_identity();
                b++;
                next = bucket_off[b + 1];
                if (next == e) {
                    const uint32_t j = bucket_upper(bucket_off, b + 1, G, e);
                    b = j - 1;
                    next = bucket_off[j];
                }
            }
            const bool neg = ent >> 31;
            const ge_niels cur = q;
            if (e + 1 < e1) {
                ent = __ldg(entries + e + 1);
                q = load_niels(rows, ent & 0x7fffffffu);
            }
            acc = ge_madd(acc, cur, neg);
        }
    } else if (active && VARIANT == 3) {
        uint32_t ent = __ldg(entries + e0);
        acc_prefetch_row(pre + threadIdx.x, rows, ent & 0x7fffffffu);
        asm volatile("cp.async.commit_group;" ::: "memory");
        uint32_t buf = 0;
#pragma unroll 1
        for (uint32_t e = e0; e < e1; e++) {
            if (e == next) {
                store_ext(partials + t + b, acc);
                acc = ge_identity();
                b++;
                next = bucket_off[b + 1];
                if (next == e) {
                    const uint32_t j = bucket_upper(bucket_off, b + 1, G, e);
                    b = j - 1;
                    next = bucket_off[j];
                }
            }
            const bool neg = ent >> 31;
            if (e + 1 < e1) {
                ent = __ldg(entries + e + 1);
                acc_prefetch_row(pre + (buf ^ 1) * 6 * ACC_THREADS + threadIdx.x, rows, ent & 0x7fffffffu);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 1;" ::: "memory");  // the row of entry e has landed (the next one may still fly)
            const uint4* slot = pre + buf * 6 * ACC_THREADS + threadIdx.x;
            const int oa = neg ? 0 : 2;
            const fe qa = acc_lds_fe(slot, oa), qb = acc_lds_fe(slot, 2 - oa), t2d = acc_lds_fe(slot, 4);
            acc = ge_madd_swapped(acc, qa, qb, t2d, neg);
            buf ^= 1;
        }
    } else if (active) {
        uint32_t ent = __ldg(entries + e0);
#pragma unroll 1
        for (uint32_t e = e0; e < e1; e++) {
            if (e == next) {  // bucket boundary inside the chunk: emit the partial of the finished bucket
                store_ext(partials + t + b, acc);
                acc = ge_identity();
                b++;
                next = bucket_off[b + 1];
                if (next == e) {  // run of empty buckets (sparse digit sets): bucket of entry e by bisection
                    const uint32_t j = bucket_upper(bucket_off, b + 1, G, e);
                    b = j - 1;
                    next = bucket_off[j];
                }
            }
            const bool neg = ent >> 31;
            const niels_swapped q = load_niels_swapped(rows, ent & 0x7fffffffu, neg);
            if (e + 1 < e1) ent = __ldg(entries + e + 1);
            acc = ge_madd_swapped(acc, q.qa, q.qb, q.t2d, neg);
        }
    }
    if (!w_uniform) {  // (a uniform CTA consists of uniform warps)
        if (active) store_ext(partials + t + b, acc);
        return;
    }
    // all 32 chunks of this warp lie in bucket bw0: one partial per warp, at the slot of the warp's first chunk
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) {
        const ge_ext other = shfl_ext(acc, (int)(lane ^ o));
        acc = ge_add(acc, other);
    }
    if (!uniform) {
        if (lane == 0) store_ext(partials + t + bw0, acc);
        return;
    }
    // all chunks of the CTA lie in bucket b0: one partial per CTA
    const uint32_t wid = threadIdx.x >> 5;
    if (lane == 0) store_ext(sh + wid, acc);
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t nw = (uint32_t)((blk_e1 - blk_e0 + (uint64_t)32 * CL - 1) / ((uint64_t)32 * CL));  // warps with work
#pragma unroll 1
        for (uint32_t k = 1; k < nw; k++) acc = ge_add(acc, load_ext(sh + k));
        store_ext(partials + blockIdx.x * ACC_THREADS + b0, acc);
    }
}

// ------------------------------------------------------------------------------------------
// stage 5 / 6: sum_b (b + 1) * S_b without scalar multiplications
// ------------------------------------------------------------------------------------------
// The partial slots of bucket gb as a virtual list of five runs (see k_accumulate): single chunks before the first
// collapsed warp, collapsed warps before the first collapsed CTA, collapsed CTAs, collapsed warps after them, single
// chunks after the last collapsed warp.
struct SlotList {
    uint32_t gb, n;
    uint32_t s0, n0;  // chunks   s0 + i
    uint32_t s1, n1;  // warps    32 * (s1 + i)
    uint32_t s2, n2;  // CTAs     ACC_THREADS * (s2 + i)
    uint32_t s3, n3;  // warps    32 * (s3 + i)
    uint32_t s4;      // chunks   s4 + i
    __device__ __forceinline__ uint32_t slot(uint32_t i) const {
        if (i < n0) return s0 + i + gb;
        i -= n0;
        if (i < n1) return 32 * (s1 + i) + gb;
        i -= n1;
        if (i < n2) return ACC_THREADS * (s2 + i) + gb;
        i -= n2;
        if (i < n3) return 32 * (s3 + i) + gb;
        return s4 + (i - n3) + gb;
    }
};
// units [u_lo, u_hi) of `span` chunks that lie completely inside the bucket's entries [lo, hi) (the last unit may be cut by E)
__device__ __forceinline__ void whole_units(uint32_t lo, uint32_t hi, uint32_t E, uint64_t span_entries, uint32_t* u_lo, uint32_t* u_hi) {
    *u_lo = (uint32_t)((lo + span_entries - 1) / span_entries);
    *u_hi = hi == E ? (uint32_t)((E + span_entries - 1) / span_entries) : (uint32_t)(hi / span_entries);
}
__device__ __forceinline__ SlotList slot_list(uint32_t gb, uint32_t lo, uint32_t hi, uint32_t E, uint32_t CL) {
    SlotList L;
    L.gb = gb;
    L.n = L.s0 = L.n0 = L.s1 = L.n1 = L.s2 = L.n2 = L.s3 = L.n3 = L.s4 = 0;
    if (hi <= lo) return L;
    const uint32_t t_lo = lo / CL, t_hi = (hi - 1) / CL;
    uint32_t w_lo, w_hi, k_lo, k_hi;
    whole_units(lo, hi, E, (uint64_t)32 * CL, &w_lo, &w_hi);
    L.s0 = t_lo;
    if (w_hi <= w_lo) {  // no collapsed warp
        L.n = L.n0 = t_hi - t_lo + 1;
        return L;
    }
    L.n0 = 32 * w_lo - t_lo;
    L.s4 = 32 * w_hi;
    const uint32_t n4 = t_hi >= L.s4 ? t_hi - L.s4 + 1 : 0;
    whole_units(lo, hi, E, (uint64_t)ACC_THREADS * CL, &k_lo, &k_hi);
    const uint32_t wpb = ACC_THREADS / 32;
    if (k_hi <= k_lo) {  // no collapsed CTA
        L.s1 = w_lo;
        L.n1 = w_hi - w_lo;
    } else {
        L.s1 = w_lo;
        L.n1 = wpb * k_lo - w_lo;
        L.s2 = k_lo;
        L.n2 = k_hi - k_lo;
        L.s3 = wpb * k_hi;
        L.n3 = w_hi > L.s3 ? w_hi - L.s3 : 0;
    }
    L.n = L.n0 + L.n1 + L.n2 + L.n3 + n4;
    return L;
}

__device__ __forceinline__ ge_ext load_ext_cg(const ge_ext* src) {  // L2 loads: data written by other SMs
    const uint4* o = reinterpret_cast<const uint4*>(src);
    ge_ext p;
    fe* f[4] = {&p.X, &p.Y, &p.Z, &p.T};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint4 a = __ldcg(o + 2 * k), b = __ldcg(o + 2 * k + 1);
        f[k]->v[0] = a.x, f[k]->v[1] = a.y, f[k]->v[2] = a.z, f[k]->v[3] = a.w;
        f[k]->v[4] = b.x, f[k]->v[5] = b.y, f[k]->v[6] = b.z, f[k]->v[7] = b.w;
    }
    return p;
}

// ---- four-warp point arithmetic on shared memory -------------------------------------------------------------------
// A lone warp runs a point addition in ~2.3 us (tools/imad_peak.cu: it is bound by its own scheduler's share of the
// multiplier pipe, however few lanes are active), and the serial part of the reduction is ~35 such operations.  Here the
// four warps of the CTA -- one per scheduler -- each compute ONE of the four independent field products of a stage, for up
// to 32 point operations at a time (lane = operation): an addition takes three product times instead of nine.
__device__ __forceinline__ fe lds_fe(const fe* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    const uint4 a = q[0], b = q[1];
    fe r;
    r.v[0] = a.x, r.v[1] = a.y, r.v[2] = a.z, r.v[3] = a.w, r.v[4] = b.x, r.v[5] = b.y, r.v[6] = b.z, r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void sts_fe(fe* p, const fe& f) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(f.v[0], f.v[1], f.v[2], f.v[3]);
    q[1] = make_uint4(f.v[4], f.v[5], f.v[6], f.v[7]);
}
struct alignas(16) CoopScratch {
    fe s[4][32];
};
// (__noinline__ throughout: every inlined copy of a point operation is ~25 KB of straight-line code that runs a handful
// of times per launch -- the first version of this kernel spent more time on instruction-cache misses than on arithmetic)
// *a += *b (shared memory, a != b) for the lanes with `live`; all BR_THREADS threads call it together
__device__ __noinline__ void coop_add(ge_ext* a, const ge_ext* b, bool live, CoopScratch* scr) {
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (live) {
        fe r;
        if (w == 0) r = fe_mul(fe_sub(lds_fe(&a->Y), lds_fe(&a->X)), fe_sub(lds_fe(&b->Y), lds_fe(&b->X)));
        else if (w == 1) r = fe_mul(fe_add(lds_fe(&a->Y), lds_fe(&a->X)), fe_add(lds_fe(&b->Y), lds_fe(&b->X)));
        else if (w == 2) r = fe_mul(fe_mul(lds_fe(&a->T), fe_2D()), lds_fe(&b->T));
        else {
            r = fe_mul(lds_fe(&a->Z), lds_fe(&b->Z));
            r = fe_add(r, r);
        }
        sts_fe(&scr->s[w][lane], r);
    }
    __syncthreads();
    if (live) {
        const fe A = lds_fe(&scr->s[0][lane]), B = lds_fe(&scr->s[1][lane]), C = lds_fe(&scr->s[2][lane]), D = lds_fe(&scr->s[3][lane]);
        if (w == 0) sts_fe(&a->X, fe_mul(fe_sub(B, A), fe_sub(D, C)));       // E F
        else if (w == 1) sts_fe(&a->Y, fe_mul(fe_add(D, C), fe_add(B, A)));  // G H
        else if (w == 2) sts_fe(&a->Z, fe_mul(fe_sub(D, C), fe_add(D, C)));  // F G
        else sts_fe(&a->T, fe_mul(fe_sub(B, A), fe_add(B, A)));              // E H
    }
    __syncthreads();
}
// *p = 2 * *p
__device__ __noinline__ void coop_dbl(ge_ext* p, bool live, CoopScratch* scr) {
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (live) {
        fe r;
        if (w == 0) r = fe_sqr(lds_fe(&p->X));
        else if (w == 1) r = fe_sqr(lds_fe(&p->Y));
        else if (w == 2) {
            r = fe_sqr(lds_fe(&p->Z));
            r = fe_add(r, r);
        } else r = fe_sqr(fe_add(lds_fe(&p->X), lds_fe(&p->Y)));
        sts_fe(&scr->s[w][lane], r);
    }
    __syncthreads();
    if (live) {
        const fe A = lds_fe(&scr->s[0][lane]), B = lds_fe(&scr->s[1][lane]), C = lds_fe(&scr->s[2][lane]), S = lds_fe(&scr->s[3][lane]);
        const fe G = fe_sub(B, A), H = fe_neg(fe_add(A, B));  // a = -1:  D = -A,  G = D + B,  H = D - B
        if (w == 0) sts_fe(&p->X, fe_mul(fe_sub(fe_sub(S, A), B), fe_sub(G, C)));  // E F
        else if (w == 1) sts_fe(&p->Y, fe_mul(G, H));
        else if (w == 2) sts_fe(&p->Z, fe_mul(fe_sub(G, C), G));                     // F G
        else sts_fe(&p->T, fe_mul(fe_sub(fe_sub(S, A), B), H));                       // E H
    }
    __syncthreads();
}

// one thread, one addition: *dst = *a + *b (any address space)
__device__ __noinline__ void pt_add(ge_ext* dst, const ge_ext* a, const ge_ext* b) {
    const ge_ext r = ge_add(load_ext(a), load_ext(b));
    store_ext(dst, r);
}

// Butterfly levels over `nseg` segment STATES of L + 1 values each, in shared memory (state s at pts + s * stride *
// REDUCE_MAXV, stride doubling with every level): joins neighbours in place until state 0 holds L + log2(nseg) + 1 values.
__device__ __noinline__ void state_levels(ge_ext* pts, uint32_t nseg, uint32_t* L_io, CoopScratch* scr) {
    uint32_t L = *L_io, stride = 1;
    const uint32_t lane = threadIdx.x & 31;
    while (nseg > 1) {
        const uint32_t nv = L + 1, items = (nseg >> 1) * nv;
        if (items > 64) {  // plenty of parallel work: one addition per thread
#pragma unroll 1
            for (uint32_t i = threadIdx.x; i < items; i += blockDim.x) {
                const uint32_t q = i / nv, v = i - q * nv;
                ge_ext* A = pts + (size_t)(2 * q) * stride * REDUCE_MAXV + v;
                pt_add(A, A, A + (size_t)stride * REDUCE_MAXV);
            }
            __syncthreads();
        } else {
#pragma unroll 1
            for (uint32_t base = 0; base < items; base += 32) {
                const uint32_t i = base + lane, q = i / nv, v = i - q * nv;
                ge_ext* A = pts + (size_t)(2 * q) * stride * REDUCE_MAXV + v;
                coop_add(A, A + (size_t)stride * REDUCE_MAXV, i < items, scr);
            }
        }
        for (uint32_t q = threadIdx.x; q < (nseg >> 1); q += blockDim.x) {  // M_L of the joined segment = R of its upper half
            ge_ext* A = pts + (size_t)(2 * q) * stride * REDUCE_MAXV;
            store_ext(A + L + 1, load_ext(A + (size_t)stride * REDUCE_MAXV));
        }
        __syncthreads();
        nseg >>= 1;
        stride <<= 1;
        L++;
    }
    *L_io = L;
}

// One thread per bucket (BR_THREADS consecutive buckets of one set per CTA; a set with fewer buckets gets one CTA).
// Phase 1: S_b = sum of the bucket's partial slots; lanes walk their own slots, the excess of an outlier bucket is
// strided over the warp.  Phase 2: butterfly over the bucket-index bits, in shared memory with the four-warp arithmetic.
// After level k every aligned segment of 2^(k+1) positions holds, at its positions 0 .. k+1:  R = sum of the segment,
// M_j = sum of its buckets with index bit j set (j <= k).  Joining the lower half A and the upper half B:
// R = R_A + R_B, M_j = M_j,A + M_j,B (j < k), M_k = R_B.  Phase 3: the CTA that arrives last in its group of `gsize`
// CTAs joins the group's states (log2 gsize more levels), and the last group of a set joins the group states and finishes
//     sum_b (b + 1) S_b = R + sum_j 2^j M_j          (value i is doubled i - 1 times, then a tree over the c values).
// No scalar multiplication anywhere; the serial part is log2(nb) additions and c - 2 doublings at four-warp latency.
struct ReduceScratch {
    ge_ext* cta;     // [br_blocks][REDUCE_MAXV] CTA states
    ge_ext* grp;     // [n_groups][REDUCE_MAXV] group states
    uint32_t* cnt;   // [n_groups + nsets] arrival counters, zero between launches
    unsigned long long* dbg;  // diagnostic mode: [gridDim.x][8] %globaltimer stamps of the phases of every CTA (or NULL)
};
__device__ __forceinline__ void phase_stamp(const ReduceScratch& sc, int phase) {
    if (sc.dbg && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        sc.dbg[(size_t)blockIdx.x * 8 + phase] = t;
    }
}
__global__ void __launch_bounds__(BR_THREADS)
    k_bucket_reduce(const ge_ext* __restrict__ partials, const uint32_t* __restrict__ bucket_off,
                    const MsmMeta* __restrict__ meta, uint32_t bpb /* buckets per CTA, power of two <= BR_THREADS */,
                    uint32_t lv /* log2(bpb) */, uint32_t nblk /* CTAs per set */, uint32_t gsize /* CTAs per group <= 16 */,
                    ReduceScratch sc, ge_ext* __restrict__ result) {
    __shared__ ge_ext pts[16 * REDUCE_MAXV];  // bucket sums of the CTA, later up to 16 segment states
    __shared__ CoopScratch scr;
    __shared__ uint32_t sh_last;
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t E = meta->E, CL = meta->CL;
    phase_stamp(sc, 0);
    ge_ext* mine = pts + tid;                 // this thread's bucket sum
    ge_ext* part = pts + BR_THREADS + tid;    // scratch of the cooperative path
    SlotList L;
    L.gb = L.n = L.s0 = L.n0 = L.s1 = L.n1 = L.s2 = L.n2 = L.s3 = L.n3 = L.s4 = 0;
    if (tid < bpb) {
        const uint32_t gb = blockIdx.x * bpb + tid;
        L = slot_list(gb, bucket_off[gb], bucket_off[gb + 1], E, CL);
    }
    // Lanes walk their own slots up to twice the warp's average (balanced buckets: every lane busy, no cooperation
    // needed); only what an outlier bucket holds beyond that is strided over the whole warp.
    uint32_t tot = L.n;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    const uint32_t cap = max((uint32_t)BR_CAP, 2 * (tot >> 5) + 4);
    {   // (the one point addition of this kernel that stays inline: every lane is busy here, the loop is hot)
        ge_ext val = L.n ? load_ext(partials + L.slot(0)) : ge_identity();
        const uint32_t own = min(L.n, cap);
#pragma unroll 1
        for (uint32_t i = 1; i < own; i++) val = ge_add(val, load_ext(partials + L.slot(i)));
        store_ext(mine, val);
    }
    uint32_t heavy = __ballot_sync(0xffffffffu, L.n > cap);
    while (heavy) {
        const int owner = __ffs(heavy) - 1;
        heavy &= heavy - 1;
        SlotList H;
        H.gb = __shfl_sync(0xffffffffu, L.gb, owner);
        H.n = __shfl_sync(0xffffffffu, L.n, owner);
        H.s0 = __shfl_sync(0xffffffffu, L.s0, owner);
        H.n0 = __shfl_sync(0xffffffffu, L.n0, owner);
        H.s1 = __shfl_sync(0xffffffffu, L.s1, owner);
        H.n1 = __shfl_sync(0xffffffffu, L.n1, owner);
        H.s2 = __shfl_sync(0xffffffffu, L.s2, owner);
        H.n2 = __shfl_sync(0xffffffffu, L.n2, owner);
        H.s3 = __shfl_sync(0xffffffffu, L.s3, owner);
        H.n3 = __shfl_sync(0xffffffffu, L.n3, owner);
        H.s4 = __shfl_sync(0xffffffffu, L.s4, owner);
        store_ext(part, ge_identity());
#pragma unroll 1
        for (uint32_t i = cap + lane; i < H.n; i += 32) pt_add(part, part, partials + H.slot(i));
        __syncwarp();
#pragma unroll 1
        for (uint32_t o = 16; o > 0; o >>= 1) {
            if (lane < o) pt_add(part, part, part + o);
            __syncwarp();
        }
        if ((int)lane == owner) pt_add(mine, mine, part - lane);
        __syncwarp();
    }
    phase_stamp(sc, 1);
    // butterfly inside the CTA: position = thread
    __syncthreads();
#pragma unroll 1
    for (uint32_t k = 0; k < lv; k++) {
        const uint32_t h = 1u << k, nv = k + 1, items = (bpb >> (k + 1)) * nv;
#pragma unroll 1
        for (uint32_t base = 0; base < items; base += 32) {
            const uint32_t i = base + lane, seg = i / nv, p = i - seg * nv;
            ge_ext* A = pts + (size_t)seg * 2 * h + p;
            coop_add(A, A + h, i < items, &scr);
        }
        if (k >= 2) {  // M_k = R of the upper half; for k < 2 position k + 1 IS the upper half's position 0
            if (tid < (bpb >> (k + 1))) store_ext(pts + (size_t)tid * 2 * h + k + 1, load_ext(pts + (size_t)tid * 2 * h + h));
            __syncthreads();
        }
    }
    if (tid <= lv) store_ext(sc.cta + (size_t)blockIdx.x * REDUCE_MAXV + tid, load_ext(pts + tid));
    phase_stamp(sc, 2);
    // group stage: the last CTA of the group to arrive joins the group's states
    const uint32_t set = blockIdx.x / nblk, ngrp = nblk / gsize;  // groups per set
    const uint32_t group = blockIdx.x / gsize;                    // global group index
    uint32_t Lc = lv;
    if (gsize > 1) {
        __threadfence();
        __syncthreads();
        if (tid == 0) sh_last = atomicAdd(sc.cnt + group, 1u) == gsize - 1;
        __syncthreads();
        if (!sh_last) return;
        if (tid == 0) sc.cnt[group] = 0;
        __threadfence();
        for (uint32_t i = tid; i < gsize * (lv + 1); i += BR_THREADS) {
            const uint32_t s_ = i / (lv + 1), v = i - s_ * (lv + 1);
            store_ext(pts + (size_t)s_ * REDUCE_MAXV + v, load_ext_cg(sc.cta + ((size_t)group * gsize + s_) * REDUCE_MAXV + v));
        }
        __syncthreads();
        state_levels(pts, gsize, &Lc, &scr);
        phase_stamp(sc, 3);
    }
    // set stage: the last group of the set joins the group states and finishes
    if (ngrp > 1) {
        if (tid <= Lc) store_ext(sc.grp + (size_t)group * REDUCE_MAXV + tid, load_ext(pts + tid));
        __threadfence();
        __syncthreads();
        const uint32_t nctr = gridDim.x / gsize;  // group counters come first
        if (tid == 0) sh_last = atomicAdd(sc.cnt + nctr + set, 1u) == ngrp - 1;
        __syncthreads();
        if (!sh_last) return;
        if (tid == 0) sc.cnt[nctr + set] = 0;
        __threadfence();
        for (uint32_t i = tid; i < ngrp * (Lc + 1); i += BR_THREADS) {
            const uint32_t s_ = i / (Lc + 1), v = i - s_ * (Lc + 1);
            store_ext(pts + (size_t)s_ * REDUCE_MAXV + v, load_ext_cg(sc.grp + ((size_t)set * ngrp + s_) * REDUCE_MAXV + v));
        }
        __syncthreads();
        state_levels(pts, ngrp, &Lc, &scr);
        phase_stamp(sc, 4);
    }
    // pts[0 .. Lc] = R, M_0 .. M_{Lc-1}:  value i >= 2 is doubled i - 1 times, then everything is added up
    for (uint32_t i = tid; i < 32; i += BR_THREADS)
        if (i > Lc) store_ext(pts + i, ge_identity());
    __syncthreads();
#pragma unroll 1
    for (uint32_t rep = 1; rep < Lc; rep++) coop_dbl(pts + lane, lane >= rep + 1 && lane <= Lc, &scr);
    phase_stamp(sc, 5);
#pragma unroll 1
    for (uint32_t o = 8; o > 0; o >>= 1) coop_add(pts + lane, pts + lane + o, lane < o, &scr);  // c <= 16 values
    if (tid < 32) {
        // (one warp copies the 128 bytes out)
        if (tid < 8) reinterpret_cast<uint4*>(result + set)[tid] = reinterpret_cast<const uint4*>(pts)[tid];
    }
    phase_stamp(sc, 6);
}

// ------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------
// One-kernel MSM for small statements (at most SMALL_KERNEL_MAX_POINTS points, 8-bit-window table, 128 buckets per set,
// one or two bucket sets).  A statement of a few hundred multipliers runs ~11 MSMs, and with six launches (+ a memset) each
// the whole proof is bound by the NUMBER of driver calls, not by the GPU (tools/gpu_timeline.py TIMELINE_MODE=c4: ~250
// calls per statement at ~3 us of a process-wide lock each).
//   CTA = a slice of SMK_SLICE scalars: digits -> counting sort of the slice's <= 32 x SMK_SLICE entries by bucket in shared
//   memory -> thread per bucket adds its rows -> the CTA's own weighted sum  T_cta = sum_b (b + 1) S_b = sum_k U_k,
//   U_k = sum_{b >= k} S_b  (suffix scan + tree over the 128 positions of a set: 14 dependent additions, no doublings).
//   The sum is linear in the bucket contents, so the CTAs never exchange buckets: the last CTA to arrive adds the T_cta.
// (A first version with one CTA per bucket, every CTA decoding every scalar, spent 8.4 M digit steps per MSM of 1026
// points: 230 us under load.)
// ------------------------------------------------------------------------------------------
#define SMALL_KERNEL_MAX_POINTS 2050u
#define SMK_THREADS 256
#define SMK_SLICE 32u                    // scalars per CTA
#define SMK_MAXE (SMK_SLICE * 32u)       // entries per CTA (K <= 32 windows)
#define SMK_HEAVY 64u                    // entries a bucket's own thread adds alone; the rest is shared by the whole CTA
#define SMK_MAX_CTAS ((SMALL_KERNEL_MAX_POINTS + SMK_SLICE - 1) / SMK_SLICE)
#define SMK_SMEM (2 * SMK_THREADS * sizeof(ge_ext) + SMK_MAXE * 4)
// digits of the slice's scalar `g` (g < segs.total): f(bucket in [0, nsets * nb), row | sign << 31) per non-zero digit
template <typename F>
__device__ __forceinline__ void smk_decode(const MsmSegments& segs, uint32_t g, int c, int K, uint32_t nb, uint32_t n_points, F&& f) {
    uint32_t si = 0, base = 0, end = 0;
#pragma unroll
    for (int k = 0; k < MSM_MAX_SEGMENTS; k++) {
        if (k < (int)segs.nseg) {
            end += segs.seg[k].count;
            if (g >= end) {
                base = end;
                si = k + 1;
            }
        }
    }
    const MsmSegment sg = segs.seg[si];
    const uint32_t i = g - base;
    uint32_t set = sg.set_id;
    if (sg.mode == 1) set += ((i % sg.period) >= (sg.period >> 1)) ? 0u : 1u;
    if (sg.mode == 2) set += ((i % sg.period) < (sg.period >> 1)) ? 0u : 1u;
    if (sg.mode == 3) set += i % sg.period;
    const uint4* sp = reinterpret_cast<const uint4*>(sg.scalars) + 2 * (size_t)i;
    const uint4 lo = __ldg(sp), hi = __ldg(sp + 1);
    uint32_t s[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    const uint32_t point = sg.point_base + i;
    const uint32_t mask = (1u << c) - 1u, half = 1u << (c - 1);
    uint32_t carry = 0;
#pragma unroll 1
    for (int w = 0; w < K; w++) {
        const uint32_t raw = (s[0] & mask) + carry;
#pragma unroll
        for (int k = 0; k < 7; k++) s[k] = __funnelshift_r(s[k], s[k + 1], c);
        s[7] >>= c;
        const uint32_t neg = raw > half;
        const uint32_t mag = neg ? ((1u << c) - raw) : raw;
        carry = neg;
        if (mag != 0) f(set * nb + (mag - 1), ((uint32_t)w * n_points + point) | (neg << 31));
        if ((s[0] | s[1] | s[2] | s[3] | s[4] | s[5] | s[6] | s[7] | carry) == 0) break;  // short scalars (0/1 witness bits)
    }
}
__global__ void __launch_bounds__(SMK_THREADS)
    k_msm_small(MsmSegments segs, const ge_niels* __restrict__ rows, uint32_t n_points, int c, int K, uint32_t nb, uint32_t nsets,
                ge_ext* __restrict__ cta_sums /* [nsets][gridDim.x] */, uint32_t* __restrict__ arrive, ge_ext* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t smk_raw[];
    ge_ext* A0 = reinterpret_cast<ge_ext*>(smk_raw);
    ge_ext* A1 = A0 + SMK_THREADS;
    uint32_t* ent = reinterpret_cast<uint32_t*>(A1 + SMK_THREADS);
    __shared__ uint32_t cnt[SMK_THREADS], off[SMK_THREADS + 1], cur[SMK_THREADS], heavy[SMK_MAXE / SMK_HEAVY + 1];
    __shared__ uint32_t n_heavy, sh_last;
    const uint32_t tid = threadIdx.x, G = nsets * nb;  // G <= SMK_THREADS
    const uint32_t g0 = blockIdx.x * SMK_SLICE;
    const uint32_t g = g0 + tid;
    const bool has = tid < SMK_SLICE && g < segs.total;
    cnt[tid] = 0;
    if (tid == 0) n_heavy = 0;
    __syncthreads();
    if (has) smk_decode(segs, g, c, K, nb, n_points, [&](uint32_t gb, uint32_t) { atomicAdd(&cnt[gb], 1u); });
    __syncthreads();
    if (tid < 32) {  // exclusive scan of the 256 counts: 8 per lane
        uint32_t v[8], sum = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            v[k] = sum;
            sum += cnt[tid * 8 + k];
        }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)tid >= o) incl += up;
        }
        const uint32_t excl = incl - sum;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            off[tid * 8 + k] = excl + v[k];
            cur[tid * 8 + k] = excl + v[k];
        }
        if (tid == 31) off[SMK_THREADS] = incl;
    }
    __syncthreads();
    if (has) smk_decode(segs, g, c, K, nb, n_points, [&](uint32_t gb, uint32_t e) { ent[atomicAdd(&cur[gb], 1u)] = e; });
    __syncthreads();
    // thread per bucket
    {
        ge_ext acc = ge_identity();
        if (tid < G) {
            const uint32_t lo = off[tid], n = cnt[tid], own = min(n, SMK_HEAVY);
#pragma unroll 1
            for (uint32_t e = 0; e < own; e++) {
                const uint32_t x = ent[lo + e];
                acc = ge_madd(acc, load_niels(rows, x & 0x7fffffffu), (x >> 31) != 0);
            }
            if (n > SMK_HEAVY) heavy[atomicAdd(&n_heavy, 1u)] = tid;
        }
        store_ext(A0 + tid, acc);
    }
    __syncthreads();
    // buckets with more than SMK_HEAVY entries in this slice (structured scalars): the whole CTA adds the remainder
    const uint32_t nh = n_heavy;
#pragma unroll 1
    for (uint32_t hk = 0; hk < nh; hk++) {
        const uint32_t gb = heavy[hk], lo = off[gb] + SMK_HEAVY, hi_ = off[gb] + cnt[gb];
        ge_ext acc = ge_identity();
#pragma unroll 1
        for (uint32_t e = lo + tid; e < hi_; e += SMK_THREADS) {
            const uint32_t x = ent[e];
            acc = ge_madd(acc, load_niels(rows, x & 0x7fffffffu), (x >> 31) != 0);
        }
        store_ext(A1 + tid, acc);
        __syncthreads();
#pragma unroll 1
        for (uint32_t o = SMK_THREADS / 2; o > 0; o >>= 1) {
            if (tid < o) pt_add(A1 + tid, A1 + tid, A1 + tid + o);
            __syncthreads();
        }
        if (tid == 0) pt_add(A0 + gb, A0 + gb, A1);
        __syncthreads();
    }
    // per set: suffix sums over the 128 bucket positions (double buffered), then their tree sum
    const uint32_t p = tid & (nb - 1);
    ge_ext *src = A0, *dst = A1;
#pragma unroll 1
    for (uint32_t o = 1; o < nb; o <<= 1) {
        if (tid < G) {
            if (p + o < nb) pt_add(dst + tid, src + tid, src + tid + o);
            else store_ext(dst + tid, load_ext(src + tid));
        }
        __syncthreads();
        ge_ext* t = src;
        src = dst;
        dst = t;
    }
#pragma unroll 1
    for (uint32_t o = nb / 2; o > 0; o >>= 1) {
        if (tid < G && p < o) pt_add(src + tid, src + tid, src + tid + o);
        __syncthreads();
    }
    if (tid < G && p == 0) store_ext(cta_sums + (size_t)(tid / nb) * gridDim.x + blockIdx.x, load_ext(src + tid));
    if (gridDim.x == 1) {
        if (tid < G && p == 0) store_ext(out + tid / nb, load_ext(src + tid));
        return;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) sh_last = atomicAdd(arrive, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!sh_last) return;
    if (tid == 0) *arrive = 0;  // zero between launches
    __threadfence();
    // the last CTA adds the CTA sums of every set: positions [set * 128 + k], k < gridDim.x <= SMK_MAX_CTAS <= 128
    if (tid < G) {
        if (p < gridDim.x) store_ext(A0 + tid, load_ext_cg(cta_sums + (size_t)(tid / nb) * gridDim.x + p));
        else store_ext(A0 + tid, ge_identity());
    }
    __syncthreads();
    uint32_t top = 1;
    while (top < gridDim.x) top <<= 1;
#pragma unroll 1
    for (uint32_t o = top / 2; o > 0; o >>= 1) {
        if (tid < G && p < o && p + o < gridDim.x) pt_add(A0 + tid, A0 + tid, A0 + tid + o);
        __syncthreads();
    }
    if (tid < G && p == 0) store_ext(out + tid / nb, load_ext(A0 + tid));
}

// Segments address points in the index space of the big table ([G | H | B | B~] with H at `capacity`).  MSMs whose points
// all lie in the small table's prefix use it instead: 8-bit windows, 128 buckets per set.
static bool to_small_table(const FixedTable& big, const FixedTable& sm, MsmSegments* segs) {
    if (!sm.rows || segs->total > 2 * SMALL_MSM_MAX_POINTS + 2) return false;
    const uint64_t cap = big.capacity, sc_cap = sm.capacity;
    MsmSegments out = *segs;
    for (uint32_t k = 0; k < segs->nseg; k++) {
        const uint64_t base = segs->seg[k].point_base, cnt = segs->seg[k].count;
        uint64_t nb;
        if (base < cap) {
            if (base + cnt > sc_cap) return false;
            nb = base;
        } else if (base < 2 * cap) {
            if (base - cap + cnt > sc_cap) return false;
            nb = sc_cap + (base - cap);
        } else {
            nb = 2 * sc_cap + (base - 2 * cap);
        }
        out.seg[k].point_base = (uint32_t)nb;
    }
    *segs = out;
    return true;
}


static uint32_t ilog2(uint32_t x) {
    uint32_t l = 0;
    while ((1u << l) < x) l++;
    return l;
}

int msm_run(bpg_ctx* ctx, const MsmSegments& segs_in, uint32_t nsets, ge_ext* d_out) {
    if (!ctx->table.rows) {
        bpg_set_error("msm_run: generator table not built");
        return BPG_E_ARG;
    }
    MsmSegments segs = segs_in;
    segs.var_base = 0;
    const bool use_small = to_small_table(ctx->table, ctx->small_table, &segs);
    return msm_run_table(ctx, use_small ? ctx->small_table : ctx->table, segs, nsets, d_out);
}

int msm_run_table(bpg_ctx* ctx, const FixedTable& tb, const MsmSegments& segs, uint32_t nsets, ge_ext* d_out) {
    if (nsets == 0 || segs.nseg > MSM_MAX_SEGMENTS || !tb.rows) return BPG_E_ARG;
    cudaStream_t st = ctx->stream;
    const uint32_t nb = 1u << (tb.c - 1);
    const uint32_t G = nsets * nb;
    if ((uint64_t)G >= (1ull << 24) || tb.c > REDUCE_MAXV) {
        bpg_set_error("msm_run: too many buckets");
        return BPG_E_ARG;
    }
    const uint64_t total = segs.total;
    const uint64_t max_entries = (uint64_t)tb.K * total;
    if (max_entries >= (1ull << 32) - (1ull << 24) || (uint64_t)(segs.var_base ? 1 : tb.K) * tb.n_points >= (1ull << 31)) {
        bpg_set_error("msm_run: problem too large for 32-bit entry indices");
        return BPG_E_ARG;
    }
    // chunk geometry: CL is fixed by the knob or derived on the device from the true entry count; the host only needs
    // an upper bound on the number of chunks to size the grid and the partial-slot array
    const uint32_t cl_fixed = (uint32_t)ctx->task_len, cl_min = (uint32_t)ctx->cl_min;
    // default: one and a half waves of resident CTAs (measured best at 2^18 and 2^20: tools/gpu_acc_sweep.py)
    const uint32_t target = ctx->target_chunks ? (uint32_t)ctx->target_chunks : (uint32_t)ctx->sm_count * 6u * ACC_THREADS;
    uint64_t max_chunks;
    if (cl_fixed) max_chunks = (max_entries + cl_fixed - 1) / cl_fixed;
    else max_chunks = std::min<uint64_t>((max_entries + cl_min - 1) / cl_min, target);
    if (max_chunks == 0) max_chunks = 1;
    const uint32_t acc_blocks = (uint32_t)((max_chunks + ACC_THREADS - 1) / ACC_THREADS);
    const uint64_t max_partials = (uint64_t)acc_blocks * ACC_THREADS + G + 1;
    const uint32_t bpb = nb < BR_THREADS ? nb : BR_THREADS, lv = ilog2(bpb);
    const uint32_t br_blocks = G / bpb, nblk = nb / bpb;
    // CTAs per group and groups per set: powers of two, at most 16 each (the states of a stage live in shared memory)
    uint32_t gsize = nblk >= 64 ? 16u : nblk >= 4 ? 4u : nblk;
    while (nblk / gsize > 16) gsize <<= 1;
    if (gsize > 16) {
        bpg_set_error("msm_run: more than 2^15 buckets per set are not supported");
        return BPG_E_ARG;
    }
    const uint32_t n_groups = br_blocks / gsize;
    MsmWork& w = ctx->work;
    int rc;
    if ((rc = w.hist.ensure(G + 16)) || (rc = w.bucket_off.ensure(G + 16)) || (rc = w.meta.ensure(1)) ||
        (rc = w.entries.ensure(max_entries + 1)) || (rc = w.partials.ensure(max_partials)) ||
        (rc = w.blockres.ensure(((size_t)br_blocks + (size_t)n_groups) * REDUCE_MAXV)) ||
        (rc = w.reduce_cnt.ensure(n_groups + nsets)))
        return rc;
    if (w.reduce_cnt.fresh) {  // arrival counters: zero between launches (the last arriver re-zeroes its counter)
        CUDA_TRY(cudaMemsetAsync(w.reduce_cnt.p, 0, w.reduce_cnt.cap * 4, st));
        w.reduce_cnt.fresh = false;
    }

    const bool timed = ctx->time_accum;
    int stage = 0;
    auto mark = [&]() -> cudaError_t { return timed ? cudaEventRecord(ctx->ev_stage[stage++], st) : cudaSuccess; };
    if (ctx->use_small_kernel && tb.c == 8 && nb == 128 && tb.K <= 32 && !segs.var_base && total > 0 && total <= SMALL_KERNEL_MAX_POINTS &&
        nsets <= 2 && !x_skip()) {
        const uint32_t ctas = (uint32_t)((total + SMK_SLICE - 1) / SMK_SLICE);
        if ((rc = w.small_cnt.ensure(64)) || (rc = w.partials.ensure(2 * (size_t)SMK_MAX_CTAS + 1))) return rc;
        if (w.small_cnt.fresh) {
            CUDA_TRY(cudaMemsetAsync(w.small_cnt.p, 0, w.small_cnt.cap * 4, st));
            w.small_cnt.fresh = false;
        }
        if (!ctx->small_attr_set) {
            CUDA_TRY(cudaFuncSetAttribute(k_msm_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMK_SMEM));
            ctx->small_attr_set = true;
        }
        for (int k = 0; k < 4; k++) CUDA_TRY(mark());
        k_msm_small<<<ctas, SMK_THREADS, SMK_SMEM, st>>>(segs, tb.rows, tb.n_points, tb.c, tb.K, nb, nsets, w.partials.p, w.small_cnt.p, d_out);
        for (int k = 0; k < 3; k++) CUDA_TRY(mark());
        ctx->launches++;
        CUDA_TRY(cudaGetLastError());
        if (timed) {
            CUDA_TRY(ctx_sync(ctx));
            float ms = 0;
            CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev_stage[3], ctx->ev_stage[4]));
            ctx->sum_stage_ms[3] += ms;
            ctx->sum_stage_ms[MSM_STAGES - 1] += ms;
            ctx->timed_msms++;
        }
        return BPG_OK;
    }
    const uint32_t nbins = G / 256, sort_ctas = (uint32_t)((total + SORT_PTS - 1) / SORT_PTS);
    const bool smem_sort = ctx->use_smem_sort && tb.K <= 16 && nb >= 256 && G <= 65536 && total > 0 &&
                           (uint64_t)nbins * sort_ctas < (1ull << 26);
    const int xs = x_skip();
    if (xs & 4) {
    } else if (smem_sort) {
        const size_t ncnt = (size_t)nbins * sort_ctas;
        if ((rc = w.sort_cnt.ensure(ncnt + 16)) || (rc = w.sort_base.ensure(ncnt + 16)) || (rc = w.sort_val.ensure(max_entries + 1)) || (rc = w.sort_fine.ensure(max_entries + 16)) ||
            (rc = w.sort_small.ensure(1024)))
            return rc;
        if (w.sort_small.fresh) {  // [0] arrival counter (zero between launches), [8 ..] bin totals, [264 ..] bin starts
            CUDA_TRY(cudaMemsetAsync(w.sort_small.p, 0, w.sort_small.cap * 4, st));
            w.sort_small.fresh = false;
        }
        uint32_t *arrive = w.sort_small.p, *bin_tot = w.sort_small.p + 8, *bin_start = w.sort_small.p + 8 + 256;

        CUDA_TRY(mark());
        k_sort_count<<<sort_ctas, 256, 0, st>>>(segs, tb.c, tb.K, nb, tb.n_points, nbins, w.sort_cnt.p);
        CUDA_TRY(mark());
        k_sort_scan<<<nbins, 256, 0, st>>>(w.sort_cnt.p, sort_ctas, nbins, w.sort_base.p, bin_tot, bin_start, arrive, target, cl_min,
                                           cl_fixed, w.meta.p);
        CUDA_TRY(mark());
        if (!ctx->sort_attr_set) {  // per device: once per context
            CUDA_TRY(cudaFuncSetAttribute(k_sort_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_PTS * 16 * 6));
            ctx->sort_attr_set = true;
        }
        k_sort_scatter<<<sort_ctas, 256, SORT_PTS * 16 * 6, st>>>(segs, tb.c, tb.K, nb, tb.n_points, nbins, w.sort_base.p, bin_start,
                                                                  w.sort_val.p, w.sort_fine.p);
        k_sort_bins<<<nbins * SORT_CLUSTER, 256, 0, st>>>(w.sort_val.p, w.sort_fine.p, bin_start, nbins, w.bucket_off.p, w.entries.p);
        CUDA_TRY(mark());
        ctx->launches += 4;
    } else {
    uint32_t* tickets = nullptr;  // rank of every entry inside its bucket (unrolled K <= 16 path of k_digits only)
    if (ctx->use_tickets && tb.K <= 16 && total > 0) {
        if ((rc = w.tickets.ensure((size_t)16 * total))) return rc;
        tickets = w.tickets.p;
    }
    if (w.hist_dirty || w.hist.fresh) {  // afterwards the scan kernel leaves it zeroed for the next MSM
        CUDA_TRY(cudaMemsetAsync(w.hist.p, 0, w.hist.cap * 4, st));
        w.hist_dirty = false;
        w.hist.fresh = false;
    }
    CUDA_TRY(mark());
    const uint32_t dblocks = (uint32_t)((total + 255) / 256);
    if (total > 0) {
        k_digits<false><<<dblocks, 256, 0, st>>>(segs, tb.c, tb.K, nb, tb.n_points, w.hist.p, nullptr, nullptr, tickets);
        ctx->launches++;
    }
    CUDA_TRY(mark());
    dev_scan_meta(st, w.hist.p, w.bucket_off.p, G, target, cl_min, cl_fixed, w.meta.p);
    ctx->launches++;
    CUDA_TRY(mark());
    if (total > 0) {
        k_digits<true><<<dblocks, 256, 0, st>>>(segs, tb.c, tb.K, nb, tb.n_points, w.hist.p, w.bucket_off.p, w.entries.p, tickets);
        ctx->launches++;
        if (!tickets) w.hist_dirty = true;  // the scatter pass used it as its cursor
    }
    CUDA_TRY(mark());
    }
    if (xs & 2) {
    } else if (ctx->acc_variant == 2)
        k_accumulate<2><<<acc_blocks, ACC_THREADS, 0, st>>>(tb.rows, w.entries.p, w.bucket_off.p, w.meta.p, G, w.partials.p);
    else if (ctx->acc_variant == 1)
        k_accumulate<1><<<acc_blocks, ACC_THREADS, 0, st>>>(tb.rows, w.entries.p, w.bucket_off.p, w.meta.p, G, w.partials.p);
    else if (ctx->acc_variant == 3)
        k_accumulate<3><<<acc_blocks, ACC_THREADS, 0, st>>>(tb.rows, w.entries.p, w.bucket_off.p, w.meta.p, G, w.partials.p);
    else {
        // acc_smem_pad > 0: unused dynamic shared memory that caps the resident CTAs per SM (57 KB -> 3), so that the sort
        // kernels of OTHER statements find registers while this kernel runs (it takes every register of the GPU otherwise)
        const size_t pad = (size_t)ctx->acc_smem_pad;
        if (pad > 48 * 1024 && !ctx->acc_attr_set) {
            CUDA_TRY(cudaFuncSetAttribute(k_accumulate<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad));
            ctx->acc_attr_set = true;
        }
        k_accumulate<0><<<acc_blocks, ACC_THREADS, pad, st>>>(tb.rows, w.entries.p, w.bucket_off.p, w.meta.p, G, w.partials.p);
    }
    CUDA_TRY(mark());
    ReduceScratch rs;
    rs.cta = w.blockres.p;
    rs.grp = rs.cta + (size_t)br_blocks * REDUCE_MAXV;
    rs.cnt = w.reduce_cnt.p;
    rs.dbg = nullptr;
    if (timed && getenv("BPG_REDUCE_TRACE")) {
        if ((rc = w.reduce_dbg.ensure((size_t)br_blocks * 8))) return rc;
        CUDA_TRY(cudaMemsetAsync(w.reduce_dbg.p, 0, (size_t)br_blocks * 64, st));
        rs.dbg = w.reduce_dbg.p;
    }
    if (!(xs & 1)) k_bucket_reduce<<<br_blocks, BR_THREADS, 0, st>>>(w.partials.p, w.bucket_off.p, w.meta.p, bpb, lv, nblk, gsize, rs, d_out);
    CUDA_TRY(mark());
    CUDA_TRY(mark());
    ctx->launches += 2;
    CUDA_TRY(cudaGetLastError());
    if (timed) {  // diagnostic mode: synchronous, reads back the entry count and the chunk length
        CUDA_TRY(ctx_sync(ctx));
        float ms[MSM_STAGES];
        for (int k = 0; k < MSM_STAGES - 1; k++) CUDA_TRY(cudaEventElapsedTime(&ms[k], ctx->ev_stage[k], ctx->ev_stage[k + 1]));
        CUDA_TRY(cudaEventElapsedTime(&ms[MSM_STAGES - 1], ctx->ev_stage[0], ctx->ev_stage[MSM_STAGES - 1]));
        for (int k = 0; k < MSM_STAGES; k++) ctx->sum_stage_ms[k] += ms[k];
        ctx->timed_msms++;
        MsmMeta hm;
        CUDA_TRY(cudaMemcpy(&hm, w.meta.p, sizeof hm, cudaMemcpyDeviceToHost));
        ctx->last_accum_ms = ms[3];
        ctx->last_entries = hm.E;
        ctx->last_chunk_len = hm.CL;
        ctx->sum_accum_ms += ms[3];
        ctx->sum_entries += hm.E;
        if (total > 0) {
            ctx->sum_scatter_ms += ms[2];
            ctx->sum_points += total;
        }
        if (rs.dbg) {  // phase profile of the reduction kernel: latest stamp of every phase relative to the earliest start
            std::vector<unsigned long long> h((size_t)br_blocks * 8);
            CUDA_TRY(cudaMemcpy(h.data(), rs.dbg, h.size() * 8, cudaMemcpyDeviceToHost));
            unsigned long long t0 = ~0ull, last[8] = {0};
            for (uint32_t b = 0; b < br_blocks; b++)
                if (h[(size_t)b * 8]) t0 = std::min(t0, h[(size_t)b * 8]);
            for (uint32_t b = 0; b < br_blocks; b++)
                for (int k = 0; k < 8; k++) last[k] = std::max(last[k], h[(size_t)b * 8 + k]);
            fprintf(stderr, "[bpg reduce] ctas %u: last start +%.1f, slots summed +%.1f, cta butterfly +%.1f, group +%.1f, set +%.1f, "
                    "doublings +%.1f, done +%.1f us\n", br_blocks, (last[0] - t0) * 1e-3, (last[1] - t0) * 1e-3, (last[2] - t0) * 1e-3,
                    last[3] ? (last[3] - t0) * 1e-3 : 0.0, last[4] ? (last[4] - t0) * 1e-3 : 0.0, (last[5] - t0) * 1e-3, (last[6] - t0) * 1e-3);
        }
        if (getenv("BPG_ACC_TRACE"))
            fprintf(stderr, "[bpg msm] sets %u points %llu entries %u CL %u | digits0 %.1f scan %.1f digits1 %.1f accumulate %.1f "
                    "reduce %.1f final %.1f | total %.1f us\n", nsets, (unsigned long long)total, hm.E, hm.CL, ms[0] * 1e3,
                    ms[1] * 1e3, ms[2] * 1e3, ms[3] * 1e3, ms[4] * 1e3, ms[5] * 1e3, ms[6] * 1e3);
    }
    return BPG_OK;
}
