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

#include <algorithm>

#include "circuit.hpp"
#include "ctx.hpp"

#define FINAL_THREADS 256
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
    const uint32_t point = sg.point_base + i;
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
                    gbv[w] = set * nb + (mag - 1);
                    entv[w] = ((uint32_t)w * n_points + point) | (neg << 31);
                }
            }
        }
        // With `tickets` the histogram pass keeps what its atomicAdd returns -- the rank of the entry inside its bucket
        // -- at [w][g] (coalesced), and the scatter pass is bucket_off + rank: no second round of atomics.
        if (!SCATTER) {
            if (tickets) {
                const uint32_t lane = threadIdx.x & 31, lt = (1u << lane) - 1u;
                uint32_t tk[16];
#pragma unroll
                for (int w = 0; w < 16; w++) {
                    const uint32_t key = gbv[w];
                    // Structured scalars (a_L in {0,1}, a_R in {0,-1}: every entry of a window falls into ONE bucket)
                    // would serialise on a single L2 address.  When neighbouring lanes collide, the lanes of a warp that
                    // share a bucket send one atomic for all of them and split the returned range by lane order.
                    const uint32_t nxt = __shfl_down_sync(0xffffffffu, key, 1);
                    if (__any_sync(0xffffffffu, key != DG_NONE && key == nxt && lane < 31)) {
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
// of them used.  A CTA whose ACC_THREADS chunks all lie inside ONE bucket (0/1-valued a_L vectors, the 16 digit buckets of
// a_R = -1: a single bucket holds 2^16 .. 2^21 entries) adds its threads' sums in shared memory and emits one partial at
// the slot of its first chunk; k_bucket_reduce derives the same predicate from bucket_off and skips the other slots.
__global__ void __launch_bounds__(ACC_THREADS, 4)
    k_accumulate(const ge_niels* __restrict__ rows, const uint32_t* __restrict__ entries,
                 const uint32_t* __restrict__ bucket_off, const MsmMeta* __restrict__ meta, uint32_t G,
                 ge_ext* __restrict__ partials) {
    __shared__ ge_ext sh[ACC_THREADS];
    __shared__ uint32_t sh_b0;
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
    ge_ext acc = ge_identity();
    if (active) {
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
            ge_niels q = load_niels(rows, ent & 0x7fffffffu);
            const bool neg = ent >> 31;
            if (e + 1 < e1) ent = __ldg(entries + e + 1);
            acc = ge_madd(acc, q, neg);
        }
    }
    if (!uniform) {
        if (active) store_ext(partials + t + b, acc);
        return;
    }
    block_tree_reduce(sh, acc, threadIdx.x, ACC_THREADS);
    if (threadIdx.x == 0) store_ext(partials + blockIdx.x * ACC_THREADS + b0, load_ext(sh));
}

// ------------------------------------------------------------------------------------------
// stage 5 / 6: sum_b (b + 1) * S_b without scalar multiplications
// ------------------------------------------------------------------------------------------
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

// The partial slots of bucket gb as a virtual list: chunks before the first collapsed CTA, one slot per collapsed CTA
// (see k_accumulate), chunks after the last one.
struct SlotList {
    uint32_t gb, t_lo, n_pre, k_lo, n_col, t_post, n;
    __device__ __forceinline__ uint32_t slot(uint32_t i) const {
        if (i < n_pre) return t_lo + i + gb;
        i -= n_pre;
        if (i < n_col) return (k_lo + i) * ACC_THREADS + gb;
        return t_post + (i - n_col) + gb;
    }
};
__device__ __forceinline__ SlotList slot_list(uint32_t gb, uint32_t lo, uint32_t hi, uint32_t E, uint32_t CL) {
    SlotList L;
    L.gb = gb;
    L.t_lo = L.n_pre = L.k_lo = L.n_col = L.t_post = L.n = 0;
    if (hi <= lo) return L;
    const uint32_t t_lo = lo / CL, t_hi = (hi - 1) / CL;
    const uint64_t bcl = (uint64_t)ACC_THREADS * CL;
    // CTA k of k_accumulate collapsed into one partial iff  lo <= k*bcl  and  min((k+1)*bcl, E) <= hi
    const uint32_t k_lo = (uint32_t)((lo + bcl - 1) / bcl);
    const uint32_t k_hi = hi == E ? (uint32_t)((E + bcl - 1) / bcl) : (uint32_t)(hi / bcl);
    L.t_lo = t_lo;
    uint32_t n_post = 0;
    if (k_hi <= k_lo) {
        L.n_pre = t_hi - t_lo + 1;
    } else {
        L.n_pre = k_lo * ACC_THREADS - t_lo;
        L.k_lo = k_lo;
        L.n_col = k_hi - k_lo;
        L.t_post = k_hi * ACC_THREADS;
        n_post = t_hi >= L.t_post ? t_hi - L.t_post + 1 : 0;
    }
    L.n = L.n_pre + L.n_col + n_post;
    return L;
}

// One thread per bucket (BR_THREADS consecutive buckets of one set per CTA; a set with fewer buckets gets one CTA).
// Phase 1: S_b = sum of the bucket's partial slots; lanes walk the first BR_CAP slots alone, the rest of a heavy bucket
// is strided over the warp.  Phase 2: butterfly over the bucket-index bits.  After level k every aligned segment of
// 2^(k+1) threads holds, at its positions 0 .. k+1:  R = sum of the segment,  M_j = sum of its buckets with index bit
// j set (j <= k).  Joining the lower half A and the upper half B: R = R_A + R_B, M_j = M_j,A + M_j,B (j < k), M_k = R_B
// -- one addition per lane.  The CTA emits lv + 1 points; k_reduce_final continues across CTAs.
__global__ void __launch_bounds__(BR_THREADS)
    k_bucket_reduce(const ge_ext* __restrict__ partials, const uint32_t* __restrict__ bucket_off,
                    const MsmMeta* __restrict__ meta, uint32_t bpb /* buckets per CTA, power of two <= BR_THREADS */,
                    uint32_t lv /* log2(bpb) */, ge_ext* __restrict__ blockres) {
    __shared__ ge_ext sh[BR_THREADS];
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t E = meta->E, CL = meta->CL;
    ge_ext val = ge_identity();
    SlotList L;
    L.gb = L.t_lo = L.n_pre = L.k_lo = L.n_col = L.t_post = L.n = 0;
    if (tid < bpb) {
        const uint32_t gb = blockIdx.x * bpb + tid;
        L = slot_list(gb, bucket_off[gb], bucket_off[gb + 1], E, CL);
        if (L.n) val = load_ext(partials + L.slot(0));
        const uint32_t own = min(L.n, (uint32_t)BR_CAP);
#pragma unroll 1
        for (uint32_t i = 1; i < own; i++) val = ge_add(val, load_ext(partials + L.slot(i)));
    }
    // heavy buckets: every lane of the warp takes a strided share of the slots beyond the first BR_CAP
    uint32_t heavy = __ballot_sync(0xffffffffu, L.n > BR_CAP);
    while (heavy) {
        const int owner = __ffs(heavy) - 1;
        heavy &= heavy - 1;
        SlotList H;
        H.gb = __shfl_sync(0xffffffffu, L.gb, owner);
        H.t_lo = __shfl_sync(0xffffffffu, L.t_lo, owner);
        H.n_pre = __shfl_sync(0xffffffffu, L.n_pre, owner);
        H.k_lo = __shfl_sync(0xffffffffu, L.k_lo, owner);
        H.n_col = __shfl_sync(0xffffffffu, L.n_col, owner);
        H.t_post = __shfl_sync(0xffffffffu, L.t_post, owner);
        H.n = __shfl_sync(0xffffffffu, L.n, owner);
        ge_ext part = ge_identity();
#pragma unroll 1
        for (uint32_t i = BR_CAP + lane; i < H.n; i += 32) part = ge_add(part, load_ext(partials + H.slot(i)));
#pragma unroll 1
        for (int o = 16; o > 0; o >>= 1) {
            const ge_ext other = shfl_ext(part, lane ^ o);
            part = ge_add(part, other);
        }
        if ((int)lane == owner) val = ge_add(val, part);
    }
    // butterfly
#pragma unroll 1
    for (uint32_t k = 0; k < lv; k++) {
        const uint32_t h = 1u << k, p = tid & (2 * h - 1);
        ge_ext other = val;
        if (k < 5) {
            const int src = p <= k ? (int)(lane + h) : (p == k + 1 ? (int)(lane - p + h) : (int)lane);
            other = shfl_ext(val, src);
        } else {
            store_ext(sh + tid, val);
            __syncthreads();
            if (p <= k) other = load_ext(sh + tid + h);
            else if (p == k + 1) other = load_ext(sh + tid - p + h);
            __syncthreads();
        }
        if (p <= k) val = ge_add(val, other);
        else if (p == k + 1 && k >= 2) val = other;
    }
    if (tid <= lv) store_ext(blockres + (size_t)blockIdx.x * REDUCE_MAXV + tid, val);
}

// One CTA per bucket set: the butterfly levels lv .. c-2 over the nblk segment states left by k_bucket_reduce (ping-pong
// between the two halves of `blockres`), then  result = R + sum_j 2^j M_j  (lane i doubles M_{i-1} i-1 times).
__global__ void __launch_bounds__(FINAL_THREADS)
    k_reduce_final(ge_ext* __restrict__ blockres, uint32_t nblk /* CTAs of k_bucket_reduce per set */, uint32_t lv,
                   size_t half /* points per ping-pong half */, ge_ext* __restrict__ result) {
    const uint32_t s = blockIdx.x, tid = threadIdx.x;
    ge_ext* cur = blockres + (size_t)s * nblk * REDUCE_MAXV;
    ge_ext* nxt = cur + half;
    uint32_t nseg = nblk, L = lv;  // every segment holds L + 1 values
    while (nseg > 1) {
        const uint32_t nv = L + 2, items = (nseg >> 1) * nv;
#pragma unroll 1
        for (uint32_t i = tid; i < items; i += FINAL_THREADS) {
            const uint32_t q = i / nv, v = i - q * nv;
            const ge_ext* A = cur + (size_t)(2 * q) * REDUCE_MAXV;
            const ge_ext* B = A + REDUCE_MAXV;
            ge_ext r;
            if (v == L + 1) r = load_ext(B);
            else r = ge_add(load_ext(A + v), load_ext(B + v));
            store_ext(nxt + (size_t)q * REDUCE_MAXV + v, r);
        }
        __syncthreads();
        ge_ext* tmp = cur;
        cur = nxt;
        nxt = tmp;
        nseg >>= 1;
        L++;
    }
    if (tid >= 32) return;
    ge_ext v = tid <= L ? load_ext(cur + tid) : ge_identity();
#pragma unroll 1
    for (uint32_t k = 1; k < tid && tid <= L; k++) v = ge_dbl(v);
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) {
        const ge_ext other = shfl_ext(v, (int)(tid ^ o));
        v = ge_add(v, other);
    }
    if (tid == 0) store_ext(result + s, v);
}

// ------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------
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
    const bool use_small = to_small_table(ctx->table, ctx->small_table, &segs);
    const FixedTable& tb = use_small ? ctx->small_table : ctx->table;
    if (nsets == 0 || segs.nseg > MSM_MAX_SEGMENTS) return BPG_E_ARG;
    cudaStream_t st = ctx->stream;
    const uint32_t nb = 1u << (tb.c - 1);
    const uint32_t G = nsets * nb;
    if ((uint64_t)G >= (1ull << 24) || tb.c > REDUCE_MAXV) {
        bpg_set_error("msm_run: too many buckets");
        return BPG_E_ARG;
    }
    const uint64_t total = segs.total;
    const uint64_t max_entries = (uint64_t)tb.K * total;
    if (max_entries >= (1ull << 32) - (1ull << 24) || (uint64_t)tb.K * tb.n_points >= (1ull << 31)) {
        bpg_set_error("msm_run: problem too large for 32-bit entry indices");
        return BPG_E_ARG;
    }
    // chunk geometry: CL is fixed by the knob or derived on the device from the true entry count; the host only needs
    // an upper bound on the number of chunks to size the grid and the partial-slot array
    const uint32_t cl_fixed = (uint32_t)ctx->task_len, cl_min = (uint32_t)ctx->cl_min;
    const uint32_t target = ctx->target_chunks ? (uint32_t)ctx->target_chunks : (uint32_t)ctx->sm_count * 4u * ACC_THREADS;
    uint64_t max_chunks;
    if (cl_fixed) max_chunks = (max_entries + cl_fixed - 1) / cl_fixed;
    else max_chunks = std::min<uint64_t>((max_entries + cl_min - 1) / cl_min, target);
    if (max_chunks == 0) max_chunks = 1;
    const uint32_t acc_blocks = (uint32_t)((max_chunks + ACC_THREADS - 1) / ACC_THREADS);
    const uint64_t max_partials = (uint64_t)acc_blocks * ACC_THREADS + G + 1;
    const uint32_t bpb = nb < BR_THREADS ? nb : BR_THREADS, lv = ilog2(bpb);
    const uint32_t br_blocks = G / bpb, nblk = nb / bpb;
    const size_t half = (size_t)br_blocks * REDUCE_MAXV;
    MsmWork& w = ctx->work;
    int rc;
    if ((rc = w.hist.ensure(G + 16)) || (rc = w.bucket_off.ensure(G + 16)) || (rc = w.meta.ensure(1)) ||
        (rc = w.entries.ensure(max_entries + 1)) || (rc = w.partials.ensure(max_partials)) ||
        (rc = w.blockres.ensure(2 * half)))
        return rc;

    uint32_t* tickets = nullptr;  // rank of every entry inside its bucket (unrolled K <= 16 path of k_digits only)
    if (ctx->use_tickets && tb.K <= 16 && total > 0) {
        if ((rc = w.tickets.ensure((size_t)16 * total))) return rc;
        tickets = w.tickets.p;
    }
    const bool timed = ctx->time_accum;
    int stage = 0;
    auto mark = [&]() -> cudaError_t { return timed ? cudaEventRecord(ctx->ev_stage[stage++], st) : cudaSuccess; };
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
    k_accumulate<<<acc_blocks, ACC_THREADS, 0, st>>>(tb.rows, w.entries.p, w.bucket_off.p, w.meta.p, G, w.partials.p);
    CUDA_TRY(mark());
    k_bucket_reduce<<<br_blocks, BR_THREADS, 0, st>>>(w.partials.p, w.bucket_off.p, w.meta.p, bpb, lv, w.blockres.p);
    CUDA_TRY(mark());
    k_reduce_final<<<nsets, FINAL_THREADS, 0, st>>>(w.blockres.p, nblk, lv, half, d_out);
    CUDA_TRY(mark());
    ctx->launches += 3;
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
        if (getenv("BPG_ACC_TRACE"))
            fprintf(stderr, "[bpg msm] sets %u points %llu entries %u CL %u | digits0 %.1f scan %.1f digits1 %.1f accumulate %.1f "
                    "reduce %.1f final %.1f | total %.1f us\n", nsets, (unsigned long long)total, hm.E, hm.CL, ms[0] * 1e3,
                    ms[1] * 1e3, ms[2] * 1e3, ms[3] * 1e3, ms[4] * 1e3, ms[5] * 1e3, ms[6] * 1e3);
    }
    return BPG_OK;
}
