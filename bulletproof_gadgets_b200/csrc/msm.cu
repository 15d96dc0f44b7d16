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
//   2 scan               bucket offsets (multi-CTA exclusive scan)
//   3 digits+scatter     counting-sort of (row | sign) entries by bucket                  [HBM/L2]
//   4 chunks             bucket of the first entry of every fixed-length chunk
//   5 accumulate         one thread per chunk of exactly `task_len` sorted entries: mixed adds of gathered
//                        Niels rows; a chunk that crosses bucket boundaries emits one partial per bucket,
//                        so every lane runs the same number of additions (no divergence)   [IMAD]
//   6 reduce             weighted running sums over bucket partials, block tree, final tree
#include <stdlib.h>

#include "circuit.hpp"
#include "ctx.hpp"

#define FINAL_THREADS 128
#define SMALL_MSM_MAX_POINTS 4096u  // MSMs up to 2 x this many points use the 8-bit-window table
#define ACC_THREADS 128

// ------------------------------------------------------------------------------------------
// stage 1 / 3: signed digit decomposition
// ------------------------------------------------------------------------------------------
template <bool SCATTER>
__global__ void __launch_bounds__(256) k_digits(MsmSegments segs, int c, int K, uint32_t nb, uint32_t n_points,
                                                uint32_t* __restrict__ hist, const uint32_t* __restrict__ bucket_off,
                                                uint32_t* __restrict__ entries, uint32_t* __restrict__ tickets) {
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
    if (K <= 16) {
        // all digits first, then all atomics back to back (16 independent L2 round trips in flight instead of a
        // dependent chain), then the scattered stores
        uint32_t gbv[16], entv[16];
#pragma unroll
        for (int w = 0; w < 16; w++) {
            gbv[w] = 0xffffffffu;
            entv[w] = 0;
            if (w < K) {
                uint32_t raw = (s[0] & mask) + carry;
#pragma unroll
                for (int k = 0; k < 7; k++) s[k] = __funnelshift_r(s[k], s[k + 1], c);
                s[7] >>= c;
                const uint32_t neg = raw > half;
                const uint32_t mag = neg ? ((1u << c) - raw) : raw;
                carry = neg;
                if (mag != 0) {
                    gbv[w] = set * nb + (mag - 1);
                    entv[w] = ((uint32_t)w * n_points + point) | (neg << 31);
                }
            }
        }
        // With `tickets` the histogram pass keeps what its atomicAdd returns -- the rank of the entry inside its bucket
        // -- at [w][g] (coalesced), and the scatter pass is bucket_off + rank: no second round of atomics.
        if (!SCATTER) {
            if (tickets) {
                uint32_t tk[16];
#pragma unroll
                for (int w = 0; w < 16; w++)
                    if (gbv[w] != 0xffffffffu) tk[w] = atomicAdd(&hist[gbv[w]], 1u);
#pragma unroll
                for (int w = 0; w < 16; w++)
                    if (gbv[w] != 0xffffffffu) tickets[(size_t)w * segs.total + g] = tk[w];
            } else {
#pragma unroll
                for (int w = 0; w < 16; w++)
                    if (gbv[w] != 0xffffffffu) atomicAdd(&hist[gbv[w]], 1u);
            }
        } else {
            uint32_t pos[16];
            if (tickets) {
#pragma unroll
                for (int w = 0; w < 16; w++)
                    if (gbv[w] != 0xffffffffu) pos[w] = bucket_off[gbv[w]] + tickets[(size_t)w * segs.total + g];
            } else {
#pragma unroll
                for (int w = 0; w < 16; w++)
                    if (gbv[w] != 0xffffffffu) pos[w] = bucket_off[gbv[w]] + atomicAdd(&hist[gbv[w]], 1u);
            }
#pragma unroll
            for (int w = 0; w < 16; w++)
                if (gbv[w] != 0xffffffffu) entries[pos[w]] = entv[w];
        }
        return;
    }
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
// stage 4: chunk t covers sorted entries [t*CL, (t+1)*CL); chunk_bucket[t] = bucket holding entry t*CL
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_chunks(const uint32_t* __restrict__ bucket_off, uint32_t* __restrict__ chunk_bucket,
                                                uint32_t G, uint32_t CL) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= G) return;
    const uint32_t lo = bucket_off[b], hi = bucket_off[b + 1];
    for (uint32_t t = (lo + CL - 1) / CL; t * CL < hi; t++) chunk_bucket[t] = b;
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

// partial slot of (chunk t, bucket b) = t + b: buckets are sorted along the entries, so the sum is unique, and the
// slots of one bucket are contiguous: chunks off[b]/CL .. (off[b+1]-1)/CL.
__global__ void __launch_bounds__(ACC_THREADS, 4)
    k_accumulate(const ge_niels* __restrict__ rows, const uint32_t* __restrict__ entries,
                 const uint32_t* __restrict__ bucket_off, const uint32_t* __restrict__ chunk_bucket, uint32_t G, uint32_t CL,
                 ge_ext* __restrict__ partials, uint32_t* __restrict__ slot_bucket) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t E = bucket_off[G];
    const uint32_t e0 = t * CL;
    if (e0 >= E) return;
    const uint32_t e1 = min(e0 + CL, E);
    uint32_t b = chunk_bucket[t];
    uint32_t next = bucket_off[b + 1];
    ge_ext acc = ge_identity();
    uint32_t ent = __ldg(entries + e0);
#pragma unroll 1
    for (uint32_t e = e0; e < e1; e++) {
        if (e == next) {  // bucket boundary inside the chunk: emit the partial of the finished bucket
            store_ext(partials + t + b, acc);
            slot_bucket[t + b] = b;
            acc = ge_identity();
            b++;
            next = bucket_off[b + 1];
            if (next == e) {  // run of empty buckets (sparse digit sets): bucket of entry e by bisection
                uint32_t lo = b + 1, hi = G;  // smallest j in (b, G] with bucket_off[j] > e; entry e lives in bucket j-1
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (bucket_off[mid] > e) hi = mid; else lo = mid + 1;
                }
                b = lo - 1;
                next = bucket_off[lo];
            }
        }
        ge_niels q = load_niels(rows, ent & 0x7fffffffu);
        const bool neg = ent >> 31;
        if (e + 1 < e1) ent = __ldg(entries + e + 1);
        acc = ge_madd(acc, q, neg);
    }
    store_ext(partials + t + b, acc);
    slot_bucket[t + b] = b;
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

// One thread per contiguous range of partial SLOTS of one set (slots are ordered by bucket; a heavy bucket -- e.g. the
// digit-1 bucket of a 0/1-valued a_L vector -- spreads over many threads).  Walking down from the top slot with
// running sums:  invariant  S + w_prev * R = sum of weight * partial over the slots seen so far.
#define SLOT_EMPTY 0xffffffffu
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
    k_reduce_chunks(const ge_ext* __restrict__ partials, const uint32_t* __restrict__ slot_bucket,
                    const uint32_t* __restrict__ bucket_off, uint32_t nb, uint32_t nsets, uint32_t CL,
                    ge_ext* __restrict__ blockres) {
    __shared__ ge_ext sh[THREADS];
    const uint32_t s = blockIdx.y;
    const uint32_t G = nb * nsets, base = s * nb;
    const uint32_t E = bucket_off[G];
    const uint32_t p0 = bucket_off[base] / CL + base;
    const uint32_t p1 = s + 1 < nsets ? bucket_off[base + nb] / CL + base + nb : (E + CL - 1) / CL + G;
    const uint32_t nthreads = gridDim.x * THREADS;
    const uint32_t per = (p1 - p0 + nthreads - 1) / nthreads;
    const uint32_t cidx = blockIdx.x * THREADS + threadIdx.x;
    const uint64_t lo64 = (uint64_t)p0 + (uint64_t)cidx * per;
    ge_ext total = ge_identity();
    if (per > 0 && lo64 < p1) {
        const uint32_t lo = (uint32_t)lo64, hi = min(lo + per, p1);
        ge_ext R = ge_identity(), S = ge_identity();
        uint32_t wprev = 0;
#pragma unroll 1
        for (uint32_t p = hi; p-- > lo;) {
            const uint32_t b = slot_bucket[p];
            if (b == SLOT_EMPTY || b < base || b >= base + nb) continue;
            const uint32_t w = b - base + 1;
            if (wprev == 0) {
                wprev = w;
                R = load_ext(partials + p);
                continue;
            }
            if (w != wprev) {
                const uint32_t gap = wprev - w;
                S = ge_add(S, gap == 1 ? R : ge_mul_small(R, gap));
                wprev = w;
            }
            R = ge_add(R, load_ext(partials + p));
        }
        if (wprev) total = ge_add(S, ge_mul_small(R, wprev));
    }
    block_tree_reduce(sh, total, threadIdx.x, THREADS);
    if (threadIdx.x == 0) store_ext(blockres + s * gridDim.x + blockIdx.x, load_ext(sh));
}

__global__ void __launch_bounds__(FINAL_THREADS) k_reduce_final(const ge_ext* __restrict__ blockres, uint32_t nblocks,
                                                                 ge_ext* __restrict__ result) {
    __shared__ ge_ext sh[FINAL_THREADS];
    const uint32_t s = blockIdx.x;
    ge_ext mine = threadIdx.x < nblocks ? load_ext(blockres + s * nblocks + threadIdx.x) : ge_identity();
#pragma unroll 1
    for (uint32_t k = threadIdx.x + FINAL_THREADS; k < nblocks; k += FINAL_THREADS)
        mine = ge_add(mine, load_ext(blockres + s * nblocks + k));
    block_tree_reduce(sh, mine, threadIdx.x, FINAL_THREADS);
    if (threadIdx.x == 0) store_ext(result + s, load_ext(sh));
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
    if ((uint64_t)G >= (1ull << 24)) {
        bpg_set_error("msm_run: too many buckets");
        return BPG_E_ARG;
    }
    const uint64_t total = segs.total;
    const uint64_t max_entries = (uint64_t)tb.K * total;
    const uint32_t T = (uint32_t)ctx->task_len;
    const uint64_t max_chunks = max_entries / T + 1;
    const uint64_t max_partials = max_chunks + G + 1;
    if (max_entries >= (1ull << 32) || (uint64_t)tb.K * tb.n_points >= (1ull << 31)) {
        bpg_set_error("msm_run: problem too large for 32-bit entry indices");
        return BPG_E_ARG;
    }
    MsmWork& w = ctx->work;
    int rc;
    if ((rc = w.hist.ensure(G + 1)) || (rc = w.bucket_off.ensure(G + 2)) || (rc = w.chunk_bucket.ensure(max_chunks + 1)) ||
        (rc = w.entries.ensure(max_entries + 1)) || (rc = w.partials.ensure(max_partials)) || (rc = w.slot_bucket.ensure(max_partials)) ||
        (rc = w.blockres.ensure((size_t)nsets * REDUCE_BLOCKS_MAX)) || (rc = w.scan_tmp.ensure(G / 2048 + 4)))
        return rc;

    uint32_t* tickets = nullptr;  // rank of every entry inside its bucket (unrolled K <= 16 path of k_digits only)
    if (ctx->use_tickets && tb.K <= 16 && total > 0) {
        if ((rc = w.tickets.ensure((size_t)16 * total))) return rc;
        tickets = w.tickets.p;
    }
    CUDA_TRY(cudaMemsetAsync(w.hist.p, 0, (size_t)G * 4, st));
    if (total > 0) {
        const uint32_t blocks = (uint32_t)((total + 255) / 256);
        k_digits<false><<<blocks, 256, 0, st>>>(segs, tb.c, tb.K, nb, tb.n_points, w.hist.p, nullptr, nullptr, tickets);
        ctx->launches++;
    }
    dev_exclusive_scan_u32(st, w.hist.p, w.bucket_off.p, G, w.scan_tmp.p);
    if (!tickets) CUDA_TRY(cudaMemsetAsync(w.hist.p, 0, (size_t)G * 4, st));  // becomes the scatter cursor
    ctx->launches += tickets ? 2 : 3;
    if (total > 0) {
        const uint32_t blocks = (uint32_t)((total + 255) / 256);
        if (ctx->time_accum) CUDA_TRY(cudaEventRecord(ctx->ev_c, st));
        k_digits<true><<<blocks, 256, 0, st>>>(segs, tb.c, tb.K, nb, tb.n_points, w.hist.p, w.bucket_off.p,
                                              w.entries.p, tickets);
        if (ctx->time_accum) CUDA_TRY(cudaEventRecord(ctx->ev_d, st));
        ctx->launches++;
    }
    k_chunks<<<(G + 255) / 256, 256, 0, st>>>(w.bucket_off.p, w.chunk_bucket.p, G, T);
    ctx->launches++;
    CUDA_TRY(cudaMemsetAsync(w.slot_bucket.p, 0xff, (size_t)max_partials * 4, st));
    {
        const uint32_t blocks = (uint32_t)((max_chunks + ACC_THREADS - 1) / ACC_THREADS);
        if (ctx->time_accum) CUDA_TRY(cudaEventRecord(ctx->ev_a, st));
        k_accumulate<<<blocks, ACC_THREADS, 0, st>>>(tb.rows, w.entries.p, w.bucket_off.p, w.chunk_bucket.p, G, T, w.partials.p,
                                                     w.slot_bucket.p);
        if (ctx->time_accum) CUDA_TRY(cudaEventRecord(ctx->ev_b, st));
        ctx->launches++;
    }
    const uint32_t rblocks = (uint32_t)ctx->reduce_blocks;
    if (ctx->reduce_threads == 32)
        k_reduce_chunks<32><<<dim3(rblocks, nsets), 32, 0, st>>>(w.partials.p, w.slot_bucket.p, w.bucket_off.p, nb, nsets, T, w.blockres.p);
    else
        k_reduce_chunks<64><<<dim3(rblocks, nsets), 64, 0, st>>>(w.partials.p, w.slot_bucket.p, w.bucket_off.p, nb, nsets, T, w.blockres.p);
    k_reduce_final<<<nsets, FINAL_THREADS, 0, st>>>(w.blockres.p, rblocks, d_out);
    ctx->launches += 2;
    CUDA_TRY(cudaGetLastError());
    if (ctx->time_accum) {  // diagnostic mode: synchronous, reads back the entry count
        CUDA_TRY(ctx_sync(ctx));
        CUDA_TRY(cudaEventElapsedTime(&ctx->last_accum_ms, ctx->ev_a, ctx->ev_b));
        uint32_t ne = 0;
        CUDA_TRY(cudaMemcpy(&ne, w.bucket_off.p + G, 4, cudaMemcpyDeviceToHost));
        ctx->last_entries = ne;
        ctx->sum_accum_ms += ctx->last_accum_ms;
        ctx->sum_entries += ne;
        if (total > 0) {
            float ms = 0.f;
            CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev_c, ctx->ev_d));
            ctx->sum_scatter_ms += ms;
            ctx->sum_points += total;
        }
        if (getenv("BPG_ACC_TRACE"))
            fprintf(stderr, "[bpg acc] sets %u points %llu entries %u accumulate %.1f us\n", nsets, (unsigned long long)total, ne,
                    ctx->last_accum_ms * 1e3);
    }
    return BPG_OK;
}
