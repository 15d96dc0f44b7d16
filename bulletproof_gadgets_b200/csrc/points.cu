// Point-side helper kernels of the R1CS driver: batched Pedersen commitments, batched ristretto
// (de)compression and the small "dynamic point" MSM of the verifier (points that are not generators:
// A_I1, A_O1, S1, V_j, T_k, L_k, R_k).
//
// Replaces dalek `PedersenGens::commit` (/root/reference/src/gadget.rs:32,
// /root/reference/src/commitments.rs:28,40 -- row a1/f2), `CompressedRistretto::decompress` inside
// `Verifier::verify` (/root/reference/src/verify.rs:71 -- row a12/K8) and the non-generator part of
// `optional_multiscalar_mul`.  Latency-bound kernels (one thread per point): kept off the critical
// path by running beside the fixed-base MSM.
#include "kernels.hpp"

// radix-16 signed digits, d[i] in [-8, 8], sum d[i] 16^i = s  (s < 2^255)
__device__ __forceinline__ void sc_radix16(const sc& s, int8_t d[64]) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) d[8 * i + k] = (int8_t)((s.v[i] >> (4 * k)) & 15u);
    }
    int carry = 0;
    for (int i = 0; i < 63; i++) {
        int v = d[i] + carry;
        carry = (v + 8) >> 4;
        d[i] = (int8_t)(v - (carry << 4));
    }
    d[63] = (int8_t)(d[63] + carry);
}

// ------------------------------------------------------------------------------------------
// Pedersen: table ped[(p*64 + w)*8 + (m-1)] = m * 16^w * P_p  (affine Niels), p in {B, B_blinding}
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_pedersen_table(const ge_ext* __restrict__ gens_ext, uint32_t idxB,
                                                        ge_niels* __restrict__ ped) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 128) return;
    const uint32_t p = t >> 6, w = t & 63;
    ge_ext base = gens_ext[idxB + p];
#pragma unroll 1
    for (uint32_t k = 0; k < 4 * w; k++) base = ge_dbl(base);
    ge_ext m = base;
#pragma unroll 1
    for (int k = 0; k < 8; k++) {
        ped[t * 8 + k] = ge_to_niels(m, fe_invert(m.Z));
        m = ge_add(m, base);
    }
}

__global__ void __launch_bounds__(64) k_pedersen(const ge_niels* __restrict__ ped, const sc* __restrict__ v,
                                                 const sc* __restrict__ r, ge_ext* __restrict__ out, uint32_t k) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    ge_ext acc = ge_identity();
#pragma unroll 1
    for (int p = 0; p < 2; p++) {
        int8_t d[64];
        sc_radix16(p ? r[i] : v[i], d);
#pragma unroll 1
        for (int w = 0; w < 64; w++) {
            int dv = d[w];
            if (dv != 0) {
                int mag = dv < 0 ? -dv : dv;
                acc = ge_madd(acc, ped[(p * 64 + w) * 8 + (mag - 1)], dv < 0);
            }
        }
    }
    out[i] = acc;
}

// A handful of commitments (the T_1 .. T_6 of a proof, the values of a small statement): one WARP per commitment.  The 128
// table additions are independent, so lane l adds its four (windows l and l + 32 of v and of r) and a shuffle tree adds the
// lanes: 4 + 5 dependent additions instead of 128 (290 us -> ~25 us for a lone warp).
__global__ void __launch_bounds__(128) k_pedersen_warp(const ge_niels* __restrict__ ped, const sc* __restrict__ v,
                                                       const sc* __restrict__ r, ge_ext* __restrict__ out, uint32_t k) {
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= k) return;  // whole warps
    ge_ext acc = ge_identity();
#pragma unroll 1
    for (int p = 0; p < 2; p++) {
        int8_t d[64];
        sc_radix16(p ? r[i] : v[i], d);
#pragma unroll 1
        for (int h = 0; h < 2; h++) {
            const int w = (int)lane + 32 * h;
            int dv = 0;
#pragma unroll
            for (int q = 0; q < 64; q++)
                if (q == w) dv = d[q];  // (keeps d[] in registers: no dynamic indexing)
            if (dv != 0) {
                const int mag = dv < 0 ? -dv : dv;
                acc = ge_madd(acc, ped[(p * 64 + w) * 8 + (mag - 1)], dv < 0);
            }
        }
    }
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) {
        ge_ext other;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            other.X.v[q] = __shfl_xor_sync(0xffffffffu, acc.X.v[q], o);
            other.Y.v[q] = __shfl_xor_sync(0xffffffffu, acc.Y.v[q], o);
            other.Z.v[q] = __shfl_xor_sync(0xffffffffu, acc.Z.v[q], o);
            other.T.v[q] = __shfl_xor_sync(0xffffffffu, acc.T.v[q], o);
        }
        acc = ge_add(acc, other);
    }
    if (lane == 0) out[i] = acc;
}

// ------------------------------------------------------------------------------------------
// codec batches
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) k_compress_batch(const ge_ext* __restrict__ in, uint8_t* __restrict__ out,
                                                       uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t b[32];
    ge_ristretto_compress(b, in[i]);
    for (int k = 0; k < 32; k++) out[32 * (size_t)i + k] = b[k];
}
__global__ void __launch_bounds__(64) k_decompress_batch(const uint8_t* __restrict__ in, ge_ext* __restrict__ out,
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

// ------------------------------------------------------------------------------------------
// dynamic-point MSM: thread per point, radix-16 signed windows, 8-entry table in local memory
// ------------------------------------------------------------------------------------------
#define DYN_THREADS 64
__device__ __forceinline__ void dyn_block_reduce(ge_ext* sh, const ge_ext& mine, uint32_t tid) {
    sh[tid] = mine;
    __syncthreads();
    for (uint32_t s = DYN_THREADS >> 1; s > 0; s >>= 1) {
        if (tid < s) sh[tid] = ge_add(sh[tid], sh[tid + s]);
        __syncthreads();
    }
}
__global__ void __launch_bounds__(DYN_THREADS) k_dyn_mul(const ge_ext* __restrict__ pts, const sc* __restrict__ s,
                                                         uint32_t n, ge_ext* __restrict__ blockres) {
    __shared__ ge_ext sh[DYN_THREADS];
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    ge_ext acc = ge_identity();
    if (i < n) {
        ge_ext tbl[8];
        tbl[0] = pts[i];
#pragma unroll 1
        for (int k = 1; k < 8; k++) tbl[k] = ge_add(tbl[k - 1], tbl[0]);
        int8_t d[64];
        sc_radix16(s[i], d);
#pragma unroll 1
        for (int w = 63; w >= 0; w--) {
            if (w != 63) {
                acc = ge_dbl(acc);
                acc = ge_dbl(acc);
                acc = ge_dbl(acc);
                acc = ge_dbl(acc);
            }
            int dv = d[w];
            if (dv != 0) {
                int mag = dv < 0 ? -dv : dv;
                ge_ext q = tbl[mag - 1];
                if (dv < 0) q = ge_neg(q);
                acc = ge_add(acc, q);
            }
        }
    }
    dyn_block_reduce(sh, acc, threadIdx.x);
    if (threadIdx.x == 0) blockres[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(DYN_THREADS) k_dyn_final(const ge_ext* __restrict__ blockres, uint32_t nblocks,
                                                           ge_ext* __restrict__ out) {
    __shared__ ge_ext sh[DYN_THREADS];
    ge_ext acc = ge_identity();
    for (uint32_t b = threadIdx.x; b < nblocks; b += DYN_THREADS) acc = ge_add(acc, blockres[b]);
    dyn_block_reduce(sh, acc, threadIdx.x);
    if (threadIdx.x == 0) *out = sh[0];
}
__global__ void k_add2(const ge_ext* a, const ge_ext* b, ge_ext* out) { *out = ge_add(*a, *b); }

// One late IPP round over FOLDED generators (dynamic points): point i < 2 nb is G'_i (i < nb) or H'_{i-nb}, its scalar
// mG[i] / mH[i-nb]; it belongs to L or to R by the round's rule (G'_j: j mod nk >= nk/2 -> L; H'_j: j mod nk < nk/2 -> L).
// Points 2 nb and 2 nb + 1 are B with the scalars cw[0] (-> L) and cw[1] (-> R).  blockres[2][gridDim.x].
__global__ void __launch_bounds__(DYN_THREADS) k_dyn_mul_lr(const ge_ext* __restrict__ gp /* [2 nb] */, const ge_ext* __restrict__ Bpt,
                                                            const sc* __restrict__ mG, const sc* __restrict__ mH,
                                                            const sc* __restrict__ cw, uint32_t nb, uint32_t nk,
                                                            ge_ext* __restrict__ blockres) {
    __shared__ ge_ext sh[DYN_THREADS];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, n = 2 * nb + 2;
    ge_ext acc = ge_identity();
    bool toL = true;
    if (i < n) {
        sc s;
        ge_ext p;
        if (i < 2 * nb) {
            const uint32_t j = (i < nb ? i : i - nb) & (nk - 1), h = nk >> 1;
            toL = i < nb ? j >= h : j < h;
            s = i < nb ? mG[i] : mH[i - nb];
            p = gp[i];
        } else {
            toL = i == 2 * nb;
            s = cw[i - 2 * nb];
            p = *Bpt;
        }
        if (!sc_is_zero(s)) {
            ge_ext tbl[8];
            tbl[0] = p;
#pragma unroll 1
            for (int k = 1; k < 8; k++) tbl[k] = ge_add(tbl[k - 1], tbl[0]);
            int8_t d[64];
            sc_radix16(s, d);
#pragma unroll 1
            for (int w = 63; w >= 0; w--) {
                if (w != 63) {
                    acc = ge_dbl_not(acc);
                    acc = ge_dbl_not(acc);
                    acc = ge_dbl_not(acc);
                    acc = ge_dbl(acc);
                }
                const int dv = d[w];
                if (dv != 0) {
                    const int mag = dv < 0 ? -dv : dv;
                    ge_ext q = tbl[mag - 1];
                    if (dv < 0) q = ge_neg(q);
                    acc = ge_add(acc, q);
                }
            }
        }
    }
    dyn_block_reduce(sh, toL ? acc : ge_identity(), threadIdx.x);
    if (threadIdx.x == 0) blockres[blockIdx.x] = sh[0];
    __syncthreads();
    dyn_block_reduce(sh, toL ? ge_identity() : acc, threadIdx.x);
    if (threadIdx.x == 0) blockres[gridDim.x + blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(DYN_THREADS) k_dyn_final_lr(const ge_ext* __restrict__ blockres, uint32_t nblocks,
                                                              ge_ext* __restrict__ out /* [2] */) {
    __shared__ ge_ext sh[DYN_THREADS];
    ge_ext acc = ge_identity();
    for (uint32_t b = threadIdx.x; b < nblocks; b += DYN_THREADS) acc = ge_add(acc, blockres[(size_t)blockIdx.x * nblocks + b]);
    dyn_block_reduce(sh, acc, threadIdx.x);
    if (threadIdx.x == 0) out[blockIdx.x] = sh[0];
}

// ------------------------------------------------------------------------------------------
// variable-base Pippenger (bpg_msm above a few hundred points): table rows straight from the encodings, window combine
// ------------------------------------------------------------------------------------------
// ristretto decoding yields an AFFINE point (Z = 1), so its Niels form (y+x, y-x, 2dxy) needs no inversion
__global__ void __launch_bounds__(64) k_decompress_niels(const uint8_t* __restrict__ in, ge_niels* __restrict__ rows,
                                                         uint32_t n, uint32_t* __restrict__ fail) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t b[32];
    for (int k = 0; k < 32; k++) b[k] = in[32 * (size_t)i + k];
    ge_ext p;
    ge_niels q = ge_niels_identity();
    if (ge_ristretto_decompress(&p, b)) {
        q.yp = fe_add(p.Y, p.X);
        q.ym = fe_sub(p.Y, p.X);
        q.t2d = fe_mul(p.T, fe_2D());
    } else {
        atomicAdd(fail, 1u);
    }
    rows[i] = q;
}
// sum_w 2^(c w) W_w by Horner from the top window (K c doublings: the serial tail of every variable-base Pippenger)
__global__ void k_window_combine(const ge_ext* __restrict__ W, int K, int c, ge_ext* __restrict__ out) {
    ge_ext acc = W[K - 1];
#pragma unroll 1
    for (int w = K - 2; w >= 0; w--) {
#pragma unroll 1
        for (int k = 0; k < c - 1; k++) acc = ge_dbl_not(acc);
        acc = ge_dbl(acc);
        acc = ge_add(acc, W[w]);
    }
    *out = acc;
}
void pk_decompress_niels(cudaStream_t st, const uint8_t* in, ge_niels* rows, uint32_t n, uint32_t* fail) {
    if (n) k_decompress_niels<<<(n + 63) / 64, 64, 0, st>>>(in, rows, n, fail);
}
void pk_window_combine(cudaStream_t st, const ge_ext* W, int K, int c, ge_ext* out) { k_window_combine<<<1, 1, 0, st>>>(W, K, c, out); }

// ------------------------------------------------------------------------------------------
void pk_pedersen_table(cudaStream_t st, const ge_ext* gens_ext, uint32_t idxB, ge_niels* ped) {
    k_pedersen_table<<<1, 128, 0, st>>>(gens_ext, idxB, ped);
}
void pk_pedersen(cudaStream_t st, const ge_niels* ped, const sc* v, const sc* r, ge_ext* out, uint32_t k) {
    if (k && k <= 256) k_pedersen_warp<<<(k + 3) / 4, 128, 0, st>>>(ped, v, r, out, k);  // latency form: a warp per commitment
    else if (k) k_pedersen<<<(k + 63) / 64, 64, 0, st>>>(ped, v, r, out, k);
}
void pk_compress(cudaStream_t st, const ge_ext* in, uint8_t* out, uint32_t n) {
    if (n) k_compress_batch<<<(n + 63) / 64, 64, 0, st>>>(in, out, n);
}
void pk_decompress(cudaStream_t st, const uint8_t* in, ge_ext* out, uint32_t n, uint32_t* fail) {
    if (n) k_decompress_batch<<<(n + 63) / 64, 64, 0, st>>>(in, out, n, fail);
}
void pk_dyn_msm(cudaStream_t st, const ge_ext* pts, const sc* s, uint32_t n, ge_ext* blockres, ge_ext* out) {
    uint32_t blocks = n ? (n + DYN_THREADS - 1) / DYN_THREADS : 1;
    k_dyn_mul<<<blocks, DYN_THREADS, 0, st>>>(pts, s, n, blockres);
    k_dyn_final<<<1, DYN_THREADS, 0, st>>>(blockres, blocks, out);
}
void pk_add2(cudaStream_t st, const ge_ext* a, const ge_ext* b, ge_ext* out) { k_add2<<<1, 1, 0, st>>>(a, b, out); }
void pk_dyn_msm_lr(cudaStream_t st, const ge_ext* gp, const ge_ext* Bpt, const sc* mG, const sc* mH, const sc* cw, uint32_t nb,
                   uint32_t nk, ge_ext* blockres, ge_ext* out2) {
    const uint32_t blocks = (2 * nb + 2 + DYN_THREADS - 1) / DYN_THREADS;
    k_dyn_mul_lr<<<blocks, DYN_THREADS, 0, st>>>(gp, Bpt, mG, mH, cw, nb, nk, blockres);
    k_dyn_final_lr<<<2, DYN_THREADS, 0, st>>>(blockres, blocks, out2);
}
