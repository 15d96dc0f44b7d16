// Batched mod-l scalar kernels of the R1CS prover / verifier (HBM-bound element-wise and reduce
// kernels; vectors never leave the device).
//
// Replaces, from the FairAds fork of dalek bulletproofs 2.1.0 (/root/reference/Cargo.lock:78-80, not
// vendored; reached from /root/reference/src/prove.rs:79 and /root/reference/src/verify.rs:71):
//   r1cs/prover.rs + verifier.rs  flattened_constraints            -> k_flatten          (row a4 / K6)
//   util.rs exp_iter, VecPoly3::special_inner_product / eval        -> k_powers, k_lr_poly, k_eval_lr (a5, a7 / K7)
//   inner_product_proof.rs create: c_L/c_R, scalar folds            -> k_ipp_* (a8)
//   inner_product_proof.rs verification_scalars + verifier.rs g/h   -> k_ver_scalars, k_ver_head (a11)
#include "kernels.hpp"

#define SK_THREADS 256

__device__ __forceinline__ sc ld_sc(const sc* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    sc r;
    r.v[0] = a.x, r.v[1] = a.y, r.v[2] = a.z, r.v[3] = a.w, r.v[4] = b.x, r.v[5] = b.y, r.v[6] = b.z, r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void st_sc(sc* p, const sc& s) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(s.v[0], s.v[1], s.v[2], s.v[3]);
    q[1] = make_uint4(s.v[4], s.v[5], s.v[6], s.v[7]);
}

// sum over the block; result valid in thread 0
__device__ sc block_sum(sc v, sc* sh /*[SK_THREADS]*/) {
    const uint32_t tid = threadIdx.x;
    sh[tid] = v;
    __syncthreads();
    for (uint32_t s = SK_THREADS >> 1; s > 0; s >>= 1) {
        if (tid < s) sh[tid] = sc_add(sh[tid], sh[tid + s]);
        __syncthreads();
    }
    sc r = sh[0];
    __syncthreads();
    return r;
}

// out[i] = base^(start + i).  Thread t computes base^(start + t) from the table of base^(2^k) (<= popcount muls) and then
// walks out[t + j*T] = out[t + (j-1)*T] * base^T with T = 2^lgT threads: coalesced stores, ~3 multiplications per output
// instead of one square-and-multiply ladder per output.
__global__ void __launch_bounds__(SK_THREADS) k_powers(sc* out, PowTable tbl, uint32_t n, uint32_t start, uint32_t lgT) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t T = 1u << lgT;
    if (t >= T || t >= n) return;
    uint32_t e = start + t;
    sc acc = sc_one();
    bool first = true;
    for (int k = 0; k < 32 && (e >> k); k++) {
        if ((e >> k) & 1u) {
            acc = first ? tbl.p[k] : sc_mul(acc, tbl.p[k]);
            first = false;
        }
    }
    const sc step = tbl.p[lgT];
    for (uint32_t i = t; i < n; i += T) {
        st_sc(out + i, acc);
        if (i + T < n) acc = sc_mul(acc, step);
    }
}

__global__ void __launch_bounds__(SK_THREADS) k_powers_multi(const __grid_constant__ PowJobs jobs) {
    const uint32_t w = blockIdx.y;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = jobs.n[w], lgT = jobs.lgT[w], T = 1u << lgT;
    if (t >= T || t >= n) return;
    uint32_t e = jobs.start[w] + t;
    sc acc = sc_one();
    bool first = true;
    for (int k = 0; k < 32 && (e >> k); k++) {
        if ((e >> k) & 1u) {
            acc = first ? jobs.tbl[w].p[k] : sc_mul(acc, jobs.tbl[w].p[k]);
            first = false;
        }
    }
    const sc step = jobs.tbl[w].p[lgT];
    sc* out = jobs.out[w];
    for (uint32_t i = t; i < n; i += T) {
        st_sc(out + i, acc);
        if (i + T < n) acc = sc_mul(acc, step);
    }
}

__global__ void __launch_bounds__(SK_THREADS) k_flatten(const uint32_t* __restrict__ col_start,
                                                        const uint32_t* __restrict__ col_row,
                                                        const sc* __restrict__ col_coef, const sc* __restrict__ zpow,
                                                        sc* __restrict__ out, uint32_t nt, uint32_t neg_from) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    sc acc = sc_zero();
    const uint32_t e0 = col_start[t], e1 = col_start[t + 1];
    if (e1 - e0 > FLATTEN_LONG) return;  // handled by k_flatten_long (one CTA per long column)
    for (uint32_t e = e0; e < e1; e++) acc = sc_add(acc, sc_mul(ld_sc(col_coef + e), ld_sc(zpow + col_row[e])));
    if (t >= neg_from) acc = sc_neg(acc);
    st_sc(out + t, acc);
}

// Columns with more than FLATTEN_LONG terms (e.g. the `One` column of a range-proof circuit: one term per bit).
// Each long column is split over FL_SPLIT CTAs; the last CTA to finish (ticket counter) adds the partial sums.
#define FL_SPLIT 64
#define FL_THREADS 256
__global__ void __launch_bounds__(FL_THREADS) k_flatten_long(const uint32_t* __restrict__ col_start,
                                                            const uint32_t* __restrict__ col_row,
                                                            const sc* __restrict__ col_coef, const sc* __restrict__ zpow,
                                                            sc* __restrict__ out, const uint32_t* __restrict__ long_targets,
                                                            uint32_t nt, uint32_t neg_from, sc* __restrict__ part,
                                                            uint32_t* __restrict__ tickets) {
    __shared__ sc sh[FL_THREADS];
    __shared__ bool last;
    const uint32_t j = blockIdx.x, sidx = blockIdx.y;
    const uint32_t t = long_targets[j];
    if (t >= nt) return;  // (uniform per block)
    const uint32_t e0 = col_start[t], e1 = col_start[t + 1], tid = threadIdx.x;
    const uint32_t len = e1 - e0, per = (len + FL_SPLIT - 1) / FL_SPLIT;
    const uint32_t a = e0 + sidx * per, b = min(a + per, e1);
    sc acc = sc_zero();
    for (uint32_t e = a + tid; e < b; e += FL_THREADS) acc = sc_add(acc, sc_mul(ld_sc(col_coef + e), ld_sc(zpow + col_row[e])));
    sh[tid] = acc;
    __syncthreads();
    for (uint32_t s = FL_THREADS / 2; s > 0; s >>= 1) {
        if (tid < s) sh[tid] = sc_add(sh[tid], sh[tid + s]);
        __syncthreads();
    }
    if (tid == 0) {
        st_sc(part + j * FL_SPLIT + sidx, sh[0]);
        __threadfence();
        last = atomicAdd(&tickets[j], 1u) == FL_SPLIT - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    sc v = tid < FL_SPLIT ? ld_sc(part + j * FL_SPLIT + tid) : sc_zero();
    sh[tid] = v;
    __syncthreads();
    for (uint32_t s = FL_THREADS / 2; s > 0; s >>= 1) {
        if (tid < s) sh[tid] = sc_add(sh[tid], sh[tid + s]);
        __syncthreads();
    }
    if (tid == 0) {
        st_sc(out + t, t >= neg_from ? sc_neg(sh[0]) : sh[0]);
        tickets[j] = 0;  // ready for the next launch
    }
}

// l1 = aL + yinv^i wR ; r0 = wO - y^i ; r1 = y^i aR + wL ; r3 = y^i sR   (l2 = aO, l3 = sL)
// partial[block][6] = block sums of the t1..t6 integrands
__global__ void __launch_bounds__(SK_THREADS)
    k_lr_poly(const sc* aL, const sc* aR, const sc* aO, const sc* sL, const sc* sR, const sc* wL, const sc* wR,
              const sc* wO, const sc* ypow, const sc* yinv, sc* l1, sc* r0, sc* r1, sc* r3, sc* partial, uint32_t n) {
    __shared__ sc sh[SK_THREADS];
    sc t[6];
#pragma unroll
    for (int k = 0; k < 6; k++) t[k] = sc_zero();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const sc yp = ld_sc(ypow + i);
        const sc L1 = sc_add(ld_sc(aL + i), sc_mul(ld_sc(yinv + i), ld_sc(wR + i)));
        const sc L2 = ld_sc(aO + i), L3 = ld_sc(sL + i);
        const sc R0 = sc_sub(ld_sc(wO + i), yp);
        const sc R1 = sc_add(sc_mul(yp, ld_sc(aR + i)), ld_sc(wL + i));
        const sc R3 = sc_mul(yp, ld_sc(sR + i));
        st_sc(l1 + i, L1);
        st_sc(r0 + i, R0);
        st_sc(r1 + i, R1);
        st_sc(r3 + i, R3);
        t[0] = sc_add(t[0], sc_mul(L1, R0));
        t[1] = sc_add(t[1], sc_add(sc_mul(L1, R1), sc_mul(L2, R0)));
        t[2] = sc_add(t[2], sc_add(sc_mul(L2, R1), sc_mul(L3, R0)));
        t[3] = sc_add(t[3], sc_add(sc_mul(L1, R3), sc_mul(L3, R1)));
        t[4] = sc_add(t[4], sc_mul(L2, R3));
        t[5] = sc_add(t[5], sc_mul(L3, R3));
    }
    for (int k = 0; k < 6; k++) {
        sc s = block_sum(t[k], sh);
        if (threadIdx.x == 0) st_sc(partial + blockIdx.x * 6 + k, s);
    }
}

// out[k] = mult * sum_b partial[b*ns + k]   (single block)
__global__ void __launch_bounds__(SK_THREADS) k_sum_sc(const sc* partial, uint32_t nblocks, uint32_t ns, sc mult,
                                                       sc* out) {
    __shared__ sc sh[SK_THREADS];
    for (uint32_t k = 0; k < ns; k++) {
        sc acc = sc_zero();
        for (uint32_t b = threadIdx.x; b < nblocks; b += SK_THREADS) acc = sc_add(acc, ld_sc(partial + b * ns + k));
        sc s = block_sum(acc, sh);
        if (threadIdx.x == 0) st_sc(out + k, sc_mul(s, mult));
    }
}

// l = x(l1 + x(l2 + x l3)), r = r0 + x(r1 + x^2 r3); padding: l = 0, r = -y^i
__global__ void __launch_bounds__(SK_THREADS)
    k_eval_lr(const sc* l1, const sc* aO, const sc* sL, const sc* r0, const sc* r1, const sc* r3, const sc* ypow, sc x,
              sc* lvec, sc* rvec, uint32_t n, uint32_t npad) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    if (i < n) {
        sc l = sc_mul(x, sc_add(ld_sc(l1 + i), sc_mul(x, sc_add(ld_sc(aO + i), sc_mul(x, ld_sc(sL + i))))));
        sc r = sc_add(ld_sc(r0 + i), sc_mul(x, sc_add(ld_sc(r1 + i), sc_mul(x, sc_mul(x, ld_sc(r3 + i))))));
        st_sc(lvec + i, l);
        st_sc(rvec + i, r);
    } else {
        st_sc(lvec + i, sc_zero());
        st_sc(rvec + i, sc_neg(ld_sc(ypow + i)));
    }
}

// IPP generator coefficients: sG_i = G_factors_i, sH_i = H_factors_i = yinv^i * G_factors_i
__global__ void __launch_bounds__(SK_THREADS) k_ipp_init(sc* sG, sc* sH, const sc* yinv, sc u, uint32_t n,
                                                         uint32_t npad) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    sc gf = i < n ? sc_one() : u;
    st_sc(sG + i, gf);
    st_sc(sH + i, i < n ? ld_sc(yinv + i) : sc_mul(ld_sc(yinv + i), u));
}

// partial[block][2] = block sums of a_lo*b_hi (c_L) and a_hi*b_lo (c_R)
__global__ void __launch_bounds__(SK_THREADS) k_ipp_cross(const sc* a, const sc* b, uint32_t h, sc* partial) {
    __shared__ sc sh[SK_THREADS];
    sc cl = sc_zero(), cr = sc_zero();
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < h; j += gridDim.x * blockDim.x) {
        cl = sc_add(cl, sc_mul(ld_sc(a + j), ld_sc(b + j + h)));
        cr = sc_add(cr, sc_mul(ld_sc(a + j + h), ld_sc(b + j)));
    }
    sc s0 = block_sum(cl, sh);
    if (threadIdx.x == 0) st_sc(partial + blockIdx.x * 2, s0);
    sc s1 = block_sum(cr, sh);
    if (threadIdx.x == 0) st_sc(partial + blockIdx.x * 2 + 1, s1);
}

// MSM scalars of the round over the ORIGINAL generators (folds are kept in sG/sH, points never move):
//   G_i: j = i mod nk;  j >= h -> a[j-h] * sG_i (term of L)   else a[j+h] * sG_i (term of R)
//   H_i:                j <  h -> b[j+h] * sH_i (term of L)   else b[j-h] * sH_i (term of R)
__global__ void __launch_bounds__(SK_THREADS) k_ipp_scalars(const sc* a, const sc* b, const sc* sG, const sc* sH,
                                                            sc* mG, sc* mH, uint32_t npad, uint32_t nk) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    const uint32_t h = nk >> 1, j = i & (nk - 1);
    const uint32_t partner = j >= h ? j - h : j + h;
    st_sc(mG + i, sc_mul(ld_sc(a + partner), ld_sc(sG + i)));
    st_sc(mH + i, sc_mul(ld_sc(b + partner), ld_sc(sH + i)));
}

// a'_j = a_j u + a_{j+h} u^-1 ; b'_j = b_j u^-1 + b_{j+h} u ; sG_i *= (j>=h ? u : u^-1) ; sH_i *= (j>=h ? u^-1 : u)
__global__ void __launch_bounds__(SK_THREADS) k_ipp_fold(sc* a, sc* b, sc* sG, sc* sH, sc u, sc uinv, uint32_t npad,
                                                         uint32_t nk) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    const uint32_t h = nk >> 1, j = i & (nk - 1);
    const bool hi = j >= h;
    st_sc(sG + i, sc_mul(ld_sc(sG + i), hi ? u : uinv));
    st_sc(sH + i, sc_mul(ld_sc(sH + i), hi ? uinv : u));
    if (i < h) {
        sc na = sc_add(sc_mul(ld_sc(a + i), u), sc_mul(ld_sc(a + i + h), uinv));
        sc nb = sc_add(sc_mul(ld_sc(b + i), uinv), sc_mul(ld_sc(b + i + h), u));
        st_sc(a + i, na);
        st_sc(b + i, nb);
    }
}

// Cross terms, their sum and the MSM scalars of a round in ONE launch (three before): the first `cb` CTAs also accumulate the
// cross terms, and the last of them to finish (ticket) adds the block sums and writes c_L w, c_R w.
__global__ void __launch_bounds__(SK_THREADS)
    k_ipp_round_big(const sc* a, const sc* b, const sc* sG, const sc* sH, sc* mG, sc* mH, sc* partial, uint32_t* ticket, sc* cw_out, sc w,
                    uint32_t cb, uint32_t npad, uint32_t nk) {
    __shared__ sc sh[SK_THREADS];
    __shared__ bool last;
    const uint32_t h = nk >> 1, tid = threadIdx.x;
    const uint32_t i = blockIdx.x * blockDim.x + tid;
    if (i < npad) {
        const uint32_t j = i & (nk - 1);
        const uint32_t partner = j >= h ? j - h : j + h;
        st_sc(mG + i, sc_mul(ld_sc(a + partner), ld_sc(sG + i)));
        st_sc(mH + i, sc_mul(ld_sc(b + partner), ld_sc(sH + i)));
    }
    if (blockIdx.x >= cb) return;  // (uniform per block)
    sc cl = sc_zero(), cr = sc_zero();
    for (uint32_t j = blockIdx.x * blockDim.x + tid; j < h; j += cb * blockDim.x) {
        cl = sc_add(cl, sc_mul(ld_sc(a + j), ld_sc(b + j + h)));
        cr = sc_add(cr, sc_mul(ld_sc(a + j + h), ld_sc(b + j)));
    }
    const sc s0 = block_sum(cl, sh);
    const sc s1 = block_sum(cr, sh);
    if (tid == 0) {
        st_sc(partial + blockIdx.x * 2, s0);
        st_sc(partial + blockIdx.x * 2 + 1, s1);
        __threadfence();
        last = atomicAdd(ticket, 1u) == cb - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    for (uint32_t k = 0; k < 2; k++) {
        sc acc = sc_zero();
        for (uint32_t q = tid; q < cb; q += SK_THREADS) acc = sc_add(acc, ld_sc(partial + q * 2 + k));
        const sc t = block_sum(acc, sh);
        if (tid == 0) st_sc(cw_out + k, sc_mul(t, w));
    }
    if (tid == 0) *ticket = 0;  // ready for the next launch
}

// Small statements (npad <= IPP_SMALL_MAX): the previous round's fold and this round's cross terms and MSM scalars in ONE
// single-CTA launch instead of four (a small proof is bound by its driver calls).  Same arithmetic, element for element.
__global__ void __launch_bounds__(SK_THREADS)
    k_ipp_round_small(sc* a, sc* b, sc* sG, sc* sH, sc* mG, sc* mH, sc* cw_out, sc w, sc u, sc uinv, int do_fold,
                      uint32_t npad, uint32_t nk) {
    __shared__ sc sh[SK_THREADS];
    if (do_fold) {  // k_ipp_fold of the round before (vector length 2 nk then)
        const uint32_t nkp = nk << 1, hp = nk;
        for (uint32_t i = threadIdx.x; i < npad; i += SK_THREADS) {
            const bool hi = (i & (nkp - 1)) >= hp;
            st_sc(sG + i, sc_mul(ld_sc(sG + i), hi ? u : uinv));
            st_sc(sH + i, sc_mul(ld_sc(sH + i), hi ? uinv : u));
        }
        // in place: element i < hp is read and written by its own thread only, elements >= hp are only read
        for (uint32_t i = threadIdx.x; i < hp; i += SK_THREADS) {
            const sc na = sc_add(sc_mul(ld_sc(a + i), u), sc_mul(ld_sc(a + i + hp), uinv));
            const sc nb_ = sc_add(sc_mul(ld_sc(b + i), uinv), sc_mul(ld_sc(b + i + hp), u));
            st_sc(a + i, na);
            st_sc(b + i, nb_);
        }
        __syncthreads();  // the block's own global stores are visible to all of its threads below
    }
    const uint32_t h = nk >> 1;
    sc cl = sc_zero(), cr = sc_zero();
    for (uint32_t j = threadIdx.x; j < h; j += SK_THREADS) {
        cl = sc_add(cl, sc_mul(ld_sc(a + j), ld_sc(b + j + h)));
        cr = sc_add(cr, sc_mul(ld_sc(a + j + h), ld_sc(b + j)));
    }
    const sc s0 = block_sum(cl, sh);
    if (threadIdx.x == 0) st_sc(cw_out, sc_mul(s0, w));
    const sc s1 = block_sum(cr, sh);
    if (threadIdx.x == 0) st_sc(cw_out + 1, sc_mul(s1, w));
    for (uint32_t i = threadIdx.x; i < npad; i += SK_THREADS) {
        const uint32_t j = i & (nk - 1);
        const uint32_t partner = j >= h ? j - h : j + h;
        st_sc(mG + i, sc_mul(ld_sc(a + partner), ld_sc(sG + i)));
        st_sc(mH + i, sc_mul(ld_sc(b + partner), ld_sc(sH + i)));
    }
}

// ---------------------------------------------------------------------------- verifier
// s_i = prod_j (bit_{lg-1-j}(i) ? u_j : u_j^-1);  g_i = uf_i (x yinv^i wR_i - a s_i);
// h_i = uf_i (yinv^i (x wL_i + wO_i - b s_{npad-1-i}) - 1);  partial[block] = sum yinv^i wR_i wL_i
__global__ void __launch_bounds__(SK_THREADS)
    k_ver_scalars(VerChallenges ch, const sc* wL, const sc* wR, const sc* wO, const sc* yinv, sc* gs, sc* hs,
                  sc* partial, uint32_t n, uint32_t npad, uint32_t lg) {
    __shared__ sc sh[SK_THREADS];
    sc dl = sc_zero();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < npad; i += gridDim.x * blockDim.x) {
        sc s = sc_one(), srev = sc_one();
        const uint32_t irev = npad - 1 - i;
        for (uint32_t j = 0; j < lg; j++) {
            const uint32_t bit = lg - 1 - j;
            s = sc_mul(s, ((i >> bit) & 1u) ? ch.u[j] : ch.uinv[j]);
            srev = sc_mul(srev, ((irev >> bit) & 1u) ? ch.u[j] : ch.uinv[j]);
        }
        const sc yi = ld_sc(yinv + i);
        sc g, hh;
        if (i < n) {
            const sc ywr = sc_mul(yi, ld_sc(wR + i));
            const sc wl = ld_sc(wL + i);
            dl = sc_add(dl, sc_mul(ywr, wl));
            g = sc_sub(sc_mul(ch.x, ywr), sc_mul(ch.a, s));
            hh = sc_sub(sc_mul(yi, sc_sub(sc_add(sc_mul(ch.x, wl), ld_sc(wO + i)), sc_mul(ch.b, srev))), sc_one());
        } else {
            g = sc_mul(ch.u_pad, sc_neg(sc_mul(ch.a, s)));
            hh = sc_mul(ch.u_pad, sc_sub(sc_neg(sc_mul(yi, sc_mul(ch.b, srev))), sc_one()));
        }
        st_sc(gs + i, g);
        st_sc(hs + i, hh);
    }
    sc d = block_sum(dl, sh);
    if (threadIdx.x == 0) st_sc(partial + blockIdx.x, d);
}

// head scalars: vs_j = wV_j * rxx ; sB = w (t_x - ab) + r (xx (wc + delta) - t_x)
__global__ void __launch_bounds__(SK_THREADS) k_ver_head(const sc* wV, const sc* wc, const sc* delta, sc rxx, sc r,
                                                         sc xx, sc w_tab, sc t_x, sc* vs, sc* sB, uint32_t m) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) st_sc(vs + j, sc_mul(ld_sc(wV + j), rxx));
    if (j == 0) {
        sc inner = sc_sub(sc_mul(xx, sc_add(ld_sc(wc), ld_sc(delta))), t_x);
        st_sc(sB, sc_add(w_tab, sc_mul(r, inner)));
    }
}

// out[j] = sum_i w_i * b_i  (t2 blinding = <wV, v_blinding>), single block
__global__ void __launch_bounds__(SK_THREADS) k_dot(const sc* a, const sc* b, uint32_t n, sc* out) {
    __shared__ sc sh[SK_THREADS];
    sc acc = sc_zero();
    for (uint32_t i = threadIdx.x; i < n; i += SK_THREADS) acc = sc_add(acc, sc_mul(ld_sc(a + i), ld_sc(b + i)));
    sc s = block_sum(acc, sh);
    if (threadIdx.x == 0) st_sc(out, s);
}

// ---------------------------------------------------------------------------- launchers
static inline uint32_t nblk(uint64_t n) { return (uint32_t)((n + SK_THREADS - 1) / SK_THREADS); }
#define SK_REDUCE_BLOCKS 296

void sk_powers(cudaStream_t st, sc* out, const PowTable& tbl, uint32_t n, uint32_t start) {
    if (!n) return;
    uint32_t lgT = 0;
    while ((1u << lgT) < n && lgT < 31) lgT++;      // T >= n: one output per thread ...
    if (lgT > 14) lgT = lgT >= 17 ? lgT - 3 : 14;   // ... until there are enough threads: then 8 outputs (or more) each
    const uint32_t T = 1u << lgT;
    k_powers<<<nblk(T < n ? T : n), SK_THREADS, 0, st>>>(out, tbl, n, start, lgT);
}
void sk_powers_multi(cudaStream_t st, PowJobs& jobs) {
    uint32_t blocks = 0;
    for (uint32_t w = 0; w < jobs.count; w++) {
        const uint32_t n = jobs.n[w];
        uint32_t lgT = 0;
        while ((1u << lgT) < n && lgT < 31) lgT++;
        if (lgT > 14) lgT = lgT >= 17 ? lgT - 3 : 14;
        jobs.lgT[w] = lgT;
        const uint32_t T = 1u << lgT;
        blocks = max(blocks, nblk(T < n ? T : n));
    }
    if (blocks) k_powers_multi<<<dim3(blocks, jobs.count), SK_THREADS, 0, st>>>(jobs);
}
void sk_flatten(cudaStream_t st, const uint32_t* col_start, const uint32_t* col_row, const sc* col_coef,
                const sc* zpow, sc* out, uint32_t nt, uint32_t neg_from, const uint32_t* long_targets,
                uint32_t n_long, sc* part, uint32_t* tickets) {
    if (nt) k_flatten<<<nblk(nt), SK_THREADS, 0, st>>>(col_start, col_row, col_coef, zpow, out, nt, neg_from);
    if (nt && n_long)
        k_flatten_long<<<dim3(n_long, FL_SPLIT), FL_THREADS, 0, st>>>(col_start, col_row, col_coef, zpow, out, long_targets, nt,
                                                                     neg_from, part, tickets);
}
void sk_lr_poly(cudaStream_t st, const sc* aL, const sc* aR, const sc* aO, const sc* sL, const sc* sR, const sc* wL,
                const sc* wR, const sc* wO, const sc* ypow, const sc* yinv, sc* l1, sc* r0, sc* r1, sc* r3,
                sc* partial, sc* t_out, uint32_t n) {
    uint32_t blocks = n ? min(nblk(n), (uint32_t)SK_REDUCE_BLOCKS) : 1;
    k_lr_poly<<<blocks, SK_THREADS, 0, st>>>(aL, aR, aO, sL, sR, wL, wR, wO, ypow, yinv, l1, r0, r1, r3, partial, n);
    k_sum_sc<<<1, SK_THREADS, 0, st>>>(partial, blocks, 6, sc_one(), t_out);
}
void sk_eval_lr(cudaStream_t st, const sc* l1, const sc* aO, const sc* sL, const sc* r0, const sc* r1, const sc* r3,
                const sc* ypow, const sc& x, sc* lvec, sc* rvec, uint32_t n, uint32_t npad) {
    k_eval_lr<<<nblk(npad), SK_THREADS, 0, st>>>(l1, aO, sL, r0, r1, r3, ypow, x, lvec, rvec, n, npad);
}
__global__ void __launch_bounds__(SK_THREADS) k_fill_one(sc* p, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_sc(p + i, sc_one());
}
void sk_fill_one(cudaStream_t st, sc* p, uint32_t n) {
    if (n) k_fill_one<<<nblk(n), SK_THREADS, 0, st>>>(p, n);
}
void sk_ipp_init(cudaStream_t st, sc* sG, sc* sH, const sc* yinv, const sc& u, uint32_t n, uint32_t npad) {
    k_ipp_init<<<nblk(npad), SK_THREADS, 0, st>>>(sG, sH, yinv, u, n, npad);
}
void sk_ipp_round_fused(cudaStream_t st, const sc* a, const sc* b, const sc* sG, const sc* sH, sc* mG, sc* mH, sc* partial,
                        uint32_t* ticket, sc* cw_out, const sc& w, uint32_t npad, uint32_t nk) {
    const uint32_t h = nk >> 1;
    const uint32_t cb = max(1u, min(nblk(h), (uint32_t)SK_REDUCE_BLOCKS));  // <= nblk(npad): h < npad
    k_ipp_round_big<<<nblk(npad), SK_THREADS, 0, st>>>(a, b, sG, sH, mG, mH, partial, ticket, cw_out, w, cb, npad, nk);
}
void sk_ipp_round_scalars(cudaStream_t st, const sc* a, const sc* b, const sc* sG, const sc* sH, sc* mG, sc* mH,
                          sc* partial, sc* cw_out, const sc& w, uint32_t npad, uint32_t nk) {
    const uint32_t h = nk >> 1;
    uint32_t blocks = min(nblk(h), (uint32_t)SK_REDUCE_BLOCKS);
    k_ipp_cross<<<blocks, SK_THREADS, 0, st>>>(a, b, h, partial);
    k_sum_sc<<<1, SK_THREADS, 0, st>>>(partial, blocks, 2, w, cw_out);
    k_ipp_scalars<<<nblk(npad), SK_THREADS, 0, st>>>(a, b, sG, sH, mG, mH, npad, nk);
}
void sk_ipp_round_small(cudaStream_t st, sc* a, sc* b, sc* sG, sc* sH, sc* mG, sc* mH, sc* cw_out, const sc& w, const sc& u,
                        const sc& uinv, bool do_fold, uint32_t npad, uint32_t nk) {
    k_ipp_round_small<<<1, SK_THREADS, 0, st>>>(a, b, sG, sH, mG, mH, cw_out, w, u, uinv, do_fold ? 1 : 0, npad, nk);
}
void sk_ipp_fold(cudaStream_t st, sc* a, sc* b, sc* sG, sc* sH, const sc& u, const sc& uinv, uint32_t npad,
                 uint32_t nk) {
    k_ipp_fold<<<nblk(npad), SK_THREADS, 0, st>>>(a, b, sG, sH, u, uinv, npad, nk);
}
void sk_ver_scalars(cudaStream_t st, const VerChallenges& ch, const sc* wL, const sc* wR, const sc* wO, const sc* yinv,
                    sc* gs, sc* hs, sc* partial, sc* delta_out, uint32_t n, uint32_t npad, uint32_t lg) {
    uint32_t blocks = min(nblk(npad), (uint32_t)SK_REDUCE_BLOCKS);
    k_ver_scalars<<<blocks, SK_THREADS, 0, st>>>(ch, wL, wR, wO, yinv, gs, hs, partial, n, npad, lg);
    k_sum_sc<<<1, SK_THREADS, 0, st>>>(partial, blocks, 1, sc_one(), delta_out);
}
void sk_ver_head(cudaStream_t st, const sc* wV, const sc* wc, const sc* delta, const sc& rxx, const sc& r,
                 const sc& xx, const sc& w_tab, const sc& t_x, sc* vs, sc* sB, uint32_t m) {
    k_ver_head<<<nblk(m ? m : 1), SK_THREADS, 0, st>>>(wV, wc, delta, rxx, r, xx, w_tab, t_x, vs, sB, m);
}
void sk_dot(cudaStream_t st, const sc* a, const sc* b, uint32_t n, sc* out) {
    k_dot<<<1, SK_THREADS, 0, st>>>(a, b, n, out);
}
