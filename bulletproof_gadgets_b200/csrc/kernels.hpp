// Launchers of the mod-l scalar kernels (scalar_kernels.cu) and of the point kernels used by the
// R1CS driver (points.cu).  All device vectors hold canonical scalars (< l), 32 bytes LE each.
#pragma once
#include <cuda_runtime.h>

#include "ge25519.cuh"
#include "sc25519.cuh"

struct PowTable {
    sc p[32];  // base^(2^k)
};
struct VerChallenges {
    sc u[32], uinv[32];  // IPP challenges and inverses, creation order
    sc x, a, b, u_pad;   // u_pad = the "u" challenge that scales the padding generators
};

void sk_powers(cudaStream_t st, sc* out, const PowTable& tbl, uint32_t n, uint32_t start);
struct PowJobs {  // up to three power vectors in one launch (y^i, y^-i, z^(1+i) of a proof)
    sc* out[3];
    PowTable tbl[3];
    uint32_t n[3], start[3], lgT[3];
    uint32_t count;
};
void sk_powers_multi(cudaStream_t st, PowJobs& jobs);
#define FLATTEN_LONG 128u  // columns with more terms than this get a whole CTA
void sk_flatten(cudaStream_t st, const uint32_t* col_start, const uint32_t* col_row, const sc* col_coef,
                const sc* zpow, sc* out, uint32_t nt, uint32_t neg_from, const uint32_t* long_targets,
                uint32_t n_long, sc* part /* [n_long * 64] */, uint32_t* tickets /* [n_long], zero */);
void sk_lr_poly(cudaStream_t st, const sc* aL, const sc* aR, const sc* aO, const sc* sL, const sc* sR, const sc* wL,
                const sc* wR, const sc* wO, const sc* ypow, const sc* yinv, sc* l1, sc* r0, sc* r1, sc* r3,
                sc* partial, sc* t_out, uint32_t n);
void sk_eval_lr(cudaStream_t st, const sc* l1, const sc* aO, const sc* sL, const sc* r0, const sc* r1, const sc* r3,
                const sc* ypow, const sc& x, sc* lvec, sc* rvec, uint32_t n, uint32_t npad);
void sk_fill_one(cudaStream_t st, sc* p, uint32_t n);
void sk_ipp_init(cudaStream_t st, sc* sG, sc* sH, const sc* yinv, const sc& u, uint32_t n, uint32_t npad);
void sk_ipp_round_scalars(cudaStream_t st, const sc* a, const sc* b, const sc* sG, const sc* sH, sc* mG, sc* mH,
                          sc* partial, sc* cw_out, const sc& w, uint32_t npad, uint32_t nk);
// the same three kernels in one launch (the last cross-term CTA to finish adds the block sums); *ticket zero between launches
void sk_ipp_round_fused(cudaStream_t st, const sc* a, const sc* b, const sc* sG, const sc* sH, sc* mG, sc* mH, sc* partial,
                        uint32_t* ticket, sc* cw_out, const sc& w, uint32_t npad, uint32_t nk);
#define IPP_SMALL_MAX 2048u  // statements up to this many (padded) multipliers use the single-launch round kernel
void sk_ipp_round_small(cudaStream_t st, sc* a, sc* b, sc* sG, sc* sH, sc* mG, sc* mH, sc* cw_out, const sc& w, const sc& u,
                        const sc& uinv, bool do_fold, uint32_t npad, uint32_t nk);
void sk_ipp_fold(cudaStream_t st, sc* a, sc* b, sc* sG, sc* sH, const sc& u, const sc& uinv, uint32_t npad,
                 uint32_t nk);
void sk_ver_scalars(cudaStream_t st, const VerChallenges& ch, const sc* wL, const sc* wR, const sc* wO, const sc* yinv,
                    sc* gs, sc* hs, sc* partial, sc* delta_out, uint32_t n, uint32_t npad, uint32_t lg);
void sk_ver_head(cudaStream_t st, const sc* wV, const sc* wc, const sc* delta, const sc& rxx, const sc& r,
                 const sc& xx, const sc& w_tab, const sc& t_x, sc* vs, sc* sB, uint32_t m);
void sk_dot(cudaStream_t st, const sc* a, const sc* b, uint32_t n, sc* out);
#define SK_PARTIAL_SCALARS (296 * 6)

// points.cu
// ped[(p*64 + w)*8 + (m-1)] = m * 16^w * P_p, p in {B, B_blinding} (1024 affine Niels rows)
void pk_pedersen_table(cudaStream_t st, const ge_ext* gens_ext, uint32_t idxB, ge_niels* ped);
// out[i] = v_i*B + r_i*B_blinding (extended), thread per commitment
void pk_pedersen(cudaStream_t st, const ge_niels* ped, const sc* v, const sc* r, ge_ext* out, uint32_t k);
void pk_compress(cudaStream_t st, const ge_ext* in, uint8_t* out, uint32_t n);
// decompress n points; *fail counts undecodable encodings
void pk_decompress(cudaStream_t st, const uint8_t* in, ge_ext* out, uint32_t n, uint32_t* fail);
// out = sum_i s_i * P_i for arbitrary (dynamic) points; blockres scratch >= ceil(n/64) points
void pk_dyn_msm(cudaStream_t st, const ge_ext* pts, const sc* s, uint32_t n, ge_ext* blockres, ge_ext* out);
// rows[i] = affine Niels form of the i-th ristretto encoding (identity + *fail++ when it does not decode)
void pk_decompress_niels(cudaStream_t st, const uint8_t* in, ge_niels* rows, uint32_t n, uint32_t* fail);
// out = sum_w 2^(c w) W[w], w < K
void pk_window_combine(cudaStream_t st, const ge_ext* W, int K, int c, ge_ext* out);
// late IPP round over folded generators gp[0 .. 2 nb) (G' then H'): out2[0] = L, out2[1] = R; blockres >= 2 * ceil((2 nb + 2) / 64)
void pk_dyn_msm_lr(cudaStream_t st, const ge_ext* gp, const ge_ext* Bpt, const sc* mG, const sc* mH, const sc* cw, uint32_t nb,
                   uint32_t nk, ge_ext* blockres, ge_ext* out2);
// out = a + b
void pk_add2(cudaStream_t st, const ge_ext* a, const ge_ext* b, ge_ext* out);
