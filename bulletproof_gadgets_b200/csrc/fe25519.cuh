// GF(2^255-19) on 8 x 32-bit saturated limbs for sm_100a.
//
// Replaces curve25519-dalek 3.2.0 `FieldElement51` / `FieldElement2625x4`
// (backend/serial/u64/field.rs, backend/vector/avx2/field.rs; crate pinned at
// /root/reference/Cargo.lock:155-157, source not vendored) -- SURVEY.md row K1 / a13.
//
// Representation: little-endian limbs, any value in [0, 2^256) congruent to the element
// ("lazy" form, 2^256 == 38 mod p).  Only fe_canon() produces the unique representative.
// Multiplication is an 8x8 schoolbook split into even/odd column accumulators so that every
// 32x32->64 product is one `mad.lo.cc / madc.hi.cc` pair riding a single carry chain
// (ptxas fuses each pair into IMAD.WIDE.U32 with predicate carry); the upper half is folded
// back with x38.  No tensor cores: this is the IMAD roofline path named by BASELINE.json.
//
// Every function is __host__ __device__: the host branch (portable u64 arithmetic) exists so
// tests/host_math can check the formulas against the oracle without a GPU; kernels always run
// the PTX branch.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define BPG_HD __host__ __device__ __forceinline__
#define BPG_D __device__ __forceinline__
#else
#define BPG_HD inline
#define BPG_D inline
#endif

struct fe {
    uint32_t v[8];
};

// ------------------------------------------------------------------------------------------
// constants (limbs little-endian)
// ------------------------------------------------------------------------------------------
#define FE_C(a0, a1, a2, a3, a4, a5, a6, a7) {{a0##u, a1##u, a2##u, a3##u, a4##u, a5##u, a6##u, a7##u}}

BPG_HD fe fe_zero() {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
}
BPG_HD fe fe_one() {
    fe r = fe_zero();
    r.v[0] = 1;
    return r;
}

// ------------------------------------------------------------------------------------------
// add / sub / neg   (lazy: result < 2^256)
// ------------------------------------------------------------------------------------------
BPG_HD fe fe_add(const fe& a, const fe& b) {
    fe r;
#ifdef __CUDA_ARCH__
    uint32_t c;
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
          "=r"(r.v[7]), "=r"(c)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    // fold the carry: +38*c, and once more if that wrapped (then the value is tiny)
    uint32_t k = c * 38u, c2;
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.cc.u32 %7, %7, 0;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]),
          "+r"(r.v[7]), "=r"(c2)
        : "r"(k));
    r.v[0] += c2 * 38u;
#else
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a.v[i] + b.v[i];
        r.v[i] = (uint32_t)c;
        c >>= 32;
    }
    uint64_t k = c * 38;
    for (int i = 0; i < 8; i++) {
        k += r.v[i];
        r.v[i] = (uint32_t)k;
        k >>= 32;
    }
    r.v[0] += (uint32_t)k * 38u;
#endif
    return r;
}

BPG_HD fe fe_sub(const fe& a, const fe& b) {
    fe r;
#ifdef __CUDA_ARCH__
    uint32_t bw;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"  // 0 or 0xffffffff
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
          "=r"(r.v[7]), "=r"(bw)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    // a-b wrapped by 2^256 == +38: subtract 38 when a borrow happened, and again if that wrapped
    uint32_t k = bw & 38u, bw2;
    asm("sub.cc.u32 %0, %0, %9;\n\t"
        "subc.cc.u32 %1, %1, 0;\n\t"
        "subc.cc.u32 %2, %2, 0;\n\t"
        "subc.cc.u32 %3, %3, 0;\n\t"
        "subc.cc.u32 %4, %4, 0;\n\t"
        "subc.cc.u32 %5, %5, 0;\n\t"
        "subc.cc.u32 %6, %6, 0;\n\t"
        "subc.cc.u32 %7, %7, 0;\n\t"
        "subc.u32 %8, 0, 0;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]),
          "+r"(r.v[7]), "=r"(bw2)
        : "r"(k));
    r.v[0] -= bw2 & 38u;
#else
    int64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (int64_t)a.v[i] - (int64_t)b.v[i];
        r.v[i] = (uint32_t)c;
        c >>= 32;  // arithmetic shift: 0 or -1
    }
    int64_t k = c ? -38 : 0;
    int64_t c2 = 0;
    for (int i = 0; i < 8; i++) {
        c2 += (int64_t)r.v[i] + (i == 0 ? k : 0);
        r.v[i] = (uint32_t)c2;
        c2 >>= 32;
    }
    if (c2) r.v[0] -= 38u;
#endif
    return r;
}

BPG_HD fe fe_neg(const fe& a) { return fe_sub(fe_zero(), a); }

// ------------------------------------------------------------------------------------------
// multiplication
// ------------------------------------------------------------------------------------------
#ifdef __CUDA_ARCH__
// acc[0..8] += (a0,a1,a2,a3) * b laid out as lo,hi pairs on consecutive limbs; acc[8] takes the carry.
#define FE_MADROW9(acc, a0, a1, a2, a3, b)                                                                    \
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"                                                                  \
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"                                                                 \
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"                                                                \
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"                                                                \
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"                                                                \
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"                                                                \
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"                                                                \
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"                                                                \
        "addc.u32 %8, %8, 0;"                                                                                 \
        : "+r"((acc)[0]), "+r"((acc)[1]), "+r"((acc)[2]), "+r"((acc)[3]), "+r"((acc)[4]), "+r"((acc)[5]),     \
          "+r"((acc)[6]), "+r"((acc)[7]), "+r"((acc)[8])                                                      \
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b))
// same without a carry-out limb (top of the product: the carry is provably zero)
#define FE_MADROW8(acc, a0, a1, a2, a3, b)                                                                    \
    asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"                                                                  \
        "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"                                                                 \
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"                                                                 \
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"                                                                 \
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"                                                                \
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"                                                                \
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"                                                                \
        "madc.hi.u32 %7, %11, %12, %7;"                                                                       \
        : "+r"((acc)[0]), "+r"((acc)[1]), "+r"((acc)[2]), "+r"((acc)[3]), "+r"((acc)[4]), "+r"((acc)[5]),     \
          "+r"((acc)[6]), "+r"((acc)[7])                                                                      \
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b))
#endif

#if defined(__CUDA_ARCH__) && defined(FE_KARATSUBA)
// acc[0..4] += (a0, a1) * b laid out as lo,hi pairs on consecutive limbs; acc[4] takes the carry
#define FE_MADROW5(acc, a0, a1, b)                                                                            \
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"                                                                  \
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"                                                                 \
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"                                                                 \
        "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"                                                                 \
        "addc.u32 %4, %4, 0;"                                                                                 \
        : "+r"((acc)[0]), "+r"((acc)[1]), "+r"((acc)[2]), "+r"((acc)[3]), "+r"((acc)[4])                      \
        : "r"(a0), "r"(a1), "r"(b))
#define FE_MADROW4(acc, a0, a1, b)                                                                            \
    asm("mad.lo.cc.u32 %0, %4, %6, %0;\n\t"                                                                  \
        "madc.hi.cc.u32 %1, %4, %6, %1;\n\t"                                                                 \
        "madc.lo.cc.u32 %2, %5, %6, %2;\n\t"                                                                 \
        "madc.hi.u32 %3, %5, %6, %3;"                                                                         \
        : "+r"((acc)[0]), "+r"((acc)[1]), "+r"((acc)[2]), "+r"((acc)[3])                                      \
        : "r"(a0), "r"(a1), "r"(b))
// r[0..8) = a[0..4) * b[0..4): 16 multiply-add pairs (IMAD.WIDE), even / odd column accumulators as in fe_mul_wide
__device__ __forceinline__ void fe_mul4(uint32_t r[8], const uint32_t a[4], const uint32_t b[4]) {
    uint32_t E[9], O[8];
#pragma unroll
    for (int k = 0; k < 9; k++) E[k] = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) O[k] = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t bi = b[i];
        if ((i & 1) == 0) {
            if (i + 4 < 8) {
                FE_MADROW5(E + i, a[0], a[2], bi);
                FE_MADROW5(O + i, a[1], a[3], bi);
            } else {
                FE_MADROW4(E + i, a[0], a[2], bi);
                FE_MADROW4(O + i, a[1], a[3], bi);
            }
        } else {
            if (i + 5 < 8) {
                FE_MADROW5(E + i + 1, a[1], a[3], bi);
            } else {
                FE_MADROW4(E + i + 1, a[1], a[3], bi);
            }
            FE_MADROW5(O + i - 1, a[0], a[2], bi);
        }
    }
    r[0] = E[0];
    asm("add.cc.u32 %0, %7, %14;\n\t"
        "addc.cc.u32 %1, %8, %15;\n\t"
        "addc.cc.u32 %2, %9, %16;\n\t"
        "addc.cc.u32 %3, %10, %17;\n\t"
        "addc.cc.u32 %4, %11, %18;\n\t"
        "addc.cc.u32 %5, %12, %19;\n\t"
        "addc.u32 %6, %13, %20;"
        : "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(O[0]), "r"(O[1]), "r"(O[2]),
          "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]));
}
// One level of Karatsuba over 128-bit halves: 48 multiply-add pairs instead of 64 (the IMAD.WIDE pipe is the bound of the
// bucket kernel); the extra additions ride the ALU pipe.  MEASURED AND NOT ENABLED (-DFE_KARATSUBA): correct (all GPU
// parity tests pass) but k_accumulate gets slower on B200, 378 us vs 339 us at 2^18 points -- ~70 extra carry-chain
// additions per product cost more issue slots than 16 IMAD.WIDE save, and the kernel goes from 108 to 126 registers.
__device__ __forceinline__ void fe_mul_wide_karatsuba(uint32_t r[16], const fe& a, const fe& b) {
    uint32_t z0[8], z2[8], zm[9], sa[4], sb[4], ca, cb;
    fe_mul4(z0, a.v, b.v);
    fe_mul4(z2, a.v + 4, b.v + 4);
    asm("add.cc.u32 %0, %5, %9;\n\t"
        "addc.cc.u32 %1, %6, %10;\n\t"
        "addc.cc.u32 %2, %7, %11;\n\t"
        "addc.cc.u32 %3, %8, %12;\n\t"
        "addc.u32 %4, 0, 0;"
        : "=r"(sa[0]), "=r"(sa[1]), "=r"(sa[2]), "=r"(sa[3]), "=r"(ca)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]));
    asm("add.cc.u32 %0, %5, %9;\n\t"
        "addc.cc.u32 %1, %6, %10;\n\t"
        "addc.cc.u32 %2, %7, %11;\n\t"
        "addc.cc.u32 %3, %8, %12;\n\t"
        "addc.u32 %4, 0, 0;"
        : "=r"(sb[0]), "=r"(sb[1]), "=r"(sb[2]), "=r"(sb[3]), "=r"(cb)
        : "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    fe_mul4(zm, sa, sb);
    // (sa + ca 2^128)(sb + cb 2^128) = sa sb + (ca sb + cb sa) 2^128 + ca cb 2^256
    const uint32_t ma = 0u - ca, mb = 0u - cb;
    asm("add.cc.u32 %0, %0, %5;\n\t"
        "addc.cc.u32 %1, %1, %6;\n\t"
        "addc.cc.u32 %2, %2, %7;\n\t"
        "addc.cc.u32 %3, %3, %8;\n\t"
        "addc.u32 %4, %9, 0;"
        : "+r"(zm[4]), "+r"(zm[5]), "+r"(zm[6]), "+r"(zm[7]), "=r"(zm[8])
        : "r"(sb[0] & ma), "r"(sb[1] & ma), "r"(sb[2] & ma), "r"(sb[3] & ma), "r"(ca & cb));
    asm("add.cc.u32 %0, %0, %5;\n\t"
        "addc.cc.u32 %1, %1, %6;\n\t"
        "addc.cc.u32 %2, %2, %7;\n\t"
        "addc.cc.u32 %3, %3, %8;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(zm[4]), "+r"(zm[5]), "+r"(zm[6]), "+r"(zm[7]), "+r"(zm[8])
        : "r"(sa[0] & mb), "r"(sa[1] & mb), "r"(sa[2] & mb), "r"(sa[3] & mb));
    // z1 = zm - z0 - z2   (0 <= z1 < 2^258)
    asm("sub.cc.u32 %0, %0, %9;\n\t"
        "subc.cc.u32 %1, %1, %10;\n\t"
        "subc.cc.u32 %2, %2, %11;\n\t"
        "subc.cc.u32 %3, %3, %12;\n\t"
        "subc.cc.u32 %4, %4, %13;\n\t"
        "subc.cc.u32 %5, %5, %14;\n\t"
        "subc.cc.u32 %6, %6, %15;\n\t"
        "subc.cc.u32 %7, %7, %16;\n\t"
        "subc.u32 %8, %8, 0;"
        : "+r"(zm[0]), "+r"(zm[1]), "+r"(zm[2]), "+r"(zm[3]), "+r"(zm[4]), "+r"(zm[5]), "+r"(zm[6]), "+r"(zm[7]), "+r"(zm[8])
        : "r"(z0[0]), "r"(z0[1]), "r"(z0[2]), "r"(z0[3]), "r"(z0[4]), "r"(z0[5]), "r"(z0[6]), "r"(z0[7]));
    asm("sub.cc.u32 %0, %0, %9;\n\t"
        "subc.cc.u32 %1, %1, %10;\n\t"
        "subc.cc.u32 %2, %2, %11;\n\t"
        "subc.cc.u32 %3, %3, %12;\n\t"
        "subc.cc.u32 %4, %4, %13;\n\t"
        "subc.cc.u32 %5, %5, %14;\n\t"
        "subc.cc.u32 %6, %6, %15;\n\t"
        "subc.cc.u32 %7, %7, %16;\n\t"
        "subc.u32 %8, %8, 0;"
        : "+r"(zm[0]), "+r"(zm[1]), "+r"(zm[2]), "+r"(zm[3]), "+r"(zm[4]), "+r"(zm[5]), "+r"(zm[6]), "+r"(zm[7]), "+r"(zm[8])
        : "r"(z2[0]), "r"(z2[1]), "r"(z2[2]), "r"(z2[3]), "r"(z2[4]), "r"(z2[5]), "r"(z2[6]), "r"(z2[7]));
    // r = z0 + z1 2^128 + z2 2^256
    r[0] = z0[0], r[1] = z0[1], r[2] = z0[2], r[3] = z0[3];
    asm("add.cc.u32 %0, %12, %24;\n\t"
        "addc.cc.u32 %1, %13, %25;\n\t"
        "addc.cc.u32 %2, %14, %26;\n\t"
        "addc.cc.u32 %3, %15, %27;\n\t"
        "addc.cc.u32 %4, %16, %28;\n\t"
        "addc.cc.u32 %5, %17, %29;\n\t"
        "addc.cc.u32 %6, %18, %30;\n\t"
        "addc.cc.u32 %7, %19, %31;\n\t"
        "addc.cc.u32 %8, %20, %32;\n\t"
        "addc.cc.u32 %9, %21, 0;\n\t"
        "addc.cc.u32 %10, %22, 0;\n\t"
        "addc.u32 %11, %23, 0;"
        : "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),
          "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(z0[4]), "r"(z0[5]), "r"(z0[6]), "r"(z0[7]), "r"(z2[0]), "r"(z2[1]), "r"(z2[2]), "r"(z2[3]), "r"(z2[4]),
          "r"(z2[5]), "r"(z2[6]), "r"(z2[7]), "r"(zm[0]), "r"(zm[1]), "r"(zm[2]), "r"(zm[3]), "r"(zm[4]), "r"(zm[5]),
          "r"(zm[6]), "r"(zm[7]), "r"(zm[8]));
}
#endif

// r[0..16) = a*b (full 512-bit product)
BPG_HD void fe_mul_wide(uint32_t r[16], const fe& a, const fe& b) {
#ifdef __CUDA_ARCH__
#ifdef FE_KARATSUBA
    fe_mul_wide_karatsuba(r, a, b);
    return;
#endif
    // E holds products with (i+j) even at limb i+j; O holds (i+j) odd, stored one limb lower
    // (O[k] has weight 2^(32(k+1))) so that lo/hi pairs stay even-aligned in both arrays.
    uint32_t E[17], O[16];
#pragma unroll
    for (int k = 0; k < 17; k++) E[k] = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) O[k] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t bi = b.v[i];
        if ((i & 1) == 0) {
            // even a-limbs -> E[i .. i+8], odd a-limbs -> O[i .. i+8]
            if (i + 8 < 16) {
                FE_MADROW9(E + i, a.v[0], a.v[2], a.v[4], a.v[6], bi);
                FE_MADROW9(O + i, a.v[1], a.v[3], a.v[5], a.v[7], bi);
            } else {
                FE_MADROW8(E + i, a.v[0], a.v[2], a.v[4], a.v[6], bi);
                FE_MADROW8(O + i, a.v[1], a.v[3], a.v[5], a.v[7], bi);
            }
        } else {
            // odd a-limbs -> E[i+1 .. i+9], even a-limbs -> O[i-1 .. i+7]
            if (i + 9 < 16) {
                FE_MADROW9(E + i + 1, a.v[1], a.v[3], a.v[5], a.v[7], bi);
            } else {
                FE_MADROW8(E + i + 1, a.v[1], a.v[3], a.v[5], a.v[7], bi);
            }
            FE_MADROW9(O + i - 1, a.v[0], a.v[2], a.v[4], a.v[6], bi);
        }
    }
    // r = E + (O << 32)
    r[0] = E[0];
    asm("add.cc.u32 %0, %15, %30;\n\t"
        "addc.cc.u32 %1, %16, %31;\n\t"
        "addc.cc.u32 %2, %17, %32;\n\t"
        "addc.cc.u32 %3, %18, %33;\n\t"
        "addc.cc.u32 %4, %19, %34;\n\t"
        "addc.cc.u32 %5, %20, %35;\n\t"
        "addc.cc.u32 %6, %21, %36;\n\t"
        "addc.cc.u32 %7, %22, %37;\n\t"
        "addc.cc.u32 %8, %23, %38;\n\t"
        "addc.cc.u32 %9, %24, %39;\n\t"
        "addc.cc.u32 %10, %25, %40;\n\t"
        "addc.cc.u32 %11, %26, %41;\n\t"
        "addc.cc.u32 %12, %27, %42;\n\t"
        "addc.cc.u32 %13, %28, %43;\n\t"
        "addc.u32 %14, %29, %44;"
        : "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(E[8]), "r"(E[9]),
          "r"(E[10]), "r"(E[11]), "r"(E[12]), "r"(E[13]), "r"(E[14]), "r"(E[15]), "r"(O[0]), "r"(O[1]),
          "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]), "r"(O[8]), "r"(O[9]), "r"(O[10]),
          "r"(O[11]), "r"(O[12]), "r"(O[13]), "r"(O[14]));
#else
    uint64_t t[16];
    for (int k = 0; k < 16; k++) t[k] = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 8; j++) {
            uint64_t p = (uint64_t)a.v[j] * b.v[i] + t[i + j] + c;
            t[i + j] = (uint32_t)p;
            c = p >> 32;
        }
        t[i + 8] = c;
    }
    for (int k = 0; k < 16; k++) r[k] = (uint32_t)t[k];
#endif
}

// fold a 512-bit value to < 2^256 using 2^256 == 38
BPG_HD fe fe_fold(const uint32_t t[16]) {
    fe r;
#ifdef __CUDA_ARCH__
    // r = t[0..8) + 38 * t[8..16): the eight 32x32->64 products 38 * t[8+k] ride two carry chains of four
    // IMAD.WIDE each (even k on limbs k, k+1; odd k likewise one limb up), as the rows of the product do
    uint32_t acc[9];
    const uint32_t k38 = 38u;
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = t[k];
    acc[8] = 0;
    FE_MADROW9(acc, t[8], t[10], t[12], t[14], k38);
    FE_MADROW8(acc + 1, t[9], t[11], t[13], t[15], k38);
#pragma unroll
    for (int k = 0; k < 8; k++) r.v[k] = acc[k];
    uint32_t r8 = acc[8];
    // r8 <= 1 + 37 + 1; fold it, then a possible final wrap (value then < 2^12)
    uint32_t k = r8 * 38u, c2;
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.cc.u32 %7, %7, 0;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]),
          "+r"(r.v[7]), "=r"(c2)
        : "r"(k));
    r.v[0] += c2 * 38u;
#else
    uint64_t c = 0;
    uint32_t lo[9];
    for (int k = 0; k < 8; k++) {
        c += (uint64_t)t[k] + (uint64_t)t[8 + k] * 38u;
        lo[k] = (uint32_t)c;
        c >>= 32;
    }
    lo[8] = (uint32_t)c;
    uint64_t k2 = (uint64_t)lo[8] * 38u;
    for (int k = 0; k < 8; k++) {
        k2 += lo[k];
        r.v[k] = (uint32_t)k2;
        k2 >>= 32;
    }
    r.v[0] += (uint32_t)k2 * 38u;
#endif
    return r;
}

BPG_HD fe fe_mul(const fe& a, const fe& b) {
    uint32_t t[16];
    fe_mul_wide(t, a, b);
    return fe_fold(t);
}

// Dedicated squaring: 28 cross products + 8 squares = 36 IMAD.WIDE instead of 64 (every doubling, the inverse-square-root
// chains of the ristretto codec and the generator-table build are squaring-heavy).  The carry-chain schedule below is
// generated and checked limb by limb by tools/gen_fe_sqr.py.
#ifdef __CUDA_ARCH__
#define FE_SQ_MADROW7(acc, a0, a1, a2, b)                                                                     \
    asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\t"                                                                  \
        "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"                                                                 \
        "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"                                                                 \
        "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"                                                                 \
        "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"                                                                 \
        "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"                                                                 \
        "addc.u32 %6, %6, 0;"                                                                                 \
        : "+r"((acc)[0]), "+r"((acc)[1]), "+r"((acc)[2]), "+r"((acc)[3]), "+r"((acc)[4]), "+r"((acc)[5]),     \
          "+r"((acc)[6])                                                                                      \
        : "r"(a0), "r"(a1), "r"(a2), "r"(b))
#define FE_SQ_MADROW5(acc, a0, a1, b)                                                                         \
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"                                                                   \
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"                                                                  \
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"                                                                  \
        "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"                                                                  \
        "addc.u32 %4, %4, 0;"                                                                                 \
        : "+r"((acc)[0]), "+r"((acc)[1]), "+r"((acc)[2]), "+r"((acc)[3]), "+r"((acc)[4])                      \
        : "r"(a0), "r"(a1), "r"(b))
#define FE_SQ_MADROW3(acc, a0, b)                                                                             \
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"                                                                   \
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"                                                                  \
        "addc.u32 %2, %2, 0;"                                                                                 \
        : "+r"((acc)[0]), "+r"((acc)[1]), "+r"((acc)[2])                                                      \
        : "r"(a0), "r"(b))
#endif

// r[0..16) = a*a
BPG_HD void fe_sqr_wide(uint32_t r[16], const fe& a) {
#ifdef __CUDA_ARCH__
    uint32_t E[17], O[16];
#pragma unroll
    for (int k = 0; k < 17; k++) E[k] = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) O[k] = 0;
    // cross products a_i a_j, i < j (even columns -> E, odd columns -> O, as in fe_mul_wide)
    FE_MADROW9(O + 0, a.v[1], a.v[3], a.v[5], a.v[7], a.v[0]);
    FE_SQ_MADROW7(E + 2, a.v[2], a.v[4], a.v[6], a.v[0]);
    FE_SQ_MADROW7(O + 2, a.v[2], a.v[4], a.v[6], a.v[1]);
    FE_SQ_MADROW7(E + 4, a.v[3], a.v[5], a.v[7], a.v[1]);
    FE_SQ_MADROW7(O + 4, a.v[3], a.v[5], a.v[7], a.v[2]);
    FE_SQ_MADROW5(E + 6, a.v[4], a.v[6], a.v[2]);
    FE_SQ_MADROW5(O + 6, a.v[4], a.v[6], a.v[3]);
    FE_SQ_MADROW5(E + 8, a.v[5], a.v[7], a.v[3]);
    FE_SQ_MADROW5(O + 8, a.v[5], a.v[7], a.v[4]);
    FE_SQ_MADROW3(E + 10, a.v[6], a.v[4]);
    FE_SQ_MADROW3(O + 10, a.v[6], a.v[5]);
    FE_SQ_MADROW3(E + 12, a.v[7], a.v[5]);
    FE_SQ_MADROW3(O + 12, a.v[7], a.v[6]);
    // C = E + (O << 32)
    r[0] = E[0];
    asm("add.cc.u32 %0, %15, %30;\n\t"
        "addc.cc.u32 %1, %16, %31;\n\t"
        "addc.cc.u32 %2, %17, %32;\n\t"
        "addc.cc.u32 %3, %18, %33;\n\t"
        "addc.cc.u32 %4, %19, %34;\n\t"
        "addc.cc.u32 %5, %20, %35;\n\t"
        "addc.cc.u32 %6, %21, %36;\n\t"
        "addc.cc.u32 %7, %22, %37;\n\t"
        "addc.cc.u32 %8, %23, %38;\n\t"
        "addc.cc.u32 %9, %24, %39;\n\t"
        "addc.cc.u32 %10, %25, %40;\n\t"
        "addc.cc.u32 %11, %26, %41;\n\t"
        "addc.cc.u32 %12, %27, %42;\n\t"
        "addc.cc.u32 %13, %28, %43;\n\t"
        "addc.u32 %14, %29, %44;"
        : "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(E[8]), "r"(E[9]),
          "r"(E[10]), "r"(E[11]), "r"(E[12]), "r"(E[13]), "r"(E[14]), "r"(E[15]), "r"(O[0]), "r"(O[1]),
          "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]), "r"(O[8]), "r"(O[9]), "r"(O[10]),
          "r"(O[11]), "r"(O[12]), "r"(O[13]), "r"(O[14]));
    // 2C
#pragma unroll
    for (int k = 15; k > 0; k--) r[k] = __funnelshift_l(r[k - 1], r[k], 1);
    r[0] <<= 1;
    // + a_i^2 at limbs 2i, 2i+1: one carry chain (the product fits 512 bits, so no carry leaves the top)
    asm("mad.lo.cc.u32 %0, %16, %16, %0;\n\t"
        "madc.hi.cc.u32 %1, %16, %16, %1;\n\t"
        "madc.lo.cc.u32 %2, %17, %17, %2;\n\t"
        "madc.hi.cc.u32 %3, %17, %17, %3;\n\t"
        "madc.lo.cc.u32 %4, %18, %18, %4;\n\t"
        "madc.hi.cc.u32 %5, %18, %18, %5;\n\t"
        "madc.lo.cc.u32 %6, %19, %19, %6;\n\t"
        "madc.hi.cc.u32 %7, %19, %19, %7;\n\t"
        "madc.lo.cc.u32 %8, %20, %20, %8;\n\t"
        "madc.hi.cc.u32 %9, %20, %20, %9;\n\t"
        "madc.lo.cc.u32 %10, %21, %21, %10;\n\t"
        "madc.hi.cc.u32 %11, %21, %21, %11;\n\t"
        "madc.lo.cc.u32 %12, %22, %22, %12;\n\t"
        "madc.hi.cc.u32 %13, %22, %22, %13;\n\t"
        "madc.lo.cc.u32 %14, %23, %23, %14;\n\t"
        "madc.hi.u32 %15, %23, %23, %15;"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
          "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]));
#else
    fe_mul_wide(r, a, a);
#endif
}

BPG_HD fe fe_sqr(const fe& a) {
    uint32_t t[16];
    fe_sqr_wide(t, a);
    return fe_fold(t);
}

// small-constant multiply (k < 2^31)
BPG_HD fe fe_mul_small(const fe& a, uint32_t k) {
    uint32_t t[16];
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a.v[i] * k;
        t[i] = (uint32_t)c;
        c >>= 32;
    }
    t[8] = (uint32_t)c;
#pragma unroll
    for (int i = 9; i < 16; i++) t[i] = 0;
    return fe_fold(t);
}

// ------------------------------------------------------------------------------------------
// canonical form, predicates, selects, bytes
// ------------------------------------------------------------------------------------------
BPG_HD fe fe_canon(const fe& a) {
    // fold bit 255 (x19), twice is enough to land in [0, 2^255), then conditionally subtract p
    fe r = a;
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
        uint64_t c = (uint64_t)(r.v[7] >> 31) * 19u;
        r.v[7] &= 0x7fffffffu;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            c += r.v[i];
            r.v[i] = (uint32_t)c;
            c >>= 32;
        }
    }
    // r < 2^255 + small; t = r + 19; if t >= 2^255 then r >= p -> result t - 2^255
    fe t;
    uint64_t c = 19;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += r.v[i];
        t.v[i] = (uint32_t)c;
        c >>= 32;
    }
    uint32_t ge = t.v[7] >> 31;
    t.v[7] &= 0x7fffffffu;
    uint32_t m = 0u - ge;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = (t.v[i] & m) | (r.v[i] & ~m);
    return r;
}

BPG_HD bool fe_is_zero(const fe& a) {
    fe c = fe_canon(a);
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= c.v[i];
    return o == 0;
}
BPG_HD bool fe_is_neg(const fe& a) { return fe_canon(a).v[0] & 1u; }
BPG_HD bool fe_eq(const fe& a, const fe& b) { return fe_is_zero(fe_sub(a, b)); }

BPG_HD fe fe_select(bool c, const fe& a, const fe& b) {  // c ? a : b
    fe r;
    uint32_t m = 0u - (uint32_t)c;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = (a.v[i] & m) | (b.v[i] & ~m);
    return r;
}
BPG_HD fe fe_cneg(const fe& a, bool c) { return fe_select(c, fe_neg(a), a); }
BPG_HD fe fe_abs(const fe& a) { return fe_cneg(a, fe_is_neg(a)); }

BPG_HD fe fe_from_bytes(const uint8_t* s) {  // 32 bytes LE, all 256 bits kept
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++)
        r.v[i] = (uint32_t)s[4 * i] | ((uint32_t)s[4 * i + 1] << 8) | ((uint32_t)s[4 * i + 2] << 16) |
                 ((uint32_t)s[4 * i + 3] << 24);
    return r;
}
BPG_HD void fe_to_bytes(uint8_t* s, const fe& a) {  // canonical
    fe c = fe_canon(a);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        s[4 * i] = (uint8_t)c.v[i];
        s[4 * i + 1] = (uint8_t)(c.v[i] >> 8);
        s[4 * i + 2] = (uint8_t)(c.v[i] >> 16);
        s[4 * i + 3] = (uint8_t)(c.v[i] >> 24);
    }
}

// ------------------------------------------------------------------------------------------
// exponentiation chains
// ------------------------------------------------------------------------------------------
BPG_HD fe fe_sqr_n(fe a, int n) {
#pragma unroll 1
    for (int i = 0; i < n; i++) a = fe_sqr(a);
    return a;
}

// returns z^(2^250-1); also z^11 through *z11
BPG_HD fe fe_pow_2_250_1(const fe& z, fe* z11) {
    fe t0 = fe_sqr(z);                 // 2
    fe t1 = fe_mul(z, fe_sqr_n(t0, 2));  // 9
    t0 = fe_mul(t0, t1);               // 11
    *z11 = t0;
    fe t2 = fe_mul(t1, fe_sqr(t0));    // 31 = 2^5-1
    t1 = fe_mul(fe_sqr_n(t2, 5), t2);  // 2^10-1
    fe t3 = fe_mul(fe_sqr_n(t1, 10), t1);  // 2^20-1
    fe t4 = fe_mul(fe_sqr_n(t3, 20), t3);  // 2^40-1
    t1 = fe_mul(fe_sqr_n(t4, 10), t1);     // 2^50-1
    t3 = fe_mul(fe_sqr_n(t1, 50), t1);     // 2^100-1
    t4 = fe_mul(fe_sqr_n(t3, 100), t3);    // 2^200-1
    return fe_mul(fe_sqr_n(t4, 50), t1);   // 2^250-1
}
BPG_HD fe fe_invert(const fe& z) {  // z^(p-2) = z^(2^255-21)
    fe z11;
    fe t = fe_pow_2_250_1(z, &z11);
    return fe_mul(fe_sqr_n(t, 5), z11);
}
BPG_HD fe fe_pow_p58(const fe& z) {  // z^((p-5)/8) = z^(2^252-3)
    fe z11;
    fe t = fe_pow_2_250_1(z, &z11);
    return fe_mul(fe_sqr_n(t, 2), z);
}
