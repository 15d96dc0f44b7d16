// Microbenchmark: sustained integer-multiply throughput on this GPU -- the roofline denominator for
// the MSM bucket stage.  Measures IMAD (32-bit), IMAD.HI.U32, IMAD.WIDE.U32, DFMA
// and the field multiplication itself (fe_mul/s).  Prints one JSON line per measurement.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../bulletproof_gadgets_b200/csrc/ge25519.cuh"

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed, int iters) {
    uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
    uint64_t x0 = threadIdx.x, x1 = a, x2 = b, x3 = a ^ b, x4 = 5, x5 = 6, x6 = 7, x7 = 8;
    uint32_t y0 = 1, y1 = 2, y2 = 3, y3 = 4, y4 = 5, y5 = 6, y6 = 7, y7 = 8;
    double d0 = a, d1 = b, d2 = 3, d3 = 4, d4 = 5, d5 = 6, d6 = 7, d7 = 8, da = 1.0000001, db = 1e-9;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (MODE == 0) {
                x0 = (uint64_t)a * (uint32_t)x0 + x0; x1 = (uint64_t)b * (uint32_t)x1 + x1; x2 = (uint64_t)a * (uint32_t)x2 + x2; x3 = (uint64_t)b * (uint32_t)x3 + x3;
                x4 = (uint64_t)a * (uint32_t)x4 + x4; x5 = (uint64_t)b * (uint32_t)x5 + x5; x6 = (uint64_t)a * (uint32_t)x6 + x6; x7 = (uint64_t)b * (uint32_t)x7 + x7;
            } else if (MODE == 1) {
                y0 = a * y0 + b; y1 = b * y1 + a; y2 = a * y2 + b; y3 = b * y3 + a; y4 = a * y4 + b; y5 = b * y5 + a; y6 = a * y6 + b; y7 = b * y7 + a;
            } else if (MODE == 2) {
                y0 = __umulhi(a, y0) + b; y1 = __umulhi(b, y1) + a; y2 = __umulhi(a, y2) + b; y3 = __umulhi(b, y3) + a;
                y4 = __umulhi(a, y4) + b; y5 = __umulhi(b, y5) + a; y6 = __umulhi(a, y6) + b; y7 = __umulhi(b, y7) + a;
            } else if (MODE == 3) {
                d0 = fma(d0, da, db); d1 = fma(d1, da, db); d2 = fma(d2, da, db); d3 = fma(d3, da, db);
                d4 = fma(d4, da, db); d5 = fma(d5, da, db); d6 = fma(d6, da, db); d7 = fma(d7, da, db);
            }
        }
    }
    uint64_t s = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
    double ds = d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7;
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)s ^ (uint32_t)(s >> 32) ^ y0 ^ y1 ^ y2 ^ y3 ^ y4 ^ y5 ^ y6 ^ y7 ^ (uint32_t)ds;
}

// 4 independent fe_mul chains per thread
__global__ void __launch_bounds__(128) k_femul(fe* out, int iters) {
    fe a[4], b;
    for (int j = 0; j < 4; j++) for (int i = 0; i < 8; i++) a[j].v[i] = threadIdx.x * 977 + blockIdx.x * 131 + i * 7 + j;
    for (int i = 0; i < 8; i++) b.v[i] = 0x9e3779b9u * (i + 1) + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 4; j++) a[j] = fe_mul(a[j], b);
    }
    fe r = fe_add(fe_add(a[0], a[1]), fe_add(a[2], a[3]));
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// mixed add throughput: one accumulator per thread, niels operand from registers
__global__ void __launch_bounds__(128, 4) k_madd(ge_ext* out, int iters) {
    ge_ext acc = ge_identity();
    ge_niels q;
    for (int i = 0; i < 8; i++) { q.yp.v[i] = threadIdx.x * 977 + i; q.ym.v[i] = blockIdx.x * 131 + i * 7; q.t2d.v[i] = 0x9e3779b9u * (i + 1) + threadIdx.x; }
    for (int it = 0; it < iters; it++) { acc = ge_madd(acc, q, it & 1); q.yp.v[0] += acc.X.v[0] & 1; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// latency of a dependent chain of point operations in one thread (what the bucket reduction's serial part is made of)
template <int OP>
__global__ void __launch_bounds__(128) k_chain(ge_ext* out, int iters) {
    ge_ext p = ge_identity(), q;
    for (int i = 0; i < 8; i++) { p.X.v[i] = threadIdx.x * 977 + i; p.Y.v[i] = blockIdx.x * 131 + i * 7 + 1; p.Z.v[i] = 3 * i + 1; p.T.v[i] = 5 * i + 2; }
    q = p;
    q.X.v[0] ^= 0x55;
    for (int it = 0; it < iters; it++) {
        if (OP == 0) p = ge_add(p, q);
        else if (OP == 1) p = ge_dbl(p);
        else p = ge_dbl_not(p);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = p;
}

int main(int argc, char** argv) {
    const bool quick = argc > 1 && !strcmp(argv[1], "--quick");  // IMAD.WIDE.U32 and IMAD only (bench.py: the roofline denominator)
    const int dev = argc > 2 ? atoi(argv[2]) : 0;                // index among the VISIBLE devices (one rank per GPU)
    if (cudaSetDevice(dev) != cudaSuccess) { printf("cannot select device %d\n", dev); return 1; }
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    void* out; cudaMalloc(&out, (size_t)sms * 16 * 256 * 128);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[4] = {"IMAD.WIDE.U32", "IMAD", "IMAD.HI.U32", "DFMA"};
    const int iters = 20000;
    for (int mode = 0; mode < (quick ? 2 : 4); mode++) {
        float best = 1e9;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<sms * 8, 256>>>((uint32_t*)out, rep + 1, iters);
            if (mode == 1) k<1><<<sms * 8, 256>>>((uint32_t*)out, rep + 1, iters);
            if (mode == 2) k<2><<<sms * 8, 256>>>((uint32_t*)out, rep + 1, iters);
            if (mode == 3) k<3><<<sms * 8, 256>>>((uint32_t*)out, rep + 1, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
        }
        double ops = (double)sms * 8 * 256 * iters * 64;
        printf("{\"op\": \"%s\", \"sms\": %d, \"max_clock_khz\": %d, \"ms\": %.3f, \"Tops\": %.3f, \"per_clk_per_sm_at_max_clock\": %.2f}\n",
               names[mode], sms, clk, best, ops / best / 1e9, ops / (best * 1e-3) / ((double)clk * 1e3) / sms);
    }
    if (!quick) {
        const int it2 = 2000; float best = 1e9;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0); k_femul<<<sms * 16, 128>>>((fe*)out, it2); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
        }
        double ops = (double)sms * 16 * 128 * it2 * 4;
        printf("{\"op\": \"fe_mul\", \"ms\": %.3f, \"Gops\": %.2f, \"sm_cycles_per_thread_op_at_max_clock\": %.3f}\n", best, ops / best / 1e6, (best * 1e-3) * clk * 1e3 * sms / ops);
        best = 1e9;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0); k_madd<<<sms * 16, 128>>>((ge_ext*)out, it2); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
        }
        ops = (double)sms * 16 * 128 * it2;
        printf("{\"op\": \"ge_madd\", \"ms\": %.3f, \"Gops\": %.2f, \"sm_cycles_per_thread_op_at_max_clock\": %.3f}\n", best, ops / best / 1e6, (best * 1e-3) * clk * 1e3 * sms / ops);
    }
    if (!quick) {
        const char* ops[3] = {"ge_add", "ge_dbl", "ge_dbl_not"};
        const int geo[4][2] = {{1, 32}, {1, 128}, {148, 128}, {148, 512}};
        const int it3 = 2000;
        for (int op = 0; op < 3; op++)
            for (int g = 0; g < 4; g++) {
                float best = 1e9;
                for (int rep = 0; rep < 4; rep++) {
                    cudaEventRecord(e0);
                    if (op == 0) k_chain<0><<<geo[g][0], geo[g][1]>>>((ge_ext*)out, it3);
                    if (op == 1) k_chain<1><<<geo[g][0], geo[g][1]>>>((ge_ext*)out, it3);
                    if (op == 2) k_chain<2><<<geo[g][0], geo[g][1]>>>((ge_ext*)out, it3);
                    cudaEventRecord(e1); cudaEventSynchronize(e1);
                    float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
                }
                printf("{\"op\": \"%s chain\", \"blocks\": %d, \"threads\": %d, \"ns_per_dependent_op\": %.1f}\n", ops[op], geo[g][0], geo[g][1], best * 1e6 / it3);
            }
    }
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
