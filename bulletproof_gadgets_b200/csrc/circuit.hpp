// Device-resident flattened constraint systems (bpg_circuit): what assign_buffer replays into the
// real prover / verifier (/root/reference/src/prove.rs:84-99, /root/reference/src/verify.rs:75-90),
// kept in HBM in the transposed (by target variable) form the flatten kernel reads.
#pragma once
#include "ctx.hpp"

// targets (columns): [wL(n) | wR(n) | wO(n) | wV(m) | wc]; column t holds (constraint row, coefficient) pairs
struct bpg_circuit {
    bpg_ctx* ctx = nullptr;  // creating context (a resident circuit may outlive it: only `device` is used to free it)
    int device = 0;
    uint32_t n = 0, m = 0, q = 0, nt = 0, nnz = 0, n_long = 0;
    uint32_t *d_col_start = nullptr, *d_col_row = nullptr, *d_long = nullptr;
    sc *d_col_coef = nullptr, *d_aL = nullptr, *d_aR = nullptr, *d_aO = nullptr;
    bool has_witness = false;
    bool pooled = false;  // storage comes from the stream-ordered pool (per-proof circuits)
    // Per-proof circuits: ONE pool allocation holds every array above plus the four status words below (small statements are
    // bound by the number of driver calls, ~250 per statement before this: tools/gpu_timeline.py TIMELINE_MODE=c4)
    void* slab = nullptr;
    uint32_t* d_flags = nullptr;  // [0] invalid variable / coefficient bits, [1] long columns, [2] invalid multiplier; zeroed by circuit_build
    uint32_t long_cap = 0;
    bool check_pending = false;   // circuit_build left its status read-back to circuit_set_witness* (one wait instead of two)
};

// Builds the transposed form ON THE DEVICE from a host CSR term list (row_start[q+1], term_var[nnz],
// 32-byte little-endian coefficients, < 2^255, reduced mod l by the kernel).  The uploads and kernels
// are queued on ctx->stream; one 8-byte read-back reports invalid variables / coefficients.
// defer_check (pooled circuits only): the caller follows up with circuit_set_witness*, which reads the status back.
int circuit_build(bpg_ctx* ctx, uint64_t n, uint64_t m, uint64_t q, const uint32_t* row_start, const uint32_t* term_var,
                  const uint8_t* term_coef32, bool pooled, bpg_circuit** out, bool defer_check = false);
// multiplier assignments: raw 32-byte scalars (< 2^255), reduced on the device; a_O = a_L * a_R
int circuit_set_witness(bpg_circuit* c, const uint8_t* aL32n, const uint8_t* aR32n);
// f3: multipliers of range-proof bit runs are generated on the device; the `h` others come as compact host arrays with
// their multiplier indices.  Runs and indices must partition [0, n).
int circuit_set_witness_bits(bpg_circuit* c, const bpg_bit_run* runs, uint64_t n_runs, const uint8_t* aL32h,
                             const uint8_t* aR32h, const uint32_t* host_index, uint64_t h);
void circuit_free(bpg_circuit* c);

// scan.cu: out[i] = sum_{j<i} in[i] for i in [0, n]  (n+1 outputs; in and out may alias); scratch >= n/2048 + 2 words
void dev_exclusive_scan_u32(cudaStream_t st, const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* scratch);
// MSM form: off[0..G] = exclusive scan of hist[0..G), hist re-zeroed, *meta = {E, CL, nchunks} with the chunk length
// CL = cl_fixed, or max(cl_min, ceil(E / target_chunks)) when cl_fixed == 0.  One launch (a single CTA).
void dev_scan_meta(cudaStream_t st, uint32_t* hist, uint32_t* off, uint32_t G, uint32_t target_chunks, uint32_t cl_min,
                   uint32_t cl_fixed, MsmMeta* meta);
