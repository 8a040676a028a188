// Host-side launch interface between the C-ABI translation unit (api.cu) and the per-limb-count
// kernel instantiation units (inst_L{1,2,3,4}.cu), which are compiled in parallel.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include "kernels.cuh"

extern std::atomic<unsigned long long> g_mpyc_launches;

// host-side description of where K2 (generate mode) puts its m share rows
struct ShareDst {
    u64* base = nullptr;
    size_t stride = 0;
    int use_rows = 0;
    u64* rows[MPYC_MAX_SHARE_ROWS] = {};
};

// persistent-grid sizing: enough CTAs to cover the items, capped at one full wave
int mpyc_grid_size(const void* kernel, size_t items, size_t dyn_smem);

template <int L>
struct Launch {
    // op in {OP_ADD, OP_SUB, OP_MUL, OP_NEG}; scal != nullptr -> broadcast second operand (host limbs)
    static cudaError_t binop(const FieldParams& fp, int op, const u64* a, const u64* b, const u64* scal, u64* out,
                             size_t n, cudaStream_t st);
    // mode 0: out = a^e; mode 1: out8 = (a^e != p-1)
    static cudaError_t pow(const FieldParams& fp, const ExpParams& ex, int mode, const u64* a, u64* out,
                           unsigned char* out8, int* zero_flag, size_t n, cudaStream_t st);
    // out[h] = a[h]^-1 (0 for a[h] == 0, reported in zero_flag); ex = p - 2; a and out must not alias
    static cudaError_t inv_batch(const FieldParams& fp, const ExpParams& ex, const u64* a, u64* out, int* zero_flag,
                                 size_t n, cudaStream_t st);
    static cudaError_t split(const FieldParams& fp, bool full, const u64* secrets, const u64* coeffs, size_t cstride,
                             u64* shares, size_t sstride, size_t n, int t, int m, const u64* gtab, u32 tab_bytes,
                             cudaStream_t st);
    // dst: strided matrix (base, stride in limbs) or explicit row pointers (use_rows; small-form tables only)
    static cudaError_t split_gen(const FieldParams& fp, bool full, const ChaChaKey& key, const u64* secrets,
                                 const ShareDst& dst, size_t n, int t, int m, const u64* gtab, u32 tab_bytes, cudaStream_t st);
    // small: 64-bit signed-magnitude lambda table
    static cudaError_t recombine(const FieldParams& fp, bool small, const RowPtrs& rows, int k, int width, const u64* gtab,
                                 u32 tab_bytes, u64* out, size_t ostride, size_t n, cudaStream_t st);
    // small: subset coefficients as 64-bit signed-magnitude constants + D^-1 (api.cu: prss_small_table)
    // simple (with small): d == 1 and the weight is 1
    static cudaError_t prss(const FieldParams& fp, bool small, bool simple, const unsigned char* bytes, size_t subset_stride, int nsub, int d,
                            int chunk_bytes, int bound_bits, const u64* gtab, u32 tab_bytes, u64* out, size_t n,
                            cudaStream_t st);
    static cudaError_t matmul(const FieldParams& fp, const u64* A, const u64* B, u64* C, size_t r, size_t k, size_t c,
                              cudaStream_t st);
    static cudaError_t fill_random(const FieldParams& fp, u64* out, size_t n, u64 base, cudaStream_t st);
    // ---- K6: protocol-local algebra on raw share values (local.cuh) ----
    // out = a*b + c (square: a*a + c, b ignored)
    static cudaError_t fma(const FieldParams& fp, bool square, const u64* a, const u64* b, const u64* c, u64* out, size_t n,
                           cudaStream_t st);
    // out = a*s + t for canonical host scalars s, t
    static cudaError_t axpb(const FieldParams& fp, const u64* a, const u64* s, const u64* t, u64* out, size_t n, cudaStream_t st);
    // out = a & (2^nbits - 1)
    static cudaError_t low_bits(const FieldParams& fp, const u64* a, int nbits, u64* out, size_t n, cudaStream_t st);
    // out8[h] = a[h] != 0 (out8 may be null); *count (device, zeroed by the caller) += non-zero elements
    static cudaError_t nonzero(const FieldParams& fp, const u64* a, unsigned char* out8, unsigned long long* count, size_t n,
                               cudaStream_t st);
    // out[i] = sum_j bits[i*f + j] 2^e(j), e(j) = j or f-1-j
    static cudaError_t bits_compose(const FieldParams& fp, const u64* bits, u64* out, size_t n, int f, bool descending, cudaStream_t st);
    // out[j*ostride + i] = bit e(j) of c[i]   (ostride in elements)
    static cudaError_t bits_decompose(const FieldParams& fp, const u64* c, u64* out, size_t ostride, size_t n, int l, bool descending,
                                      cudaStream_t st);
    // out (C, R) = in (R, C) transposed; out (R, C) = running sums of in down the rows; out = a (R, C) (op) b (C) broadcast
    // over the rows (reflected: b (op) a)
    static cudaError_t transpose(const FieldParams& fp, const u64* in, u64* out, size_t R, size_t C, cudaStream_t st);
    static cudaError_t cumsum_rows(const FieldParams& fp, const u64* in, u64* out, size_t R, size_t C, cudaStream_t st);
    static cudaError_t binop_rows(const FieldParams& fp, int op, bool reflected, const u64* a, const u64* b, u64* out, size_t R,
                                  size_t C, cudaStream_t st);
    // Y (k, v, m, n) = 'same' correlation of X (k, r, m, n) with W (v, r, s, s) over the r input channels + B (v); s odd
    static cudaError_t conv2d(const FieldParams& fp, const u64* X, const u64* W, const u64* B, u64* Y, int k, int r, int m, int n,
                              int v, int s, cudaStream_t st);
};

extern template struct Launch<1>;
extern template struct Launch<2>;
extern template struct Launch<3>;
extern template struct Launch<4>;
