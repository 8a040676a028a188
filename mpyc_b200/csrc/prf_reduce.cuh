// PRF values for an ARBITRARY bound (thresha.PRF.__call__, mpyc/thresha.py:257-261):
//     value = int.from_bytes(chunk, 'little') % bound
// K4 (kernels.cuh) folds the two bounds the secure-randomness protocols use -- the field order and 2^b <= p -- into
// the combine kernel itself.  Every other bound (runtime._convert's (1 << (k+l)) // comb(m,t) + 1, runtime.py:735-739;
// the source field's order applied to a smaller target field, runtime.py:758-760; a power of two above p) goes
// through this kernel first: one thread per chunk, Horner over the chunk's 64-bit limbs from the top with one
// Barrett step per limb against `bound` (any modulus >= 3 that is not a power of two; odd or even), or a mask for
// bound = 2^b.  The values are written as fixed-width little-endian integers of 8*LB bytes, which K4 then consumes as
// chunks with the bound "field order" (a true reduction mod p when bound > p, the identity otherwise).
// For a GF(2^8) field the value is the GF(2)[X] polynomial whose integer encoding is the reduced chunk (gfpx's
// int -> polynomial coercion, mpyc/gfpx.py:73-81), reduced modulo the field polynomial: one byte per value.
#pragma once
#include "kernels.cuh"

// fb: Barrett constants of `bound` (field_setup.h: bound_params_init), LB = limbs of bound.
// pow2_bits > 0: bound = 2^pow2_bits.  gf_poly != 0: one output byte per value (see above), else 8*LB bytes.
template <int LB>
__global__ void __launch_bounds__(MPYC_THREADS)
k_prf_reduce(FieldParams fb, int pow2_bits, const unsigned char* __restrict__ bytes, size_t subset_stride, int nsub,
             size_t count, int chunk_bytes, unsigned char* __restrict__ out, size_t out_stride, unsigned gf_poly) {
    typedef Fp<LB, KIND_GENERIC> F;
    constexpr int N = 2 * LB;
    const int nl = (chunk_bytes + 7) >> 3;
    const size_t total = (size_t)nsub * count;
    const size_t nth = (size_t)gridDim.x * blockDim.x;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += nth) {
        const size_t S = idx / count, i = idx % count;
        const unsigned char* src = bytes + S * subset_stride + i * (size_t)chunk_bytes;
        u32 v[N];
        zero_n<N>(v);
        if (pow2_bits > 0) {
#pragma unroll
            for (int l = 0; l < LB; l++) {
                u64 limb = prss_chunk_limb(src, l, chunk_bytes, false);
                const int top = pow2_bits - 64 * l;
                if (top < 64) limb &= top > 0 ? ((1ull << top) - 1) : 0ull;
                set64(v, l, limb);
            }
        } else {
            for (int w = nl - 1; w >= 0; w--) {     // v <- (v * 2^64 + limb) mod bound
                u32 x[N + 2];
                set64(x, 0, prss_chunk_limb(src, w, chunk_bytes, false));
#pragma unroll
                for (int l = 0; l < N; l++) x[l + 2] = v[l];
                F::barrett_small(v, x, fb);
            }
        }
        if (gf_poly) {
            unsigned r = 0;
            for (int b = 64 * LB - 1; b >= 0; b--) {
                r = (r << 1) | ((v[b >> 5] >> (b & 31)) & 1u);
                if (r & 0x100u) r ^= gf_poly;
            }
            out[S * out_stride + i] = (unsigned char)r;
        } else {
            u64* dst = reinterpret_cast<u64*>(out + S * out_stride) + i * LB;
#pragma unroll
            for (int l = 0; l < LB; l++) dst[l] = get64(v, l);
        }
    }
}
