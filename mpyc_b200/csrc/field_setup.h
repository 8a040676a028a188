// Host-side construction of FieldParams from a modulus: picks the reduction family and computes
// the Montgomery constants.  Shared by the C-ABI (api.cu) and the host-only algorithm checker
// (tests/native/host_check.cpp).
#pragma once
#include <string.h>
#include "ff_arith.cuh"

static inline int bit_length(const u64* x, int n) {
    for (int i = n - 1; i >= 0; i--)
        if (x[i]) return 64 * i + 64 - __builtin_clzll(x[i]);
    return 0;
}

// q[0..5) = low 320 bits of floor(2^e / d), d = nd limbs (nd <= 4, d != 0): restoring division, bit by bit
static inline void pow2_div(int e, const u64* d, int nd, u64* q) {
    u64 rem[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 5; i++) q[i] = 0;
    for (int bit = e; bit >= 0; bit--) {
        for (int i = 4; i > 0; i--) rem[i] = (rem[i] << 1) | (rem[i - 1] >> 63);   // rem < d <= 2^256: no overflow
        rem[0] = (rem[0] << 1) | (bit == e ? 1u : 0u);
        bool ge = true;   // rem >= d ?
        for (int i = 4; i >= 0; i--) {
            const u64 di = i < nd ? d[i] : 0;
            if (rem[i] != di) {
                ge = rem[i] > di;
                break;
            }
        }
        if (ge) {
            u64 bw = 0;
            for (int i = 0; i < 5; i++) {
                const u64 di = i < nd ? d[i] : 0;
                const u64 t = rem[i] - di - bw;
                bw = (rem[i] < di + bw) || (bw && di == ~0ull);
                rem[i] = t;
            }
            if (bit < 320) q[bit >> 6] |= 1ull << (bit & 63);
        }
    }
}

// Barrett constants of a GENERIC field (see FieldParams)
static inline void barrett_constants(FieldParams& fp) {
    const int L = (int)fp.L;
    u64 q[5];
    fp.nsh = 64u * L - fp.k;
    for (int i = 0; i <= L; i++) fp.p2[i] = ((i < L ? fp.p[i] : 0) << 1) | (i ? fp.p[i - 1] >> 63 : 0);
    pow2_div((int)fp.k + 64, fp.p, L, q);           // 2^64 <= quotient < 2^65
    fp.mus = q[0];
    pow2_div((int)fp.k + 32, fp.p, L, q);           // 2^32 <= quotient < 2^33
    fp.mus32 = (u32)q[0];
    for (int i = L - 1; i >= 0; i--)
        fp.pn[i] = fp.nsh ? ((fp.p[i] << fp.nsh) | (i ? fp.p[i - 1] >> (64 - fp.nsh) : 0)) : fp.p[i];
    for (int i = 0; i <= L; i++) fp.pn2[i] = ((i < L ? fp.pn[i] : 0) << 1) | (i ? fp.pn[i - 1] >> 63 : 0);
    pow2_div(128 * L, fp.pn, L, q);                 // 2^(64L) <= quotient < 2^(64L+1)
    for (int i = 0; i < L; i++) fp.muf[i] = q[i];
}

template <int LL>
static inline void montgomery_constants(FieldParams& fp) {
    typedef Fp<LL, KIND_GENERIC> F;
    u32 x[2 * LL] = {0};
    x[0] = 1;   // 1 mod p (p >= 3)
    for (int i = 0; i < 64 * (LL + 1); i++) F::add(x, x, x, fp);   // R' = 2^(64(L+1)) mod p by doubling
    for (int i = 0; i < LL; i++) fp.r1[i] = get64(x, i);
    for (int i = 0; i < 64 * (LL + 1); i++) F::add(x, x, x, fp);   // R'^2 mod p
    for (int i = 0; i < LL; i++) fp.r2[i] = get64(x, i);
}

// Barrett constants only, for a modulus that need not be prime or odd: the `bound` of a PRF (prf_reduce.cuh).
// bound: nlimbs limbs (top limb non-zero), bound >= 3 and not a power of two (then 2^64 <= floor(2^(k+64)/bound) < 2^65).
static inline void bound_params_init(const uint64_t* bound, int nlimbs, FieldParams* out) {
    FieldParams& fp = *out;
    memset(&fp, 0, sizeof fp);
    for (int i = 0; i < nlimbs; i++) fp.p[i] = bound[i];
    fp.L = nlimbs;
    fp.k = bit_length(fp.p, nlimbs);
    fp.s = fp.k & 63;
    fp.kind = KIND_GENERIC;
    barrett_constants(fp);
}

// modulus: nlimbs (already stripped of leading zero limbs, 1..4) limbs of an odd p >= 3
static inline void field_params_init(const uint64_t* modulus, int nlimbs, FieldParams* out) {
    FieldParams& fp = *out;
    memset(&fp, 0, sizeof fp);
    const int L = nlimbs;
    for (int i = 0; i < L; i++) fp.p[i] = modulus[i];
    fp.L = L;
    fp.k = bit_length(fp.p, L);
    fp.s = fp.k & 63;
    // pseudo-Mersenne test: p = 2^k - c with c < 2^16 and k >= 56 (all bits above limb 0 set)
    fp.kind = KIND_GENERIC;
    bool high_ones = true;
    for (int i = 1; i < L; i++) {
        u64 want = (i == L - 1 && fp.s) ? ((1ull << fp.s) - 1) : ~0ull;
        if (fp.p[i] != want) high_ones = false;
    }
    const u64 low_full = (L == 1 && fp.s) ? ((1ull << fp.s) - 1) : ~0ull;   // 2^k - 1 restricted to limb 0
    const u64 c = low_full - fp.p[0] + 1;
    if (high_ones && fp.k >= 56 && c >= 1 && c < (1ull << 16)) {
        fp.c = c;
        fp.kind = fp.s ? KIND_PM_SHIFT : KIND_PM_ALIGNED;
    }
    u64 inv = fp.p[0];   // Newton iteration: p^-1 mod 2^64
    for (int i = 0; i < 6; i++) inv *= 2 - fp.p[0] * inv;
    fp.pinv = 0 - inv;
    barrett_constants(fp);
    switch (L) {
        case 1: montgomery_constants<1>(fp); break;
        case 2: montgomery_constants<2>(fp); break;
        case 3: montgomery_constants<3>(fp); break;
        case 4: montgomery_constants<4>(fp); break;
    }
}
