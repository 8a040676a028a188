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
    switch (L) {
        case 1: montgomery_constants<1>(fp); break;
        case 2: montgomery_constants<2>(fp); break;
        case 3: montgomery_constants<3>(fp); break;
        case 4: montgomery_constants<4>(fp); break;
    }
}
