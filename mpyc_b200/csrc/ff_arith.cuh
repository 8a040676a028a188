// Multi-limb prime-field arithmetic for sm_100a (and, for table preparation and
// algorithm tests, the host).  Little-endian 64-bit limbs, L in {1,2,3,4}.
//
// Two reduction families, chosen per modulus when the field context is created:
//   * pseudo-Mersenne p = 2^k - c with c < 2^16 and k >= 56  (MPyC's default primes from
//     finfields.find_prime_root, reference mpyc/finfields.py:325-331: 2^61-1, 2^64-189,
//     2^69-93, 2^128-173, 2^256-189): fold the high part down with a multiply by c.
//       KIND_PM_ALIGNED  k == 64 L     (fold at a limb boundary)
//       KIND_PM_SHIFT    k % 64 != 0   (fold at bit k with funnel shifts)
//   * KIND_GENERIC: any odd p < 2^(64 L).  Montgomery reduction with ONE GUARD LIMB,
//     R' = 2^(64 (L+1)), so that a lazily accumulated sum of up to 2^20 full products
//     (< K p^2 < p R') reduces with a single REDC and one conditional subtraction.
//     Constants ("table form") are pre-multiplied by R' on the host, data stays canonical:
//     REDC(sum_i lambda_i R' * share_i) = sum_i lambda_i share_i mod p.
//
// Everything at the kernel boundary is a canonical residue in [0, p); Montgomery form never
// leaves a kernel.  The arithmetic replaces the reference's Python-int `(a op b) % p` inside
// NumPy object loops (mpyc/finfields.py:717-725,1056-1124; mpyc/thresha.py:63,129).
#pragma once
#include <stdint.h>

typedef unsigned long long u64;
typedef unsigned int u32;

#if defined(__CUDACC__)
#define FF_HD __host__ __device__ __forceinline__
#else
#define FF_HD inline
#endif

enum { KIND_GENERIC = 0, KIND_PM_ALIGNED = 1, KIND_PM_SHIFT = 2 };

// Upper bound on the number of full-width products that may be accumulated lazily
// (in a 2L+1 limb accumulator) before a reduction is required.
#define FF_MAX_LAZY_TERMS (1u << 20)

struct FieldParams {
    u64 p[4];    // modulus
    u64 r1[4];   // GENERIC: R' mod p      (table form of 1)
    u64 r2[4];   // GENERIC: R'^2 mod p
    u64 pinv;    // GENERIC: -p^-1 mod 2^64
    u64 c;       // PM: 2^k - p
    u32 k;       // bit length of p (PM: the exponent)
    u32 s;       // k % 64
    u32 L;       // limbs
    u32 kind;
};

// ---------------------------------------------------------------------------------------
// n-limb primitives.  Device: PTX carry chains (add.cc / madc.hi.cc ...), one instruction per
// asm statement, fully unrolled so limbs stay in registers.  Host: unsigned __int128.
// ---------------------------------------------------------------------------------------

template <int N>
FF_HD u64 add_n(u64* r, const u64* a, const u64* b) {
#ifdef __CUDA_ARCH__
    asm volatile("add.cc.u64 %0, %1, %2;" : "=l"(r[0]) : "l"(a[0]), "l"(b[0]));
#pragma unroll
    for (int i = 1; i < N; i++)
        asm volatile("addc.cc.u64 %0, %1, %2;" : "=l"(r[i]) : "l"(a[i]), "l"(b[i]));
    u64 cy;
    asm volatile("addc.u64 %0, 0, 0;" : "=l"(cy));
    return cy;
#else
    unsigned __int128 acc = 0;
    for (int i = 0; i < N; i++) {
        acc += (unsigned __int128)a[i] + b[i];
        r[i] = (u64)acc;
        acc >>= 64;
    }
    return (u64)acc;
#endif
}

// r = a - b, returns 1 on borrow
template <int N>
FF_HD u64 sub_n(u64* r, const u64* a, const u64* b) {
#ifdef __CUDA_ARCH__
    asm volatile("sub.cc.u64 %0, %1, %2;" : "=l"(r[0]) : "l"(a[0]), "l"(b[0]));
#pragma unroll
    for (int i = 1; i < N; i++)
        asm volatile("subc.cc.u64 %0, %1, %2;" : "=l"(r[i]) : "l"(a[i]), "l"(b[i]));
    u64 bw;
    asm volatile("subc.u64 %0, 0, 0;" : "=l"(bw));
    return bw & 1ull;
#else
    u64 bw = 0;
    for (int i = 0; i < N; i++) {
        unsigned __int128 d = (unsigned __int128)a[i] - b[i] - bw;
        r[i] = (u64)d;
        bw = (u64)(d >> 64) & 1ull;
    }
    return bw;
#endif
}

// acc[0..REM) += a[0..N) * b   (REM >= N+1; a carry out of acc[REM-1] must be impossible)
template <int N, int REM>
FF_HD void mac_1(u64* acc, const u64* a, u64 b) {
    static_assert(REM >= N + 1, "accumulator too short");
#ifdef __CUDA_ARCH__
    asm volatile("mad.lo.cc.u64 %0, %1, %2, %0;" : "+l"(acc[0]) : "l"(a[0]), "l"(b));
#pragma unroll
    for (int i = 1; i < N; i++)
        asm volatile("madc.lo.cc.u64 %0, %1, %2, %0;" : "+l"(acc[i]) : "l"(a[i]), "l"(b));
#pragma unroll
    for (int i = N; i < REM; i++)
        asm volatile("addc.cc.u64 %0, %0, 0;" : "+l"(acc[i]));
    asm volatile("mad.hi.cc.u64 %0, %1, %2, %0;" : "+l"(acc[1]) : "l"(a[0]), "l"(b));
#pragma unroll
    for (int i = 1; i < N; i++)
        asm volatile("madc.hi.cc.u64 %0, %1, %2, %0;" : "+l"(acc[i + 1]) : "l"(a[i]), "l"(b));
#pragma unroll
    for (int i = N + 1; i < REM; i++)
        asm volatile("addc.cc.u64 %0, %0, 0;" : "+l"(acc[i]));
#else
    unsigned __int128 cy = 0;
    for (int i = 0; i < N; i++) {
        cy += (unsigned __int128)a[i] * b + acc[i];
        acc[i] = (u64)cy;
        cy >>= 64;
    }
    for (int i = N; i < REM; i++) {
        cy += acc[i];
        acc[i] = (u64)cy;
        cy >>= 64;
    }
#endif
}

// acc[0..W) += a[0..LA) * b[0..LB)      (W >= LA + LB); row J of the schoolbook product
template <int LA, int LB, int W, int J = 0>
FF_HD void mac_n(u64* acc, const u64* a, const u64* b) {
    static_assert(W >= LA + LB, "accumulator too short");
    if constexpr (J < LB) {
        mac_1<LA, W - J>(acc + J, a, b[J]);
        mac_n<LA, LB, W, J + 1>(acc, a, b);
    }
}

template <int N>
FF_HD void zero_n(u64* r) {
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = 0;
}

template <int N>
FF_HD void copy_n(u64* r, const u64* a) {
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = a[i];
}

template <int N>
FF_HD bool is_zero_n(const u64* a) {
    u64 t = 0;
#pragma unroll
    for (int i = 0; i < N; i++) t |= a[i];
    return t == 0;
}

// r = take ? t : r   (branch-free select)
template <int N>
FF_HD void select_n(u64* r, const u64* t, bool take) {
    u64 m = take ? ~0ull : 0ull;
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = (t[i] & m) | (r[i] & ~m);
}

// ---------------------------------------------------------------------------------------
// Field<L, KIND>
// ---------------------------------------------------------------------------------------

template <int L, int KIND>
struct Fp {
    static constexpr int WACC = 2 * L + 1;   // lazy accumulator width (full products)
    static constexpr int WSM = L + 1;        // lazy accumulator width (element * 64-bit constant)

    // a in [0, 2p) (as L limbs plus an explicit carry bit `hi`)  ->  [0, p)
    static FF_HD void csub(u64* a, u64 hi, const FieldParams& f) {
        u64 t[L];
        u64 bw = sub_n<L>(t, a, f.p);
        select_n<L>(a, t, (hi != 0) | (bw == 0));
    }

    static FF_HD void add(u64* r, const u64* a, const u64* b, const FieldParams& f) {
        u64 cy = add_n<L>(r, a, b);
        csub(r, cy, f);
    }

    static FF_HD void sub(u64* r, const u64* a, const u64* b, const FieldParams& f) {
        u64 t[L];
        u64 bw = sub_n<L>(r, a, b);
        add_n<L>(t, r, f.p);
        select_n<L>(r, t, bw != 0);
    }

    static FF_HD void neg(u64* r, const u64* a, const FieldParams& f) {
        u64 t[L];
        bool z = is_zero_n<L>(a);
        sub_n<L>(t, f.p, a);
        zero_n<L>(r);
        select_n<L>(r, t, !z);
    }

    // ---- pseudo-Mersenne: true reduction of a W-limb value ---------------------------
    // Precondition (met by every caller in this library): x < 2^(2k+20), k = bit length of p
    // (covers products < p^2, lazy sums of < 2^20 products, and 64-bit-constant sums < 2^64 p).
    template <int W>
    static FF_HD void pm_reduce(u64* r, const u64* x, const FieldParams& f) {
        static_assert(W > L && W <= 2 * L + 1, "unsupported width");
        const u64 c = f.c;
        u64 r1[L + 1];
        if constexpr (KIND == KIND_PM_ALIGNED) {
            constexpr int H = W - L;                     // limbs above the fold point
            copy_n<L>(r1, x);
            r1[L] = 0;
            constexpr int HM = H <= L ? H : L;
            mac_1<HM, L + 1>(r1, x + L, c);
            if constexpr (H == L + 1) r1[L] += x[2 * L] * c;       // top limb of a lazy sum is < 2^20
            u64 u = r1[L] * c;                           // < 2^64 by the preconditions
            u64 uu[L];
            zero_n<L>(uu);
            uu[0] = u;
            u64 cy = add_n<L>(r, r1, uu);
            zero_n<L>(uu);
            uu[0] = cy ? c : 0;                          // 2^(64L) = c (mod p); cannot carry again
            add_n<L>(r, r, uu);
            csub(r, 0, f);
        } else {
            const u32 s = f.s;                           // 1..63, p = 2^(64(L-1)+s) - c
            const u64 mask = (1ull << s) - 1;
            u64 hi[L + 1];
#pragma unroll
            for (int i = 0; i <= L; i++) {
                u64 lo_part = (L - 1 + i < W) ? (x[L - 1 + i] >> s) : 0;
                u64 hi_part = (L + i < W) ? (x[L + i] << (64 - s)) : 0;
                hi[i] = lo_part | hi_part;
            }
            copy_n<L>(r1, x);
            r1[L - 1] &= mask;
            r1[L] = 0;
            mac_1<L, L + 1>(r1, hi, c);
            if constexpr (2 * L - 1 < W) r1[L] += hi[L] * c;
            u64 top = (r1[L - 1] >> s) | (r1[L] << (64 - s));
            r1[L - 1] &= mask;
            u64 uu[L];
            zero_n<L>(uu);
            uu[0] = top * c;
            add_n<L>(r, r1, uu);
            csub(r, 0, f);
        }
    }

    // L+1 Montgomery rounds on T (2L+2 limbs): round I clears limb I
    template <int I>
    static FF_HD void redc_rounds(u64* T, const FieldParams& f) {
        if constexpr (I <= L) {
            u64 m = T[I] * f.pinv;
            mac_1<L, 2 * L + 2 - I>(T + I, f.p, m);
            redc_rounds<I + 1>(T, f);
        }
    }

    // ---- generic: Montgomery REDC with guard limb: returns x / R' mod p, canonical -----
    // Precondition: x < p * R'  (R' = 2^(64(L+1))).
    template <int W>
    static FF_HD void redc(u64* r, const u64* x, const FieldParams& f) {
        static_assert(W >= 1 && W <= 2 * L + 1, "unsupported width");
        u64 T[2 * L + 2];
#pragma unroll
        for (int i = 0; i < 2 * L + 2; i++) T[i] = i < W ? x[i] : 0;
        redc_rounds<0>(T, f);
        // result = T[L+1 .. 2L+2) < 2p
        copy_n<L>(r, T + L + 1);
        csub(r, T[2 * L + 1], f);
    }

    // acc (WACC limbs) += a * tab        (tab: L limbs, table form)
    static FF_HD void mac(u64* acc, const u64* a, const u64* tab) { mac_n<L, L, WACC>(acc, a, tab); }

    // r = lazily accumulated sum, reduced.  GENERIC: the table-form factor R' is divided out.
    static FF_HD void finish(u64* r, const u64* acc, const FieldParams& f) {
        if constexpr (KIND == KIND_GENERIC) redc<WACC>(r, acc, f);
        else pm_reduce<WACC>(r, acc, f);
    }

    // "domain" multiplication: PM: plain modular product.  GENERIC: Montgomery product ab/R'.
    static FF_HD void dmul(u64* r, const u64* a, const u64* b, const FieldParams& f) {
        u64 x[2 * L];
        zero_n<2 * L>(x);
        mac_n<L, L, 2 * L>(x, a, b);
        if constexpr (KIND == KIND_GENERIC) redc<2 * L>(r, x, f);
        else pm_reduce<2 * L>(r, x, f);
    }

    static FF_HD void to_dom(u64* r, const u64* a, const FieldParams& f) {
        if constexpr (KIND == KIND_GENERIC) dmul(r, a, f.r2, f);
        else copy_n<L>(r, a);
    }

    static FF_HD void from_dom(u64* r, const u64* a, const FieldParams& f) {
        if constexpr (KIND == KIND_GENERIC) redc<L>(r, a, f);
        else copy_n<L>(r, a);
    }

    static FF_HD void dom_one(u64* r, const FieldParams& f) {
        if constexpr (KIND == KIND_GENERIC) copy_n<L>(r, f.r1);
        else {
            zero_n<L>(r);
            r[0] = 1;
        }
    }

    // canonical * canonical -> canonical
    static FF_HD void mul(u64* r, const u64* a, const u64* b, const FieldParams& f) {
        if constexpr (KIND == KIND_GENERIC) {
            u64 t[L];
            dmul(t, a, b, f);        // ab / R'
            dmul(r, t, f.r2, f);     // ab
        } else {
            dmul(r, a, b, f);
        }
    }

    // true reduction of an arbitrary (L+1)-limb value
    static FF_HD void reduce_small(u64* r, const u64* x, const FieldParams& f) {
        if constexpr (KIND == KIND_GENERIC) {
            u64 t[L];
            redc<L + 1>(t, x, f);    // x / R'
            dmul(r, t, f.r2, f);     // x
        } else {
            pm_reduce<L + 1>(r, x, f);
        }
    }

    // r = a^e in the domain; e = little-endian limbs, ebits = bit length of e (>= 0)
    static FF_HD void dpow(u64* r, const u64* a, const u64* e, int ebits, const FieldParams& f) {
        u64 acc[L];
        dom_one(acc, f);
        for (int i = ebits - 1; i >= 0; i--) {
            u64 sq[L];
            dmul(sq, acc, acc, f);
            u64 mu[L];
            dmul(mu, sq, a, f);
            bool bit = (e[i >> 6] >> (i & 63)) & 1;   // per-thread exponent: stay branch-free
            copy_n<L>(acc, sq);
            select_n<L>(acc, mu, bit);
        }
        copy_n<L>(r, acc);
    }

    // same, for an exponent that is identical for every thread (public: p-2, (p+1)/4, ...):
    // the branch on the exponent bit is warp-uniform, so zero bits cost one squaring only.
    static FF_HD void dpow_uniform(u64* r, const u64* a, const u64* e, int ebits, const FieldParams& f) {
        u64 acc[L];
        dom_one(acc, f);
        for (int i = ebits - 1; i >= 0; i--) {
            dmul(acc, acc, acc, f);
            if ((e[i >> 6] >> (i & 63)) & 1) dmul(acc, acc, a, f);
        }
        copy_n<L>(r, acc);
    }
};
