// See shake128_x8.h.  Written from FIPS 202 (same round structure as the scalar sponge in shake128.h); compiled by
// the host compiler only, every function that touches zmm registers carries its own target attribute so that the
// translation unit needs no global -mavx512f (the library must load on hosts without AVX-512).
#include "shake128_x8.h"

#include <immintrin.h>
#include <stdlib.h>
#include <string.h>

namespace mpyc_shake {

bool x8_available() {
    static const bool ok = __builtin_cpu_supports("avx512f") && getenv("MPYC_B200_NO_AVX512") == nullptr;
    return ok;
}

#define X8_TARGET __attribute__((target("avx512f")))
#define V __m512i
#define XOR3(a, b, c) _mm512_ternarylogic_epi64(a, b, c, 0x96)
#define CHI(a, b, c) _mm512_ternarylogic_epi64(a, b, c, 0xD2)   // a ^ (~b & c)
#define ROL(v, s) _mm512_rol_epi64(v, s)

X8_TARGET static void keccak_f1600_x8(uint64_t (*A)[8]) {
    static const uint64_t RC[24] = {
        0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808Aull, 0x8000000080008000ull, 0x000000000000808Bull,
        0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008Aull, 0x0000000000000088ull,
        0x0000000080008009ull, 0x000000008000000Aull, 0x000000008000808Bull, 0x800000000000008Bull, 0x8000000000008089ull,
        0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800Aull, 0x800000008000000Aull,
        0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
#define LD(i) _mm512_load_si512((const void*)A[i])
    V a00 = LD(0), a01 = LD(1), a02 = LD(2), a03 = LD(3), a04 = LD(4), a05 = LD(5), a06 = LD(6), a07 = LD(7), a08 = LD(8),
      a09 = LD(9), a10 = LD(10), a11 = LD(11), a12 = LD(12), a13 = LD(13), a14 = LD(14), a15 = LD(15), a16 = LD(16),
      a17 = LD(17), a18 = LD(18), a19 = LD(19), a20 = LD(20), a21 = LD(21), a22 = LD(22), a23 = LD(23), a24 = LD(24);
#undef LD
    for (int round = 0; round < 24; round++) {
        // theta
        const V c0 = XOR3(XOR3(a00, a05, a10), a15, a20), c1 = XOR3(XOR3(a01, a06, a11), a16, a21),
                c2 = XOR3(XOR3(a02, a07, a12), a17, a22), c3 = XOR3(XOR3(a03, a08, a13), a18, a23),
                c4 = XOR3(XOR3(a04, a09, a14), a19, a24);
        const V d0 = _mm512_xor_si512(c4, ROL(c1, 1)), d1 = _mm512_xor_si512(c0, ROL(c2, 1)), d2 = _mm512_xor_si512(c1, ROL(c3, 1)),
                d3 = _mm512_xor_si512(c2, ROL(c4, 1)), d4 = _mm512_xor_si512(c3, ROL(c0, 1));
#define X(a, d) a = _mm512_xor_si512(a, d)
        X(a00, d0); X(a05, d0); X(a10, d0); X(a15, d0); X(a20, d0);
        X(a01, d1); X(a06, d1); X(a11, d1); X(a16, d1); X(a21, d1);
        X(a02, d2); X(a07, d2); X(a12, d2); X(a17, d2); X(a22, d2);
        X(a03, d3); X(a08, d3); X(a13, d3); X(a18, d3); X(a23, d3);
        X(a04, d4); X(a09, d4); X(a14, d4); X(a19, d4); X(a24, d4);
#undef X
        // rho + pi: lane (x, y) rotated by r[x][y] moves to (y, 2x + 3y)
        const V b00 = a00,           b10 = ROL(a01, 1),  b20 = ROL(a02, 62), b05 = ROL(a03, 28), b15 = ROL(a04, 27),
                b16 = ROL(a05, 36),  b01 = ROL(a06, 44), b11 = ROL(a07, 6),  b21 = ROL(a08, 55), b06 = ROL(a09, 20),
                b07 = ROL(a10, 3),   b17 = ROL(a11, 10), b02 = ROL(a12, 43), b12 = ROL(a13, 25), b22 = ROL(a14, 39),
                b23 = ROL(a15, 41),  b08 = ROL(a16, 45), b18 = ROL(a17, 15), b03 = ROL(a18, 21), b13 = ROL(a19, 8),
                b14 = ROL(a20, 18),  b24 = ROL(a21, 2),  b09 = ROL(a22, 61), b19 = ROL(a23, 56), b04 = ROL(a24, 14);
        // chi (+ iota on lane 0)
        a00 = _mm512_xor_si512(CHI(b00, b01, b02), _mm512_set1_epi64((long long)RC[round]));
        a01 = CHI(b01, b02, b03); a02 = CHI(b02, b03, b04); a03 = CHI(b03, b04, b00); a04 = CHI(b04, b00, b01);
        a05 = CHI(b05, b06, b07); a06 = CHI(b06, b07, b08); a07 = CHI(b07, b08, b09); a08 = CHI(b08, b09, b05); a09 = CHI(b09, b05, b06);
        a10 = CHI(b10, b11, b12); a11 = CHI(b11, b12, b13); a12 = CHI(b12, b13, b14); a13 = CHI(b13, b14, b10); a14 = CHI(b14, b10, b11);
        a15 = CHI(b15, b16, b17); a16 = CHI(b16, b17, b18); a17 = CHI(b17, b18, b19); a18 = CHI(b18, b19, b15); a19 = CHI(b19, b15, b16);
        a20 = CHI(b20, b21, b22); a21 = CHI(b21, b22, b23); a22 = CHI(b22, b23, b24); a23 = CHI(b23, b24, b20); a24 = CHI(b24, b20, b21);
    }
#define ST(i, v) _mm512_store_si512((void*)A[i], v)
    ST(0, a00); ST(1, a01); ST(2, a02); ST(3, a03); ST(4, a04); ST(5, a05); ST(6, a06); ST(7, a07); ST(8, a08); ST(9, a09);
    ST(10, a10); ST(11, a11); ST(12, a12); ST(13, a13); ST(14, a14); ST(15, a15); ST(16, a16); ST(17, a17); ST(18, a18);
    ST(19, a19); ST(20, a20); ST(21, a21); ST(22, a22); ST(23, a23); ST(24, a24);
#undef ST
}

void Shake128x8::reset(int n) {
    memset(st, 0, sizeof st);
    pos = 0;
    squeezing = false;
    count = n < 1 ? 1 : (n > 8 ? 8 : n);
}

void Shake128x8::absorb(const uint8_t* const* in, size_t len) {
    for (size_t i = 0; i < len; i++) {
        for (int s = 0; s < count; s++) st[pos >> 3][s] ^= (uint64_t)in[s][i] << (8 * (pos & 7));
        if (++pos == RATE) {
            keccak_f1600_x8(st);
            pos = 0;
        }
    }
}

void Shake128x8::finish() {     // SHAKE domain separation 1111 + pad10*1, on every active sponge
    for (int s = 0; s < count; s++) {
        st[pos >> 3][s] ^= (uint64_t)0x1F << (8 * (pos & 7));
        st[(RATE - 1) >> 3][s] ^= (uint64_t)0x80 << (8 * ((RATE - 1) & 7));
    }
    keccak_f1600_x8(st);
    pos = 0;
    squeezing = true;
}

void Shake128x8::squeeze(uint8_t* const* out, size_t len) {
    if (!squeezing) finish();
    size_t done = 0;
    while (done < len) {
        if (pos == RATE) {
            keccak_f1600_x8(st);
            pos = 0;
        }
        const size_t take = RATE - pos < len - done ? RATE - pos : len - done;
        for (int s = 0; s < count; s++) {
            uint8_t* dst = out[s] + done;
            size_t p = pos, left = take;
            while (left) {                    // byte p of a sponge's block is byte p & 7 of lane p >> 3
                const size_t off = p & 7, chunk = 8 - off < left ? 8 - off : left;
                const uint64_t lane = st[p >> 3][s];
                memcpy(dst, reinterpret_cast<const uint8_t*>(&lane) + off, chunk);
                dst += chunk;
                p += chunk;
                left -= chunk;
            }
        }
        pos += take;
        done += take;
    }
}

}   // namespace mpyc_shake
