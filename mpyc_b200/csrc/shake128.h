// SHAKE128 (FIPS 202) on the host: the XOF behind thresha.PRF (reference mpyc/thresha.py:220-266, which
// calls hashlib.shake_128(key + s).digest(n * byte_length)).  One sponge is strictly sequential -- squeezing
// block i+1 needs the permutation of block i -- so a PRSS call is parallel only across its C(m-1, t) key
// subsets: the library runs one sponge per host thread (hashlib does not release the GIL while squeezing),
// squeezing chunk by chunk into pinned staging buffers while the previous chunk is copied and combined on
// the GPU (api.cu: mpyc_b200_prss_host).  Written from the FIPS 202 specification; checked against
// hashlib.shake_128 in tests/test_shake128.py.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

namespace mpyc_shake {

// Keccak-f[1600]: 24 rounds of theta, rho+pi, chi, iota on 25 little-endian 64-bit lanes A[x + 5y].
// Every lane has its own variable and the steps of a round are straight-line code, so the state stays in
// registers for all 24 rounds.
#define MPYC_KECCAK_ROL(v, s) (((v) << (s)) | ((v) >> (64 - (s))))
static inline void keccak_f1600(uint64_t* A) {
    static const uint64_t RC[24] = {
        0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808Aull, 0x8000000080008000ull, 0x000000000000808Bull,
        0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008Aull, 0x0000000000000088ull,
        0x0000000080008009ull, 0x000000008000000Aull, 0x000000008000808Bull, 0x800000000000008Bull, 0x8000000000008089ull,
        0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800Aull, 0x800000008000000Aull,
        0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
    uint64_t a00 = A[0], a01 = A[1], a02 = A[2], a03 = A[3], a04 = A[4], a05 = A[5], a06 = A[6], a07 = A[7], a08 = A[8],
             a09 = A[9], a10 = A[10], a11 = A[11], a12 = A[12], a13 = A[13], a14 = A[14], a15 = A[15], a16 = A[16],
             a17 = A[17], a18 = A[18], a19 = A[19], a20 = A[20], a21 = A[21], a22 = A[22], a23 = A[23], a24 = A[24];
    for (int round = 0; round < 24; round++) {
        // theta
        const uint64_t c0 = a00 ^ a05 ^ a10 ^ a15 ^ a20, c1 = a01 ^ a06 ^ a11 ^ a16 ^ a21, c2 = a02 ^ a07 ^ a12 ^ a17 ^ a22,
                       c3 = a03 ^ a08 ^ a13 ^ a18 ^ a23, c4 = a04 ^ a09 ^ a14 ^ a19 ^ a24;
        const uint64_t d0 = c4 ^ MPYC_KECCAK_ROL(c1, 1), d1 = c0 ^ MPYC_KECCAK_ROL(c2, 1), d2 = c1 ^ MPYC_KECCAK_ROL(c3, 1),
                       d3 = c2 ^ MPYC_KECCAK_ROL(c4, 1), d4 = c3 ^ MPYC_KECCAK_ROL(c0, 1);
        a00 ^= d0; a05 ^= d0; a10 ^= d0; a15 ^= d0; a20 ^= d0;
        a01 ^= d1; a06 ^= d1; a11 ^= d1; a16 ^= d1; a21 ^= d1;
        a02 ^= d2; a07 ^= d2; a12 ^= d2; a17 ^= d2; a22 ^= d2;
        a03 ^= d3; a08 ^= d3; a13 ^= d3; a18 ^= d3; a23 ^= d3;
        a04 ^= d4; a09 ^= d4; a14 ^= d4; a19 ^= d4; a24 ^= d4;
        // rho + pi: lane (x, y) rotated by r[x][y] moves to (y, 2x + 3y)   (FIPS 202 tables 2 and 3.2.3)
        const uint64_t b00 = a00,                        b10 = MPYC_KECCAK_ROL(a01, 1),  b20 = MPYC_KECCAK_ROL(a02, 62),
                       b05 = MPYC_KECCAK_ROL(a03, 28),   b15 = MPYC_KECCAK_ROL(a04, 27),
                       b16 = MPYC_KECCAK_ROL(a05, 36),   b01 = MPYC_KECCAK_ROL(a06, 44), b11 = MPYC_KECCAK_ROL(a07, 6),
                       b21 = MPYC_KECCAK_ROL(a08, 55),   b06 = MPYC_KECCAK_ROL(a09, 20),
                       b07 = MPYC_KECCAK_ROL(a10, 3),    b17 = MPYC_KECCAK_ROL(a11, 10), b02 = MPYC_KECCAK_ROL(a12, 43),
                       b12 = MPYC_KECCAK_ROL(a13, 25),   b22 = MPYC_KECCAK_ROL(a14, 39),
                       b23 = MPYC_KECCAK_ROL(a15, 41),   b08 = MPYC_KECCAK_ROL(a16, 45), b18 = MPYC_KECCAK_ROL(a17, 15),
                       b03 = MPYC_KECCAK_ROL(a18, 21),   b13 = MPYC_KECCAK_ROL(a19, 8),
                       b14 = MPYC_KECCAK_ROL(a20, 18),   b24 = MPYC_KECCAK_ROL(a21, 2),  b09 = MPYC_KECCAK_ROL(a22, 61),
                       b19 = MPYC_KECCAK_ROL(a23, 56),   b04 = MPYC_KECCAK_ROL(a24, 14);
        // chi (+ iota on lane 0)
        a00 = b00 ^ (~b01 & b02) ^ RC[round]; a01 = b01 ^ (~b02 & b03); a02 = b02 ^ (~b03 & b04); a03 = b03 ^ (~b04 & b00); a04 = b04 ^ (~b00 & b01);
        a05 = b05 ^ (~b06 & b07); a06 = b06 ^ (~b07 & b08); a07 = b07 ^ (~b08 & b09); a08 = b08 ^ (~b09 & b05); a09 = b09 ^ (~b05 & b06);
        a10 = b10 ^ (~b11 & b12); a11 = b11 ^ (~b12 & b13); a12 = b12 ^ (~b13 & b14); a13 = b13 ^ (~b14 & b10); a14 = b14 ^ (~b10 & b11);
        a15 = b15 ^ (~b16 & b17); a16 = b16 ^ (~b17 & b18); a17 = b17 ^ (~b18 & b19); a18 = b18 ^ (~b19 & b15); a19 = b19 ^ (~b15 & b16);
        a20 = b20 ^ (~b21 & b22); a21 = b21 ^ (~b22 & b23); a22 = b22 ^ (~b23 & b24); a23 = b23 ^ (~b24 & b20); a24 = b24 ^ (~b20 & b21);
    }
    A[0] = a00; A[1] = a01; A[2] = a02; A[3] = a03; A[4] = a04; A[5] = a05; A[6] = a06; A[7] = a07; A[8] = a08; A[9] = a09;
    A[10] = a10; A[11] = a11; A[12] = a12; A[13] = a13; A[14] = a14; A[15] = a15; A[16] = a16; A[17] = a17; A[18] = a18;
    A[19] = a19; A[20] = a20; A[21] = a21; A[22] = a22; A[23] = a23; A[24] = a24;
}
#undef MPYC_KECCAK_ROL

struct Shake128 {
    static constexpr size_t RATE = 168;   // bytes: 1600 - 2*128 bits of capacity
    uint64_t st[25];
    size_t pos;          // absorbing: bytes absorbed into the current block; squeezing: bytes already read from it
    bool squeezing;

    Shake128() { reset(); }
    void reset() {
        memset(st, 0, sizeof st);
        pos = 0;
        squeezing = false;
    }
    void xor_byte(size_t i, uint8_t v) { st[i >> 3] ^= (uint64_t)v << (8 * (i & 7)); }
    void absorb(const uint8_t* in, size_t len) {
        for (size_t i = 0; i < len; i++) {
            xor_byte(pos++, in[i]);
            if (pos == RATE) {
                keccak_f1600(st);
                pos = 0;
            }
        }
    }
    void finish() {      // SHAKE domain separation 1111 + pad10*1
        xor_byte(pos, 0x1F);
        xor_byte(RATE - 1, 0x80);
        keccak_f1600(st);
        pos = 0;
        squeezing = true;
    }
    void squeeze(uint8_t* out, size_t len) {
        if (!squeezing) finish();
        while (len) {
            if (pos == RATE) {
                keccak_f1600(st);
                pos = 0;
            }
            size_t take = RATE - pos < len ? RATE - pos : len;
            // little-endian host: the state bytes are the lane bytes in memory order
            memcpy(out, reinterpret_cast<const uint8_t*>(st) + pos, take);
            out += take;
            pos += take;
            len -= take;
        }
    }
};

}   // namespace mpyc_shake
