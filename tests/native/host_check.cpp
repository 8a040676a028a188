// Host-only checker for the algorithms in mpyc_b200/csrc/ff_arith.cuh (compiled with g++; the
// __int128 host path of the n-limb primitives, everything above them shared with the kernels).
// Reads commands from stdin, one per line, all integers in hex without 0x:
//   field <p>                         -> prints "kind L k"
//   mul|add|sub <a> <b> ; neg <a>
//   lazy <K> a1 b1 ... aK bK          -> sum a_i*b_i mod p via mac (b in table form) + finish
//   small <K> <s> a1 v1 ... aK vK     -> (s + sum a_i*v_i) mod p via mac_const + reduce_small
//   redsmall <x>                      -> x mod p for x < 2^(k+64) via reduce_small
//   redsmall32 <x>                    -> x mod p for x < 2^(k+31) via reduce_small_q32 with q32 set
//   pow <a> <e>
// TEST INFRASTRUCTURE: not part of the shipped library.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <iostream>
#include <sstream>
#include "../../mpyc_b200/csrc/field_setup.h"

// n = number of 32-bit limbs
static void parse_hex(const std::string& s, u32* out, int n) {
    for (int i = 0; i < n; i++) out[i] = 0;
    int pos = 0;
    for (int i = (int)s.size() - 1; i >= 0; i--, pos++) {
        char c = s[i];
        u32 v = c <= '9' ? c - '0' : (c | 32) - 'a' + 10;
        if (pos / 8 < n) out[pos / 8] |= v << (4 * (pos % 8));
    }
}
static std::string to_hex(const u32* x, int n) {
    char buf[32];
    std::string s;
    bool started = false;
    for (int i = n - 1; i >= 0; i--) {
        if (!started) {
            if (x[i] == 0 && i > 0) continue;
            snprintf(buf, sizeof buf, "%x", x[i]);
            started = true;
        } else snprintf(buf, sizeof buf, "%08x", x[i]);
        s += buf;
    }
    return s;
}

static FieldParams fp;

template <int L, int K>
static std::string run(const std::string& cmd, std::istringstream& in) {
    typedef Fp<L, K> F;
    constexpr int N = 2 * L;
    std::string t;
    u32 a[N], b[N], r[N];
    if (cmd == "mul" || cmd == "add" || cmd == "sub") {
        in >> t; parse_hex(t, a, N);
        in >> t; parse_hex(t, b, N);
        if (cmd == "mul") F::mul(r, a, b, fp);
        else if (cmd == "add") F::add(r, a, b, fp);
        else F::sub(r, a, b, fp);
        return to_hex(r, N);
    }
    if (cmd == "neg") {
        in >> t; parse_hex(t, a, N);
        F::neg(r, a, fp);
        return to_hex(r, N);
    }
    if (cmd == "lazy") {
        int cnt; in >> cnt;
        u32 acc[F::WACC];
        zero_n<F::WACC>(acc);
        for (int i = 0; i < cnt; i++) {
            in >> t; parse_hex(t, a, N);
            in >> t; parse_hex(t, b, N);
            u32 tb[N];
            F::to_dom(tb, b, fp);
            F::mac(acc, a, tb);
        }
        F::finish(r, acc, fp);
        return to_hex(r, N);
    }
    if (cmd == "small") {
        int cnt; in >> cnt;
        u32 acc[F::WSM];
        in >> t; parse_hex(t, acc, N);
        acc[N] = acc[N + 1] = 0;
        for (int i = 0; i < cnt; i++) {
            u32 v[2];
            in >> t; parse_hex(t, a, N);
            in >> t; parse_hex(t, v, 2);
            F::mac_const(acc, a, (u64)v[0] | ((u64)v[1] << 32));
        }
        F::reduce_small(r, acc, fp);
        return to_hex(r, N);
    }
    if (cmd == "redsmall32") {           // x < 2^(k+31): the one-limb-quotient / single-fold forms (f.q32 set)
        u32 x[N + 2];
        in >> t; parse_hex(t, x, N + 2);
        FieldParams f2 = fp;
        f2.q32 = 1;
        F::reduce_small_q32(r, x, f2);
        return to_hex(r, N);
    }
    if (cmd == "redsmall") {
        u32 x[N + 2];
        in >> t; parse_hex(t, x, N + 2);
        F::reduce_small(r, x, fp);
        return to_hex(r, N);
    }
    if (cmd == "pow") {
        u32 e32[16];
        u64 e[8];
        in >> t; parse_hex(t, a, N);
        in >> t; parse_hex(t, e32, 16);
        for (int i = 0; i < 8; i++) e[i] = get64(e32, i);
        u32 x[N];
        F::to_dom(x, a, fp);
        F::dpow_uniform(x, x, e, bit_length(e, 8), fp);
        F::from_dom(r, x, fp);
        std::string r1 = to_hex(r, N);
        F::to_dom(x, a, fp);
        F::dpow(x, x, e, bit_length(e, 8), fp);
        F::from_dom(r, x, fp);
        return r1 == to_hex(r, N) ? r1 : (std::string("MISMATCH:") + r1 + "/" + to_hex(r, N));
    }
    return "?";
}

template <int L>
static std::string run_kind(const std::string& cmd, std::istringstream& in) {
    switch (fp.kind) {
        case KIND_GENERIC: return run<L, KIND_GENERIC>(cmd, in);
        case KIND_PM_ALIGNED: return run<L, KIND_PM_ALIGNED>(cmd, in);
        default: return run<L, KIND_PM_SHIFT>(cmd, in);
    }
}

int main() {
    std::string line;
    while (std::getline(std::cin, line)) {
        std::istringstream in(line);
        std::string cmd, t;
        in >> cmd;
        if (cmd.empty()) continue;
        if (cmd == "field") {
            u32 p32[8];
            in >> t; parse_hex(t, p32, 8);
            uint64_t p[4];
            for (int i = 0; i < 4; i++) p[i] = get64(p32, i);
            int n = 4;
            while (n > 1 && p[n - 1] == 0) n--;
            field_params_init(p, n, &fp);
            printf("%u %u %u\n", fp.kind, fp.L, fp.k);
            continue;
        }
        std::string out;
        switch (fp.L) {
            case 1: out = run_kind<1>(cmd, in); break;
            case 2: out = run_kind<2>(cmd, in); break;
            case 3: out = run_kind<3>(cmd, in); break;
            default: out = run_kind<4>(cmd, in); break;
        }
        printf("%s\n", out.c_str());
    }
    return 0;
}
