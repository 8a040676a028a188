// Device-side twin of host_check.cpp: runs the same commands through the SAME templates in
// mpyc_b200/csrc/ff_arith.cuh, but inside a <<<1,1>>> kernel, so the PTX carry-chain primitives are
// exercised directly.  Built and run by tests/test_gpu_arith.py on the GPU box.  TEST INFRASTRUCTURE.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include "../../mpyc_b200/csrc/field_setup.h"

struct Cmd {
    int op;            // 0 mul 1 add 2 sub 3 neg 4 lazy 5 small 6 redsmall 7 pow 8 redsmall32
    int cnt;
    u32 a[64][8];      // operands (up to 64 terms)
    u32 b[64][8];
    u32 x[18];
    u64 e[8];
    int ebits;
};

template <int L, int K>
__global__ void run(FieldParams fp, const Cmd* cmd, u32* out) {
    typedef Fp<L, K> F;
    constexpr int N = 2 * L;
    u32 r[N];
    zero_n<N>(r);
    const Cmd& c = *cmd;
    if (c.op == 0) F::mul(r, c.a[0], c.b[0], fp);
    else if (c.op == 1) F::add(r, c.a[0], c.b[0], fp);
    else if (c.op == 2) F::sub(r, c.a[0], c.b[0], fp);
    else if (c.op == 3) F::neg(r, c.a[0], fp);
    else if (c.op == 4) {
        u32 acc[F::WACC];
        zero_n<F::WACC>(acc);
        for (int i = 0; i < c.cnt; i++) {
            u32 tb[N];
            F::to_dom(tb, c.b[i], fp);
            F::mac(acc, c.a[i], tb);
        }
        F::finish(r, acc, fp);
    } else if (c.op == 5) {
        u32 acc[F::WSM];
        copy_n<N>(acc, c.x);
        acc[N] = acc[N + 1] = 0;
        for (int i = 0; i < c.cnt; i++) F::mac_const(acc, c.a[i], (u64)c.b[i][0] | ((u64)c.b[i][1] << 32));
        F::reduce_small(r, acc, fp);
    } else if (c.op == 6) {
        F::reduce_small(r, c.x, fp);
    } else if (c.op == 8) {                 // x < 2^(k+31): one-limb-quotient Barrett / single-multiply fold
        FieldParams f2 = fp;
        f2.q32 = 1;
        F::reduce_small_q32(r, c.x, f2);
    } else if (c.op == 7) {
        u32 x[N], r2[N];
        F::to_dom(x, c.a[0], fp);
        F::dpow_uniform(r, x, c.e, c.ebits, fp);
        F::from_dom(r, r, fp);
        F::dpow(r2, x, c.e, c.ebits, fp);
        F::from_dom(r2, r2, fp);
        for (int i = 0; i < N; i++) if (r[i] != r2[i]) r[i] = 0xDEADBEEF;
    }
    for (int i = 0; i < N; i++) out[i] = r[i];
}

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
static Cmd* d_cmd;
static u32* d_out;

template <int L>
static void launch(const Cmd& c, u32* out) {
    cudaMemcpy(d_cmd, &c, sizeof c, cudaMemcpyHostToDevice);
    switch (fp.kind) {
        case KIND_GENERIC: run<L, KIND_GENERIC><<<1, 1>>>(fp, d_cmd, d_out); break;
        case KIND_PM_ALIGNED: run<L, KIND_PM_ALIGNED><<<1, 1>>>(fp, d_cmd, d_out); break;
        default: run<L, KIND_PM_SHIFT><<<1, 1>>>(fp, d_cmd, d_out); break;
    }
    cudaError_t e = cudaMemcpy(out, d_out, 8 * sizeof(u32), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { fprintf(stderr, "cuda error %s\n", cudaGetErrorString(e)); exit(2); }
}

int main() {
    cudaMalloc(&d_cmd, sizeof(Cmd));
    cudaMalloc(&d_out, 8 * sizeof(u32));
    std::string line;
    static Cmd c;
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
        memset(&c, 0, sizeof c);
        const int N = 2 * fp.L;
        if (cmd == "mul" || cmd == "add" || cmd == "sub") {
            c.op = cmd == "mul" ? 0 : (cmd == "add" ? 1 : 2);
            in >> t; parse_hex(t, c.a[0], N);
            in >> t; parse_hex(t, c.b[0], N);
        } else if (cmd == "neg") {
            c.op = 3;
            in >> t; parse_hex(t, c.a[0], N);
        } else if (cmd == "lazy") {
            c.op = 4;
            in >> c.cnt;
            if (c.cnt > 64) { printf("skip\n"); continue; }
            for (int i = 0; i < c.cnt; i++) {
                in >> t; parse_hex(t, c.a[i], N);
                in >> t; parse_hex(t, c.b[i], N);
            }
        } else if (cmd == "small") {
            c.op = 5;
            in >> c.cnt;
            in >> t; parse_hex(t, c.x, N);
            for (int i = 0; i < c.cnt; i++) {
                in >> t; parse_hex(t, c.a[i], N);
                in >> t; parse_hex(t, c.b[i], 2);
            }
        } else if (cmd == "redsmall") {
            c.op = 6;
            in >> t; parse_hex(t, c.x, N + 2);
        } else if (cmd == "redsmall32") {
            c.op = 8;
            in >> t; parse_hex(t, c.x, N + 2);
        } else if (cmd == "pow") {
            c.op = 7;
            u32 e32[16];
            in >> t; parse_hex(t, c.a[0], N);
            in >> t; parse_hex(t, e32, 16);
            for (int i = 0; i < 8; i++) c.e[i] = get64(e32, i);
            c.ebits = bit_length(c.e, 8);
        } else { printf("?\n"); continue; }
        u32 out[8];
        switch (fp.L) {
            case 1: launch<1>(c, out); break;
            case 2: launch<2>(c, out); break;
            case 3: launch<3>(c, out); break;
            default: launch<4>(c, out); break;
        }
        printf("%s\n", to_hex(out, N).c_str());
    }
    return 0;
}
