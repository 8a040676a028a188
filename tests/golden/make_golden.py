"""Generate golden fixtures by running the UNMODIFIED reference (lschoe/mpyc) in the build container.

    python tests/golden/make_golden.py          # needs /root/reference (read-only); writes tests/golden/*.json

The reference draws Shamir coefficients from secrets.randbelow (thresha.py:37,58-60);
here that is replaced, for the duration of each call, by a deterministic stream so the
share values become reproducible.  Everything else is the stock reference code path.
Fixtures store inputs AND outputs, so the GPU box (which has no reference) can replay them.
All integers are stored as hex strings.
"""

import json
import os
import sys

REF = os.environ.get('MPYC_REFERENCE', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
_argv, sys.argv = sys.argv, [sys.argv[0], '--no-log']   # mpyc parses sys.argv at import
from mpyc import finfields, thresha, gfpx, gmpy   # noqa: E402
import numpy as np   # noqa: E402
sys.argv = _argv

from oracle import shamir_oracle as orc   # noqa: E402  (only for the seeded input recipe)

SEED = 20260923

P61 = 2**61 - 1
P64 = 2**64 - 189
P69 = 2**69 - 93
P127 = 2**127 - 1
P128 = 2**128 - 173
P256 = 2**256 - 189
P64G = 9409569905028393239          # bnnmnist's prime (docs/demos.rst:1139): generic 64-bit
KEY = int('0x00112233445566778899aabbccddeeff', 16).to_bytes(16, 'little')   # tests/test_thresha.py:43


def hx(v):
    if isinstance(v, (list, tuple)):
        return [hx(x) for x in v]
    if isinstance(v, np.ndarray):
        return hx(v.tolist())
    if isinstance(v, (bool, np.bool_)):
        return bool(v)
    return hex(int(v))


class injected:
    """Context manager: secrets.randbelow -> next value of a fixed stream (restored on exit)."""

    def __init__(self, stream):
        self.it = iter(stream)

    def __enter__(self):
        self.orig = thresha.secrets.randbelow
        thresha.secrets.randbelow = lambda order: next(self.it)
        return self

    def __exit__(self, *exc):
        thresha.secrets.randbelow = self.orig


def find_generic_primes():
    """Generic (non pseudo-Mersenne) primes of 96/128/192/255/256 bits via the reference's find_prime_root."""
    out = {}
    for l, n in ((96, 3), (128, 7), (192, 5), (250, 11), (256, 3)):
        p, _, _ = finfields.find_prime_root(l, n=n)
        out[str(l)] = p
    return out


def split_cases(primes):
    cases = []
    mts = [(1, 0), (3, 1), (4, 1), (5, 2), (7, 3), (9, 4), (17, 8), (13, 2)]
    for p in primes:
        F = finfields.GF(p)
        for (m, t) in mts:
            base = orc.edge_block(p) + orc.synth_elements(p, 5, SEED, stream=1)
            n = len(base)
            stream = orc.synth_elements(p, t * n, SEED + m * 100 + t, stream=2)
            # np order
            with injected(stream):
                sh_np = thresha.np_random_split(F, F.array(np.array(base, dtype=object)), t, m)
            # list order
            with injected(stream):
                sh_li = thresha.random_split(F, list(base), t, m)
            # recombine from parties 1..t+1 and from a scattered subset, at x_r = 0 and at several points
            xs_a = list(range(1, t + 2))
            xs_b = sorted(set([m - j for j in range(t + 1)]))
            rec = []
            for xs in (xs_a, xs_b, list(range(1, min(m, 2 * t + 1) + 1))):
                pts = [(x, sh_np[x - 1]) for x in xs]
                y0 = thresha.np_recombine(F, pts)
                x_rs = [0, m + 1, 1]
                yw = thresha.np_recombine(F, pts, x_rs)
                rec.append({'xs': xs, 'x_rs': x_rs, 'y0': hx(y0.value), 'yw': hx(yw.value),
                            'lambda0': hx(thresha._recombination_vector(F, tuple(xs), 0))})
            cases.append({'p': hex(p), 'm': m, 't': t, 'secrets': hx(base), 'stream': hx(stream),
                          'shares_np': hx(sh_np), 'shares_list': hx(sh_li), 'recombine': rec})
    return cases


def prf_cases(primes):
    out = []
    for bound in [1, 2, 100, 256, 2**32, 2**61, 2**64] + list(primes):
        for s, n in ((b'test uci', 3), (b'', 1), (b'\x07\x00\x00\x00\x00\x00\x00\x00', 17)):
            F = thresha.PRF(KEY, bound)
            out.append({'bound': hex(bound), 's': s.hex(), 'n': n, 'l': F.byte_length,
                        'values': hx(F(s, n))})
    return out


def prss_cases(primes):
    from itertools import combinations
    out = []
    for p in primes:
        F = finfields.GF(p)
        for (m, t) in ((1, 0), (3, 1), (5, 2), (7, 3)):
            n = 6
            uci = (12345).to_bytes(8, 'little')
            # key layout as in Runtime.__init__ (runtime.py:110-121): one key per subset of size m-t containing i
            keys = {}
            for S in combinations(range(m), m - t):
                keys[S] = bytes((sum(S) * 17 + j * 3 + len(S)) & 0xFF for j in range(16))
            per_party = []
            for i in range(m):
                prfs = {S: thresha.PRF(k, p) for S, k in keys.items() if i in S}
                a_np = thresha.np_pseudorandom_share(F, m, i, prfs, uci, n)
                a_li = thresha.pseudorandom_share(F, m, i, prfs, uci, n)
                z_np = thresha.np_pseudorandom_share_0(F, m, i, prfs, uci, n) if t else None
                z_li = thresha.pseudorandom_share_zero(F, m, i, prfs, uci, n)
                fs = {','.join(map(str, S)): hex(int(thresha._f_S_i(F, m, i, S))) for S in prfs}
                per_party.append({'i': i, 'f_S_i': fs,
                                  'share_np': hx(a_np.value), 'share_list': hx([x.value for x in a_li]),
                                  'zero_np': hx(z_np.value) if z_np is not None else None,
                                  'zero_list': hx([x.value for x in z_li])})
            out.append({'p': hex(p), 'm': m, 't': t, 'n': n, 'uci': uci.hex(),
                        'keys': {','.join(map(str, S)): k.hex() for S, k in keys.items()},
                        'parties': per_party})
    return out


def prss_bound_cases():
    """PRSS with PRF bounds other than the field order (thresha.py:257-261 reduces chunk % bound for ANY bound):
    prfs(2) of runtime.random_bits / np_random_bits (runtime.py:4138,4218), also on GF(2^8); the power-of-two bounds of
    _randoms / _np_randoms (runtime.py:4050-4056,4090-4098); runtime._convert's (1 << (k+l)) // comb(m,t) + 1 and the
    source field's order applied to another (smaller or larger) target field (runtime.py:735-739,758-760)."""
    from itertools import combinations
    from math import comb
    f283 = gfpx.GFpX(2)(283)
    fields = [('gf', 283, finfields.GF(f283)), ('p', P61, finfields.GF(P61)), ('p', P64G, finfields.GF(P64G)),
              ('p', P128, finfields.GF(P128)), ('p', 101, finfields.GF(101))]
    out = []
    for kind, mod, F in fields:
        for (m, t) in ((1, 0), (3, 1), (5, 2)):
            bounds = [2, 16, 256, 1 << 10, 1000, (1 << 62) // comb(m, t) + 1, P61, P69, 1 << 70, P256, (1 << 200) + 12345, 1 << 256]
            if kind == 'p':
                bounds.append(mod - 1)
            for bound in bounds:
                n = 5
                uci = (777 + m).to_bytes(8, 'little')
                keys = {}
                for S in combinations(range(m), m - t):
                    keys[S] = bytes((sum(S) * 29 + j * 5 + len(S) + 1) & 0xFF for j in range(16))
                per_party = []
                for i in range(m):
                    prfs = {S: thresha.PRF(k, bound) for S, k in keys.items() if i in S}
                    a_np = thresha.np_pseudorandom_share(F, m, i, prfs, uci, n)
                    a_li = thresha.pseudorandom_share(F, m, i, prfs, uci, n)
                    z_np = thresha.np_pseudorandom_share_0(F, m, i, prfs, uci, n) if t else None
                    z_li = thresha.pseudorandom_share_zero(F, m, i, prfs, uci, n)
                    per_party.append({'i': i,
                                      'share_np': hx([int(v) for v in a_np.value]),
                                      'share_list': hx([int(x.value) for x in a_li]),
                                      'zero_np': hx([int(v) for v in z_np.value]) if z_np is not None else None,
                                      'zero_list': hx([int(x.value) for x in z_li])})
                out.append({'field': kind, 'modulus': hex(mod), 'bound': hex(bound), 'm': m, 't': t, 'n': n,
                            'uci': uci.hex(), 'keys': {','.join(map(str, S)): k.hex() for S, k in keys.items()},
                            'parties': per_party})
    return out


def ff_cases(primes):
    out = []
    for p in primes:
        F = finfields.GF(p)
        a = orc.edge_block(p) + orc.synth_elements(p, 24, SEED, stream=3)
        b = list(reversed(orc.edge_block(p))) + orc.synth_elements(p, 24, SEED, stream=4)
        A, B = F.array(np.array(a, dtype=object)), F.array(np.array(b, dtype=object))
        nz = [x if x else 1 for x in b]
        NZ = F.array(np.array(nz, dtype=object))
        case = {'p': hex(p), 'a': hx(a), 'b': hx(b), 'nz': hx(nz),
                'add': hx((A + B).value), 'sub': hx((A - B).value), 'mul': hx((A * B).value),
                'neg': hx((-A).value), 'inv_nz': hx(NZ.reciprocal().value),
                'div': hx((A / NZ).value), 'lshift7': hx((A << 7).value), 'rshift7': hx((A >> 7).value),
                'pow5': hx((A ** 5).value), 'powm3_nz': hx((NZ ** -3).value),
                'pow_big': hx((A ** ((p - 1) // 2 + 3)).value),
                'is_sqr': hx(A.is_sqr()), 'signed': [str(int(x)) for x in A.signed_()],
                'mul_scalar': hx((A * 12345678901234567890123).value),
                'add_scalar': hx((A + (p - 5)).value)}
        if p & 3 == 3:
            sq = (A * A)
            case['sqrt_of_sq'] = hx(sq.sqrt().value)
            nzs = (NZ * NZ)
            case['invsqrt_of_nzsq'] = hx(nzs.sqrt(INV=True).value)
            case['sqrt_a'] = hx(A.sqrt().value)     # defined as a^((p+1)/4) even for non-residues
        M1 = F.array(np.array(a[:12], dtype=object).reshape(3, 4))
        M2 = F.array(np.array(b[:20], dtype=object).reshape(4, 5))
        case['matmul_3x4_4x5'] = hx((M1 @ M2).value)
        out.append(case)
    return out


def gf256_cases():
    f = gfpx.GFpX(2)(283)
    F = finfields.GF(f)
    poly = type(f)
    tab = bytearray(65536)
    for a in range(256):
        for b in range(256):
            tab[a * 256 + b] = int((F(a) * F(b)).value)
    inv = [0] + [int((1 / F(a)).value) for a in range(1, 256)]
    cases = []
    for (m, t) in ((1, 0), (3, 1), (5, 2), (7, 3), (17, 8)):
        base = [0, 1, 2, 0x53, 0xCA, 0xFF, 0x80, 0x1B] + [orc.splitmix64(SEED + i) & 0xFF for i in range(8)]
        n = len(base)
        stream = [orc.splitmix64(SEED + 1000 * m + i) & 0xFF for i in range(t * n)]
        with injected(stream):
            sh_np = thresha.np_random_split(F, F.array(np.array([poly(x) for x in base], dtype=object)), t, m)
        with injected(stream):
            sh_li = thresha.random_split(F, [F(x) for x in base], t, m)
        xs = list(range(1, min(m, 2 * t + 1) + 1))
        pts = [(x, sh_np[x - 1]) for x in xs]
        y0 = thresha.np_recombine(F, pts)
        cases.append({'m': m, 't': t, 'secrets': hx(base), 'stream': hx(stream),
                      'shares_np': hx([[int(v) for v in row] for row in sh_np]),
                      'shares_list': hx([[int(v) for v in row] for row in sh_li]),
                      'xs': xs, 'y0': hx([int(v) for v in y0.value]),
                      'lambda0': hx([int(v) for v in thresha._recombination_vector(F, tuple(xs), 0)])})
    return {'modulus': 283, 'mul_table_hex': bytes(tab).hex(), 'inv': inv, 'split': cases}


def main():
    gen = find_generic_primes()
    defaults = {str(l): finfields.find_prime_root(l)[0] for l in (61, 64, 69, 128, 256)}
    assert defaults == {'61': P61, '64': P64, '69': P69, '128': P128, '256': P256}
    primes_all = [19, 101, P61, P64, P64G, P69, gen['96'], P127, P128, gen['128'], gen['192'],
                  gen['250'], P256, gen['256']]
    meta = {'reference': 'lschoe/mpyc v0.11.2 (unmodified, gmpy2 stubs=%s)' % (not hasattr(gmpy, 'mpz') or gmpy.__name__),
            'seed': SEED, 'default_primes': {k: hex(v) for k, v in defaults.items()},
            'generic_primes': {k: hex(v) for k, v in gen.items()}}
    files = {
        'split_recombine.json': {'meta': meta, 'cases': split_cases(primes_all)},
        'prf.json': {'meta': meta, 'key': KEY.hex(), 'cases': prf_cases([P61, P64, P69, P128, P256])},
        'prss.json': {'meta': meta, 'cases': prss_cases([P61, P69, P128, P256, P64G])},
        'finfields.json': {'meta': meta, 'cases': ff_cases(primes_all)},
        'gf256.json': {'meta': meta, **gf256_cases()},
        'prss_bounds.json': {'meta': meta, 'cases': prss_bound_cases()},
    }
    for name, obj in files.items():
        with open(os.path.join(HERE, name), 'w') as fh:
            json.dump(obj, fh, separators=(',', ':'))
        print(name, os.path.getsize(os.path.join(HERE, name)), 'bytes')


if __name__ == '__main__':
    main()
