"""Command streams for the arithmetic checkers (tests/native/host_check.cpp on the CPU,
tests/native/dev_check.cu on the GPU) with the expected answers computed on Python integers."""
import random

from golden_util import load

PRIMES = sorted({int(c['p'], 16) for c in load('finfields.json')['cases']} | {2**255 - 19, 2**192 - 237, 2**89 - 1, 65537, 3, 2**64 - 59})


def expected_kind(p):
    k = p.bit_length()
    c = (1 << k) - p
    if c < (1 << 16) and k >= 56:
        return 1 if k % 64 == 0 else 2
    return 0



def commands(p, lazy_sizes=(1, 2, 3, 7, 17, 64, 1000)):
    """(lines, want): stdin lines for a checker and the expected output values."""
    rnd = random.Random(p)
    L = (p.bit_length() + 63) // 64
    k = p.bit_length()
    edge = [0, 1, 2, p - 1, p - 2, p >> 1, (p >> 1) + 1, ((1 << (64 * L)) - 1) % p]
    vals = edge + [rnd.randrange(p) for _ in range(24)]
    lines, want = [f'field {p:x}'], []
    for a in vals:
        for b in (vals[0], vals[3], rnd.choice(vals), rnd.randrange(p)):
            lines += [f'mul {a:x} {b:x}', f'add {a:x} {b:x}', f'sub {a:x} {b:x}']
            want += [a * b % p, (a + b) % p, (a - b) % p]
        lines.append(f'neg {a:x}')
        want.append(-a % p)
    # lazy sums of full products: worst case all p-1, and long sums
    for K in lazy_sizes:
        for mode in ('max', 'rand'):
            terms = [(p - 1, p - 1) if mode == 'max' else (rnd.randrange(p), rnd.randrange(p)) for _ in range(K)]
            lines.append(f'lazy {K} ' + ' '.join(f'{a:x} {b:x}' for a, b in terms))
            want.append(sum(a * b for a, b in terms) % p)
    for K in (0, 1, 2, 8):
        for mode in ('max', 'rand', 'rand32'):
            s = p - 1 if mode == 'max' else rnd.randrange(p)
            vmax = 1 << (32 if mode == 'rand32' else 59)
            terms = [(p - 1, (1 << 59) - 1) if mode == 'max' else (rnd.randrange(p), rnd.randrange(vmax)) for _ in range(K)]
            lines.append(f'small {K} {s:x} ' + ' '.join(f'{a:x} {v:x}' for a, v in terms))
            want.append((s + sum(a * v for a, v in terms)) % p)
    for x in [0, 1, p, p - 1, (1 << (k + 64)) - 1, (p << 64) - 1] + [rnd.randrange(1 << (k + 64)) for _ in range(16)]:
        lines.append(f'redsmall {x:x}')
        want.append(x % p)
    # share-generation sums (1 + m + ... + m^t) p <= 2^31 p: the 32-bit-quotient Barrett step of generic fields and
    # the single-multiply fold of 2^64 - c (Fp::reduce_small_q32 with q32 set)
    for x in [0, 1, p, p - 1, 2 * p - 1, 2 * p, 3 * p + 5, (1 << (k + 31)) - 1, (p << 31) - 1, ((1 << 31) - 1) * p, 31 * (p - 1), 400 * (p - 1)] \
            + [rnd.randrange(1 << (k + 31)) for _ in range(24)] + [rnd.randrange(1 << 9) * p + rnd.randrange(p) for _ in range(24)]:
        if x < 1 << (k + 31):
            lines.append(f'redsmall32 {x:x}')
            want.append(x % p)
    for a in vals[:12]:
        for e in (0, 1, 2, 5, p - 2, (p - 1) // 2, (3 * p - 5) // 4 if p > 3 else 1, rnd.randrange(1 << 300)):
            lines.append(f'pow {a:x} {e:x}')
            want.append(pow(a, e, p))
    return lines, want, (L, k)
