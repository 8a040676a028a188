"""An MPyC program written against MPyC's public API (not a copy of any demo): a batch of secure-array operations whose
OPENED results are deterministic, printed as one digest line per step.  tests/test_reference_runtime.py runs it through
tests/run_installed.py with and without the engine installed and requires identical output, and each step also checks
itself against plain NumPy integer arithmetic.  Exercises, per run: input (np_random_split), output (np_recombine),
secure multiplication / matmul (_reshare: split + exchange + recombine), random bits and random values (PRSS with
bounds 2 and 2^b), truncation of fixed-point products (PRSS zero-shares), conversion between secure types (PRSS with
a general bound), GF(2^8) inversion chains (the np_aes S-box core) and field division (batched inverse).

    python tests/run_installed.py tests/programs/secure_ops.py [-M3] [size]
"""
import hashlib
import sys

import numpy as np
from mpyc.runtime import mpc


def digest(label, a):
    a = np.asarray(a)
    flat = [int(v) if not isinstance(v, float) else round(v, 6) for v in a.reshape(-1).tolist()]
    h = hashlib.sha256(repr(flat).encode()).hexdigest()[:16]
    print(f'{label}: shape={tuple(a.shape)} digest={h} head={flat[:4]}')


async def main():
    n = int(sys.argv[1]) if sys.argv[1:] else 64
    await mpc.start()
    rng = np.random.default_rng(12345)

    # ---- secure integers (prime field of 32 + sec_param + 2 bits) ----------------------------------------
    secint = mpc.SecInt(32)
    a = rng.integers(-1000, 1000, size=n)
    b = rng.integers(-1000, 1000, size=n)
    x = mpc.input(secint.array(a), senders=0)
    y = mpc.input(secint.array(b), senders=0)
    z = x * y + x - 3 * y
    got = await mpc.output(z)
    assert (np.asarray(got) == a * b + a - 3 * b).all()
    digest('int mul/add', got)
    k = max(n // 8, 1)
    A = rng.integers(-50, 50, size=(4, k))
    B = rng.integers(-50, 50, size=(k, 8))
    X = mpc.input(secint.array(A), senders=0)
    Y = mpc.input(secint.array(B), senders=0)
    got = await mpc.output(X @ Y)
    assert (np.asarray(got) == A @ B).all()
    digest('int matmul', got)
    got = await mpc.output((x < y) * x + (x >= y) * y)          # comparisons: random bits, PRSS, many reshares
    assert (np.asarray(got) == np.minimum(a, b)).all()
    digest('int min via compare', got)
    bits = await mpc.output(mpc.np_random_bits(secint, n))       # values are random: only the range is checked
    assert set(np.asarray(bits).tolist()) <= {0, 1}
    print('random bits ok', len(bits))
    r = await mpc.output(mpc.np_randoms(secint, n, bound=1 << 10) if hasattr(mpc, 'np_randoms') else mpc.np_random_bits(secint, n))
    assert all(0 <= int(v) < (1 << 10) for v in np.asarray(r).tolist())
    print('random values ok', len(r))

    # ---- fixed point (products are truncated: PRSS zero shares + random bits) ---------------------------
    secfxp = mpc.SecFxp(32, 16)
    u = rng.integers(-200, 200, size=n) / 8
    v = rng.integers(-200, 200, size=n) / 16
    U = mpc.input(secfxp.array(u), senders=0)
    V = mpc.input(secfxp.array(v), senders=0)
    got = np.asarray(await mpc.output(U * V), dtype=float)
    assert np.abs(got - u * v).max() < 2**-14, np.abs(got - u * v).max()   # probabilistic rounding of the last bit
    print('fxp mul ok', got.shape)

    # ---- conversion between secure types (PRF bound (1 << (k+l)) // comb(m, t) + 1) ----------------------
    secint64 = mpc.SecInt(48)
    w = mpc.convert(list(mpc.input(secint.array(a[:8]), senders=0)), secint64)
    got = await mpc.output(w)
    assert [int(g) for g in got] == [int(t) for t in a[:8]]
    digest('convert', got)

    # ---- GF(2^8): inversion by x^254 (multiplications + resharing on byte shares), random bits -----------
    secfld = mpc.SecFld(2**8)
    f256 = secfld.field
    g = rng.integers(1, 256, size=min(n, 32))
    G = mpc.input(secfld.array(f256.array(g)), senders=0)
    inv = G**254
    got = await mpc.output(inv * G)
    assert all(int(e) == 1 for e in got)
    digest('gf256 inverse', await mpc.output(inv))
    gb = await mpc.output(mpc.np_random_bits(secfld, 16))
    assert all(int(e) in (0, 1) for e in gb)
    print('gf256 random bits ok', len(gb))
    gbits = await mpc.output(mpc.np_to_bits(G))
    assert [[int(e) for e in row] for row in gbits] == [[(int(t) >> i) & 1 for i in range(8)] for t in g]
    digest('gf256 to_bits', [[int(e) for e in row] for row in gbits])

    # ---- field division in a prime field (batched modular inverses on opened values) ---------------------
    secp = mpc.SecFld(2**61 - 1)
    c = rng.integers(1, 10**6, size=n)
    d = rng.integers(1, 10**6, size=n)
    C = mpc.input(secp.array(secp.field.array(c)), senders=0)
    D = mpc.input(secp.array(secp.field.array(d)), senders=0)
    q = await mpc.output(C / D)
    p = 2**61 - 1
    assert [int(e) for e in q] == [int(s) * pow(int(t), -1, p) % p for s, t in zip(c, d)]
    digest('field division', [int(e) for e in q])
    await mpc.shutdown()

mpc.run(main())
