"""MPyC program for the protocol-local algebra (mpyc_b200.resident.ModValue, K6 kernels): a secure comparison and a secure
fixed-point product of n-element arrays -- `x < y` goes through Runtime.np_sgn (random bits, bit composition, an opened
masked value, the (l, n) bit-matrix algebra, a product tree), `u * v` through np_trunc -- timed after a warm-up pass, with
the engine's conversion counters for the timed part.

    python tests/run_installed.py tests/programs/resident_compare.py [-M3] [n]

Prints one JSON line.  Works without the engine too; the opened results are checked against NumPy either way.
"""
import json
import sys
import time

import numpy as np
from mpyc.runtime import mpc


def counters():
    try:
        from mpyc_b200 import resident
    except ImportError:
        return None
    return dict(resident.calls)


async def main():
    n = int(sys.argv[1]) if sys.argv[1:] else 4096
    secint = mpc.SecInt(37)                     # np_cnnmnist's type: 69-bit field, two limbs
    secfxp = mpc.SecFxp(32, 16)
    await mpc.start()
    rng = np.random.default_rng(11)
    a = rng.integers(-2**20, 2**20, size=n)
    b = rng.integers(-2**20, 2**20, size=n)
    u = rng.integers(-2000, 2000, size=n) / 16
    v = rng.integers(-2000, 2000, size=n) / 8
    x = mpc.input(secint.array(a), senders=0)
    y = mpc.input(secint.array(b), senders=0)
    U = mpc.input(secfxp.array(u), senders=0)
    V = mpc.input(secfxp.array(v), senders=0)
    await mpc.gather(x, y, U, V)
    await mpc.gather(x[:64] < y[:64], U[:64] * V[:64])        # warm-up: CUDA context, tables, PRSS keys, staging buffers
    before = counters()
    t0 = time.perf_counter()
    lt = x < y
    await mpc.gather(lt)
    t1 = time.perf_counter()
    prod = U * V
    await mpc.gather(prod)
    t2 = time.perf_counter()
    after = counters()
    got_lt = np.asarray(await mpc.output(lt))
    got_prod = np.asarray(await mpc.output(prod), dtype=float)
    ok = bool((got_lt == (a < b)).all()) and bool(np.abs(got_prod - u * v).max() < 2**-14)
    rec = {'n': n, 'parties': len(mpc.parties), 'ok': ok, 'seconds_compare': round(t1 - t0, 4), 'seconds_fxp_mul': round(t2 - t1, 4),
           'comparisons_per_s': round(n / (t1 - t0), 1), 'fxp_products_per_s': round(n / (t2 - t1), 1)}
    if before is not None:
        rec['engine_counters'] = {k: after[k] - before[k] for k in before}
    await mpc.shutdown()
    if mpc.pid == 0:
        print(json.dumps(rec))

mpc.run(main())
