"""MPyC program for the limb-resident path (mpyc_b200.resident): two chained secure multiplications of n-element secure
integer arrays over a 128-bit prime -- input, c = a * b, _reshare, d = c * a, _reshare, output -- and a report of how many
int <-> limb conversions happened BETWEEN input and output (must be zero in resident mode), plus timing.

    python tests/run_installed.py tests/programs/resident_chain.py [-M3] [n]

Prints one JSON line: {"n":..., "parties":..., "ok": true, "conversions_between": {...}, "seconds_chain": ..., ...}.
Works without the engine too (conversions are then not counted): the opened result is checked against NumPy either way.
"""
import json
import sys
import time

import numpy as np
from mpyc.runtime import mpc

P128 = 2**128 - 173


def counters():
    try:
        from mpyc_b200 import codec, resident
    except ImportError:
        return None
    c = dict(resident.calls)
    c['pycodec'] = getattr(codec, '_calls', [0])[0]
    return c


def count_pycodec():
    """Wrap the C codec's entry points with a call counter (real conversions, whoever asks for them)."""
    try:
        from mpyc_b200 import codec
    except ImportError:
        return
    inner = codec._pycodec
    codec._calls = [0]

    class Counting:
        def pack(self, *a):
            codec._calls[0] += 1
            return inner.pack(*a)

        def unpack_into(self, *a):
            codec._calls[0] += 1
            return inner.unpack_into(*a)
    codec._pycodec = Counting()


async def main():
    n = int(sys.argv[1]) if sys.argv[1:] else 4096
    count_pycodec()
    secint = mpc.SecInt(32, p=P128)
    await mpc.start()
    rng = np.random.default_rng(7)
    a = rng.integers(-2**15, 2**15, size=n)
    b = rng.integers(-2**15, 2**15, size=n)
    t0 = time.perf_counter()
    x = mpc.input(secint.array(a), senders=0)
    y = mpc.input(secint.array(b), senders=0)
    await mpc.gather(x, y)                      # shares of the inputs have arrived
    await mpc.gather((x * y) * x)               # warm-up pass of the same chain (CUDA context, tables, staging buffers)
    t1 = time.perf_counter()
    before = counters()
    z = x * y                                   # local product + _reshare
    w = z * x                                   # consumes the reshared product: second product + _reshare
    await mpc.gather(w)
    t2 = time.perf_counter()
    after = counters()
    got = await mpc.output(w)
    t3 = time.perf_counter()
    ok = bool((np.asarray(got) == a.astype(object) * b * a).all())
    between = {k: after[k] - before[k] for k in after} if after else None
    print(json.dumps({'n': n, 'parties': len(mpc.parties), 'pid': mpc.pid, 'ok': ok, 'conversions_between': between,
                      'seconds_input': round(t1 - t0, 4), 'seconds_chain': round(t2 - t1, 4),
                      'seconds_output': round(t3 - t2, 4),
                      'multiplications_per_s': round(2 * n / (t2 - t1), 1)}))
    await mpc.shutdown()
    assert ok

mpc.run(main())
