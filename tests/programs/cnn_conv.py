"""An MPyC program with a secure convolution layer written the way the reference's CNN demo writes its own
(gather the shares, correlate locally on raw values, `field.array`, `_reshare`), with the local step bound to the engine
when it is installed: one `mpyc_b200.resident.conv2d` call instead of the NumPy object loops.
tests/test_reference_runtime.py runs it with and without the engine and requires identical opened outputs.

    python tests/run_installed.py tests/programs/cnn_conv.py [-M3] [batch]
"""
import hashlib
import os
import sys

import numpy as np
from mpyc.runtime import mpc

ENGINE = 'off' not in os.environ.get('MPYC_B200_HARNESS', 'off').split(',') and 'resident' in os.environ.get('MPYC_B200_HARNESS', '')


def local_correlation(x, W, b, field):
    """Y[i, j] = b[j] + sum_l correlate2d(x[i, l], W[j, l], 'same') on raw values, reduced into a field array."""
    k, r, m, n = x.shape
    v, _, s, _ = W.shape
    if ENGINE:
        from mpyc_b200 import resident
        return resident.conv2d(x, W, b)
    xv, Wv, bv = x.value, W.value, b.value
    s2 = (s - 1) // 2
    Y = np.zeros((k, v, m, n), dtype=object)
    for i in range(k):
        for j in range(v):
            for l in range(r):
                for I in range(m):
                    for i_d in range(max(0, I - s2), min(I - s2 + s, m)):
                        Y[i, j, I] += np.correlate(xv[i, l, i_d], Wv[j, l, i_d - I + s2], mode='same')
    Y += bv[:, np.newaxis, np.newaxis]
    return field.array(Y)


@mpc.coroutine
async def conv_layer(x, W, b):
    stype = type(x)
    k, r, m, n = x.shape
    v = W.shape[0]
    await mpc.returnType((stype, (k, v, m, n)))
    x, W, b = await mpc.gather(x, W, b)
    Y = local_correlation(x, W, b, stype.sectype.field)
    Y = mpc._reshare(Y)
    return Y


async def main():
    batch = int(sys.argv[1]) if sys.argv[1:] else 2
    await mpc.start()
    rng = np.random.default_rng(2026)
    secint = mpc.SecInt(37)
    for (r, m, n, v, s) in ((1, 12, 12, 4, 5), (4, 7, 9, 3, 3)):
        x = rng.integers(-60, 60, size=(batch, r, m, n))
        W = rng.integers(-30, 30, size=(v, r, s, s))
        b = rng.integers(-1000, 1000, size=v)
        X = mpc.input(secint.array(x), senders=0)
        Ws = mpc.input(secint.array(W), senders=0)
        Bs = mpc.input(secint.array(b), senders=0)
        Y = conv_layer(X, Ws, Bs)
        Y = Y * Y                                   # a secure product after the layer: the reshared shares are real shares
        got = np.asarray(await mpc.output(Y))
        want = np.zeros((batch, v, m, n), dtype=np.int64)
        s2 = (s - 1) // 2
        xp = np.pad(x, ((0, 0), (0, 0), (s2, s2), (s2, s2)))
        for dy in range(s):
            for dx in range(s):
                want += np.einsum('krmn,vr->kvmn', xp[:, :, dy:dy + m, dx:dx + n], W[:, :, dy, dx])
        want += b[None, :, None, None]
        assert (got == want * want).all()
        flat = [int(t) for t in got.reshape(-1)]
        print(f'conv {r}x{m}x{n} -> {v}, {s}x{s}: digest={hashlib.sha256(repr(flat).encode()).hexdigest()[:16]} head={flat[:3]}')
    await mpc.shutdown()

mpc.run(main())
