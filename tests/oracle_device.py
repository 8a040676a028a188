"""TEST INFRASTRUCTURE: the engine's device round trips answered by the CPU oracle.

The build container has no GPU.  To exercise the drop-in adapter under the reference's own runtime there
(tests/test_reference_runtime.py, tests/run_installed.py) the four functions of mpyc_b200.thresha that cross the C ABI
into CUDA -- _split_limbs, _split_generate, _recombine_limbs, _prss_device -- are replaced by the functions below, which
do the same arithmetic with oracle/shamir_oracle.py on the very limb arrays the adapter built.  Everything above them
(argument handling, codecs, draw orders, bound classification, wrapping, install()) is the product code under test.
Never imported by the package; on a box with a GPU the tests leave the real calls in place.
"""
import ctypes
import os

import numpy as np

from oracle import shamir_oracle as orc
from mpyc_b200 import _cabi, codec, thresha


def _empty(ctx, rows):
    return np.zeros((rows, 0) if ctx.binary else (rows, 0, ctx.nlimbs), dtype=np.uint8 if ctx.binary else np.uint64)


def split_limbs(ctx, sec, C, t, m):
    F = orc.field_of(ctx.modulus, binary=ctx.binary)
    s = [int(v) for v in codec.limbs_to_ints(sec, ctx)]
    rows = [[int(v) for v in codec.limbs_to_ints(C[j], ctx)] for j in range(t)]
    shares = orc.split_np_order(F, s, rows, m)
    return np.stack([codec.ints_to_limbs(r, ctx) for r in shares]) if s else _empty(ctx, m)


def split_generate(ctx, sec, t, m):
    """Generate mode: coefficients from the OS CSPRNG (the kernel's ChaCha20 stream is not reproduced; shares of the
    generate path are never claimed bit-exact, only that they recombine)."""
    n = sec.shape[0]
    order = ctx.order
    nb = (order.bit_length() + 64 + 7) // 8
    raw = os.urandom(t * n * nb)
    vals = [int.from_bytes(raw[i:i + nb], 'little') % order for i in range(0, t * n * nb, nb)]
    C = codec.ints_to_limbs(vals, ctx, reduce=False).reshape(t, n, ctx.nlimbs)
    return split_limbs(ctx, sec, C, t, m)


def recombine_limbs(ctx, xs, rows, pts):
    thresha._check_rows(ctx, rows)
    F = orc.field_of(ctx.modulus, binary=ctx.binary)
    vals = orc.recombine(F, list(xs), [[int(v) for v in codec.limbs_to_ints(r, ctx)] for r in rows], list(pts))
    n = rows[0].shape[0]
    return np.stack([codec.ints_to_limbs(v, ctx) for v in vals]) if n else _empty(ctx, len(pts))


def prss_device(ctx, keys, uci, d, width, bound_bits, coef, weights, n, general=None):
    """out[h] = sum_S coef_S * sum_j (chunk_{S,h,j} mod bound) * w_j  in the field (include/mpyc_b200.h);
    the XOF is the library's own host SHAKE128 (pinned against hashlib in tests/test_shake128.py)."""
    F = orc.field_of(ctx.modulus, binary=ctx.binary)
    L = max(ctx.nlimbs, 1)
    cs = [_cabi.limbs_to_int(coef[i * L:(i + 1) * L]) for i in range(len(keys))]
    ws = [_cabi.limbs_to_int(weights[j * L:(j + 1) * L]) for j in range(d)]
    full = 256 if ctx.binary else ctx.modulus
    bound = general if general is not None else (1 << bound_bits if bound_bits else full)
    acc = [0] * n
    for key, c in zip(keys, cs):
        raw = ctypes.create_string_buffer(n * d * width)
        msg = bytes(key) + bytes(uci)
        _cabi.check(_cabi.lib.mpyc_b200_shake128(msg, len(msg), raw, n * d * width))
        raw = raw.raw
        for h in range(n):
            y = 0
            for j in range(d):
                v = int.from_bytes(raw[(h * d + j) * width:(h * d + j + 1) * width], 'little') % bound
                y = F.add(y, F.mul(v, ws[j]))
            acc[h] = F.add(acc[h], F.mul(c, y))
    return codec.ints_to_limbs([F.red(a) for a in acc], ctx)


def patch(monkeypatch=None):
    """Route the adapter's device calls to the oracle (monkeypatch: pytest fixture, or None for a permanent patch)."""
    repl = {'_split_limbs': split_limbs, '_split_generate': split_generate, '_recombine_limbs': recombine_limbs,
            '_prss_device': prss_device}
    for name, fn in repl.items():
        if monkeypatch is not None:
            monkeypatch.setattr(thresha, name, fn)
        else:
            setattr(thresha, name, fn)
