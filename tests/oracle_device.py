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


# ---- mpyc_b200.resident backend on host limb arrays (oracle arithmetic) -------------------------------------------

class OracleBackend:
    """Stand-in for resident.CudaBackend: stores are host limb arrays, the arithmetic is the oracle's."""

    device = 0

    def _ints(self, ctx, a):
        return [int(v) for v in codec.limbs_to_ints(np.ascontiguousarray(a), ctx)]

    def to_store(self, ctx, limbs):
        return limbs

    def to_host(self, ctx, store):
        return store

    def _op(self, ctx, op, x, y):
        if not ctx.binary:                       # exact integer arithmetic on object arrays, then % p (what the oracle's ff_* do)
            X, Y, p = np.array(x, dtype=object), np.array(y, dtype=object), ctx.modulus
            r = (X + Y) if op == _cabi.OP_ADD else ((X - Y) if op == _cabi.OP_SUB else (X * Y))
            return codec.ints_to_limbs(r % p, ctx)
        F = orc.field_of(ctx.modulus, binary=ctx.binary)
        f = {_cabi.OP_ADD: F.add, _cabi.OP_SUB: F.sub, _cabi.OP_MUL: F.mul}[op]
        return codec.ints_to_limbs([F.red(f(a, b)) for a, b in zip(x, y)], ctx)

    def binop(self, ctx, op, a, b):
        return self._op(ctx, op, self._ints(ctx, a), self._ints(ctx, b))

    def binop_scalar(self, ctx, op, a, scalar):
        x = self._ints(ctx, a)
        s = int(scalar) if ctx.binary else int(scalar) % ctx.modulus
        return self._op(ctx, op, x, [s] * len(x))

    def neg(self, ctx, a):
        x = self._ints(ctx, a)
        return self._op(ctx, _cabi.OP_SUB, [0] * len(x), x)

    def matmul(self, ctx, a, b, r, k, c):
        F = orc.field_of(ctx.modulus, binary=ctx.binary)
        A, B = self._ints(ctx, a), self._ints(ctx, b)
        out = []
        for i in range(r):
            for j in range(c):
                acc = 0
                for l in range(k):
                    acc = F.add(acc, F.mul(A[i * k + l], B[l * c + j]))
                out.append(F.red(acc))
        return codec.ints_to_limbs(out, ctx)

    def split(self, ctx, sec, t, m, coeffs=None):
        return split_generate(ctx, sec, t, m) if coeffs is None else split_limbs(ctx, sec, coeffs, t, m)

    # ---- protocol-local algebra (the K6 kernels' stand-ins: oracle.local_*) ----
    def fma(self, ctx, a, b, c):
        x = self._ints(ctx, a)
        return codec.ints_to_limbs(orc.local_fma(ctx.modulus, x, x if b is None else self._ints(ctx, b), self._ints(ctx, c)), ctx)

    def axpb(self, ctx, a, s, t):
        p = ctx.modulus
        return codec.ints_to_limbs(orc.local_axpb(p, self._ints(ctx, a), int(s) % p, int(t) % p), ctx)

    def low_bits(self, ctx, a, nbits):
        return codec.ints_to_limbs(orc.local_low_bits(self._ints(ctx, a), nbits), ctx)

    def nonzero(self, ctx, a):
        mask = np.array(orc.local_nonzero(self._ints(ctx, a)), dtype=bool)
        return mask, int(mask.sum())

    def bits_compose(self, ctx, bits, n, f, descending):
        return codec.ints_to_limbs(orc.local_bits_compose(ctx.modulus, self._ints(ctx, bits), n, f, descending), ctx)

    def sqrt(self, ctx, a, INV):
        return codec.ints_to_limbs(orc.ff_sqrt(ctx.modulus, self._ints(ctx, a), INV=INV), ctx)

    def bits_decompose(self, ctx, c, l, descending):
        rows = orc.local_bits_decompose(self._ints(ctx, c), l, descending)
        return codec.ints_to_limbs([b for row in rows for b in row], ctx)

    def transpose(self, ctx, a, rows, cols):
        a = np.ascontiguousarray(a)
        return np.ascontiguousarray(a.reshape((rows, cols) + a.shape[1:]).swapaxes(0, 1)).reshape(a.shape)

    def cumsum_rows(self, ctx, a, rows, cols):
        x = np.array(self._ints(ctx, a), dtype=object).reshape(rows, cols)
        return codec.ints_to_limbs((np.cumsum(x, axis=0) % ctx.modulus).reshape(-1), ctx)

    def binop_rows(self, ctx, op, a, b, rows, cols, reflected):
        x, y = self._ints(ctx, a), self._ints(ctx, b)
        yy = y * rows
        return self._op(ctx, op, yy, x) if reflected else self._op(ctx, op, x, yy)

    def conv2d(self, ctx, X, W, B, k, r, m, n, v, s):
        if s % 2 == 0 or n < s:
            raise _cabi.UnsupportedFieldError('conv2d: even or oversized filters are not covered')
        return codec.ints_to_limbs(orc.local_conv2d(ctx.modulus, self._ints(ctx, X), self._ints(ctx, W), self._ints(ctx, B),
                                                    k, r, m, n, v, s), ctx)

    def slice(self, ctx, store, start, stop):
        return store[start:stop]

    def concat(self, ctx, stores):
        return np.concatenate([np.ascontiguousarray(st) for st in stores], axis=0)

    def recombine(self, ctx, xs, rows, pts):
        out = recombine_limbs(ctx, xs, rows, pts)
        return [out[r] for r in range(len(pts))]


def patch_resident(monkeypatch=None):
    from mpyc_b200 import resident
    if monkeypatch is not None:
        monkeypatch.setattr(resident, 'backend', OracleBackend())
    else:
        resident.backend = OracleBackend()


def patch_finfields(monkeypatch=None):
    """mpyc_b200.finfields' batched inverse / pow / sqrt / is_sqr / matmul answered by the oracle."""
    from mpyc_b200 import finfields as ff

    def shaped(fn):
        def call(cls, a, *args, **kwargs):
            a = np.asarray(a, dtype=object)
            p = cls.field.modulus
            vals = fn(p, [int(v) for v in a.reshape(-1)], *args, **kwargs)
            out = np.empty(len(vals), dtype=object)
            out[:] = vals
            return out.reshape(a.shape)
        return call

    def recip(p, vals):
        if any(v % p == 0 for v in vals):
            raise ZeroDivisionError('inverse of zero')
        return orc.ff_inv(p, vals)

    def power(cls, a, b, _fallback=None):
        if not isinstance(b, (int, np.integer)):
            return _fallback(a, b)
        return shaped(lambda p, vals: orc.ff_pow(p, vals, int(b)))(cls, a)

    def sqrt(cls, a, INV=False, _fallback=None):
        if cls.field.modulus & 3 != 3:
            return _fallback(a, INV=INV)
        return shaped(lambda p, vals: orc.ff_sqrt(p, vals, INV=INV))(cls, a)

    def is_sqr(cls, a):
        a = np.asarray(a, dtype=object)
        p = cls.field.modulus
        return np.array(orc.ff_is_sqr(p, [int(v) for v in a.reshape(-1)]), dtype=bool).reshape(a.shape)

    repl = {'reciprocal': shaped(recip), 'power': power, 'sqrt': sqrt, 'is_sqr': is_sqr}
    for name, fn in repl.items():
        if monkeypatch is not None:
            monkeypatch.setattr(ff, name, fn)
        else:
            setattr(ff, name, fn)


# ---- cross-check mode: the REAL device calls run, and every result is compared with the oracle ------------------------

def verify(log_path):
    """Wrap the adapter's device round trips so that each real result is checked against the oracle on the same limb
    arrays (split: the shares must lie on one polynomial of degree <= t through the secret; recombine and PRSS: equal
    to the oracle's output).  Mismatches are appended to log_path and raise."""
    real = {name: getattr(thresha, name) for name in ('_split_limbs', '_split_generate', '_recombine_limbs', '_prss_device')}

    def fail(msg):
        with open(log_path, 'a') as fh:
            fh.write(f'{os.getpid()} {msg}\n')
        raise AssertionError(msg)

    def ints(ctx, a):
        return [int(v) for v in codec.limbs_to_ints(np.ascontiguousarray(a), ctx)]

    def check_sharing(ctx, sec, shares, t, m, what):
        F = orc.field_of(ctx.modulus, binary=ctx.binary)
        s = ints(ctx, sec)
        rows = [ints(ctx, shares[i]) for i in range(m)]
        if t + 1 <= m:
            xs = list(range(1, t + 2))
            if orc.recombine(F, xs, [rows[x - 1] for x in xs], [0])[0] != s:
                fail(f'{what}: shares 1..t+1 do not recombine to the secrets (n={len(s)}, bits={ctx.bits})')
            for i in range(t + 1, m):            # every further share lies on the same polynomial
                if orc.recombine(F, xs, [rows[x - 1] for x in xs], [i + 1])[0] != rows[i]:
                    fail(f'{what}: share {i + 1} is not on the polynomial through shares 1..t+1 (n={len(s)}, bits={ctx.bits})')

    def split_limbs_v(ctx, sec, C, t, m):
        out = real['_split_limbs'](ctx, sec, C, t, m)
        want = split_limbs(ctx, sec, C, t, m)
        if not np.array_equal(out, want):
            fail(f'_split_limbs differs from the oracle (n={sec.shape[0]}, bits={ctx.bits}, t={t}, m={m})')
        return out

    def split_generate_v(ctx, sec, t, m):
        out = real['_split_generate'](ctx, sec, t, m)
        check_sharing(ctx, sec, out, t, m, '_split_generate')
        return out

    def recombine_limbs_v(ctx, xs, rows, pts):
        out = real['_recombine_limbs'](ctx, xs, rows, pts)
        want = recombine_limbs(ctx, xs, rows, pts)
        if not np.array_equal(out, want):
            fail(f'_recombine_limbs differs from the oracle (n={rows[0].shape[0]}, bits={ctx.bits}, xs={list(xs)}, pts={list(pts)})')
        return out

    def prss_device_v(*args, **kwargs):
        out = real['_prss_device'](*args, **kwargs)
        want = prss_device(*args, **kwargs)
        if not np.array_equal(out, want):
            fail(f'_prss_device differs from the oracle (n={args[-1] if not kwargs else "?"}, bits={args[0].bits})')
        return out

    thresha._split_limbs = split_limbs_v
    thresha._split_generate = split_generate_v
    thresha._recombine_limbs = recombine_limbs_v
    thresha._prss_device = prss_device_v
