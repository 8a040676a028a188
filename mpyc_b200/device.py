"""Device-resident field arrays: limb tensors that stay in HBM between operations.

torch is used for what it is good at here -- device memory, streams, (later) torch.distributed --
while every computation goes through the C ABI (mpyc_b200._cabi) on the tensor's data pointer and
torch's current CUDA stream.  A DeviceArray mirrors the operator surface of the reference's
PrimeFieldArray / BinaryFieldArray (mpyc/finfields.py:1056-1281,1371-1470) for 1-D arrays.
"""
import ctypes

import numpy as np
import torch

from mpyc_b200 import _cabi, codec
from mpyc_b200._cabi import lib, check
from mpyc_b200.field import FieldContext


def _stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError('mpyc_b200: no CUDA device available (there is no CPU fallback)')


def _alloc(ctx, rows, n, device):
    """Uninitialised limb storage: int64 (rows, n, L) or uint8 (rows, n); row stride padded to 32 bytes
    so that every row start allows 256-bit vector accesses."""
    _require_cuda()
    if ctx.binary:
        stride = (n + 31) // 32 * 32
        return torch.empty((rows, stride), dtype=torch.uint8, device=device)[:, :n]
    L = ctx.nlimbs
    q = {1: 4, 2: 2, 3: 4, 4: 1}[L]
    stride = (n + q - 1) // q * q
    return torch.empty((rows, stride, L), dtype=torch.int64, device=device)[:, :n]


class DeviceArray:
    """n elements of one field in device memory (1-D)."""

    __slots__ = ('ctx', 't')

    def __init__(self, ctx: FieldContext, tensor):
        self.ctx = ctx
        self.t = tensor   # int64 (n, L) contiguous, or uint8 (n,)

    # ---- construction / extraction ---------------------------------------------------------------
    @classmethod
    def empty(cls, ctx, n, device='cuda'):
        return cls(ctx, _alloc(ctx, 1, n, device)[0])

    @classmethod
    def from_limbs(cls, ctx, limbs, device='cuda'):
        _require_cuda()
        a = np.ascontiguousarray(limbs)
        src = torch.from_numpy(a if ctx.binary else a.view(np.int64))
        out = cls.empty(ctx, a.shape[0], device)
        out.t.copy_(src)
        return out

    @classmethod
    def from_ints(cls, ctx, values, device='cuda', reduce=True):
        return cls.from_limbs(ctx, codec.ints_to_limbs(values, ctx, reduce=reduce), device)

    @classmethod
    def random(cls, ctx, n, seed, stream_id=0, device='cuda'):
        """Deterministic synthetic residues (same recipe as oracle.synth_elements)."""
        out = cls.empty(ctx, n, device)
        check(lib.mpyc_b200_fill_random(ctx.handle, out.ptr, n, seed & (2**64 - 1), stream_id, _stream_ptr()))
        return out

    def to_limbs(self):
        a = self.t.contiguous().cpu().numpy()
        return a if self.ctx.binary else a.view(np.uint64)

    def to_ints(self):
        return codec.limbs_to_ints(self.to_limbs(), self.ctx)

    @property
    def ptr(self):
        return ctypes.c_void_p(self.t.data_ptr())

    def __len__(self):
        return self.t.shape[0]

    @property
    def n(self):
        return self.t.shape[0]

    def _like(self):
        return DeviceArray.empty(self.ctx, self.n, self.t.device)

    def _check_contiguous(self):
        if not self.t.is_contiguous():
            self.t = self.t.contiguous()

    # ---- arithmetic (finfields.py:1056-1124,1189-1281) ---------------------------------------------
    def _binop(self, other, op):
        self._check_contiguous()
        out = self._like()
        if isinstance(other, DeviceArray):
            if other.ctx is not self.ctx or other.n != self.n:
                raise ValueError('operands must share field and length')
            other._check_contiguous()
            check(lib.mpyc_b200_ff_binop(self.ctx.handle, op, self.ptr, other.ptr, out.ptr, self.n, _stream_ptr()))
        elif isinstance(other, (int, np.integer)):
            check(lib.mpyc_b200_ff_binop_scalar(self.ctx.handle, op, self.ptr, self.ctx.scalar_limbs(other), out.ptr,
                                                self.n, _stream_ptr()))
        else:
            return NotImplemented
        return out

    def __add__(self, other):
        return self._binop(other, _cabi.OP_ADD)

    __radd__ = __add__

    def __sub__(self, other):
        return self._binop(other, _cabi.OP_SUB)

    def __rsub__(self, other):
        return (-self)._binop(other, _cabi.OP_ADD)

    def __mul__(self, other):
        return self._binop(other, _cabi.OP_MUL)

    __rmul__ = __mul__

    def __neg__(self):
        self._check_contiguous()
        out = self._like()
        check(lib.mpyc_b200_ff_neg(self.ctx.handle, self.ptr, out.ptr, self.n, _stream_ptr()))
        return out

    def __pow__(self, e):
        """a ** e for one public integer exponent (negative allowed: powmod semantics, finfields.py:1408-1414)."""
        if not isinstance(e, (int, np.integer)):
            return NotImplemented
        e = int(e)
        base = self
        if e < 0:
            base, e = self.reciprocal(), -e
        base._check_contiguous()
        q1 = self.ctx.order - 1
        if e > q1:   # a^(q-1) = 1 for a != 0, and 0^e = 0 for e > 0: keep the exponent in 1..q-1
            e = e % q1 or q1
        out = base._like()
        nl = max(1, (e.bit_length() + 63) // 64)
        check(lib.mpyc_b200_ff_pow(self.ctx.handle, base.ptr, _cabi.u64_array(_cabi.int_to_limbs(e, nl)), nl, out.ptr,
                                   self.n, _stream_ptr()))
        return out

    def reciprocal(self):
        """Elementwise inverse; ZeroDivisionError if any element is 0 (gmpy2.invert, finfields.py:1416-1422)."""
        self._check_contiguous()
        out = self._like()
        check(lib.mpyc_b200_ff_inv(self.ctx.handle, self.ptr, out.ptr, self.n, _stream_ptr()))
        return out

    def __truediv__(self, other):
        if isinstance(other, DeviceArray):
            return self * other.reciprocal()
        if isinstance(other, (int, np.integer)) and not self.ctx.binary:
            return self * pow(int(other), -1, self.ctx.order)
        return NotImplemented

    def __lshift__(self, k):
        if self.ctx.binary:
            return NotImplemented   # polynomial shifts of GF(2^8) arrays are not part of the batched surface
        return self * pow(2, int(k), self.ctx.order)

    def __rshift__(self, k):
        """'>> k' multiplies by (2^k)^-1 mod p (finfields.py:1250-1259) -- not a bit shift."""
        if self.ctx.binary:
            return NotImplemented
        return self * pow(pow(2, int(k), self.ctx.order), -1, self.ctx.order)

    def sqrt(self, INV=False):
        self._check_contiguous()
        out = self._like()
        check(lib.mpyc_b200_ff_sqrt(self.ctx.handle, self.ptr, 1 if INV else 0, out.ptr, self.n, _stream_ptr()))
        return out

    def is_sqr(self):
        self._check_contiguous()
        out = torch.empty(self.n, dtype=torch.uint8, device=self.t.device)
        check(lib.mpyc_b200_ff_is_sqr(self.ctx.handle, self.ptr, ctypes.c_void_p(out.data_ptr()), self.n, _stream_ptr()))
        return out.bool()

    def signed_(self):
        """Host-side signed representatives (finfields.py:1395-1398)."""
        p = self.ctx.modulus
        v = self.to_ints()
        return np.where(v > p >> 1, v - p, v)

    def count_mismatch(self, other):
        cnt = torch.zeros(1, dtype=torch.int64, device=self.t.device)
        self._check_contiguous()
        other._check_contiguous()
        check(lib.mpyc_b200_count_mismatch(self.ctx.handle, self.ptr, other.ptr, self.n,
                                           ctypes.c_void_p(cnt.data_ptr()), _stream_ptr()))
        return int(cnt.item())


class DeviceMatrix:
    """rows x n elements (row-major, row stride in elements) -- share or coefficient matrices."""

    __slots__ = ('ctx', 't')

    def __init__(self, ctx, tensor):
        self.ctx = ctx
        self.t = tensor   # int64 (rows, n, L) with stride(0) = row stride * L, or uint8 (rows, n)

    @classmethod
    def empty(cls, ctx, rows, n, device='cuda'):
        return cls(ctx, _alloc(ctx, rows, n, device))

    @classmethod
    def from_ints(cls, ctx, rows_of_values, device='cuda'):
        rows = [codec.ints_to_limbs(r, ctx) for r in rows_of_values]
        out = cls.empty(ctx, len(rows), rows[0].shape[0] if rows else 0, device)
        for i, r in enumerate(rows):
            out.t[i].copy_(torch.from_numpy(r if ctx.binary else r.view(np.int64)))
        return out

    @property
    def rows(self):
        return self.t.shape[0]

    @property
    def n(self):
        return self.t.shape[1]

    @property
    def stride(self):
        """row stride in elements"""
        if self.t.shape[0] <= 1:
            return max(self.n, 1)
        return self.t.stride(0) // (1 if self.ctx.binary else self.ctx.nlimbs)

    @property
    def ptr(self):
        return ctypes.c_void_p(self.t.data_ptr())

    def row(self, i):
        return DeviceArray(self.ctx, self.t[i])

    def to_ints(self):
        return [self.row(i).to_ints() for i in range(self.rows)]


def shamir_split(ctx, secrets: DeviceArray, coeffs, t, m, out=None):
    """Device-resident np_random_split with explicit coefficients (mpyc/thresha.py:47-64).

    coeffs: DeviceMatrix with t rows (row j-1 = coefficient of X^j), or None when t == 0."""
    n = secrets.n
    secrets._check_contiguous()
    if out is None:
        out = DeviceMatrix.empty(ctx, m, n, secrets.t.device)
    cptr, cstride = (coeffs.ptr, coeffs.stride) if (t > 0 and coeffs is not None) else (ctypes.c_void_p(0), n)
    check(lib.mpyc_b200_shamir_split(ctx.handle, secrets.ptr, cptr, cstride, out.ptr, out.stride, n, t, m, _stream_ptr()))
    return out


def shamir_recombine(ctx, xs, rows, x_rs=0, out=None):
    """Device-resident np_recombine (mpyc/thresha.py:119-132). rows: list of DeviceArray (one per x in xs)."""
    single = not isinstance(x_rs, (list, tuple))
    pts = [x_rs] if single else list(x_rs)
    n = rows[0].n
    for r in rows:
        r._check_contiguous()
    if out is None:
        out = DeviceMatrix.empty(ctx, len(pts), n, rows[0].t.device)
    check(lib.mpyc_b200_shamir_recombine(ctx.handle, _cabi.ptr_array([r.t.data_ptr() for r in rows]),
                                         _cabi.i64_array([int(x) for x in xs]), len(rows),
                                         _cabi.i64_array([int(x) for x in pts]), len(pts), out.ptr, out.stride, n,
                                         _stream_ptr()))
    return out.row(0) if single else out


def shamir_split_generate(ctx, secrets: DeviceArray, t, m, key=None, nonce=0, out=None):
    """Share generation with coefficients drawn INSIDE the kernel from a ChaCha20 keystream (never written
    to memory).  key: 32 bytes of fresh CSPRNG output (default: os.urandom(32)); nonce: 63-bit call counter.
    The role of secrets.randbelow in mpyc/thresha.py:58-60; not bit-reproducible against anything unless the
    key is fixed (tests do that)."""
    import os
    key = os.urandom(32) if key is None else bytes(key)
    if len(key) != 32:
        raise ValueError('key must be 32 bytes')
    n = secrets.n
    secrets._check_contiguous()
    if out is None:
        out = DeviceMatrix.empty(ctx, m, n, secrets.t.device)
    kbuf = (ctypes.c_uint8 * 32).from_buffer_copy(key)
    check(lib.mpyc_b200_shamir_split_generate(ctx.handle, secrets.ptr, out.ptr, out.stride, n, t, m, kbuf,
                                              int(nonce) & (2**63 - 1), _stream_ptr()))
    return out


def shamir_split_generate_rows(ctx, secrets: DeviceArray, t, m, row_ptrs, key=None, nonce=0):
    """As shamir_split_generate, every share row written to its own destination: row_ptrs = m device addresses
    (ints), each with room for n elements, 32-byte aligned for the vector path.  A row may live on another GPU
    (memory mapped through CUDA IPC with peer access enabled): the kernel then stores it there over NVLink."""
    import os
    key = os.urandom(32) if key is None else bytes(key)
    if len(key) != 32 or len(row_ptrs) != m:
        raise ValueError('key must be 32 bytes and row_ptrs must hold m addresses')
    secrets._check_contiguous()
    kbuf = (ctypes.c_uint8 * 32).from_buffer_copy(key)
    check(lib.mpyc_b200_shamir_split_generate_rows(ctx.handle, secrets.ptr, _cabi.ptr_array([int(a) for a in row_ptrs]),
                                                   secrets.n, t, m, kbuf, int(nonce) & (2**63 - 1), _stream_ptr()))


def matmul(ctx, A, B, r, k, c):
    """C = A @ B mod p for row-major DeviceArrays A (r*k elements) and B (k*c elements): DeviceArray of r*c
    elements (FiniteFieldArray.__matmul__, mpyc/finfields.py:1126-1146)."""
    if A.n != r * k or B.n != k * c:
        raise ValueError('matmul: shapes do not match the buffers')
    A._check_contiguous()
    B._check_contiguous()
    out = DeviceArray.empty(ctx, r * c, A.t.device)
    check(lib.mpyc_b200_ff_matmul(ctx.handle, A.ptr, B.ptr, out.ptr, r, k, c, _stream_ptr()))
    return out


# ---- protocol-local algebra on raw share values (csrc/local.cuh) -----------------------------

def fma(a: DeviceArray, b, c: DeviceArray):
    """a*b + c, or a*a + c when b is None -- np_random_bits' `_r.value**2 + z.value` (mpyc/runtime.py:4252)."""
    if c.ctx is not a.ctx or c.n != a.n or (b is not None and (b.ctx is not a.ctx or b.n != a.n)):
        raise ValueError('operands must share field and length')
    for x in (a, b, c):
        if x is not None:
            x._check_contiguous()
    out = a._like()
    check(lib.mpyc_b200_ff_fma(a.ctx.handle, a.ptr, b.ptr if b is not None else ctypes.c_void_p(0), c.ptr, out.ptr, a.n,
                               _stream_ptr()))
    return out


def axpb(a: DeviceArray, s, t):
    """a*s + t for public integers s, t (reduced here) -- the affine steps on raw values: `bits += 1; bits *= (p+1)>>1;
    bits <<= f` (mpyc/runtime.py:4267-4271), `(r_bits << 1) - 1` (:3646), `x + (1 << l)` (:3656)."""
    a._check_contiguous()
    q = a.ctx.order
    out = a._like()
    check(lib.mpyc_b200_ff_axpb(a.ctx.handle, a.ptr, a.ctx.scalar_limbs(int(s) % q), a.ctx.scalar_limbs(int(t) % q), out.ptr,
                                a.n, _stream_ptr()))
    return out


def low_bits(a: DeviceArray, nbits):
    """a & (2^nbits - 1) on canonical residues -- `c.value & ((1<<f) - 1)` (mpyc/runtime.py:870, 3657)."""
    a._check_contiguous()
    out = a._like()
    check(lib.mpyc_b200_ff_low_bits(a.ctx.handle, a.ptr, int(nbits), out.ptr, a.n, _stream_ptr()))
    return out


def nonzero(a: DeviceArray, want_mask=True):
    """(bool tensor a != 0 or None, number of non-zero elements) -- `_r2.value != 0` (mpyc/runtime.py:4254-4255)."""
    a._check_contiguous()
    mask = torch.empty(a.n, dtype=torch.uint8, device=a.t.device) if want_mask else None
    cnt = torch.zeros(1, dtype=torch.int64, device=a.t.device)
    check(lib.mpyc_b200_ff_nonzero(a.ctx.handle, a.ptr, ctypes.c_void_p(mask.data_ptr() if want_mask else 0),
                                   ctypes.c_void_p(cnt.data_ptr()), a.n, _stream_ptr()))
    return (mask.bool() if want_mask else None), int(cnt.item())


def bits_compose(bits: DeviceArray, n, f, descending=False):
    """out[i] = sum_j bits[i*f + j] << e(j) mod p, e(j) = j or f-1-j: `np.sum(r_bits.reshape((n, f)) << shifts, axis=1)`
    (mpyc/runtime.py:860, 3650-3651, 4415).  bits: n*f residues, row-major."""
    if bits.n != n * f:
        raise ValueError('bits_compose: buffer does not hold n*f elements')
    bits._check_contiguous()
    out = DeviceArray.empty(bits.ctx, n, bits.t.device)
    if n and f:
        check(lib.mpyc_b200_ff_bits_compose(bits.ctx.handle, bits.ptr, n, f, 1 if descending else 0, out.ptr, _stream_ptr()))
    elif n:
        out.t.zero_()
    return out


def bits_decompose(c: DeviceArray, l, descending=False):
    """DeviceMatrix (l, n): row j = bit e(j) of every c[i] as a field element: `np.right_shift.outer(c, shifts).T & 1`
    (mpyc/runtime.py:3659-3660, 4422-4423)."""
    c._check_contiguous()
    out = DeviceMatrix.empty(c.ctx, l, c.n, c.t.device)
    if l and c.n:
        check(lib.mpyc_b200_ff_bits_decompose(c.ctx.handle, c.ptr, c.n, l, 1 if descending else 0, out.ptr, out.stride,
                                              _stream_ptr()))
    return out


def conv2d(X: DeviceArray, W: DeviceArray, B: DeviceArray, k, r, m, n, v, s):
    """np_cnnmnist's convolvetensor body (demos/np_cnnmnist.py:69-81) mod p: X (k,r,m,n), W (v,r,s,s), B (v) row-major
    -> Y (k,v,m,n)."""
    if X.n != k * r * m * n or W.n != v * r * s * s or B.n != v:
        raise ValueError('conv2d: shapes do not match the buffers')
    for x in (X, W, B):
        x._check_contiguous()
    out = DeviceArray.empty(X.ctx, k * v * m * n, X.t.device)
    check(lib.mpyc_b200_ff_conv2d(X.ctx.handle, X.ptr, W.ptr, B.ptr, out.ptr, k, r, m, n, v, s, _stream_ptr()))
    return out


def bits_decompose_flat(c: DeviceArray, l, descending=False):
    """As bits_decompose, into ONE contiguous buffer of l*n elements (row-major (l, n), no row padding)."""
    c._check_contiguous()
    out = DeviceArray.empty(c.ctx, l * c.n, c.t.device)
    if l and c.n:
        check(lib.mpyc_b200_ff_bits_decompose(c.ctx.handle, c.ptr, c.n, l, 1 if descending else 0, out.ptr, c.n, _stream_ptr()))
    return out


def transpose(a: DeviceArray, rows, cols):
    """(rows, cols) row-major -> (cols, rows): `r_bits.T` (mpyc/runtime.py:3661)."""
    if a.n != rows * cols:
        raise ValueError('transpose: buffer does not hold rows*cols elements')
    a._check_contiguous()
    out = a._like()
    check(lib.mpyc_b200_ff_transpose(a.ctx.handle, a.ptr, rows, cols, out.ptr, _stream_ptr()))
    return out


def cumsum_rows(a: DeviceArray, rows, cols):
    """Running sums down the rows of a (rows, cols) matrix: `np.cumsum(x, axis=0)` mod p (mpyc/runtime.py:3667)."""
    if a.n != rows * cols:
        raise ValueError('cumsum_rows: buffer does not hold rows*cols elements')
    a._check_contiguous()
    out = a._like()
    check(lib.mpyc_b200_ff_cumsum_rows(a.ctx.handle, a.ptr, rows, cols, out.ptr, _stream_ptr()))
    return out


def binop_rows(a: DeviceArray, b: DeviceArray, op, rows, cols, reflected=False):
    """a (rows, cols) (op) b (cols) broadcast over the rows; reflected: b (op) a (mpyc/runtime.py:3670)."""
    if a.n != rows * cols or b.n != cols or a.ctx is not b.ctx:
        raise ValueError('binop_rows: shapes do not match the buffers')
    a._check_contiguous()
    b._check_contiguous()
    out = a._like()
    check(lib.mpyc_b200_ff_binop_rows(a.ctx.handle, op, 1 if reflected else 0, a.ptr, b.ptr, out.ptr, rows, cols, _stream_ptr()))
    return out
