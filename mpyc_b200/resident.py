"""Limb-resident field arrays behind the reference's own FiniteFieldArray objects.

MPyC's field arrays keep a NumPy dtype=object array of Python ints in their `value` slot
(mpyc/finfields.py:695-725) and every protocol step of runtime.py -- local product (`c = a * b`,
runtime.py:1133), `_reshare` (:603-689: np_random_split, pickle, np_recombine), the next product -- goes through
that representation, i.e. through ~40 ns per element of int <-> limb conversion on each side of every kernel.
This module lets the SAME objects carry their elements as limb buffers instead:

  * `LimbValue` is what `array.value` holds while the elements live as limbs (in HBM as a DeviceArray, or in host
    memory as the uint64 (n, L) / uint8 (n,) array the wire delivered).  It answers the questions runtime.py asks
    of `.value` on the protocol path -- shape, reshape(-1), len, pickle -- without creating a single Python int;
    ANY other use (indexing, arithmetic on the raw value as in runtime.py:860-872,3656-3671, iteration, NumPy
    functions through __array__) materialises the object array once and from then on the LimbValue is a plain
    delegate to it: behaviour identical to the reference's, at the reference's cost.
  * install(operators=True) (mpyc_b200/install.py) turns FiniteFieldArray.value into a property over the original
    slot: only the code that is known to treat `.value` as an opaque handle -- runtime._reshare / output /
    _distribute and FiniteFieldArray.reshape / ravel / flatten / copy / shape / ndim / size / __len__ -- ever sees a
    LimbValue; for any other reader the property first settles the slot to the real object ndarray, so raw values
    that escape into isinstance checks, in-place updates or C-level NumPy calls are exactly what the reference has.
    It also wraps FiniteFieldArray.__init__ so that a LimbValue / ShareRow is stored as is, and `+ - * neg @` (finfields.py:1056-1146) so that limb-backed operands -- or plain object
    arrays of at least `min_size` elements -- are combined by the K1 / K1c kernels, the result being limb-backed
    again.  Broadcasting, mixed operand kinds and everything below the threshold stay with the reference's code.
  * with `resident = True` np_recombine / np_pseudorandom_share return limb-backed arrays and np_random_split
    accepts them, so a chain input -> multiply -> reshare -> multiply -> ... -> output creates Python ints only at
    input and at output.
  * `ModValue` (with `resident` and `local_algebra`): Runtime.np_random_bits / np_trunc / np_sgn / np_to_bits -- recognised
    by the hash of their source -- compute on RAW share values between two openings (runtime.py:856-872, 3644-3694,
    4243-4273, 4413-4433).  Their `.value` reads get a ModValue, which evaluates those NumPy expressions mod p on the
    K1 / K6 kernels (csrc/local.cuh) instead of settling; see the class docstring for why that is exact.  The hooks
    around them (<< >>, contiguous __getitem__, np.concatenate(axis=0), == / != with a scalar, matrix/vector row
    broadcast) keep np_prod's halving loop and np_is_zero_public limb-backed as well.

All arithmetic runs on the GPU through the C ABI (`backend` below); there is no CPU path.  tests/oracle_device.py
replaces `backend` by the oracle to exercise this module's host logic where no GPU exists.
"""
import numpy as np

from mpyc_b200 import _cabi, codec
from mpyc_b200.field import context_for, context_of_field
from mpyc_b200.wire import ShareRow, _poly_type

import sys

resident = False      # np_recombine / PRSS results stay limb-backed (set by install(resident=True))
min_size = 1024       # plain object-array operands of fewer elements are left to the reference's operators
calls = {'limb_ops': 0, 'materialised': 0, 'packed': 0, 'mod_values': 0}    # counters (tests, profiling)


# ---------------------------------------------------------------------------------------------
# backend: where the limbs live and which kernels combine them
# ---------------------------------------------------------------------------------------------

class CudaBackend:
    """Stores are mpyc_b200.device.DeviceArray (1-D, n elements) on `device`; host limb arrays are uploaded on
    first use.  torch is imported lazily (device memory and streams only)."""

    def __init__(self, device=0):
        self.device = device

    def _dev(self):
        import torch
        from mpyc_b200 import device as dev
        return dev, torch.device('cuda', self.device)

    def to_store(self, ctx, limbs):
        dev, where = self._dev()
        if isinstance(limbs, dev.DeviceArray):
            return limbs
        return dev.DeviceArray.from_limbs(ctx, np.ascontiguousarray(limbs), where)

    def to_host(self, ctx, store):
        if isinstance(store, np.ndarray):
            return store
        return np.ascontiguousarray(store.to_limbs())

    def binop(self, ctx, op, a, b):
        a, b = self.to_store(ctx, a), self.to_store(ctx, b)
        return a._binop(b, op)

    def binop_scalar(self, ctx, op, a, scalar):
        return self.to_store(ctx, a)._binop(int(scalar), op)

    def neg(self, ctx, a):
        return -self.to_store(ctx, a)

    def matmul(self, ctx, a, b, r, k, c):
        dev, _ = self._dev()
        return dev.matmul(ctx, self.to_store(ctx, a), self.to_store(ctx, b), r, k, c)

    def split(self, ctx, sec, t, m, coeffs=None):
        """sec: store of n secrets -> host limb array (m, n, L) / (m, n): the rows leave for the wire anyway."""
        dev, where = self._dev()
        sec = self.to_store(ctx, sec)
        if coeffs is not None:
            C = dev.DeviceMatrix.empty(ctx, t, sec.n, where) if t else None
            import torch
            for j in range(t):
                src = np.ascontiguousarray(coeffs[j])
                C.t[j].copy_(torch.from_numpy(src if ctx.binary else src.view(np.int64)))
            out = dev.shamir_split(ctx, sec, C, t, m)
        else:
            out = dev.shamir_split_generate(ctx, sec, t, m)
        a = out.t.contiguous().cpu().numpy()
        return a if ctx.binary else a.view(np.uint64)

    def recombine(self, ctx, xs, rows, pts):
        """rows: stores (host limb arrays or DeviceArrays) -> list of len(pts) stores."""
        dev, _ = self._dev()
        rows = [self.to_store(ctx, r) for r in rows]
        out = dev.shamir_recombine(ctx, xs, rows, list(pts))
        return [out.row(r) for r in range(len(pts))]

    # ---- protocol-local algebra (csrc/local.cuh), used by ModValue ----
    def fma(self, ctx, a, b, c):
        dev, _ = self._dev()
        return dev.fma(self.to_store(ctx, a), None if b is None else self.to_store(ctx, b), self.to_store(ctx, c))

    def axpb(self, ctx, a, s, t):
        dev, _ = self._dev()
        return dev.axpb(self.to_store(ctx, a), s, t)

    def low_bits(self, ctx, a, nbits):
        dev, _ = self._dev()
        return dev.low_bits(self.to_store(ctx, a), nbits)

    def nonzero(self, ctx, a):
        """(bool ndarray a != 0, count)"""
        dev, _ = self._dev()
        mask, count = dev.nonzero(self.to_store(ctx, a))
        return mask.cpu().numpy(), count

    def bits_compose(self, ctx, bits, n, f, descending):
        dev, _ = self._dev()
        return dev.bits_compose(self.to_store(ctx, bits), n, f, descending)

    def sqrt(self, ctx, a, INV):
        return self.to_store(ctx, a).sqrt(INV=INV)

    def bits_decompose(self, ctx, c, l, descending):
        """store of l*n elements, row-major (l, n)"""
        dev, _ = self._dev()
        return dev.bits_decompose_flat(self.to_store(ctx, c), l, descending)

    def transpose(self, ctx, a, rows, cols):
        dev, _ = self._dev()
        return dev.transpose(self.to_store(ctx, a), rows, cols)

    def cumsum_rows(self, ctx, a, rows, cols):
        dev, _ = self._dev()
        return dev.cumsum_rows(self.to_store(ctx, a), rows, cols)

    def binop_rows(self, ctx, op, a, b, rows, cols, reflected):
        dev, _ = self._dev()
        return dev.binop_rows(self.to_store(ctx, a), self.to_store(ctx, b), op, rows, cols, reflected)

    def conv2d(self, ctx, X, W, B, k, r, m, n, v, s):
        dev, _ = self._dev()
        return dev.conv2d(self.to_store(ctx, X), self.to_store(ctx, W), self.to_store(ctx, B), k, r, m, n, v, s)

    def slice(self, ctx, store, start, stop):
        if isinstance(store, np.ndarray):
            return store[start:stop]
        dev, _ = self._dev()
        return dev.DeviceArray(ctx, store.t[start:stop])

    def concat(self, ctx, stores):
        """One store holding the elements of `stores` back to back (device-to-device copies; torch is the allocator)."""
        dev, _ = self._dev()
        parts = [self.to_store(ctx, st) for st in stores]
        out = dev.DeviceArray.empty(ctx, sum(p.n for p in parts), parts[0].t.device)
        at = 0
        for p in parts:                       # contiguous device-to-device copies (cudaMemcpyAsync)
            out.t[at:at + p.n].copy_(p.t)
            at += p.n
        return out


backend = CudaBackend()


# ---------------------------------------------------------------------------------------------
# LimbValue
# ---------------------------------------------------------------------------------------------

def _norm_shape(n, shape):
    if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
        shape = tuple(shape[0])
    shape = tuple(int(s) for s in shape)
    if shape.count(-1) > 1:
        raise ValueError('can only specify one unknown dimension')
    if -1 in shape:
        known = 1
        for s in shape:
            if s != -1:
                known *= s
        if known == 0 or n % known:
            raise ValueError(f'cannot reshape array of size {n} into shape {shape}')
        shape = tuple(n // known if s == -1 else s for s in shape)
    size = 1
    for s in shape:
        size *= s
    if size != n:
        raise ValueError(f'cannot reshape array of size {n} into shape {shape}')
    return shape


class LimbValue:
    """The `value` of a field array whose elements live as limbs (see module docstring).

    store: DeviceArray or host limb array holding prod(shape) elements in C order; immutable by convention (every
    operation produces a new store), so reshape / ravel / copy share it."""

    __slots__ = ('ctx', 'shape', 'store', '_ints', '_poly')
    __array_priority__ = 0.0
    __hash__ = None

    def __init__(self, ctx, store, shape=None, poly_type=None):
        self.ctx = ctx
        self.store = store
        n = len(store)
        self.shape = (n,) if shape is None else tuple(shape)
        self._ints = None
        self._poly = poly_type

    # ---- what the protocol path of runtime.py asks of `.value` (no Python ints are created) -------------
    @property
    def limb_backed(self):
        return self.store is not None

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        n = 1
        for s in self.shape:
            n *= s
        return n

    dtype = np.dtype(object)

    def __len__(self):
        if not self.shape:
            raise TypeError('len() of unsized object')
        return self.shape[0]

    def reshape(self, *shape, **kwargs):
        if self.store is None or kwargs:
            return self._materialise().reshape(*shape, **kwargs)
        return LimbValue(self.ctx, self.store, _norm_shape(self.size, shape), self._poly)

    def ravel(self, *args, **kwargs):
        if self.store is None or args or kwargs:
            return self._materialise().ravel(*args, **kwargs)
        return LimbValue(self.ctx, self.store, (self.size,), self._poly)

    def flatten(self, *args, **kwargs):
        if self.store is None or args or kwargs:
            return self._materialise().flatten(*args, **kwargs)
        return LimbValue(self.ctx, self.store, (self.size,), self._poly)

    def copy(self, *args, **kwargs):
        if self.store is None:
            return self._ints.copy(*args, **kwargs)
        return LimbValue(self.ctx, self.store, self.shape, self._poly)

    def host_limbs(self):
        """The elements as a host limb array (n, L) / (n,) (one D2H copy when they are in HBM)."""
        return backend.to_host(self.ctx, self.store)

    def __reduce__(self):
        if self.store is None:
            return _from_ints, (self._ints,)
        return _from_wire, (self.ctx.modulus, self.ctx.binary, self.shape, codec.limbs_to_wire(self.host_limbs(), self.ctx))

    # ---- everything else: become the object array the reference would have had ---------------------------
    def _materialise(self):
        if self._ints is None:
            vals = codec.limbs_to_ints(self.host_limbs(), self.ctx)
            if self.ctx.binary:
                tp = self._poly or _poly_type()
                out = np.empty(len(vals), dtype=object)
                out[:] = [tp(int(v)) for v in vals]
                vals = out
            self._ints = vals.reshape(self.shape)
            self.store = None        # the ints may be modified in place from here on: the limbs are dropped
            calls['materialised'] += 1
        return self._ints

    def __array__(self, dtype=None, copy=None):
        a = self._materialise()
        if dtype is not None and dtype != object:
            return a.astype(dtype)
        return a.copy() if copy else a

    def __getattr__(self, name):
        if name.startswith('__') and name.endswith('__'):
            raise AttributeError(name)
        return getattr(self._materialise(), name)

    def __repr__(self):
        if self.store is None:
            return repr(self._ints)
        where = 'host' if isinstance(self.store, np.ndarray) else 'device'
        return f'LimbValue(shape={self.shape}, {where}, {self.ctx!r})'


def _delegate(name):
    def method(self, *args, **kwargs):
        return getattr(self._materialise(), name)(*args, **kwargs)
    method.__name__ = name
    return method


for _name in ('getitem', 'setitem', 'iter', 'contains', 'add', 'radd', 'iadd', 'sub', 'rsub', 'isub', 'mul', 'rmul', 'imul',
              'matmul', 'rmatmul', 'imatmul', 'mod', 'rmod', 'imod', 'floordiv', 'rfloordiv', 'ifloordiv', 'truediv',
              'rtruediv', 'itruediv', 'pow', 'rpow', 'ipow', 'neg', 'pos', 'abs', 'invert', 'lshift', 'rlshift', 'ilshift',
              'rshift', 'rrshift', 'irshift', 'and', 'rand', 'iand', 'or', 'ror', 'ior', 'xor', 'rxor', 'ixor', 'lt', 'le',
              'gt', 'ge', 'eq', 'ne', 'bool', 'int', 'index', 'float', 'divmod', 'rdivmod', 'str', 'format'):
    setattr(LimbValue, f'__{_name}__', _delegate(f'__{_name}__'))



# ---------------------------------------------------------------------------------------------
# ModValue: raw-value algebra of whitelisted protocol functions, mod p on the device
# ---------------------------------------------------------------------------------------------

class ModValue:
    """What `array.value` returns to np_random_bits / np_trunc / np_sgn / np_to_bits (`_MOD_SOURCES`) while the elements are
    limbs.

    Those functions compute on raw share values with NumPy object arithmetic -- `_r.value**2 + z.value`,
    `np.sum(r_bits.value.reshape((n, f)) << np.arange(f), axis=1)`, `ar_modf + (1 << l-1) + (r_divf << f)`,
    `c.value & ((1<<f) - 1)`, `bits += 1; bits *= (p+1) >> 1` (mpyc/runtime.py:856-872, 3644-3658, 4252-4271) -- over
    the integers, and reduce when the result enters `Zp.array(...)`.  A ModValue runs the same expressions mod p on the
    K1 / K6 kernels: reduction is a ring homomorphism, so the field array that comes out is identical.  Operations that
    are not ring operations (`&`, `!= 0`, `%` by a power of two) are only accepted on EXACT values -- residues read
    straight from a field array (an opened value) -- where the canonical residue IS the reference's integer; on a
    modular intermediate they raise instead of returning something that could differ.

    Scalar steps are kept as a pending affine map (value = store * s + t) and applied by one k_axpb launch when the
    elements are needed.  Anything not covered here turns the value into the settled object array (canonical
    residues) and continues with NumPy, like LimbValue."""

    __slots__ = ('ctx', 'shape', 'store', '_s', '_t', '_sq', 'exact', '_ints')
    __hash__ = None
    dtype = np.dtype(object)
    # sectypes.SecureArray subclasses accept raw values only as `isinstance(value, np.ndarray)` (sectypes.py:1402,1440)
    # and then hand them to `field.array(value)` -- which is hooked and takes the limbs.  isinstance() consults
    # __class__; type() -- which every check in this package uses -- does not.  NumPy's C code is not fooled either
    # (PyArray_Check looks at the type), it goes through __array__ / __array_ufunc__ / __array_function__ below.
    __class__ = property(lambda self: np.ndarray)

    def __init__(self, ctx, store, shape, exact=False, s=1, t=0, sq=False):
        self.ctx, self.store, self.shape = ctx, store, tuple(shape)
        self._s, self._t, self._sq = s, t, sq
        self.exact = exact and s == 1 and t == 0 and not sq
        self._ints = None

    # ---- plumbing ---------------------------------------------------------------------------------
    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        n = 1
        for d in self.shape:
            n *= d
        return n

    def __len__(self):
        if not self.shape:
            raise TypeError('len() of unsized object')
        return self.shape[0]

    def _flush(self):
        """Apply the pending square / affine map; afterwards store holds the canonical residues of the value."""
        if self._sq:
            self.store = backend.binop(self.ctx, _cabi.OP_MUL, self.store, self.store)
            self._sq = False
            calls['limb_ops'] += 1
        if self._s != 1 or self._t != 0:
            self.store = backend.axpb(self.ctx, self.store, self._s, self._t)
            self._s, self._t = 1, 0
            calls['limb_ops'] += 1
        return self.store

    def limb_value(self):
        """The value as a LimbValue (for FiniteFieldArray.__init__), or None once it has become an object array."""
        if self.store is None:
            return None
        return LimbValue(self.ctx, self._flush(), self.shape)

    def _new(self, store, shape=None, exact=False, s=1, t=0, sq=False):
        return ModValue(self.ctx, store, self.shape if shape is None else shape, exact, s, t, sq)

    def _materialise(self):
        if self._ints is None:
            lv = LimbValue(self.ctx, self._flush(), self.shape)
            self._ints = lv._materialise()
            self.store = None
        return self._ints

    def __array__(self, dtype=None, copy=None):
        a = self._materialise()
        if dtype is not None and dtype != object:
            return a.astype(dtype)
        return a.copy() if copy else a

    def __getattr__(self, name):
        if name.startswith('__') and name.endswith('__'):
            raise AttributeError(name)
        return getattr(self._materialise(), name)

    def __repr__(self):
        if self.store is None:
            return repr(self._ints)
        return f'ModValue(shape={self.shape}, exact={self.exact}, {self.ctx!r})'

    def __reduce__(self):
        return _from_ints, (self._materialise(),)

    def reshape(self, *shape, **kwargs):
        if self.store is None or kwargs:
            return self._materialise().reshape(*shape, **kwargs)
        return self._new(self.store, _norm_shape(self.size, shape), self.exact, self._s, self._t, self._sq)

    @property
    def T(self):
        """`r_bits.T` (runtime.py:3661): a 2-D value is transposed on the device (k_transpose)."""
        if self.store is not None and len(self.shape) == 2:
            r, c = self.shape
            calls['limb_ops'] += 1
            return self._new(backend.transpose(self.ctx, self._flush(), r, c), (c, r), self.exact)
        return self._materialise().T

    def ravel(self, *args, **kwargs):
        if self.store is None or args or kwargs:
            return self._materialise().ravel(*args, **kwargs)
        return self.reshape(-1)

    flatten = ravel

    def copy(self, *args, **kwargs):
        if self.store is None:
            return self._ints.copy(*args, **kwargs)
        return self._new(self.store, None, self.exact, self._s, self._t, self._sq)

    def __getitem__(self, key):
        if (self.store is not None and len(self.shape) == 1 and isinstance(key, slice) and key.step in (None, 1)):
            start, stop, _ = key.indices(self.shape[0])
            stop = max(stop, start)
            return self._new(backend.slice(self.ctx, self._flush(), start, stop), (stop - start,), self.exact)
        return self._materialise()[key]

    def __setitem__(self, key, value):
        self._materialise()[key] = value

    def __iter__(self):
        return iter(self._materialise())

    # ---- ring operations ----------------------------------------------------------------------------
    def _scalar(self, x):
        return int(x) if isinstance(x, (int, np.integer)) and not isinstance(x, (bool, np.bool_)) else None

    def _other_store(self, other, shape=None):
        """Store of an array operand of the given shape (default: this value's), packed from ints if need be; else None."""
        shape = self.shape if shape is None else tuple(shape)
        if type(other) is ModValue:
            if other.store is None:
                other = other._ints
            elif other.ctx is self.ctx and other.shape == shape:
                return other._flush()
            else:
                return None
        if type(other) is LimbValue:
            if other.store is not None:
                return other.store if (other.ctx is self.ctx and other.shape == shape) else None
            other = other._ints
        if type(other) is np.ndarray and other.shape == shape and other.dtype.kind in 'Oiu' and other.size:
            calls['packed'] += 1
            if other.dtype != object:
                other = other.astype(object)
            return codec.ints_to_limbs(other.reshape(-1), self.ctx)      # reduces: raw values may be negative / >= p
        return None

    def _ring(self, other, op, reflected=False):
        """self (op) other, op in '+', '-', '*'; reflected: other (op) self."""
        if self.store is None:
            return _NUMPY_OPS[op](other, self._ints) if reflected else _NUMPY_OPS[op](self._ints, other)
        p = self.ctx.modulus
        k = self._scalar(other)
        if k is not None:
            s, t = self._s, self._t
            if self._sq:
                self._flush()
                s, t = 1, 0
            if op == '+':
                t = (t + k) % p
            elif op == '-':
                s, t = ((-s) % p, (k - t) % p) if reflected else (s, (t - k) % p)
            else:
                s, t = (s * k) % p, (t * k) % p
            return self._new(self.store, None, False, s, t)
        b = self._other_store(other)
        if b is None:
            r = self._ring_rows(other, op, reflected)
            if r is not None:
                return r
            a = self._materialise()
            return _NUMPY_OPS[op](other, a) if reflected else _NUMPY_OPS[op](a, other)
        calls['limb_ops'] += 1
        if op == '+' and self._sq and self._s == 1 and self._t == 0:                      # a*a + c in one pass
            return self._new(backend.fma(self.ctx, self.store, None, b))
        a = self._flush()
        code = {'+': _cabi.OP_ADD, '-': _cabi.OP_SUB, '*': _cabi.OP_MUL}[op]
        x, y = (b, a) if reflected else (a, b)
        return self._new(backend.binop(self.ctx, code, x, y))

    def _ring_rows(self, other, op, reflected):
        """NumPy's row broadcast between a vector (C,) and a matrix (R, C) -- `s_sign - <matrix>` (runtime.py:3670)."""
        oshape = getattr(other, 'shape', None)
        if oshape is None or not self.shape:
            return None
        if len(self.shape) == 1 and len(oshape) == 2 and oshape[1] == self.shape[0]:
            vec, mat, vec_is_self = self, other, True
        elif len(self.shape) == 2 and len(oshape) == 1 and oshape[0] == self.shape[1]:
            vec, mat, vec_is_self = other, self, False
        else:
            return None
        R, C = mat.shape
        if R == 0 or C == 0:
            return None
        other_store = self._other_store(other, oshape)
        if other_store is None:
            return None
        mine = self._flush()
        mstore, vstore = (other_store, mine) if vec_is_self else (mine, other_store)
        code = {'+': _cabi.OP_ADD, '-': _cabi.OP_SUB, '*': _cabi.OP_MUL}[op]
        # kernel computes matrix (op) vector, or vector (op) matrix when `refl`
        refl = (vec_is_self and not reflected) or (not vec_is_self and reflected)
        calls['limb_ops'] += 1
        return ModValue(self.ctx, backend.binop_rows(self.ctx, code, mstore, vstore, R, C, refl), (R, C))

    def __add__(self, other):
        return self._ring(other, '+')

    __radd__ = __iadd__ = __add__

    def __sub__(self, other):
        return self._ring(other, '-')

    __isub__ = __sub__

    def __rsub__(self, other):
        return self._ring(other, '-', reflected=True)

    def __mul__(self, other):
        return self._ring(other, '*')

    __rmul__ = __imul__ = __mul__

    def __neg__(self):
        return self._ring(-1, '*')

    def __pos__(self):
        return self

    def __pow__(self, e):
        if self.store is not None and self._scalar(e) == 2 and not self._sq:
            self._flush()
            return self._new(self.store, None, False, sq=True)
        return self._materialise() ** e

    def __lshift__(self, k):
        if self.store is None:
            return self._ints << k
        n = self._scalar(k)
        if n is not None and n >= 0:
            return self._ring(pow(2, n, self.ctx.modulus), '*')
        if (isinstance(k, np.ndarray) and type(k) is np.ndarray and k.ndim == 1 and k.dtype.kind in 'iu' and len(self.shape) >= 2
                and k.shape[0] == self.shape[-1] and k.shape[0] > 0 and self.size):
            f = k.shape[0]
            if np.array_equal(k, np.arange(f)):
                return _ShiftedBits(self, k, False)
            if np.array_equal(k, np.arange(f - 1, -1, -1)):
                return _ShiftedBits(self, k, True)
        return self._materialise() << k

    __ilshift__ = __lshift__

    def sqrt(self, INV=False):
        """PrimeFieldArray._sqrt on the limbs (Blum primes; finfields.py:1424-1438): an exact value again."""
        return self._new(backend.sqrt(self.ctx, self._flush(), INV), None, True)

    # ---- operations that are not ring operations: exact values only -----------------------------------------
    def _need_exact(self, what):
        if not self.exact:
            raise AssertionError(f'mpyc_b200: {what} on a modular intermediate (only opened / stored residues are exact)')

    def __and__(self, mask):
        if self.store is None:
            return self._ints & mask
        m = self._scalar(mask)
        if m is not None and m >= 0 and (m + 1) & m == 0:           # 2^b - 1
            self._need_exact('bit mask')
            calls['limb_ops'] += 1
            return self._new(backend.low_bits(self.ctx, self.store, m.bit_length()), None, True)
        self._need_exact('bitwise and')
        return self._materialise() & mask

    __rand__ = __and__

    def __mod__(self, q):
        if self.store is None:
            return self._ints % q
        m = self._scalar(q)
        if m == self.ctx.modulus:
            return self if not self._sq and self._s == 1 and self._t == 0 else self._new(self._flush())
        if m is not None and m > 0 and m & (m - 1) == 0:
            return self & (m - 1)
        self._need_exact('%')
        return self._materialise() % q

    __imod__ = __mod__

    def __ne__(self, other):
        if self.store is not None and self._scalar(other) == 0:
            self._need_exact('!= 0')
            mask, _ = backend.nonzero(self.ctx, self.store)
            return np.asarray(mask, dtype=bool).reshape(self.shape)
        if self.store is not None:
            self._need_exact('comparison')
        return self._materialise() != other

    def __eq__(self, other):
        if self.store is not None:
            self._need_exact('comparison')
        return self._materialise() == other

    # ---- NumPy protocol: `ndarray (op) ModValue` and direct ufunc calls -----------------------------------------
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method == '__call__' and len(inputs) == 2 and not kwargs and ufunc in _UFUNC_OPS:
            a, b = inputs
            if a is self:
                return self._ring(b, _UFUNC_OPS[ufunc]) if ufunc is not np.left_shift else self.__lshift__(b)
            if ufunc is not np.left_shift:
                return self._ring(a, _UFUNC_OPS[ufunc], reflected=True)
        if (method == 'outer' and ufunc is np.right_shift and len(inputs) == 2 and not kwargs and inputs[0] is self
                and self.store is not None and len(self.shape) == 1 and self.exact):
            k = inputs[1]
            if isinstance(k, np.ndarray) and type(k) is np.ndarray and k.ndim == 1 and k.dtype.kind in 'iu' and k.shape[0] > 0:
                l = k.shape[0]
                if np.array_equal(k, np.arange(l)):
                    return _OuterBits(self, k, False, False)
                if np.array_equal(k, np.arange(l - 1, -1, -1)):
                    return _OuterBits(self, k, True, False)
        if method in ('__call__', 'outer') and ufunc in (np.right_shift, np.bitwise_and, np.not_equal, np.equal, np.remainder):
            for x in inputs:
                if type(x) is ModValue and x.store is not None:
                    x._need_exact(ufunc.__name__)
        args = [x._materialise() if type(x) in (ModValue, LimbValue) else x for x in inputs]
        if 'out' in kwargs:
            kwargs['out'] = tuple(x._materialise() if type(x) in (ModValue, LimbValue) else x for x in kwargs['out'])
        return getattr(ufunc, method)(*args, **kwargs)

    def __array_function__(self, func, types, args, kwargs):
        """np.vstack / np.cumsum(axis=0) of 2-D values stay on the device (runtime.py:3667-3670); every other NumPy
        function gets the settled object arrays."""
        if func is np.vstack and len(args) == 1 and not kwargs:
            r = _vstack(self.ctx, args[0])
            if r is not None:
                return r
        elif (func is np.cumsum and len(args) == 1 and args[0] is self and kwargs == {'axis': 0} and self.store is not None
              and len(self.shape) == 2 and self.size):
            calls['limb_ops'] += 1
            return self._new(backend.cumsum_rows(self.ctx, self._flush(), self.shape[0], self.shape[1]))
        return func(*_settled(args), **_settled(kwargs))


def _settled(x):
    """x with every ModValue / LimbValue inside (tuples, lists, dicts) replaced by its object array."""
    if type(x) in (ModValue, LimbValue):
        return x._materialise()
    if type(x) is _OuterBits or type(x) is _ShiftedBits:
        return x._eval()
    if isinstance(x, tuple):
        return tuple(_settled(v) for v in x)
    if isinstance(x, list):
        return [_settled(v) for v in x]
    if isinstance(x, dict):
        return {k: _settled(v) for k, v in x.items()}
    return x


def _vstack(ctx, parts):
    """np.vstack of 2-D pieces with the same number of columns, at least one limb-backed: ModValue, else None."""
    parts = list(parts)
    stores, rows, cols = [], 0, None
    for part in parts:
        shape = getattr(part, 'shape', None)
        if shape is None or len(shape) != 2 or (cols is not None and shape[1] != cols) or shape[0] == 0 or shape[1] == 0:
            return None
        cols = shape[1]
        if type(part) is ModValue and part.store is not None and part.ctx is ctx:
            stores.append(part._flush())
        else:
            if type(part) in (ModValue, LimbValue):
                part = part._materialise()
            if type(part) is not np.ndarray or part.dtype.kind not in 'Oiu':
                return None
            calls['packed'] += 1
            stores.append(codec.ints_to_limbs(part.astype(object).reshape(-1), ctx))
        rows += shape[0]
    if not any(type(part) is ModValue for part in parts):
        return None
    calls['limb_ops'] += 1
    return ModValue(ctx, backend.concat(ctx, stores), (rows, cols))


import operator as _operator   # noqa: E402

_NUMPY_OPS = {'+': _operator.add, '-': _operator.sub, '*': _operator.mul}
_UFUNC_OPS = {np.add: '+', np.subtract: '-', np.multiply: '*', np.left_shift: '<<'}


def _delegate_mod(name):
    def method(self, *args, **kwargs):
        return getattr(self._materialise(), name)(*args, **kwargs)
    method.__name__ = name
    return method


for _name in ('contains', 'matmul', 'rmatmul', 'imatmul', 'rmod', 'floordiv', 'rfloordiv', 'ifloordiv', 'truediv', 'rtruediv',
              'itruediv', 'rpow', 'ipow', 'abs', 'invert', 'rlshift', 'rshift', 'rrshift', 'irshift', 'iand', 'or', 'ror', 'ior',
              'xor', 'rxor', 'ixor', 'lt', 'le', 'gt', 'ge', 'bool', 'int', 'index', 'float', 'divmod', 'rdivmod', 'str',
              'format'):
    setattr(ModValue, f'__{_name}__', _delegate_mod(f'__{_name}__'))


class _ShiftedBits:
    """`value << shifts` for a value of shape (..., f) and shifts = arange(f) or arange(f-1, -1, -1): summed along the last
    axis by the k_bits_compose kernel (runtime.py:860, 3650-3651, 4415); any other use evaluates the shift on the object
    array."""

    def __init__(self, base, shifts, descending):
        self.base, self.shifts, self.descending = base, shifts, descending

    def sum(self, axis=None, out=None, **kwargs):
        b = self.base
        if b.store is not None and axis in (len(b.shape) - 1, -1) and out is None and not kwargs:
            f = b.shape[-1]                              # the bit positions are the last axis (runtime.py:860, 4415)
            n = b.size // f
            calls['limb_ops'] += 1
            return b._new(backend.bits_compose(b.ctx, b._flush(), n, f, self.descending), b.shape[:-1])
        return self._eval().sum(axis=axis, out=out, **kwargs)

    def _eval(self):
        return self.base._materialise() << self.shifts

    def __array__(self, dtype=None, copy=None):
        return self._eval()

    def __getattr__(self, name):
        if name.startswith('__') and name.endswith('__'):
            raise AttributeError(name)
        return getattr(self._eval(), name)


class _OuterBits:
    """`np.right_shift.outer(c, shifts)` for an opened value c and shifts = arange(l) or arange(l-1, -1, -1): `.T & 1`
    (runtime.py:3660) is the k_bits_decompose kernel; any other use evaluates the shifts on the object array."""

    def __init__(self, base, shifts, descending, transposed):
        self.base, self.shifts, self.descending, self.transposed = base, shifts, descending, transposed

    @property
    def shape(self):
        n, l = self.base.shape[0], self.shifts.shape[0]
        return (l, n) if self.transposed else (n, l)

    @property
    def T(self):
        return _OuterBits(self.base, self.shifts, self.descending, not self.transposed)

    def __and__(self, mask):
        b = self.base
        if b.store is not None and isinstance(mask, (int, np.integer)) and not isinstance(mask, (bool, np.bool_)) and mask == 1:
            n, l = b.shape[0], self.shifts.shape[0]
            calls['limb_ops'] += 1
            bits = backend.bits_decompose(b.ctx, b._flush(), l, self.descending)         # (l, n)
            if self.transposed:
                return b._new(bits, (l, n), True)
            calls['limb_ops'] += 1
            return b._new(backend.transpose(b.ctx, bits, l, n), (n, l), True)
        return self._eval() & mask

    __rand__ = __and__

    def _eval(self):
        a = np.right_shift.outer(self.base._materialise(), self.shifts)
        return a.T if self.transposed else a

    def __array__(self, dtype=None, copy=None):
        return self._eval()

    def __getattr__(self, name):
        if name.startswith('__') and name.endswith('__'):
            raise AttributeError(name)
        return getattr(self._eval(), name)


def _from_wire(modulus, binary, shape, data):
    ctx = context_for(modulus, binary=binary)
    limbs = codec.wire_to_limbs(data, ctx)
    n = 1
    for s in shape:
        n *= s
    if limbs.shape[0] != n:
        raise ValueError('LimbValue: wire data does not hold the announced number of elements')
    return LimbValue(ctx, limbs, shape)


def _from_ints(ints):
    return ints


def as_limb_value(value):
    """LimbValue for a limb-backed LimbValue / ShareRow / ModValue, else None."""
    if type(value) is LimbValue:
        return value if value.store is not None else None
    if type(value) is ModValue:
        return value.limb_value()
    if type(value) is ShareRow and value._ints is None:
        return LimbValue(value.ctx, value.limbs, (len(value),), value._poly)
    return None


# ---------------------------------------------------------------------------------------------
# FiniteFieldArray.value as a property over the original slot (installed by mpyc_b200.install)
# ---------------------------------------------------------------------------------------------

_slot = None            # the member descriptor FiniteFieldArray.__dict__['value'] of the patched finfields module
_lazy_codes = set()     # code objects that may see a LimbValue (see module docstring)
_mod_codes = set()      # code objects that get a ModValue: their arithmetic on raw values runs mod p on the device
_runtime_seen = [False]
local_algebra = True    # hand ModValues to the functions below (install(local_algebra=...))

# Runtime methods whose raw-value algebra is known to be ring operations ending in `Zp.array(...)` (plus `&` / `!= 0` on
# opened, canonical values): sha256 of their source in lschoe/mpyc v0.11.2.  A function whose source differs is not
# whitelisted -- its `.value` reads settle to the reference's object arrays as for every other reader.
_MOD_SOURCES = {
    'np_random_bits': '26b092e612b3091d4ac89237593b4069fa5eee590abcbe21077d13b992bf3868',     # runtime.py:4187-4273
    'np_trunc': 'e6354dbb71d6369d8bef9b065bef0cb65b1dc55631634954792f7945e087e3ad',           # runtime.py:838-872
    'np_sgn': '936eb16da844a41f636537f44306b80fff9aea3834ba205e187e6ba5a24547ad',             # runtime.py:3622-3694
    'np_to_bits': 'fb87bacf263e7c49f50422bba78d6bbf24fb200bc068983a511619da6bd16c1a',         # runtime.py:4391-4441
}


def raw_value(arr):
    """The content of arr's value slot without settling it (arr: a patched FiniteFieldArray, or any stand-in)."""
    if _slot is not None and isinstance(arr, _slot.__objclass__):
        return _slot.__get__(arr)
    return arr.value


def _collect_runtime_codes(finfields_module):
    """runtime._reshare / output / _distribute of the MPyC the patched finfields belongs to, once it is imported."""
    if _runtime_seen[0]:
        return
    pkg = finfields_module.__name__.rsplit('.', 1)[0]
    rt = sys.modules.get(pkg + '.runtime')
    if rt is None or not hasattr(rt, 'Runtime'):
        return
    for name in ('_reshare', 'output', '_distribute'):
        fn = getattr(rt.Runtime, name, None)
        fn = getattr(fn, '__wrapped__', fn)
        if fn is not None and hasattr(fn, '__code__'):
            _lazy_codes.add(fn.__code__)
    st = sys.modules.get(pkg + '.sectypes')
    fn = getattr(getattr(st, 'SecureArray', None), '__init__', None)
    if fn is not None and hasattr(fn, '__code__') and fn.__code__.co_names.count('value') and 'shape' in fn.__code__.co_names:
        _lazy_codes.add(fn.__code__)          # `shape = value.value.shape` (sectypes.py:1019): asks the shape only
    if local_algebra:
        import hashlib
        import inspect
        for name, digest in _MOD_SOURCES.items():
            fn = getattr(rt.Runtime, name, None)
            fn = getattr(fn, '__wrapped__', fn)
            try:
                if hashlib.sha256(inspect.getsource(fn).encode()).hexdigest() == digest:
                    _mod_codes.add(fn.__code__)
            except (OSError, TypeError):
                pass
    _runtime_seen[0] = True


def make_value_property(finfields_module):
    global _slot
    cls = finfields_module.FiniteFieldArray
    _slot = cls.__dict__['value']
    slot_get, slot_set = _slot.__get__, _slot.__set__
    _lazy_codes.clear()
    _mod_codes.clear()
    _runtime_seen[0] = False
    for name in ('reshape', 'ravel', 'flatten', 'copy', '__len__'):
        fn = cls.__dict__.get(name)
        if fn is not None and hasattr(fn, '__code__'):
            _lazy_codes.add(fn.__code__)
    for name in ('shape', 'ndim', 'size'):
        prop = cls.__dict__.get(name)
        fn = getattr(prop, 'fget', None)
        if fn is not None and hasattr(fn, '__code__'):
            _lazy_codes.add(fn.__code__)

    def get(self):
        v = slot_get(self)
        if type(v) is not LimbValue:
            return v
        if v.store is not None:
            if not _runtime_seen[0]:
                _collect_runtime_codes(finfields_module)
            code = sys._getframe(1).f_code
            if code in _lazy_codes:
                return v
            if code in _mod_codes and not v.ctx.binary:
                calls['mod_values'] += 1
                return ModValue(v.ctx, v.store, v.shape, exact=True)
        a = v._materialise()
        slot_set(self, a)           # settled: from here on this array is an ordinary reference array
        return a

    return property(get, slot_set, doc='array of field element values (mpyc_b200: limb-backed until looked at)')


def restore_value_slot(finfields_module):
    global _slot
    if _slot is not None:
        finfields_module.FiniteFieldArray.value = _slot
        _slot = None


# ---------------------------------------------------------------------------------------------
# operator hooks for FiniteFieldArray (installed by mpyc_b200.install)
# ---------------------------------------------------------------------------------------------

_MISS = object()


def _ctx_of(cls):
    try:
        return context_of_field(cls.field)
    except _cabi.UnsupportedFieldError:
        return None


def _operand(ctx, cls, x, want_shape=None):
    """(store, shape) of an array operand as limbs, or None when it should stay with the reference's code."""
    if isinstance(x, cls):
        x = raw_value(x)
    lv = as_limb_value(x)
    if lv is not None:
        return (lv.store, lv.shape) if lv.ctx is ctx else None
    if type(x) is LimbValue or type(x) is ModValue:
        x = x._ints
    if isinstance(x, np.ndarray) and x.dtype == object and x.size >= min_size and not ctx.binary:
        calls['packed'] += 1
        return codec.ints_to_limbs(x.reshape(-1), ctx), x.shape
    return None


def _gf2_mod(a, f):
    """a mod f in GF(2)[X] (integer encodings): an int operand of a binary-field array is the polynomial with that
    encoding (gfpx's int coercion, mpyc/gfpx.py:73-81), reduced when the result is (finfields.py:724)."""
    fb = f.bit_length()
    while a.bit_length() >= fb:
        a ^= f << (a.bit_length() - fb)
    return a


def _scalar(ctx, cls, x):
    if isinstance(x, cls.field):
        x = x.value
    if ctx.binary:
        if type(x).__name__.endswith('Polynomial') or (isinstance(x, (int, np.integer)) and not isinstance(x, bool) and x >= 0):
            return _gf2_mod(int(x), ctx.modulus)
        return None
    if isinstance(x, (int, np.integer)) and not isinstance(x, bool):
        return int(x)
    return None


def _wrap(cls, ctx, store, shape):
    calls['limb_ops'] += 1
    poly = type(cls.field.modulus) if ctx.binary else None
    return cls(LimbValue(ctx, store, shape, poly), check=False)


def binop(self, other, op, reflected=False):
    """self (op) other on limbs when self is limb-backed or large; _MISS -> caller uses the reference's method.
    op: _cabi.OP_ADD / OP_SUB / OP_MUL.  reflected: other - self."""
    cls = type(self)
    ctx = _ctx_of(cls)
    if ctx is None:
        return _MISS
    mine = raw_value(self)
    big_or_limbs = as_limb_value(mine) is not None or (isinstance(mine, np.ndarray) and mine.size >= min_size)
    if not big_or_limbs and not (isinstance(other, cls) and as_limb_value(raw_value(other)) is not None):
        return _MISS
    a = _operand(ctx, cls, self)
    if a is None:
        return _MISS
    s = _scalar(ctx, cls, other)
    if s is not None:
        if reflected:          # s - a = (-a) + s
            neg = backend.neg(ctx, a[0])
            return _wrap(cls, ctx, backend.binop_scalar(ctx, _cabi.OP_ADD, neg, s), a[1])
        return _wrap(cls, ctx, backend.binop_scalar(ctx, op, a[0], s), a[1])
    if not isinstance(other, (cls, np.ndarray, LimbValue)):
        return _MISS
    b = _operand(ctx, cls, other)
    if b is None:
        return _MISS
    if b[1] != a[1]:
        # NumPy's row broadcast between a matrix (R, C) and a vector (C,) runs on k_binop_rows; any other broadcast stays
        # with NumPy
        if ctx.binary:
            return _MISS
        if len(a[1]) == 2 and len(b[1]) == 1 and a[1][1] == b[1][0] and a[1][0] and a[1][1]:
            mat, vec, vec_first = a, b, reflected
        elif len(a[1]) == 1 and len(b[1]) == 2 and b[1][1] == a[1][0] and b[1][0] and b[1][1]:
            mat, vec, vec_first = b, a, not reflected
        else:
            return _MISS
        R, C = mat[1]
        return _wrap(cls, ctx, backend.binop_rows(ctx, op, mat[0], vec[0], R, C, vec_first), (R, C))
    x, y = (b[0], a[0]) if reflected else (a[0], b[0])
    return _wrap(cls, ctx, backend.binop(ctx, op, x, y), a[1])


def getitem(self, key):
    """self[key] for a limb-backed array when the selection is one contiguous run of the C-ordered elements: an int or a
    step-1 slice along axis 0 (what np_prod's halving loop, runtime.py:2198-2204, and row selections use).  The result
    shares the limbs (FiniteFieldArray.__getitem__ returns a no-copy view, finfields.py:1004-1009; limb stores are
    immutable, a later write settles the array that is written to).  _MISS otherwise."""
    cls = type(self)
    lv = as_limb_value(raw_value(self))
    if lv is None:
        return _MISS
    shape = lv.shape
    if not shape:
        return _MISS
    if isinstance(key, tuple):
        if len(key) != 1:
            return _MISS
        key = key[0]
    inner = 1
    for d in shape[1:]:
        inner *= d
    if isinstance(key, slice):
        if key.step not in (None, 1):
            return _MISS
        start, stop, _ = key.indices(shape[0])
        stop = max(stop, start)
        new_shape = (stop - start,) + shape[1:]
    elif isinstance(key, (int, np.integer)) and not isinstance(key, (bool, np.bool_)) and len(shape) >= 2:
        start = int(key) + (shape[0] if key < 0 else 0)
        if not 0 <= start < shape[0]:
            return _MISS                 # NumPy raises the IndexError
        stop, new_shape = start + 1, shape[1:]
    else:
        return _MISS
    store = backend.slice(lv.ctx, lv.store, start * inner, stop * inner)
    return cls(LimbValue(lv.ctx, store, new_shape, lv._poly), check=False)


def concatenate(cls, arrays, axis):
    """np.concatenate(arrays, axis=0) of limb-backed arrays of one field (np_prod's odd step, runtime.py:2202); _MISS otherwise."""
    if axis != 0 or not arrays:
        return _MISS
    lvs = [as_limb_value(raw_value(a)) if isinstance(a, cls) else None for a in arrays]
    if any(v is None or v.ctx is not lvs[0].ctx or v.shape[1:] != lvs[0].shape[1:] or not v.shape for v in lvs):
        return _MISS
    ctx = lvs[0].ctx
    calls['limb_ops'] += 1
    store = backend.concat(ctx, [v.store for v in lvs])
    return cls(LimbValue(ctx, store, (sum(v.shape[0] for v in lvs),) + lvs[0].shape[1:], lvs[0]._poly), check=False)


def equals(self, other, negate):
    """self == other / self != other (FiniteFieldArray.__eq__ / __ne__, finfields.py:1030-1042) for a limb-backed array and
    an integer: a bool ndarray from the k_nonzero kernel (np_is_zero_public's `a == 0`, runtime.py:966); _MISS otherwise."""
    cls = type(self)
    ctx = _ctx_of(cls)
    lv = as_limb_value(raw_value(self))
    if ctx is None or ctx.binary or lv is None or lv.ctx is not ctx or not lv.size:
        return _MISS
    if not isinstance(other, (int, np.integer)) or isinstance(other, (bool, np.bool_)):
        return _MISS
    k = int(other) % ctx.modulus
    store = lv.store if k == 0 else backend.binop_scalar(ctx, _cabi.OP_SUB, lv.store, k)
    calls['limb_ops'] += 1
    mask, _ = backend.nonzero(ctx, store)
    mask = np.asarray(mask, dtype=bool).reshape(lv.shape)
    return mask if negate else ~mask


def shift(self, other, right):
    """self << n / self >> n for a limb-backed array and an integer n >= 0; _MISS otherwise."""
    cls = type(self)
    ctx = _ctx_of(cls)
    if ctx is None or ctx.binary or as_limb_value(raw_value(self)) is None:
        return _MISS
    if not isinstance(other, (int, np.integer)) or isinstance(other, (bool, np.bool_)) or other < 0:
        return _MISS
    k = pow(2, int(other), ctx.modulus)
    if right:
        k = pow(k, -1, ctx.modulus)
    a = _operand(ctx, cls, self)
    return _wrap(cls, ctx, backend.binop_scalar(ctx, _cabi.OP_MUL, a[0], k), a[1])


def negate(self):
    cls = type(self)
    ctx = _ctx_of(cls)
    if ctx is None or as_limb_value(raw_value(self)) is None:
        return _MISS
    a = _operand(ctx, cls, self)
    return _wrap(cls, ctx, backend.neg(ctx, a[0]), a[1])


def matmul(self, other, reflected=False):
    """self @ other (reflected: other @ self) for 1-D / 2-D operands through K1c."""
    cls = type(self)
    ctx = _ctx_of(cls)
    if ctx is None:
        return _MISS
    A, B = (other, self) if reflected else (self, other)

    def shape_of(x):
        v = raw_value(x) if isinstance(x, cls) else x
        return getattr(v, 'shape', None)
    sa, sb = shape_of(A), shape_of(B)
    if sa is None or sb is None or not (1 <= len(sa) <= 2 and 1 <= len(sb) <= 2):
        return _MISS
    r, k = (1, sa[0]) if len(sa) == 1 else sa
    k2, c = (sb[0], 1) if len(sb) == 1 else sb
    if k != k2:
        return _MISS                      # let NumPy raise its own error
    limbish = any(isinstance(x, cls) and as_limb_value(raw_value(x)) is not None for x in (A, B))
    if not limbish and r * k * c < min_size * 8:
        return _MISS
    saved = min_size
    try:
        globals()['min_size'] = 0         # both operands are packed whatever their size
        a = _operand(ctx, cls, A)
        b = _operand(ctx, cls, B)
    finally:
        globals()['min_size'] = saved
    if a is None or b is None:
        return _MISS
    out = backend.matmul(ctx, a[0], b[0], r, k, c)
    if len(sa) == 1 and len(sb) == 1:     # inner product: a field element (finfields.py:1134)
        calls['limb_ops'] += 1
        poly = type(cls.field.modulus) if ctx.binary else None
        return cls.field(LimbValue(ctx, out, (1,), poly)._materialise()[0])
    shape = (c,) if len(sa) == 1 else ((r,) if len(sb) == 1 else (r, c))
    return _wrap(cls, ctx, out, shape)


def conv2d(x, W, b):
    """The local step of np_cnnmnist's convolvetensor (demos/np_cnnmnist.py:69-82) as one kernel call: x (k, r, m, n),
    W (v, r, s, s), b (v) -- field arrays of one covered prime field, as `mpc.gather` returns them -- to the field array
    Y (k, v, m, n) with Y[i, j] = b[j] + sum_l correlate2d(x[i, l], W[j, l], 'same'), limb-backed, ready for
    `mpc._reshare(Y)`.  The demo is caller code that install() cannot reach; INTEGRATION.md section 5 shows the edit."""
    cls = type(x)
    ctx = _ctx_of(cls)
    if ctx is None or ctx.binary or type(W) is not cls or type(b) is not cls:
        raise _cabi.UnsupportedFieldError('conv2d: operands must be arrays of one covered prime field')
    sx, sw, sb = (getattr(raw_value(a), 'shape', None) for a in (x, W, b))
    if len(sx) != 4 or len(sw) != 4 or len(sb) != 1 or sx[1] != sw[1] or sw[2] != sw[3] or sb[0] != sw[0]:
        raise ValueError(f'conv2d: shapes {sx}, {sw}, {sb} do not fit (k,r,m,n), (v,r,s,s), (v,)')
    saved = min_size
    try:
        globals()['min_size'] = 0
        ops = [_operand(ctx, cls, a) for a in (x, W, b)]
    finally:
        globals()['min_size'] = saved
    if any(o is None for o in ops):
        raise _cabi.UnsupportedFieldError('conv2d: operands must be limb-backed or object arrays of the field')
    k, r, m, n = sx
    v, s = sw[0], sw[2]
    out = backend.conv2d(ctx, ops[0][0], ops[1][0], ops[2][0], k, r, m, n, v, s)
    return _wrap(cls, ctx, out, (k, v, m, n))
