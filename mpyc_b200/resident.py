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
calls = {'limb_ops': 0, 'materialised': 0, 'packed': 0}    # counters (tests, profiling)


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
    """LimbValue for a limb-backed LimbValue / ShareRow, else None."""
    if type(value) is LimbValue:
        return value if value.store is not None else None
    if type(value) is ShareRow and value._ints is None:
        return LimbValue(value.ctx, value.limbs, (len(value),), value._poly)
    return None


# ---------------------------------------------------------------------------------------------
# FiniteFieldArray.value as a property over the original slot (installed by mpyc_b200.install)
# ---------------------------------------------------------------------------------------------

_slot = None            # the member descriptor FiniteFieldArray.__dict__['value'] of the patched finfields module
_lazy_codes = set()     # code objects that may see a LimbValue (see module docstring)
_runtime_seen = [False]


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
    _runtime_seen[0] = True


def make_value_property(finfields_module):
    global _slot
    cls = finfields_module.FiniteFieldArray
    _slot = cls.__dict__['value']
    slot_get, slot_set = _slot.__get__, _slot.__set__
    _lazy_codes.clear()
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
            if sys._getframe(1).f_code in _lazy_codes:
                return v
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
    if type(x) is LimbValue:
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
    if b is None or b[1] != a[1]:      # broadcasting stays with NumPy
        return _MISS
    x, y = (b[0], a[0]) if reflected else (a[0], b[0])
    return _wrap(cls, ctx, backend.binop(ctx, op, x, y), a[1])


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
