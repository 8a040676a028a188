"""Limb wire format: shares that travel between parties without ever becoming Python ints.

The reference pickles each party's row of shares -- a NumPy dtype=object array of Python ints -- to send it
(`marshal = pickle.dumps`, mpyc/runtime.py:484-485,571-572,655-656: 34.5 MB per 10^6 256-bit elements) and
the receiver unpickles it back into ints before recombining.  With `mpyc_b200.thresha.limb_wire = True`
(install(limb_wire=True)) np_random_split hands runtime.py `ShareRow` objects instead: a row is the limb
buffer the kernel wrote, it pickles as the fixed-width little-endian byte string of
FiniteFieldElement.to_bytes (mpyc/finfields.py:91-102; `byte_length` bytes per element, the --mix32-64bit
wire layout) and unpickles into a ShareRow again, which np_recombine feeds to the GPU as is.  Only the values
a party actually keeps (its recombined result, its own input share) are turned into ints, once.

runtime.py itself is untouched: it only iterates over the rows, pickles them and passes what it unpickled
back to np_recombine / field.array.  A ShareRow still behaves like the object array it replaces for every
other consumer (len, iteration, indexing, reshape(-1), np.array(row, dtype=object) via __array__), at the
cost of the conversion it otherwise avoids.  Both ends of a connection need mpyc_b200 importable (the
pickle refers to mpyc_b200.wire._row_from_wire), which is why the format is opt-in.
"""
import numpy as np

from mpyc_b200 import codec
from mpyc_b200.field import context_for


def _poly_type():
    """gfpx polynomial type over GF(2) for rows that were unpickled without a field at hand."""
    try:
        from mpyc import gfpx
        return gfpx.GFpX(2)
    except ImportError:
        return int


class ShareRow:
    """n field elements as a limb buffer: uint64 (n, L) for prime fields, uint8 (n,) for GF(2^8)."""

    __slots__ = ('ctx', 'limbs', '_poly', '_ints')
    __array_priority__ = 0.0

    def __init__(self, ctx, limbs, poly_type=None):
        self.ctx = ctx
        self.limbs = limbs
        self._poly = poly_type
        self._ints = None

    # ---- what runtime.py and np_recombine touch ------------------------------------------------------
    def __len__(self):
        return self.limbs.shape[0]

    @property
    def shape(self):
        return (self.limbs.shape[0],)

    ndim = 1
    dtype = np.dtype(object)

    @property
    def size(self):
        return self.limbs.shape[0]

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        if shape in ((-1,), (len(self),)):
            return self
        return self.__array__().reshape(*shape)

    def __reduce__(self):
        return _row_from_wire, (self.ctx.modulus, self.ctx.binary, len(self), codec.limbs_to_wire(self.limbs, self.ctx))

    # ---- everything else: materialise the ints (once) ----------------------------------------------------
    def __array__(self, dtype=None, copy=None):
        if self._ints is None:
            vals = codec.limbs_to_ints(self.limbs, self.ctx)
            if self.ctx.binary:
                tp = self._poly or _poly_type()
                out = np.empty(len(vals), dtype=object)
                out[:] = [tp(int(v)) for v in vals]
                vals = out
            self._ints = vals
        return self._ints.copy() if copy else self._ints

    def tolist(self):
        return self.__array__().tolist()

    def __iter__(self):
        return iter(self.__array__())

    def __getitem__(self, key):
        return self.__array__()[key]

    def __repr__(self):
        return f'ShareRow(n={len(self)}, {self.ctx!r})'


def _row_from_wire(modulus, binary, n, data):
    ctx = context_for(modulus, binary=binary)
    limbs = codec.wire_to_limbs(data, ctx)
    if limbs.shape[0] != n:
        raise ValueError('ShareRow: wire data does not hold the announced number of elements')
    return ShareRow(ctx, limbs)


class ShareRows:
    """The (m, n) result of np_random_split in limb form: iterating / indexing yields one ShareRow per party."""

    __slots__ = ('ctx', 'limbs', '_poly')

    def __init__(self, ctx, limbs, poly_type=None):
        self.ctx = ctx
        self.limbs = limbs          # uint64 (m, n, L) / uint8 (m, n)
        self._poly = poly_type

    def __len__(self):
        return self.limbs.shape[0]

    @property
    def shape(self):
        return self.limbs.shape[:2]

    ndim = 2
    dtype = np.dtype(object)

    def __getitem__(self, i):
        if isinstance(i, (int, np.integer)):
            return ShareRow(self.ctx, self.limbs[i], self._poly)
        return self.__array__()[i]

    def __iter__(self):
        return (ShareRow(self.ctx, self.limbs[i], self._poly) for i in range(len(self)))

    def __array__(self, dtype=None, copy=None):
        m, n = self.shape
        out = np.empty((m, n), dtype=object)
        for i in range(m):
            out[i] = ShareRow(self.ctx, self.limbs[i], self._poly).__array__()
        return out

    def __repr__(self):
        return f'ShareRows(m={self.shape[0]}, n={self.shape[1]}, {self.ctx!r})'
