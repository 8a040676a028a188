"""install(): put the B200 engine behind an imported MPyC without touching MPyC's code.

mpyc/runtime.py looks the secret-sharing functions up as attributes of the `mpyc.thresha` module at
call time (runtime.py:478-485,565-572,647-656,963,982,1280,4057,4099,4219,4248); assigning the drop-in
functions of mpyc_b200.thresha over those attributes makes `_distribute`, `output`, `_reshare` and the
PRSS helpers run on the GPU while runtime.py, sectypes.py and the demos stay byte-for-byte unchanged.

Fields the engine does not cover (GF(2), GF(p^n) with n > 1, binary fields other than GF(2^8), primes
wider than 256 bits) are NOT intercepted: for those the wrapper hands the call to the reference's own
function, exactly as if install() had not been called (strict=True raises UnsupportedFieldError
instead).  The same holds per call: an argument outside the kernels' range on a covered field (e.g. a PRF bound
above 2^256) goes to the reference's function.  That hand-over is the only one: a covered call never runs on a CPU
restatement of the kernels, and without a CUDA device it raises.
"""
import functools

from mpyc_b200 import thresha as engine
from mpyc_b200._cabi import UnsupportedFieldError
from mpyc_b200.field import context_of_field

_NAMES = ('random_split', 'recombine', 'np_random_split', 'np_recombine',
          'pseudorandom_share', 'pseudorandom_share_zero', 'np_pseudorandom_share', 'np_pseudorandom_share_0',
          '_recombination_vector', '_f_S_i')
_saved = {}


def _covered(field):
    try:
        context_of_field(field)
        return True
    except UnsupportedFieldError:
        return False


def _batch_size(name, args):
    """Number of elements a call works on (for the min_size routing of install())."""
    try:
        if name in ('random_split', 'np_random_split'):
            return len(args[0])
        if name in ('recombine', 'np_recombine'):
            return len(args[0][0][1])
        return int(args[-1])            # PRSS functions: (m, i, prfs, uci, n)
    except Exception:   # noqa: BLE001
        return None


def _wrap(name, ours, theirs, strict, min_size=0):
    @functools.wraps(theirs)
    def call(field, *args, **kwargs):
        if _covered(field):
            if min_size:
                n = _batch_size(name, args)
                if n is not None and n < min_size:
                    return theirs(field, *args, **kwargs)     # the reference's own code, as without install()
            if strict:
                return ours(field, *args, **kwargs)
            try:
                return ours(field, *args, **kwargs)
            except UnsupportedFieldError:
                # coverage is decided per CALL, not per field: a covered field with an argument outside the kernels'
                # range (a PRF bound above 2^256, more than 64 recombination points, a constant table beyond shared
                # memory ...) is computed by the reference's own function, exactly as without install() -- never
                # raise where the reference computes.  These functions are pure, so nothing is half-done.
                return theirs(field, *args, **kwargs)
        if strict:
            raise UnsupportedFieldError(f'mpyc_b200 does not cover field {getattr(field, "__name__", field)}')
        return theirs(field, *args, **kwargs)
    call.__mpyc_b200__ = True
    return call


def _install_finfields(finfields_module, min_size):
    """PrimeFieldArray._reciprocal/_pow/_sqrt/_is_sqr -> batched kernels for covered primes and arrays of
    at least `min_size` elements (below that the per-call overhead of a launch outweighs the gain)."""
    from mpyc_b200 import finfields as ours
    from mpyc_b200 import resident as _res
    cls = finfields_module.PrimeFieldArray
    orig = {name: cls.__dict__[name] for name in ('_reciprocal', '_pow', '_sqrt', '_is_sqr')}
    _saved_ff.update({'cls': cls, **orig})

    def use_gpu(klass, a):
        import numpy as np
        return np.size(a) >= min_size and klass.field.modulus % 2 == 1 and klass.field.modulus.bit_length() <= 256

    def _reciprocal(klass, a):
        return ours.reciprocal(klass, a) if use_gpu(klass, a) else orig['_reciprocal'].__func__(klass, a)

    def _pow(klass, a, b):
        theirs = lambda x, y: orig['_pow'].__func__(klass, x, y)   # noqa: E731
        return ours.power(klass, a, b, _fallback=theirs) if use_gpu(klass, a) else theirs(a, b)

    def _sqrt(klass, a, INV=False):
        theirs = lambda x, INV=False: orig['_sqrt'].__func__(klass, x, INV=INV)   # noqa: E731
        if type(a) is _res.ModValue:
            # np_random_bits: `field.array._sqrt(r2, INV=True)` on an opened value (runtime.py:4265) stays on the device
            if a.store is not None and a.exact and klass.field.modulus & 3 == 3:
                return a.sqrt(INV=INV)
            a = a._materialise()
        return ours.sqrt(klass, a, INV=INV, _fallback=theirs) if use_gpu(klass, a) else theirs(a, INV=INV)

    def _is_sqr(klass, a):
        return ours.is_sqr(klass, a) if use_gpu(klass, a) else orig['_is_sqr'].__func__(klass, a)

    cls._reciprocal = classmethod(_reciprocal)
    cls._pow = classmethod(_pow)
    cls._sqrt = classmethod(_sqrt)
    cls._is_sqr = classmethod(_is_sqr)


_saved_ff = {}
_saved_ops = {}


def _install_operators(finfields_module, min_size):
    """FiniteFieldArray.__init__ and + - * neg @ (mpyc/finfields.py:717-725,1056-1146): limb-backed operands, and plain
    object arrays of at least `min_size` elements, are combined by the K1 / K1c kernels (mpyc_b200.resident); everything
    else -- scalars with small arrays, broadcasting, uncovered fields -- runs the reference's own method unchanged."""
    from mpyc_b200 import _cabi, resident
    resident.min_size = int(min_size)
    cls = finfields_module.FiniteFieldArray
    names = ('__init__', '__add__', '__radd__', '__sub__', '__rsub__', '__mul__', '__rmul__', '__neg__', '__matmul__',
             '__rmatmul__', '__lshift__', '__rshift__', '__ilshift__', '__irshift__', '__getitem__', '__array_function__', '__eq__', '__ne__')
    orig = {name: cls.__dict__[name] for name in names}
    _saved_ops.update({'cls': cls, 'module': finfields_module, **orig})
    MISS = resident._MISS
    orig_hash = cls.__hash__          # assigning __eq__ on a class does not touch __hash__, kept explicit for clarity
    value_property = resident.make_value_property(finfields_module)
    slot_set = resident._slot.__set__

    def __init__(self, value, check=True, copy=False):
        lv = resident.as_limb_value(value)
        if lv is not None:
            slot_set(self, lv)         # canonical residues by construction: nothing to check, nothing to copy
            return
        if type(value) is resident.LimbValue or type(value) is resident.ModValue:
            value = value._ints        # already an object array (canonical residues)
        orig['__init__'](self, value, check=check, copy=copy)

    def make(op, name, reflected=False):
        def method(self, other):
            r = resident.binop(self, other, op, reflected)
            return orig[name](self, other) if r is MISS else r
        method.__name__ = name
        method.__doc__ = orig[name].__doc__
        return method

    def __neg__(self):
        r = resident.negate(self)
        return orig['__neg__'](self) if r is MISS else r

    def __matmul__(self, other):
        r = resident.matmul(self, other)
        return orig['__matmul__'](self, other) if r is MISS else r

    def __rmatmul__(self, other):
        r = resident.matmul(self, other, reflected=True)
        return orig['__rmatmul__'](self, other) if r is MISS else r

    def make_shift(name, right, inplace):
        # `a << n` = a * 2^n, `a >> n` = a * (2^n)^-1 for an integer n (finfields.py:1227-1271); limb-backed arrays only
        def method(self, other):
            r = resident.shift(self, other, right)
            if r is MISS:
                return orig[name](self, other)
            if inplace:
                slot_set(self, resident.raw_value(r))
                return self
            return r
        method.__name__ = name
        method.__doc__ = orig[name].__doc__
        return method

    def __getitem__(self, key):
        r = resident.getitem(self, key)
        return orig['__getitem__'](self, key) if r is MISS else r

    def __array_function__(self, func, types, args, kwargs):
        import numpy as np
        if func is np.concatenate and isinstance(self, cls) and len(args) == 1 and set(kwargs) <= {'axis'}:
            r = resident.concatenate(type(self), list(args[0]), kwargs.get('axis', 0))
            if r is not MISS:
                return r
        return orig['__array_function__'](self, func, types, args, kwargs)

    def __eq__(self, other):
        r = resident.equals(self, other, False)
        return orig['__eq__'](self, other) if r is MISS else r

    def __ne__(self, other):
        r = resident.equals(self, other, True)
        return orig['__ne__'](self, other) if r is MISS else r

    cls.__eq__ = __eq__
    cls.__ne__ = __ne__
    cls.__hash__ = orig_hash
    cls.__getitem__ = __getitem__
    cls.__array_function__ = __array_function__
    cls.__lshift__ = make_shift('__lshift__', False, False)
    cls.__rshift__ = make_shift('__rshift__', True, False)
    cls.__ilshift__ = make_shift('__ilshift__', False, True)
    cls.__irshift__ = make_shift('__irshift__', True, True)
    cls.value = value_property
    cls.__init__ = __init__
    cls.__add__ = make(_cabi.OP_ADD, '__add__')
    cls.__radd__ = make(_cabi.OP_ADD, '__radd__')
    cls.__sub__ = make(_cabi.OP_SUB, '__sub__')
    cls.__rsub__ = make(_cabi.OP_SUB, '__rsub__', reflected=True)
    cls.__mul__ = make(_cabi.OP_MUL, '__mul__')
    cls.__rmul__ = make(_cabi.OP_MUL, '__rmul__')
    cls.__neg__ = __neg__
    cls.__matmul__ = __matmul__
    cls.__rmatmul__ = __rmatmul__


def install(thresha_module=None, strict=False, device=0, finfields_module=None, finfields_min_size=256,
            limb_wire=False, min_size=0, operators=False, operators_min_size=1024, resident=False, local_algebra=True):
    """Patch `mpyc.thresha` (or the module passed in); with finfields_module also the batched
    inverse/pow/sqrt/is_sqr of PrimeFieldArray.  limb_wire=True: shares travel between parties as limb
    buffers (mpyc_b200.wire; every party must run mpyc_b200).  min_size > 0: calls on fewer elements are left to
    the reference's own functions -- a GPU round trip costs ~35 us per call, which the reference beats below a few
    dozen 64-bit elements; the default 0 sends every covered call to the GPU.
    operators=True: FiniteFieldArray's + - * neg @ (finfields.py:1056-1146) run on the K1 / K1c kernels for operands
    of at least operators_min_size elements (mpyc_b200.resident).  resident=True (implies operators and limb_wire):
    recombined / pseudorandom shares stay limb-backed in HBM between protocol steps; Python ints are created only
    where a value is actually looked at (input, output, raw-value arithmetic).  local_algebra (with resident): the raw-value
    arithmetic of Runtime.np_random_bits / np_trunc / np_sgn runs mod p on the device as well (resident.ModValue, K6 kernels).
    Returns the list of patched names."""
    if thresha_module is None:
        import mpyc.thresha as thresha_module
    if _saved:
        uninstall()
    if finfields_module is not None:
        _install_finfields(finfields_module, finfields_min_size)
    if operators or resident:
        if finfields_module is None:
            import mpyc.finfields as finfields_module
        _install_operators(finfields_module, operators_min_size)
    from mpyc_b200 import resident as res
    res.resident = bool(resident)
    res.local_algebra = bool(local_algebra and resident)
    res.backend.device = device
    engine.device = device
    engine.limb_wire = bool(limb_wire or resident)
    for name in _NAMES:
        theirs = getattr(thresha_module, name)
        _saved[name] = (thresha_module, theirs)
        if name in ('_recombination_vector', '_f_S_i'):
            # keep the reference's functools.cache behaviour: results are cached per argument tuple
            ours = functools.cache(getattr(engine, name))
        else:
            ours = getattr(engine, name)
        routed = min_size if name not in ('_recombination_vector', '_f_S_i') else 0
        setattr(thresha_module, name, _wrap(name, ours, theirs, strict, routed))
    return list(_NAMES)


def uninstall():
    for name, (module, theirs) in list(_saved.items()):
        setattr(module, name, theirs)
    _saved.clear()
    engine.limb_wire = False
    from mpyc_b200 import resident as res
    res.resident = False
    if _saved_ops:
        cls = _saved_ops.pop('cls')
        res.restore_value_slot(_saved_ops.pop('module'))
        for name, member in _saved_ops.items():
            setattr(cls, name, member)
        _saved_ops.clear()
    if _saved_ff:
        cls = _saved_ff.pop('cls')
        for name, member in _saved_ff.items():
            setattr(cls, name, member)
        _saved_ff.clear()
