"""Field contexts: the device-side counterpart of finfields.GF(p) / GF(2^8).

Reference: mpyc/finfields.py:23-42 (GF), :347-363 (pGF) build one element class per modulus and
cache it; here one FieldContext per modulus owns the C-ABI handle (reduction family, Montgomery
constants, cached Vandermonde/Lagrange tables on the device).
"""
import ctypes
import threading

from mpyc_b200 import _cabi
from mpyc_b200._cabi import lib, check

_lock = threading.Lock()
_contexts = {}


class FieldContext:
    """Handle on one finite field: prime p (odd, < 2^256) or GF(2^8) given by its modulus polynomial."""

    def __init__(self, modulus, binary=False):
        self.modulus = int(modulus)
        self.binary = bool(binary)
        handle = ctypes.c_void_p()
        if binary:
            if self.modulus.bit_length() != 9:
                raise _cabi.UnsupportedFieldError('only GF(2^8) binary fields are supported')
            check(lib.mpyc_b200_field_create_gf256(self.modulus, ctypes.byref(handle)))
            self.order = 256
        else:
            if self.modulus < 3 or self.modulus % 2 == 0:
                raise _cabi.UnsupportedFieldError('modulus must be an odd prime >= 3')
            nl = (self.modulus.bit_length() + 63) // 64
            if nl > _cabi.MAX_LIMBS:
                raise _cabi.UnsupportedFieldError('modulus wider than 256 bits')
            check(lib.mpyc_b200_field_create(_cabi.u64_array(_cabi.int_to_limbs(self.modulus, nl)), nl,
                                             ctypes.byref(handle)))
            self.order = self.modulus
        self.handle = handle
        nlimbs, kind, bits, eb = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_size_t()
        check(lib.mpyc_b200_field_info(handle, ctypes.byref(nlimbs), ctypes.byref(kind), ctypes.byref(bits),
                                       ctypes.byref(eb)))
        self.nlimbs = nlimbs.value          # 0 for GF(2^8)
        self.kind = kind.value
        self.bits = bits.value
        self.elem_bytes = eb.value
        self.byte_length = (self.order.bit_length() + 7) >> 3 if not binary else 1   # finfields.py:359

    def __del__(self):
        h, self.handle = getattr(self, 'handle', None), None
        if h:
            try:
                lib.mpyc_b200_field_destroy(h)
            except Exception:   # interpreter shutdown
                pass

    def __repr__(self):
        kind = {0: 'generic-montgomery', 1: 'pseudo-mersenne-aligned', 2: 'pseudo-mersenne-shift', 3: 'gf256'}[self.kind]
        return f'FieldContext({hex(self.modulus)}, {kind}, limbs={self.nlimbs})'

    # ---- host-only helpers ---------------------------------------------------------------------
    def recombination_vector(self, xs, x_rs):
        """Lagrange coefficients lambda[r][i] as Python ints (role of thresha._recombination_vector,
        mpyc/thresha.py:67-85).  Runs on the host inside the library; no GPU needed."""
        xs, x_rs = [int(x) for x in xs], [int(x) for x in x_rs]
        k, width = len(xs), len(x_rs)
        nl = max(self.nlimbs, 1)
        out = (ctypes.c_uint64 * (k * width * nl))()
        check(lib.mpyc_b200_recombination_vector(self.handle, _cabi.i64_array(xs), k, _cabi.i64_array(x_rs), width, out))
        return [[_cabi.limbs_to_int(out[(r * k + i) * nl:(r * k + i + 1) * nl]) for i in range(k)]
                for r in range(width)]

    def scalar_limbs(self, value):
        return _cabi.u64_array(_cabi.int_to_limbs(int(value) % self.order if not self.binary else int(value) & 0xFF,
                                                  max(self.nlimbs, 1)))


def context_for(modulus, binary=False):
    """Cached FieldContext (one per modulus, like functools.cache on pGF, finfields.py:347)."""
    key = (int(modulus), bool(binary))
    with _lock:
        ctx = _contexts.get(key)
        if ctx is None:
            ctx = _contexts[key] = FieldContext(*key)
        return ctx


def context_of_field(field):
    """FieldContext for an MPyC field class (finfields.GF(...)): prime fields by .modulus int,
    GF(2^8) by a gfpx.BinaryPolynomial modulus (integer encoding via int())."""
    modulus = field.modulus
    if getattr(field, 'characteristic', None) == 2 and getattr(field, 'ext_deg', None) == 8:
        return context_for(int(modulus), binary=True)
    if isinstance(modulus, int) and getattr(field, 'ext_deg', 1) == 1:
        return context_for(modulus)
    raise _cabi.UnsupportedFieldError(f'field {getattr(field, "__name__", field)} is not supported by mpyc_b200')
