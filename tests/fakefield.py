"""Minimal stand-ins for MPyC's field classes (finfields.GF(p), GF(2^8)) for tests that run where the
reference is not importable (the GPU box).  Only the attributes the drop-in functions touch:
field.modulus / .order / .characteristic / .ext_deg / .array, field(value).value, array(value, check).value."""
import numpy as np


class Poly(int):
    """Stand-in for gfpx.BinaryPolynomial: an int subclass (int(poly) is the bit encoding)."""
    __slots__ = ()


def make_prime_field(p):
    class Elt:
        __slots__ = ('value',)
        modulus = p
        order = p
        characteristic = p
        ext_deg = 1

        def __init__(self, value):
            self.value = int(value) % p

        def __eq__(self, other):
            return isinstance(other, Elt) and other.value == self.value

        def __hash__(self):
            return hash(self.value)

    class Arr:
        __slots__ = ('value',)
        field = Elt

        def __init__(self, value, check=True, copy=False):
            value = np.array(value, dtype=object)
            if check:
                value %= p
            self.value = value

    Elt.array = Arr
    Elt.__name__ = f'GF({p})'
    return Elt


def make_gf256(poly=283):
    class Elt:
        __slots__ = ('value',)
        modulus = Poly(poly)
        order = 256
        characteristic = 2
        ext_deg = 8

        def __init__(self, value):
            self.value = Poly(int(value))

        def __eq__(self, other):
            return isinstance(other, Elt) and int(other.value) == int(self.value)

        def __hash__(self):
            return hash(int(self.value))

    class Arr:
        __slots__ = ('value',)
        field = Elt

        def __init__(self, value, check=True, copy=False):
            self.value = np.array(value, dtype=object)

    Elt.array = Arr
    Elt.__name__ = 'GF(2^8)'
    return Elt
