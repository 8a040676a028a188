"""The shared library builds, loads, and exports exactly the C ABI that include/mpyc_b200.h declares
(no compute calls: this runs without a GPU).  Host-only entry points are exercised for real."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'mpyc_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(mpyc_b200_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from mpyc_b200 import _cabi
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(_cabi.lib, name), f'{name} declared in include/mpyc_b200.h but not exported'
    assert sorted(_cabi.EXPORTED) == names, 'ctypes binding and header disagree'
    assert _cabi.lib.mpyc_b200_version() >= 100


def test_field_classification_and_host_only_entry_points():
    import mpyc_b200
    from oracle import shamir_oracle as orc
    kinds = {2**61 - 1: 2, 2**64 - 189: 1, 2**69 - 93: 2, 2**127 - 1: 2, 2**128 - 173: 1, 2**256 - 189: 1,
             9409569905028393239: 0, 19: 0, 2**255 - 19: 2, 2**64 - 2**32 + 1: 0}
    for p, kind in kinds.items():
        ctx = mpyc_b200.context_for(p)
        assert (ctx.kind, ctx.nlimbs, ctx.bits) == (kind, (p.bit_length() + 63) // 64, p.bit_length())
        F = orc.field_of(p)
        for xs in ([1], [1, 2, 3], [2, 3, 5, 7, 11], list(range(1, 18))):
            assert ctx.recombination_vector(xs, [0, 20]) == [orc.recombination_vector(F, xs, r) for r in (0, 20)]
    g = mpyc_b200.context_for(283, binary=True)
    assert g.kind == 3 and g.elem_bytes == 1
    Fb = orc.field_of(283, binary=True)
    assert g.recombination_vector([1, 2, 3, 4, 5], [0, 9]) == [orc.recombination_vector(Fb, [1, 2, 3, 4, 5], r) for r in (0, 9)]
    with pytest.raises(ZeroDivisionError):
        mpyc_b200.context_for(2**61 - 1).recombination_vector([3, 3], [0])
    with pytest.raises(mpyc_b200.UnsupportedFieldError):
        mpyc_b200.context_for(2**256 + 297)
    with pytest.raises(mpyc_b200.UnsupportedFieldError):
        mpyc_b200.context_for(0x11d ^ 0x1, binary=True)   # reducible degree-8 polynomial (divisible by x)


def test_compute_fails_loudly_without_gpu():
    """No CPU fallback: with no CUDA device the compute entry points return an error, never a result."""
    import numpy as np
    torch = pytest.importorskip('torch')
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    import mpyc_b200
    from mpyc_b200 import thresha
    import fakefield
    F = fakefield.make_prime_field(2**61 - 1)
    with pytest.raises(RuntimeError):
        thresha.np_random_split(F, np.array([1, 2, 3], dtype=object), 1, 3)
    with pytest.raises(RuntimeError):
        thresha.np_recombine(F, [(1, np.array([1], dtype=object)), (2, np.array([2], dtype=object))])


def test_codec_roundtrip():
    import numpy as np
    import mpyc_b200
    from mpyc_b200 import codec
    for p in (2**61 - 1, 2**69 - 93, 2**128 - 173, 2**256 - 189, 101):
        ctx = mpyc_b200.context_for(p)
        vals = [0, 1, p - 1, p // 2, -1, p, p + 5, 3 * p + 2]
        limbs = codec.ints_to_limbs(vals, ctx)
        assert limbs.shape == (len(vals), ctx.nlimbs)
        assert codec.limbs_to_ints(limbs, ctx).tolist() == [v % p for v in vals]
        wire = codec.limbs_to_wire(limbs, ctx)
        assert len(wire) == len(vals) * ctx.byte_length
        assert wire == b''.join((v % p).to_bytes(ctx.byte_length, 'little') for v in vals)
        assert np.array_equal(codec.wire_to_limbs(wire, ctx), limbs)


def test_argument_errors_and_no_gpu_errors_of_the_newer_entry_points():
    """Bad arguments are rejected before any CUDA call (ValueError / TypeError, also without a GPU); with valid arguments
    and no CUDA device every compute / peer entry point fails with MPYC_B200_ECUDA (RuntimeError) -- PRSS included:
    the library's host SHAKE128 never produces a result on its own."""
    import numpy as np
    import mpyc_b200
    from mpyc_b200 import _cabi, thresha
    from mpyc_b200._cabi import lib, check
    import fakefield
    torch = pytest.importorskip('torch')
    ctx = mpyc_b200.context_for(2**61 - 1)
    one = _cabi.u64_array([1])
    out = np.zeros((4, 1), dtype=np.uint64)
    with pytest.raises(ValueError):      # nsub = 0
        check(lib.mpyc_b200_prss_host(ctx.handle, b'k' * 16, 16, b'u', 1, 0, 1, 24, 0, one, one, out.ctypes.data, 4, 0, 0))
    with pytest.raises(ValueError):      # null row table
        check(lib.mpyc_b200_shamir_split_generate_rows(ctx.handle, out.ctypes.data, None, 4, 1, 3,
                                                       (ctypes.c_uint8 * 32)(), 0, None))
    with pytest.raises(ValueError):      # t >= m
        check(lib.mpyc_b200_shamir_split_generate_host(ctx.handle, out.ctypes.data, out.ctypes.data, 4, 4, 3, 3,
                                                       (ctypes.c_uint8 * 32)(), 0, 0))
    if torch.cuda.is_available():
        return
    F = fakefield.make_prime_field(2**61 - 1)
    prfs = {(0, 1): thresha.PRF(b'k' * 16, 2**61 - 1), (0, 2): thresha.PRF(b'j' * 16, 2**61 - 1)}
    with pytest.raises(RuntimeError):
        thresha.np_pseudorandom_share(F, 3, 0, prfs, b'uci', 5)
    ptr, handle = ctypes.c_void_p(), (ctypes.c_uint8 * 64)()
    with pytest.raises(RuntimeError):
        check(lib.mpyc_b200_peer_alloc(1024, ctypes.byref(ptr), handle))
