"""mpyc_b200.resident.ModValue on its own (no reference, no GPU; the oracle stands in for the kernels): the raw-value
expressions of np_random_bits / np_trunc / np_sgn (mpyc/runtime.py:856-872, 3644-3658, 4252-4271), evaluated on
ModValues, give the field array the same expressions give on NumPy object arrays -- and operations that are not ring
operations are refused on modular intermediates."""
import numpy as np
import pytest

import mpyc_b200
from mpyc_b200 import codec, resident
from mpyc_b200.resident import ModValue, LimbValue
import oracle_device
from oracle import shamir_oracle as orc

PRIMES = [2**61 - 1, 2**64 - 189, 9409569905028393239, 2**128 - 173, 2**256 - 189]


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except ImportError:
        return False


@pytest.fixture(autouse=True, params=['oracle', pytest.param('cuda', marks=pytest.mark.gpu)])
def backend(request, monkeypatch):
    """'oracle': tests/oracle_device.py answers the kernels (CPU suite); 'cuda' (-m gpu): the real K1 / K6 kernels."""
    if request.param == 'cuda':
        if not _has_cuda():
            pytest.skip('no CUDA device')
        monkeypatch.setattr(resident, 'backend', resident.CudaBackend())
    else:
        if _has_cuda():
            pytest.skip('GPU present: the cuda variant runs the real kernels')
        oracle_device.patch_resident(monkeypatch)


def mv(ctx, vals, shape=None, exact=True):
    vals = list(vals)
    return ModValue(ctx, codec.ints_to_limbs(vals, ctx), shape or (len(vals),), exact=exact)


def obj(vals, shape=None):
    a = np.array(list(vals), dtype=object)
    return a.reshape(shape) if shape else a


def reduced(ctx, x):
    """What Zp.array(x).value holds: x as canonical residues (x: ModValue or object array)."""
    lv = resident.as_limb_value(x) if type(x) is ModValue else None
    if lv is not None:
        return [int(v) for v in codec.limbs_to_ints(lv.host_limbs(), ctx)]
    if type(x) is ModValue:
        x = x._ints
    return [int(v) % ctx.modulus for v in np.asarray(x, dtype=object).reshape(-1)]


@pytest.mark.parametrize('p', PRIMES, ids=lambda p: f'p{p.bit_length()}')
def test_np_trunc_expressions(p):
    ctx = mpyc_b200.context_for(p)
    n, f, l = 7, 6, 20
    bits = orc.synth_elements(p, n * f, 3, stream=1)
    a = orc.synth_elements(p, n, 3, stream=2)
    rdiv = orc.synth_elements(p, n, 3, stream=3)
    c_open = orc.synth_elements(p, n, 3, stream=4)
    before = resident.calls['materialised']

    def run(rb, av, rd, cv):
        ar_modf = np.sum(rb.reshape((n, f)) << np.arange(f), axis=1)          # runtime.py:860
        ar_modf = ar_modf.reshape((n,))
        rd = rd.reshape((n,))
        ar_modf += av                                                          # :867
        opened = ar_modf + (1 << l - 1) + (rd << f)                            # :868
        c = cv & ((1 << f) - 1)                                                # :870
        return opened, ar_modf - c                                             # :871
    want = run(obj(bits), obj(a), obj(rdiv), obj(c_open))
    got = run(mv(ctx, bits), mv(ctx, a), mv(ctx, rdiv), mv(ctx, c_open))
    assert all(type(g) is ModValue and g.store is not None for g in got)
    assert [reduced(ctx, g) for g in got] == [reduced(ctx, w) for w in want]
    assert resident.calls['materialised'] == before                            # no Python ints on the way
    # mixed operands: a plain object array where the reference has one (an input that never was limb-backed)
    got = run(mv(ctx, bits), obj(a), mv(ctx, rdiv), mv(ctx, c_open))
    assert [reduced(ctx, g) for g in got] == [reduced(ctx, w) for w in want]


@pytest.mark.parametrize('p', PRIMES, ids=lambda p: f'p{p.bit_length()}')
def test_np_sgn_prefix_expressions(p):
    ctx = mpyc_b200.context_for(p)
    n, l = 5, 37
    r_bits = orc.synth_elements(p, (l + 1) * n, 5, stream=1)
    a = orc.synth_elements(p, n, 5, stream=2)
    rdiv = orc.synth_elements(p, n, 5, stream=3)
    c_open = orc.synth_elements(p, n, 5, stream=4)

    def run(rb, av, rd, cv):
        s_sign = (rb[-n:] << 1) - 1                                            # runtime.py:3646
        rb = rb[:l * n].reshape((n, l))
        shifts = np.arange(l - 1, -1, -1)
        r_modl = np.sum(rb << shifts, axis=1)                                  # :3651
        a_r = av.reshape((n,)) + (1 << l) + r_modl                             # :3656
        opened = a_r + (rd << l)
        c = cv & ((1 << l) - 1)                                                # :3658
        z = c - a_r
        c_bits = np.right_shift.outer(c, shifts).T & 1                         # :3660 (an object array from here on)
        rbT = rb.T
        Xor = c_bits + rbT - (c_bits * rbT << 1)
        e = s_sign - np.vstack((c_bits - rbT, np.ones((1, n), dtype=object))) + 3 * np.cumsum(np.vstack((np.zeros((1, n), dtype=object), Xor)), axis=0)
        h = (1 - (np.array([0, 1, 0, 1, 1], dtype=object) << 1)) * s_sign + 3
        z2 = z + (h << l - 1)
        return s_sign, opened, z, Xor, e, z2
    want = run(obj(r_bits), obj(a), obj(rdiv), obj(c_open))
    before = resident.calls['materialised']
    got = run(mv(ctx, r_bits), mv(ctx, a), mv(ctx, rdiv), mv(ctx, c_open))
    assert resident.calls['materialised'] == before                              # the whole bit-matrix algebra ran on limbs
    assert all(type(g) is ModValue and g.store is not None for g in got)
    assert got[3].shape == (l, n) and got[4].shape == (l + 1, n)
    for g, w in zip(got, want):
        assert reduced(ctx, g) == reduced(ctx, w)


@pytest.mark.parametrize('p', [q for q in PRIMES if q & 3 == 3], ids=lambda p: f'p{p.bit_length()}')
@pytest.mark.parametrize('signed', [False, True])
def test_np_random_bits_expressions(p, signed):
    ctx = mpyc_b200.context_for(p)
    n, f = 9, 4
    r = orc.synth_elements(p, n, 7, stream=1)
    z = orc.synth_elements(p, n, 7, stream=2)
    r2_open = [(x * x) % p for x in r]                                         # what output() returns for r^2 + (shares of 0)

    def sqrt_inv(x):
        if type(x) is ModValue:
            return x.sqrt(INV=True)
        return obj(orc.ff_sqrt(p, [int(v) for v in x], INV=True))

    def run(rv, zv, r2v):
        sq = rv**2 + zv                                                        # runtime.py:4252
        mask = r2v != 0                                                        # :4254
        bits = rv * sqrt_inv(r2v)                                              # :4265
        if not signed:
            bits %= p
            bits += 1
            bits *= (p + 1) >> 1
        bits <<= f
        return sq, mask, bits
    want = run(obj(r), obj(z), obj(r2_open))
    got = run(mv(ctx, r), mv(ctx, z), mv(ctx, r2_open))
    assert reduced(ctx, got[0]) == reduced(ctx, want[0]) and type(got[0]) is ModValue
    assert got[1].tolist() == want[1].tolist() and got[1].dtype == bool
    assert reduced(ctx, got[2]) == reduced(ctx, want[2]) and type(got[2]) is ModValue
    if not signed:
        assert set(reduced(ctx, got[2])) <= {0, (1 << f) % p}                 # they are bits, scaled by 2^f


def test_non_ring_operations_need_exact_values():
    ctx = mpyc_b200.context_for(PRIMES[3])
    x = mv(ctx, [5, 6, 7])
    y = x + x                                                                   # a modular intermediate
    assert not y.exact and (x & 3).exact
    with pytest.raises(AssertionError):
        y & 3
    with pytest.raises(AssertionError):
        y != 0
    with pytest.raises(AssertionError):
        np.right_shift.outer(y, np.arange(3))
    assert (y % ctx.modulus) is y                                               # `bits %= modulus` is a no-op mod p
    assert reduced(ctx, x % 4) == [1, 2, 3]


def test_unknown_uses_fall_back_to_the_object_array():
    ctx = mpyc_b200.context_for(PRIMES[1])
    vals = [3, 4, 5, 6]
    x = mv(ctx, vals, (2, 2))
    xt = x.T
    assert type(xt) is ModValue and xt.shape == (2, 2) and reduced(ctx, xt) == [3, 5, 4, 6]    # 2-D transpose: k_transpose
    assert x[1, 0] == 5 and x.store is None                                     # fancy indexing: settled
    assert (x + 1).tolist() == [[4, 5], [6, 7]]                                 # a plain array from here on
    assert isinstance(x, np.ndarray) and type(x) is ModValue                    # what sectypes' isinstance checks see
    y = mv(ctx, vals)
    assert (y + np.array([[1], [2]], dtype=object)).shape == (2, 4)             # broadcasting: NumPy's job
    w = mv(ctx, vals)
    assert np.concatenate([w, np.array([9], dtype=object)]).tolist() == vals + [9]
    assert resident.as_limb_value(w) is None and resident.as_limb_value(mv(ctx, vals)) is not None
    v = mv(ctx, vals)
    assert type(v[1:3]) is ModValue and reduced(ctx, v[1:3]) == [4, 5]
    assert v[np.array([True, False, True, False])].tolist() == [3, 5]
    lv = (mv(ctx, vals) * 2 + 1).limb_value()
    assert type(lv) is LimbValue and [int(t) for t in codec.limbs_to_ints(lv.host_limbs(), ctx)] == [7, 9, 11, 13]



@pytest.mark.parametrize('p', [PRIMES[1], PRIMES[3]], ids=lambda p: f'p{p.bit_length()}')
def test_matrix_steps_against_numpy(p):
    """vstack / cumsum(axis=0) / row broadcast / transpose / right_shift.outer(..).T & 1 on ModValues == NumPy on objects."""
    ctx = mpyc_b200.context_for(p)
    R, C = 5, 7
    m = orc.synth_elements(p, R * C, 9, stream=1)
    v = orc.synth_elements(p, C, 9, stream=2)
    M, V = mv(ctx, m, (R, C)), mv(ctx, v)
    Mo, Vo = obj(m, (R, C)), obj(v)
    ones = np.ones((1, C), dtype=object)
    for got, want in ((np.vstack((ones, M)), np.vstack((ones, Mo))), (np.cumsum(M, axis=0), np.cumsum(Mo, axis=0)),
                      (V - M, Vo - Mo), (M - V, Mo - Vo), (M * V + 3 * M, Mo * Vo + 3 * Mo), (M.T, Mo.T),
                      (np.vstack((M, M * 2, ones)), np.vstack((Mo, Mo * 2, ones)))):
        assert type(got) is ModValue and got.shape == want.shape
        assert reduced(ctx, got) == reduced(ctx, want)
    for shifts in (np.arange(9), np.arange(8, -1, -1)):
        assert reduced(ctx, np.right_shift.outer(V, shifts).T & 1) == reduced(ctx, np.right_shift.outer(Vo, shifts).T & 1)
        assert reduced(ctx, np.right_shift.outer(V, shifts) & 1) == reduced(ctx, np.right_shift.outer(Vo, shifts) & 1)
    # anything else is NumPy's: settled arrays in, plain arrays out
    assert np.cumsum(M, axis=1).tolist() == np.cumsum(Mo, axis=1).tolist()
    assert np.where(np.array([True] * C), mv(ctx, v), 0).tolist() == v
    assert np.hstack((mv(ctx, v), Vo)).tolist() == v + v
    import pickle
    assert pickle.loads(pickle.dumps(mv(ctx, v))).tolist() == v


@pytest.mark.parametrize('p', [PRIMES[1], PRIMES[3]], ids=lambda p: f'p{p.bit_length()}')
def test_np_to_bits_expressions(p):
    """runtime.py:4413-4433: bit positions on the LAST axis of an N-d value; `c % (1<<l)`; public bits through np.int8."""
    ctx = mpyc_b200.context_for(p)
    shape, l, bl = (3, 2), 9, 20
    n = 6
    r_bits = orc.synth_elements(p, n * l, 21, stream=1)
    rdiv = orc.synth_elements(p, n, 21, stream=2)
    c_open = orc.synth_elements(p, n, 21, stream=3)
    shifts = np.arange(l)

    def run(rb, rd, cv):
        r_modl = np.sum(rb.reshape(shape + (l,)) << shifts, axis=-1)                 # :4415
        masked = (1 << bl) + (rd.reshape(shape) << l) - r_modl                        # :4431
        c = cv % (1 << l)                                                             # :4432
        c_bits = np.int8(np.right_shift.outer(c, shifts) & 1)                         # :4433
        return r_modl, masked, c, c_bits
    want = run(obj(r_bits), obj(rdiv), obj(c_open))
    got = run(mv(ctx, r_bits), mv(ctx, rdiv), mv(ctx, c_open))
    assert type(got[0]) is ModValue and got[0].shape == shape and type(got[1]) is ModValue and type(got[2]) is ModValue
    for g, w in zip(got[:3], want[:3]):
        assert reduced(ctx, g) == reduced(ctx, w)
    assert got[3].dtype == np.int8 and got[3].tolist() == want[3].tolist()


def test_np_random_bits_retry_path_with_a_zero_square():
    """runtime.py:4254-4264: when an opened square is 0 the function keeps the non-zero entries (`_r.value[mask]`,
    np.append) and draws again -- boolean-mask indexing and np.append turn ModValues into plain arrays, and the rest of
    the function continues on those exactly as the reference does."""
    p = PRIMES[0]
    ctx = mpyc_b200.context_for(p)
    r1 = [5, 0, 7, 11]                                     # second entry: r = 0, so r^2 opens to 0
    r2 = [3]                                               # the redraw for the one missing bit
    sq = lambda v: [(x * x) % p for x in v]   # noqa: E731

    def run(R1, S1, R2, S2):
        r = np.array([], dtype='O')
        r2_acc = np.array([], dtype='O')
        h = 4
        mask = S1 != 0
        h -= np.count_nonzero(mask)
        assert h == 1
        r = np.append(r, R1[mask])
        r2_acc = np.append(r2_acc, S1[mask])
        mask = S2 != 0
        h -= np.count_nonzero(mask)
        assert h == 0
        r = np.append(r, R2)
        r2_acc = np.append(r2_acc, S2)
        inv = np.array(orc.ff_sqrt(p, [int(v) for v in r2_acc], INV=True), dtype=object)
        bits = r * inv
        bits %= p
        bits += 1
        bits *= (p + 1) >> 1
        return bits
    want = run(obj(r1), obj(sq(r1)), obj(r2), obj(sq(r2)))
    got = run(mv(ctx, r1), mv(ctx, sq(r1)), mv(ctx, r2), mv(ctx, sq(r2)))
    assert isinstance(got, np.ndarray) and type(got) is np.ndarray
    assert [int(v) % p for v in got] == [int(v) % p for v in want] and set(int(v) % p for v in got) <= {0, 1}
