"""Host logic of the drop-in adapter (mpyc_b200/thresha.py) WITHOUT a GPU: the two device round trips
(`_split_limbs`, `_recombine_limbs`) are replaced by the CPU oracle working on the same limb arrays, and the
reference-generated golden cases are replayed through np_random_split / random_split / np_recombine / recombine.
What this pins on CPU: argument handling (lists, field elements, field arrays), the int <-> limb codec, the
consumption order of the coefficient stream (np: (t, n) row-major; list: element-major with Horner reversal,
thresha.py:37-43,58-60), shapes and return types -- BASELINE configs[0] ("bit-exact plumbing, no GPU").  The
kernels themselves are covered by the -m gpu tests with the very same fixtures."""
import itertools

import numpy as np
import pytest

import fakefield
from golden_util import load, unhex
from oracle import shamir_oracle as orc
from mpyc_b200 import codec, thresha, wire

SR = load('split_recombine.json')


@pytest.fixture
def oracle_device(monkeypatch):
    """thresha's device calls answered by the oracle on the limb arrays the adapter built."""
    def empty(ctx, rows):
        return np.zeros((rows, 0) if ctx.binary else (rows, 0, ctx.nlimbs), dtype=np.uint8 if ctx.binary else np.uint64)

    def split_limbs(ctx, sec, C, t, m):
        F = orc.field_of(ctx.modulus, binary=ctx.binary)
        s = [int(v) for v in codec.limbs_to_ints(sec, ctx)]
        rows = [[int(v) for v in codec.limbs_to_ints(C[j], ctx)] for j in range(t)]
        shares = orc.split_np_order(F, s, rows, m)
        return np.stack([codec.ints_to_limbs(r, ctx) for r in shares]) if s else empty(ctx, m)

    def recombine_limbs(ctx, xs, rows, pts):
        F = orc.field_of(ctx.modulus, binary=ctx.binary)
        vals = orc.recombine(F, list(xs), [[int(v) for v in codec.limbs_to_ints(r, ctx)] for r in rows], list(pts))
        n = rows[0].shape[0]
        return np.stack([codec.ints_to_limbs(v, ctx) for v in vals]) if n else empty(ctx, len(pts))

    monkeypatch.setattr(thresha, '_split_limbs', split_limbs)
    monkeypatch.setattr(thresha, '_recombine_limbs', recombine_limbs)

    def inject(stream):
        it = iter(stream)
        monkeypatch.setattr(thresha, 'coefficient_source', lambda order, count: list(itertools.islice(it, count)))
    return inject


@pytest.mark.parametrize('case', SR['cases'], ids=lambda c: f"p{int(c['p'],16).bit_length()}_m{c['m']}t{c['t']}")
def test_golden_cases_through_the_adapter(case, oracle_device):
    p, m, t = int(case['p'], 16), case['m'], case['t']
    F = fakefield.make_prime_field(p)
    s, stream = unhex(case['secrets']), unhex(case['stream'])
    oracle_device(stream)
    sh_np = thresha.np_random_split(F, F.array(np.array(s, dtype=object)), t, m)
    assert sh_np.dtype == object and sh_np.shape == (m, len(s)) and sh_np.tolist() == unhex(case['shares_np'])
    oracle_device(stream)
    sh_li = thresha.random_split(F, list(s), t, m)
    assert sh_li == unhex(case['shares_list'])
    oracle_device(stream)
    assert thresha.random_split(F, [F(x) for x in s], t, m) == unhex(case['shares_list'])     # field elements in
    for rec in case['recombine']:
        xs = rec['xs']
        y0 = thresha.np_recombine(F, [(x, sh_np[x - 1]) for x in xs])
        assert isinstance(y0, F.array) and y0.value.tolist() == unhex(rec['y0'])
        assert thresha.np_recombine(F, [(x, sh_np[x - 1]) for x in xs], rec['x_rs']).value.tolist() == unhex(rec['yw'])
        assert thresha.recombine(F, [(x, unhex(case['shares_np'])[x - 1]) for x in xs]) == unhex(rec['y0'])
        ye = thresha.recombine(F, [(x, [F(v) for v in sh_li[x - 1]]) for x in xs], [0])
        assert all(isinstance(v, F) for v in ye[0])


def test_c1_plumbing_and_limb_wire(oracle_device, monkeypatch):
    """BASELINE configs[0]: 1024 secrets, p = 2^61-1, m = 3, t = 1 -- adapter == oracle bit for bit on an injected
    coefficient stream, for the np and the list variant; and the limb-wire form of the same call recombines to the
    same values after a pickle round trip of every row."""
    import pickle
    import random
    p, m, t, n = 2**61 - 1, 3, 1, 1024
    F, Fo = fakefield.make_prime_field(p), orc.field_of(p)
    rnd = random.Random(1)
    s = [rnd.randrange(p) for _ in range(n)]
    stream = [rnd.randrange(p) for _ in range(t * n)]
    oracle_device(stream)
    sh = thresha.np_random_split(F, np.array(s, dtype=object), t, m)
    assert sh.tolist() == orc.split_np_order(Fo, s, [stream[:n]], m)
    oracle_device(stream)
    sl = thresha.random_split(F, s, t, m)
    assert sl == orc.split_np_order(Fo, s, [stream], m)            # t = 1: element-major == row-major
    assert thresha.np_recombine(F, [(1, sh[0]), (3, sh[2])]).value.tolist() == s
    monkeypatch.setattr(thresha, 'limb_wire', True)
    oracle_device(stream)
    rows = thresha.np_random_split(F, np.array(s, dtype=object), t, m)
    assert isinstance(rows, wire.ShareRows) and len(rows) == m
    shipped = [pickle.loads(pickle.dumps(r)) for r in rows]
    assert all(isinstance(r, wire.ShareRow) for r in shipped)
    assert [r.tolist() for r in shipped] == sh.tolist()
    assert thresha.np_recombine(F, [(2, shipped[1]), (3, shipped[2])]).value.tolist() == s


def test_gf256_golden_through_the_adapter(oracle_device):
    """GF(2^8) (np_aes's field): values cross the Python surface as polynomial objects, points are the field elements
    with integer encoding i+1 (thresha.py:54,61); np and list draw orders as for prime fields."""
    G = load('gf256.json')
    F = fakefield.make_gf256(G['modulus'])
    for case in G['split']:
        m, t = case['m'], case['t']
        s, stream = unhex(case['secrets']), unhex(case['stream'])
        oracle_device(stream)
        sh = thresha.np_random_split(F, np.array([fakefield.Poly(x) for x in s], dtype=object), t, m)
        assert all(isinstance(v, fakefield.Poly) for v in sh[0])
        assert [[int(v) for v in row] for row in sh] == unhex(case['shares_np'])
        oracle_device(stream)
        sl = thresha.random_split(F, [F(x) for x in s], t, m)
        assert [[int(v) for v in row] for row in sl] == unhex(case['shares_list'])
        xs = case['xs']
        assert [int(v) for v in thresha._recombination_vector(F, tuple(xs), 0)] == unhex(case['lambda0'])
        y = thresha.np_recombine(F, [(x, sh[x - 1]) for x in xs])
        assert [int(v) for v in y.value] == unhex(case['y0'])


# ---- PRSS adapter: keys, coefficients f_S(i) (computed by the library's HOST code), weights and their order ----------

PRSS = load('prss.json')


@pytest.fixture
def oracle_prss(monkeypatch):
    """thresha._prss_device answered on the CPU: SHAKE128 streams via the library's own host XOF (checked against hashlib
    in test_shake128.py), chunk -> value and the linear combination on Python ints, straight from the formula of
    include/mpyc_b200.h (out = sum_S coef_S * sum_j value_{S,h,j} * w_j)."""
    import ctypes
    from mpyc_b200 import _cabi

    def prss_device(ctx, keys, uci, d, width, bound_bits, coef, weights, n, general=None):
        F = orc.field_of(ctx.modulus, binary=ctx.binary)
        L = max(ctx.nlimbs, 1)
        cs = [_cabi.limbs_to_int(coef[i * L:(i + 1) * L]) for i in range(len(keys))]
        ws = [_cabi.limbs_to_int(weights[j * L:(j + 1) * L]) for j in range(d)]
        full = 256 if ctx.binary else ctx.modulus
        bound = general if general is not None else (1 << bound_bits if bound_bits else full)
        acc = [0] * n
        for key, c in zip(keys, cs):
            raw = ctypes.create_string_buffer(n * d * width)
            _cabi.check(_cabi.lib.mpyc_b200_shake128(key + uci, len(key) + len(uci), raw, n * d * width))
            raw = raw.raw
            for h in range(n):
                y = 0
                for j in range(d):
                    v = int.from_bytes(raw[(h * d + j) * width:(h * d + j + 1) * width], 'little') % bound
                    y = F.add(y, F.mul(v, ws[j]))
                acc[h] = F.add(acc[h], F.mul(c, y))
        return codec.ints_to_limbs([F.red(a) for a in acc], ctx)

    monkeypatch.setattr(thresha, '_prss_device', prss_device)


@pytest.mark.parametrize('case', PRSS['cases'], ids=lambda c: f"p{int(c['p'],16).bit_length()}_m{c['m']}t{c['t']}")
def test_golden_prss_through_the_adapter(case, oracle_prss):
    p, m, t, n = int(case['p'], 16), case['m'], case['t'], case['n']
    F = fakefield.make_prime_field(p)
    uci = bytes.fromhex(case['uci'])
    keys = {tuple(int(x) for x in k.split(',')): bytes.fromhex(v) for k, v in case['keys'].items()}
    for party in case['parties']:
        i = party['i']
        prfs = {S: thresha.PRF(k, p) for S, k in keys.items() if i in S}
        for S in prfs:
            assert thresha._f_S_i(F, m, i, S) == int(party['f_S_i'][','.join(map(str, S))], 16) % p
        a_np = thresha.np_pseudorandom_share(F, m, i, prfs, uci, n)
        assert isinstance(a_np, F.array) and a_np.value.tolist() == unhex(party['share_np'])
        assert [x.value for x in thresha.pseudorandom_share(F, m, i, prfs, uci, n)] == unhex(party['share_list'])
        assert [x.value for x in thresha.pseudorandom_share_zero(F, m, i, prfs, uci, n)] == unhex(party['zero_list'])
        if t:
            assert thresha.np_pseudorandom_share_0(F, m, i, prfs, uci, n).value.tolist() == unhex(party['zero_np'])


BOUNDS = load('prss_bounds.json')


@pytest.mark.parametrize('case', BOUNDS['cases'],
                         ids=lambda c: f"{c['field']}{int(c['modulus'], 16).bit_length()}_b{int(c['bound'], 16).bit_length()}_m{c['m']}t{c['t']}")
def test_golden_prss_any_bound_through_the_adapter(case, oracle_prss):
    """Host side of PRSS with PRF bounds other than the field order (which form of the bound reaches the device call,
    chunk width, GF(2^8) polynomial wrapping), on the reference-generated cases."""
    mod, bound, m, t, n = int(case['modulus'], 16), int(case['bound'], 16), case['m'], case['t'], case['n']
    F = fakefield.make_gf256(mod) if case['field'] == 'gf' else fakefield.make_prime_field(mod)
    uci = bytes.fromhex(case['uci'])
    keys = {tuple(int(x) for x in k.split(',')): bytes.fromhex(v) for k, v in case['keys'].items()}
    for party in case['parties']:
        i = party['i']
        prfs = {S: thresha.PRF(k, bound) for S, k in keys.items() if i in S}
        a_np = thresha.np_pseudorandom_share(F, m, i, prfs, uci, n)
        assert isinstance(a_np, F.array) and [int(v) for v in a_np.value] == unhex(party['share_np'])
        assert [int(x.value) for x in thresha.pseudorandom_share(F, m, i, prfs, uci, n)] == unhex(party['share_list'])
        assert [int(x.value) for x in thresha.pseudorandom_share_zero(F, m, i, prfs, uci, n)] == unhex(party['zero_list'])
        if t:
            assert [int(v) for v in thresha.np_pseudorandom_share_0(F, m, i, prfs, uci, n).value] == unhex(party['zero_np'])


def test_bound_form_handed_to_the_device(monkeypatch):
    """bound == order -> bound_bits 0; 2^b <= order -> bound_bits b (also GF(2^8): the bug of round 1 was dropping it);
    anything else up to 2^256 -> general; wider -> UnsupportedFieldError (install() then defers to the reference)."""
    import mpyc_b200
    seen = []

    def spy(ctx, keys, uci, d, width, bound_bits, coef, weights, n, general=None):
        seen.append((width, bound_bits, general))
        return np.zeros((n,) if ctx.binary else (n, ctx.nlimbs), dtype=np.uint8 if ctx.binary else np.uint64)
    monkeypatch.setattr(thresha, '_prss_device', spy)
    G, P = fakefield.make_gf256(283), fakefield.make_prime_field(2**61 - 1)
    key = bytes(16)
    for F, bound, want in ((G, 256, (1, 0, None)), (G, 2, (1, 1, None)), (G, 16, (1, 4, None)), (G, 1 << 9, (2, 0, 1 << 9)),
                           (G, 1000, (2 + 16, 0, 1000)), (P, 2**61 - 1, (8 + 16, 0, None)), (P, 2, (1, 1, None)),
                           (P, 1 << 60, (8, 60, None)), (P, 1 << 61, (8, 0, 1 << 61)), (P, 12345, (2 + 16, 0, 12345)),
                           (P, 2**69 - 93, (9 + 16, 0, 2**69 - 93)), (P, 1 << 256, (32, 0, 1 << 256))):
        thresha.np_pseudorandom_share(F, 1, 0, {(0,): thresha.PRF(key, bound)}, b'u', 3)
        assert seen[-1] == want, (bound, seen[-1])
    with pytest.raises(mpyc_b200.UnsupportedFieldError):
        thresha.np_pseudorandom_share(P, 1, 0, {(0,): thresha.PRF(key, (1 << 256) + 1)}, b'u', 3)


def test_ragged_share_rows_are_rejected(oracle_device):
    """A share row of another length (a buggy or malicious peer) must raise before the C ABI copies n elements from it."""
    F = fakefield.make_prime_field(2**61 - 1)
    good, short = np.array([1, 2, 3, 4], dtype=object), np.array([1, 2, 3], dtype=object)
    with pytest.raises(ValueError):
        thresha.np_recombine(F, [(1, good), (2, short)])
    with pytest.raises(ValueError):
        thresha.recombine(F, [(1, [1, 2, 3, 4]), (2, [1, 2, 3])])
    G = fakefield.make_gf256(283)
    with pytest.raises(ValueError):
        thresha.np_recombine(G, [(1, np.array([fakefield.Poly(1)] * 4, dtype=object)), (2, np.array([fakefield.Poly(1)] * 5, dtype=object))])


def test_small_integer_form_of_the_prss_coefficients():
    """The host detection behind K4's small form (api.cu: prss_small_table): for every party i of every (m, t) up to 9
    parties, the coefficients f_S(i) of the party's key subsets times some k! are exactly the small signed integers the
    rational formula prod_{j not in S} (i - j) / (-(j + 1)) predicts -- checked with Fractions, no GPU involved."""
    import ctypes
    import math
    from fractions import Fraction
    import mpyc_b200
    from mpyc_b200 import _cabi
    for p in (2**61 - 1, 2**128 - 173, 2**256 - 189, 9409569905028393239):
        ctx = mpyc_b200.context_for(p)
        L = ctx.nlimbs
        F = fakefield.make_prime_field(p)
        for m, t in ((3, 1), (5, 2), (7, 3), (9, 4), (4, 1), (6, 2)):
            for i in range(m):
                subsets = [S for S in itertools.combinations(range(m), m - t) if i in S]
                coef = []
                for S in subsets:
                    coef.extend(_cabi.int_to_limbs(int(thresha._f_S_i(F, m, i, S)), L))
                num = (ctypes.c_int64 * len(subsets))()
                inv = (ctypes.c_uint64 * L)()
                rc = _cabi.lib.mpyc_b200_prss_small_form(ctx.handle, len(subsets), 1, _cabi.u64_array(coef),
                                                         _cabi.u64_array(_cabi.int_to_limbs(1, L)), num, inv)
                if len(subsets) > 64:          # more than 64 subsets (m = 9, t = 4: 70): the kernel keeps full products
                    assert rc == _cabi.EUNSUPPORTED
                    continue
                _cabi.check(rc)
                scale_inv = _cabi.limbs_to_int(list(inv))
                D = pow(scale_inv, -1, p)
                assert any(D == math.factorial(k) % p for k in range(1, 21))
                for S, n_S in zip(subsets, num):
                    exact = Fraction(1)
                    for j in range(m):
                        if j not in S:
                            exact *= Fraction(i - j, -(j + 1))
                    assert n_S * scale_inv % p == exact.numerator * pow(exact.denominator, -1, p) % p
                    assert abs(n_S) < 2**57
    # coefficients that are not small rationals have no small form
    ctx = mpyc_b200.context_for(2**128 - 173)
    junk = _cabi.u64_array(_cabi.int_to_limbs(0x123456789abcdef0123456789abcdef % (2**128 - 173), 2))
    with pytest.raises(mpyc_b200.UnsupportedFieldError):
        _cabi.check(_cabi.lib.mpyc_b200_prss_small_form(ctx.handle, 1, 1, junk, _cabi.u64_array([1, 0]), (ctypes.c_int64 * 1)(), None))
