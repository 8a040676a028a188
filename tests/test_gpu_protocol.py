"""Caller-level check: the [GRR98] secure multiplication that mpyc/runtime.py builds from the hot path
(`np_multiply` -> local product -> `_reshare`, runtime.py:1096-1141,603-689) and `output`
(runtime.py:511-601), replayed for m simulated parties with the drop-in functions only.  What the parties
reconstruct must be the plain product -- for prime fields and for GF(2^8)."""
import random

import numpy as np
import pytest

import fakefield

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
if not torch.cuda.is_available():
    pytest.skip('no CUDA device', allow_module_level=True)

from mpyc_b200 import thresha   # noqa: E402


def _share(F, values, t, m):
    rows = thresha.np_random_split(F, np.array(values, dtype=object), t, m)
    return [F.array(rows[i], check=False) for i in range(m)]


def _reshare(F, prod_shares, t, m):
    """every one of the 2t+1 parties splits its degree-2t share; party j recombines what it received"""
    k = 2 * t + 1
    dealt = [thresha.np_random_split(F, prod_shares[i].value, t, m) for i in range(k)]
    return [thresha.np_recombine(F, [(i + 1, dealt[i][j]) for i in range(k)]) for j in range(m)]


def _open(F, shares, t, who):
    """`output`: party `who` recombines its own share with those of its t predecessors (wrapping)"""
    m = len(shares)
    pts = [((who - t + j) % m + 1, shares[(who - t + j) % m].value) for j in range(t)] + [(who + 1, shares[who].value)]
    return thresha.np_recombine(F, pts).value.tolist()


@pytest.mark.parametrize('p,m,t', [(2**61 - 1, 3, 1), (2**64 - 189, 5, 2), (2**69 - 93, 3, 1), (2**128 - 173, 5, 2),
                                   (2**256 - 189, 7, 3), (9409569905028393239, 3, 1)])
def test_secure_multiplication_prime_fields(p, m, t):
    F = fakefield.make_prime_field(p)
    rnd = random.Random(p % 1000 + m)
    n = 257
    a = [rnd.randrange(p) for _ in range(n)]
    b = [rnd.randrange(p) for _ in range(n)]
    sa, sb = _share(F, a, t, m), _share(F, b, t, m)
    prod = [F.array(sa[i].value * sb[i].value) for i in range(m)]        # local products: degree 2t
    sc = _reshare(F, prod, t, m)
    want = [x * y % p for x, y in zip(a, b)]
    for who in range(m):
        assert _open(F, sc, t, who) == want
    # a second multiplication on the reshared (degree-t) result still works: (ab)*a
    prod2 = [F.array(sc[i].value * sa[i].value) for i in range(m)]
    sd = _reshare(F, prod2, t, m)
    assert _open(F, sd, t, 0) == [x * y % p for x, y in zip(want, a)]


def test_secure_multiplication_gf256():
    """np_aes.py's field: GF(2^8) with modulus 283, m=3, t=1, tiny batches (n in {4, 16, 32})."""
    from oracle import shamir_oracle as orc
    F = fakefield.make_gf256(283)
    Poly = fakefield.Poly
    m, t = 3, 1
    rnd = random.Random(5)
    for n in (4, 16, 32, 1000):
        a = [rnd.randrange(256) for _ in range(n)]
        b = [rnd.randrange(256) for _ in range(n)]
        sa = thresha.np_random_split(F, np.array([Poly(x) for x in a], dtype=object), t, m)
        sb = thresha.np_random_split(F, np.array([Poly(x) for x in b], dtype=object), t, m)
        prod = [np.array([Poly(v) for v in orc.bf_mul(283, [int(x) for x in sa[i]], [int(y) for y in sb[i]])], dtype=object)
                for i in range(m)]
        dealt = [thresha.np_random_split(F, prod[i], t, m) for i in range(2 * t + 1)]
        sc = [thresha.np_recombine(F, [(i + 1, dealt[i][j]) for i in range(2 * t + 1)]) for j in range(m)]
        opened = thresha.np_recombine(F, [(2, sc[1].value), (3, sc[2].value)])
        assert [int(v) for v in opened.value] == orc.bf_mul(283, a, b)


@pytest.mark.parametrize('p,m,t', [(2**61 - 1, 3, 1), (2**69 - 93, 3, 1), (2**128 - 173, 5, 2), (2**256 - 189, 7, 3),
                                   (9409569905028393239, 3, 1)])
def test_secure_multiplication_over_the_limb_wire(p, m, t, monkeypatch):
    """Same protocol with thresha.limb_wire: rows are ShareRow objects, every row that crosses to another
    party goes through pickle.dumps / pickle.loads as in runtime.py:655-656,665,676, and np_recombine
    consumes the unpickled limbs directly.  Results equal the object-array path bit for bit."""
    import pickle
    from mpyc_b200 import wire
    F = fakefield.make_prime_field(p)
    rnd = random.Random(p % 1000 + m + 1)
    n = 301
    a = [rnd.randrange(p) for _ in range(n)]
    b = [rnd.randrange(p) for _ in range(n)]
    seq = iter(rnd.randrange(p) for _ in range(10**7))
    monkeypatch.setattr(thresha, 'coefficient_source', lambda order, count: [next(seq) for _ in range(count)])

    def run(limb_wire):
        nonlocal seq
        seq = iter(random.Random(99).randrange(p) for _ in range(10**7))
        monkeypatch.setattr(thresha, 'limb_wire', limb_wire)
        ship = (lambda row: pickle.loads(pickle.dumps(row))) if limb_wire else (lambda row: row)
        sa = [F.array(ship(r), check=False) for r in thresha.np_random_split(F, np.array(a, dtype=object), t, m)]
        sb = [F.array(ship(r), check=False) for r in thresha.np_random_split(F, np.array(b, dtype=object), t, m)]
        prod = [F.array(sa[i].value * sb[i].value) for i in range(m)]
        k = 2 * t + 1
        dealt = [thresha.np_random_split(F, prod[i].value, t, m) for i in range(k)]
        if limb_wire:
            assert all(isinstance(dealt[i][j], wire.ShareRow) for i in range(k) for j in range(m))
        sc = [thresha.np_recombine(F, [(i + 1, dealt[i][j] if i == j else ship(dealt[i][j])) for i in range(k)])
              for j in range(m)]
        return [s.value.tolist() for s in sc]

    plain, limbs = run(False), run(True)
    assert plain == limbs
    monkeypatch.setattr(thresha, 'limb_wire', False)
    sc = [F.array(np.array(v, dtype=object), check=False) for v in limbs]
    assert _open(F, sc, t, 0) == [x * y % p for x, y in zip(a, b)]
