"""Limb wire format (mpyc_b200/wire.py): a ShareRow pickles as the fixed-width little-endian
byte string of FiniteFieldElement.to_bytes (mpyc/finfields.py:91-102), unpickles into limbs again, and still
behaves like the object array it replaces for consumers other than np_recombine.  Host logic only: no GPU."""
import pickle

import numpy as np
import pytest

import fakefield
import mpyc_b200
from mpyc_b200 import codec, wire

PRIMES = [2**61 - 1, 2**64 - 189, 2**69 - 93, 2**128 - 173, 2**256 - 189, 9409569905028393239, 101]


@pytest.mark.parametrize('p', PRIMES, ids=lambda p: f'p{p.bit_length()}')
def test_row_pickles_as_fixed_width_bytes(p):
    ctx = mpyc_b200.context_for(p)
    vals = [0, 1, 2, p - 1, p - 2, p >> 1, 1234567890123456789012345678901234567890 % p]
    row = wire.ShareRow(ctx, codec.ints_to_limbs(vals, ctx))
    fn, (modulus, binary, n, data) = row.__reduce__()
    r = (p.bit_length() + 7) >> 3                                  # field.byte_length, finfields.py:359
    assert (modulus, binary, n) == (p, False, len(vals))
    assert data == b''.join(v.to_bytes(r, 'little') for v in vals)   # == field.to_bytes(vals)
    back = pickle.loads(pickle.dumps(row))
    assert isinstance(back, wire.ShareRow) and back.ctx is ctx
    assert back.limbs.dtype == np.uint64 and back.limbs.shape == (len(vals), ctx.nlimbs)
    assert back.tolist() == vals


def test_row_behaves_like_the_object_array_it_replaces():
    p = 2**128 - 173
    ctx = mpyc_b200.context_for(p)
    vals = [5, 7, p - 1, 0]
    row = wire.ShareRow(ctx, codec.ints_to_limbs(vals, ctx))
    assert len(row) == 4 and row.shape == (4,) and row.reshape(-1) is row and row.reshape(4) is row
    assert row.reshape(2, 2).tolist() == [[5, 7], [p - 1, 0]]
    assert list(row) == vals and row[2] == p - 1
    a = np.array(row, dtype=object, copy=None)                     # what FiniteFieldArray.__init__ does (finfields.py:717-725)
    assert a.dtype == object and a.tolist() == vals
    F = fakefield.make_prime_field(p)
    assert F.array(row, check=False).value.tolist() == vals


def test_rows_matrix_iterates_to_rows():
    p = 2**61 - 1
    ctx = mpyc_b200.context_for(p)
    limbs = np.arange(3 * 5, dtype=np.uint64).reshape(3, 5, 1)
    rows = wire.ShareRows(ctx, limbs)
    assert len(rows) == 3 and rows.shape == (3, 5)
    got = [r.tolist() for r in rows]
    assert got == [[0, 1, 2, 3, 4], [5, 6, 7, 8, 9], [10, 11, 12, 13, 14]]
    assert np.array(rows, dtype=object).tolist() == got
    assert rows[1].tolist() == got[1]


def test_gf256_row():
    ctx = mpyc_b200.context_for(283, binary=True)
    row = wire.ShareRow(ctx, np.array([0, 1, 2, 255], dtype=np.uint8), poly_type=fakefield.Poly)
    fn, (modulus, binary, n, data) = row.__reduce__()
    assert (modulus, binary, n, data) == (283, True, 4, bytes([0, 1, 2, 255]))
    assert all(isinstance(v, fakefield.Poly) for v in row.tolist())
    back = pickle.loads(pickle.dumps(row))
    assert [int(v) for v in back.tolist()] == [0, 1, 2, 255]


def test_truncated_wire_data_is_rejected():
    with pytest.raises(ValueError):
        wire._row_from_wire(2**61 - 1, False, 3, b'\x00' * 16)
