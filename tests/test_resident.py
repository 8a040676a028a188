"""mpyc_b200.resident.LimbValue on its own (no reference, no GPU): the opaque-handle operations stay limb-backed, any
other use turns it into the object array the reference would have had, and it pickles as fixed-width wire bytes."""
import pickle

import numpy as np
import pytest

import mpyc_b200
from mpyc_b200 import codec, resident
from mpyc_b200.resident import LimbValue
from mpyc_b200.wire import ShareRow

P = 2**128 - 173


@pytest.fixture
def lv():
    ctx = mpyc_b200.context_for(P)
    vals = [0, 1, P - 1, 2**64, 12345678901234567890123, 7]
    return ctx, vals, LimbValue(ctx, codec.ints_to_limbs(vals, ctx), (2, 3))


def test_handle_operations_keep_the_limbs(lv):
    ctx, vals, v = lv
    before = resident.calls['materialised']
    assert v.shape == (2, 3) and v.ndim == 2 and v.size == 6 and len(v) == 2 and v.dtype == object
    flat = v.reshape(-1)
    assert type(flat) is LimbValue and flat.shape == (6,) and flat.store is v.store
    assert v.reshape(3, 2).shape == (3, 2) and v.reshape((6, 1)).shape == (6, 1) and v.ravel().shape == (6,)
    assert v.flatten().shape == (6,) and v.copy().shape == (2, 3)
    with pytest.raises(ValueError):
        v.reshape(4, 2)
    back = pickle.loads(pickle.dumps(v))
    assert type(back) is LimbValue and back.shape == (2, 3) and back.limb_backed
    assert (back.host_limbs() == v.host_limbs()).all()
    assert resident.calls['materialised'] == before          # no Python ints so far
    assert resident.as_limb_value(v) is v
    row = ShareRow(ctx, codec.ints_to_limbs(vals, ctx))
    assert resident.as_limb_value(row).shape == (6,)


def test_any_other_use_materialises_once(lv):
    ctx, vals, v = lv
    before = resident.calls['materialised']
    assert v[1, 2] == 7                                        # indexing -> object array from here on
    assert resident.calls['materialised'] == before + 1 and not v.limb_backed
    assert resident.as_limb_value(v) is None
    assert np.asarray(v).tolist() == [vals[:3], vals[3:]]
    assert (v + 1).tolist() == [[x + 1 for x in vals[:3]], [x + 1 for x in vals[3:]]]
    assert (2 * v)[0, 1] == 2 and (v % 5).shape == (2, 3)
    v[0, 0] = 99                                               # in-place update reaches the ints
    assert np.asarray(v)[0, 0] == 99 and v.tolist()[0][0] == 99
    assert isinstance(v.reshape(-1), np.ndarray) and v.T.shape == (3, 2)
    assert resident.calls['materialised'] == before + 1
    w = pickle.loads(pickle.dumps(v))                          # a settled value pickles as the plain object array
    assert isinstance(w, np.ndarray) and w[0, 0] == 99


def test_numpy_functions_see_the_values(lv):
    ctx, vals, v = lv
    assert np.concatenate([v.reshape(-1), np.array([5], dtype=object)]).tolist() == vals + [5]
    x = LimbValue(ctx, codec.ints_to_limbs(vals, ctx), (6,))
    assert (np.array([1, 1, 1, 1, 1, 1], dtype=object) * x).tolist() == vals      # reflected operand of an ndarray
    y = LimbValue(ctx, codec.ints_to_limbs(vals, ctx), (6,))
    y += 1                                                     # in-place dunder returns the real array
    assert isinstance(y, np.ndarray) and y.tolist() == [a + 1 for a in vals]


def test_gf256_values_come_back_as_polynomial_objects():
    import fakefield
    ctx = mpyc_b200.context_for(283, binary=True)
    v = LimbValue(ctx, np.array([1, 2, 0x53], dtype=np.uint8), (3,), fakefield.Poly)
    assert type(pickle.loads(pickle.dumps(v))) is LimbValue
    assert all(isinstance(e, fakefield.Poly) for e in v) and [int(e) for e in v] == [1, 2, 0x53]
