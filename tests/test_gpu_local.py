"""GPU parity of the protocol-local kernels (csrc/local.cuh) through the C ABI: the reference-generated
fixtures of tests/golden/local.json, seeded inputs against the oracle at ragged sizes, and size-independent properties
at sizes the oracle cannot reach.  Bit-exact.
"""
import numpy as np
import pytest

from golden_util import load, unhex
from oracle import shamir_oracle as orc

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
if not torch.cuda.is_available():
    pytest.skip('no CUDA device', allow_module_level=True)

import mpyc_b200                                   # noqa: E402
from mpyc_b200 import device as dev                # noqa: E402
from mpyc_b200.device import DeviceArray           # noqa: E402

LOCAL = load('local.json')
P61, P64, P69, P128, P256 = 2**61 - 1, 2**64 - 189, 2**69 - 93, 2**128 - 173, 2**256 - 189
P64G = 9409569905028393239
GEN = {int(c['p'], 16).bit_length(): int(c['p'], 16) for c in LOCAL['algebra'] if int(c['p'], 16) not in (101, P61, P64, P64G, P69, P128, P256)}
PRIMES = [P61, P64, P64G, P69, P128, GEN[128], GEN[192], P256, GEN[256]]


def ints(x):
    return [int(v) for v in x.to_ints()]


def matrix_ints(M):
    return [[int(v) for v in row] for row in M.to_ints()]


@pytest.mark.parametrize('case', LOCAL['algebra'], ids=lambda c: f"p{int(c['p'],16).bit_length()}_{c['p'][-4:]}")
def test_golden_algebra(case):
    p = int(case['p'], 16)
    ctx = mpyc_b200.context_for(p)
    a, b, c = (DeviceArray.from_ints(ctx, unhex(case[k])) for k in 'abc')
    assert ints(dev.fma(a, None, c)) == unhex(case['square_add'])
    assert ints(dev.fma(a, b, c)) == unhex(case['mul_add'])
    mask, count = dev.nonzero(a)
    assert mask.cpu().tolist() == case['nonzero'] and count == sum(case['nonzero'])
    for f in (0, 6):
        s = (p + 1) >> 1 << f
        assert ints(dev.axpb(a, s, s)) == unhex(case[f'bits_tail_f{f}'])
    assert ints(dev.axpb(a, 2, -1)) == unhex(case['s_sign'])
    for key in case:
        if key.startswith('low_bits_'):
            assert ints(dev.low_bits(c, int(key[9:]))) == unhex(case[key])
    for comp in case['compose']:
        bits = DeviceArray.from_ints(ctx, unhex(comp['bits']))
        assert ints(dev.bits_compose(bits, comp['n'], comp['f'])) == unhex(comp['ascending'])
        assert ints(dev.bits_compose(bits, comp['n'], comp['f'], descending=True)) == unhex(comp['descending'])
    for dec in case['decompose']:
        assert matrix_ints(dev.bits_decompose(c, dec['l'])) == unhex(dec['ascending'])
        assert matrix_ints(dev.bits_decompose(c, dec['l'], descending=True)) == unhex(dec['descending'])


@pytest.mark.parametrize('case', LOCAL['conv'], ids=lambda c: f"p{int(c['p'],16).bit_length()}_{'x'.join(map(str, c['shape']))}")
def test_golden_conv2d(case):
    p = int(case['p'], 16)
    ctx = mpyc_b200.context_for(p)
    k, r, m, n, v, s = case['shape']
    X, W, B = (DeviceArray.from_ints(ctx, unhex(case[key])) for key in 'XWB')
    assert ints(dev.conv2d(X, W, B, k, r, m, n, v, s)) == unhex(case['Y'])


@pytest.mark.parametrize('p', PRIMES, ids=lambda p: f'p{p.bit_length()}_{p % 10000}')
@pytest.mark.parametrize('n', [0, 1, 3, 255, 256, 257, 1027])
def test_elementwise_vs_oracle(p, n):
    ctx = mpyc_b200.context_for(p)
    a = orc.synth_elements(p, n, 11, stream=1)
    b = orc.synth_elements(p, n, 11, stream=2)
    c = orc.synth_elements(p, n, 11, stream=3)
    if n > 2:
        a[1], c[2] = 0, p - 1
    A, B, C = (DeviceArray.from_ints(ctx, x) for x in (a, b, c))
    assert ints(dev.fma(A, B, C)) == orc.local_fma(p, a, b, c)
    assert ints(dev.fma(A, None, C)) == orc.local_fma(p, a, a, c)
    for s, t in ((1, 5), (0, 7), (p - 1, p - 1), ((p + 1) >> 1, (p + 1) >> 1), (1 << 40, -(1 << 41))):
        assert ints(dev.axpb(A, s, t)) == orc.local_axpb(p, a, s % p, t % p)
    for nb in (0, 5, 63, 64, 65, p.bit_length() - 1, 300):
        assert ints(dev.low_bits(C, nb)) == orc.local_low_bits(c, nb)
    mask, count = dev.nonzero(A)
    assert mask.cpu().tolist() == orc.local_nonzero(a) and count == sum(orc.local_nonzero(a))
    assert dev.nonzero(A, want_mask=False) == (None, count)


@pytest.mark.parametrize('p', PRIMES, ids=lambda p: f'p{p.bit_length()}_{p % 10000}')
@pytest.mark.parametrize('n,f', [(1, 1), (3, 4), (256, 6), (257, 8), (300, 16), (515, 17), (70, 37), (33, 64), (40, 70), (9, 130), (1000, 5)])
def test_bits_compose_vs_oracle(p, n, f):
    ctx = mpyc_b200.context_for(p)
    bits = orc.synth_elements(p, n * f, 13 + f, stream=4)
    bits[0] = bits[-1] = p - 1
    Bt = DeviceArray.from_ints(ctx, bits)
    assert ints(dev.bits_compose(Bt, n, f)) == orc.local_bits_compose(p, bits, n, f)
    assert ints(dev.bits_compose(Bt, n, f, descending=True)) == orc.local_bits_compose(p, bits, n, f, descending=True)


@pytest.mark.parametrize('p', [P61, P64G, GEN[192], P128], ids=lambda p: f'p{p.bit_length()}_{p % 10000}')
@pytest.mark.parametrize('f', [1, 2, 5, 6, 16, 17, 37])
@pytest.mark.parametrize('skip', [1, 2, 3])
def test_bits_compose_on_a_view_that_starts_mid_buffer(p, f, skip):
    """Rows of 8- and 24-byte elements start on odd 8-byte boundaries depending on row, f and where the view begins:
    the head / tail pieces of the staged copy (local.cuh) must agree with the plain result."""
    ctx = mpyc_b200.context_for(p)
    n = 519
    bits = orc.synth_elements(p, n * f + skip, 41 + f, stream=4)
    whole = DeviceArray.from_ints(ctx, bits)
    view = DeviceArray(ctx, whole.t[skip:])
    assert ints(dev.bits_compose(view, n, f)) == orc.local_bits_compose(p, bits[skip:], n, f)
    assert ints(dev.bits_compose(view, n, f, descending=True)) == orc.local_bits_compose(p, bits[skip:], n, f, descending=True)


@pytest.mark.parametrize('p', PRIMES, ids=lambda p: f'p{p.bit_length()}_{p % 10000}')
@pytest.mark.parametrize('n,l', [(1, 1), (5, 6), (256, 37), (259, 16), (1026, 3)])
def test_bits_decompose_vs_oracle(p, n, l):
    ctx = mpyc_b200.context_for(p)
    c = orc.synth_elements(p, n, 17, stream=5)
    c[0] = p - 1
    C = DeviceArray.from_ints(ctx, c)
    assert matrix_ints(dev.bits_decompose(C, l)) == orc.local_bits_decompose(c, l)
    assert matrix_ints(dev.bits_decompose(C, l, descending=True)) == orc.local_bits_decompose(c, l, descending=True)
    full = p.bit_length()
    assert matrix_ints(dev.bits_decompose(C, full)) == orc.local_bits_decompose(c, full)


@pytest.mark.parametrize('p', [P61, P64G, P69, P128, GEN[128], P256, GEN[256]], ids=lambda p: f'p{p.bit_length()}_{p % 10000}')
@pytest.mark.parametrize('shape', [(1, 1, 28, 28, 2, 5), (2, 3, 7, 9, 3, 3), (1, 4, 14, 14, 3, 5), (3, 2, 1, 5, 2, 1), (1, 2, 3, 300, 2, 3)])
def test_conv2d_vs_oracle(p, shape):
    k, r, m, n, v, s = shape
    ctx = mpyc_b200.context_for(p)
    X = orc.synth_elements(p, k * r * m * n, 19, stream=6)
    W = orc.synth_elements(p, v * r * s * s, 19, stream=7)
    B = orc.synth_elements(p, v, 19, stream=8)
    X[0] = W[0] = B[0] = p - 1
    got = dev.conv2d(*(DeviceArray.from_ints(ctx, x) for x in (X, W, B)), k, r, m, n, v, s)
    assert ints(got) == orc.local_conv2d(p, X, W, B, k, r, m, n, v, s)


def test_conv2d_rejects_even_filters():
    ctx = mpyc_b200.context_for(P61)
    z = DeviceArray.from_ints(ctx, [1] * 64)
    with pytest.raises(mpyc_b200.UnsupportedFieldError):
        dev.conv2d(z, DeviceArray.from_ints(ctx, [1] * 4), DeviceArray.from_ints(ctx, [1]), 1, 1, 8, 8, 1, 2)
    with pytest.raises(mpyc_b200.UnsupportedFieldError):       # rows shorter than the filter: the demo's loop raises there
        dev.conv2d(z, DeviceArray.from_ints(ctx, [1] * 25), DeviceArray.from_ints(ctx, [1]), 1, 1, 16, 4, 1, 5)


@pytest.mark.parametrize('p', PRIMES, ids=lambda p: f'p{p.bit_length()}_{p % 10000}')
@pytest.mark.parametrize('R,C', [(1, 1), (3, 5), (38, 33), (32, 64), (5, 1000), (70, 31), (1, 300)])
def test_matrix_kernels_vs_numpy(p, R, C):
    """k_transpose / k_cumsum_rows / k_binop_rows / contiguous decompose against NumPy object arithmetic mod p."""
    ctx = mpyc_b200.context_for(p)
    m = orc.synth_elements(p, R * C, 43, stream=1)
    v = orc.synth_elements(p, C, 43, stream=2)
    m[0] = v[0] = p - 1
    M, V = DeviceArray.from_ints(ctx, m), DeviceArray.from_ints(ctx, v)
    Mo, Vo = np.array(m, dtype=object).reshape(R, C), np.array(v, dtype=object)
    flat = lambda a: [int(t) % p for t in a.reshape(-1)]   # noqa: E731
    assert ints(dev.transpose(M, R, C)) == flat(Mo.T)
    assert ints(dev.cumsum_rows(M, R, C)) == flat(np.cumsum(Mo, axis=0))
    from mpyc_b200 import _cabi
    assert ints(dev.binop_rows(M, V, _cabi.OP_ADD, R, C)) == flat(Mo + Vo)
    assert ints(dev.binop_rows(M, V, _cabi.OP_SUB, R, C)) == flat(Mo - Vo)
    assert ints(dev.binop_rows(M, V, _cabi.OP_SUB, R, C, reflected=True)) == flat(Vo - Mo)
    assert ints(dev.binop_rows(M, V, _cabi.OP_MUL, R, C)) == flat(Mo * Vo)
    l = min(p.bit_length(), 37)
    for desc in (False, True):
        rows = orc.local_bits_decompose(v, l, descending=desc)
        assert ints(dev.bits_decompose_flat(V, l, descending=desc)) == [b for row in rows for b in row]


def test_transpose_twice_and_cumsum_of_ones_at_size():
    ctx = mpyc_b200.context_for(P128)
    R, C = 38, 200_003
    M = DeviceArray.random(ctx, R * C, seed=47, stream_id=1)
    assert dev.transpose(dev.transpose(M, R, C), C, R).count_mismatch(M) == 0
    ones = dev.axpb(M, 0, 1)
    cs = dev.cumsum_rows(ones, R, C)
    assert ints(DeviceArray(ctx, cs.t[(R - 1) * C:(R - 1) * C + 3])) == [R, R, R]
    last = DeviceArray(ctx, dev.cumsum_rows(M, R, C).t[(R - 1) * C:])
    total = DeviceArray(ctx, M.t[:C])
    for j in range(1, R):
        total = total + DeviceArray(ctx, M.t[j * C:(j + 1) * C])
    assert last.count_mismatch(total) == 0


# ---- size-independent properties at sizes the oracle does not reach --------------------------------------------

@pytest.mark.parametrize('p,l', [(P64, 37), (P128, 64), (P256, 40), (P69, 38), (GEN[128], 16)])
def test_compose_of_decompose_is_low_bits(p, l):
    """sum_j bit_j(c) 2^j = c & (2^l - 1): decompose -> transpose -> compose against low_bits, both bit orders."""
    ctx = mpyc_b200.context_for(p)
    n = 300_001
    C = DeviceArray.random(ctx, n, seed=23, stream_id=1)
    want = dev.low_bits(C, l)
    for desc in (False, True):
        M = dev.bits_decompose(C, l, descending=desc)               # (l, n)
        flat = M.t.permute(1, 0, 2).contiguous().reshape(n * l, -1)    # (n, l) row-major
        got = dev.bits_compose(DeviceArray(ctx, flat), n, l, descending=desc)
        assert got.count_mismatch(want) == 0


@pytest.mark.parametrize('p', [P64, P128, GEN[128], P256])
def test_compose_is_linear(p):
    """compose(x + y) = compose(x) + compose(y) and compose(s x) = s compose(x) on share-like (arbitrary) residues."""
    ctx = mpyc_b200.context_for(p)
    n, f = 200_003, 37
    X = DeviceArray.random(ctx, n * f, seed=29, stream_id=1)
    Y = DeviceArray.random(ctx, n * f, seed=29, stream_id=2)
    lhs = dev.bits_compose(X + Y, n, f)
    rhs = dev.bits_compose(X, n, f) + dev.bits_compose(Y, n, f)
    assert lhs.count_mismatch(rhs) == 0
    s = 0x1234567890ABCDEF1234567 % p
    assert dev.bits_compose(X * s, n, f, descending=True).count_mismatch(dev.bits_compose(X, n, f, descending=True) * s) == 0


@pytest.mark.parametrize('p', [P64, P128, P64G, P256])
def test_fma_axpb_agree_with_binops(p):
    ctx = mpyc_b200.context_for(p)
    n = 1_000_003
    A, B, C = (DeviceArray.random(ctx, n, seed=31, stream_id=i) for i in (1, 2, 3))
    assert dev.fma(A, B, C).count_mismatch(A * B + C) == 0
    assert dev.fma(A, None, C).count_mismatch(A * A + C) == 0
    s, t = (p + 1) >> 1, p - 12345
    assert dev.axpb(A, s, t).count_mismatch(A * s + t) == 0
    assert dev.axpb(A, 1, t).count_mismatch(A + t) == 0
    _, nz = dev.nonzero(A - A, want_mask=False)
    assert nz == 0
    _, nz = dev.nonzero(A * A + 1 - A * A, want_mask=False)
    assert nz == n


@pytest.mark.parametrize('p', [P69, P128])
def test_conv2d_with_a_delta_filter_is_the_identity_plus_bias(p):
    ctx = mpyc_b200.context_for(p)
    k, r, m, n, v, s = 2, 3, 28, 28, 3, 5
    X = DeviceArray.random(ctx, k * r * m * n, seed=37, stream_id=1)
    W = [0] * (v * r * s * s)
    for j in range(v):                       # output channel j copies input channel j
        W[((j * r + j) * s + s // 2) * s + s // 2] = 1
    bias = [5, p - 1, 0]
    Y = dev.conv2d(X, DeviceArray.from_ints(ctx, W), DeviceArray.from_ints(ctx, bias), k, r, m, n, v, s)
    x = np.array(ints(X), dtype=object).reshape(k, r, m, n)
    want = (x + np.array(bias, dtype=object)[None, :, None, None]) % p
    assert ints(Y) == [int(t) for t in want.reshape(-1)]
