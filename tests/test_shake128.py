"""The library's host SHAKE128 (mpyc_b200/csrc/shake128.h, the XOF of thresha.PRF, mpyc/thresha.py:257) against
hashlib.shake_128: FIPS 202 known answers, every input length around the 168-byte rate, long outputs."""
import ctypes
import hashlib

import pytest

from mpyc_b200 import _cabi


def shake(msg, n):
    out = ctypes.create_string_buffer(max(n, 1))
    _cabi.check(_cabi.lib.mpyc_b200_shake128(msg, len(msg), out, n))
    return out.raw[:n]


def test_known_answer_empty_message():
    # FIPS 202 / NIST CAVP: SHAKE128(""), first 32 bytes
    assert shake(b'', 32).hex() == '7f9c2ba4e88f827d616045507605853ed73b8093f6efbc88eb1a6eacfa66ef26'


@pytest.mark.parametrize('inlen', [0, 1, 15, 16, 24, 135, 136, 166, 167, 168, 169, 335, 336, 337, 1000])
def test_against_hashlib_around_the_rate(inlen):
    msg = bytes((7 * i + inlen) & 0xFF for i in range(inlen))
    for outlen in (0, 1, 16, 167, 168, 169, 336, 1000, 5000):
        assert shake(msg, outlen) == hashlib.shake_128(msg).digest(outlen)


def test_long_stream_like_a_prss_call():
    key, uci = bytes(range(16)), b'\x00\x01\x02uci'
    n = 213_248 * 48 // 16          # a sixteenth of the largest np_cnnmnist PRSS stream (n = 213,248, 48-byte chunks)
    assert shake(key + uci, n) == hashlib.shake_128(key + uci).digest(n)


def test_lock_step_sponges_match_hashlib():
    """mpyc_b200_shake128_multi: the XOF streams of one PRSS call, eight sponges at a time in the AVX-512 lock-step form
    where the host has it (scalar otherwise) -- every stream equals hashlib.shake_128(key_i + suffix), for sponge
    counts around the group size and output lengths around the rate."""
    import ctypes
    import hashlib
    from mpyc_b200 import _cabi
    for count in (1, 2, 3, 7, 8, 9, 16, 20):
        for key_bytes, suffix in ((16, b'uci-0001'), (16, b''), (0, b'x' * 200), (7, bytes(range(167)))):
            keys = bytes((31 * i + 7) & 0xFF for i in range(count * key_bytes))
            for outlen in (1, 167, 168, 169, 335, 336, 337, 5000):
                stride = outlen + 5
                out = ctypes.create_string_buffer(count * stride)
                wide = ctypes.c_int(-1)
                _cabi.check(_cabi.lib.mpyc_b200_shake128_multi(keys, key_bytes, suffix, len(suffix), count, out, stride, outlen,
                                                               ctypes.byref(wide)))
                assert wide.value in (0, 1)
                for i in range(count):
                    want = hashlib.shake_128(keys[i * key_bytes:(i + 1) * key_bytes] + suffix).digest(outlen)
                    assert out.raw[i * stride:i * stride + outlen] == want, (count, key_bytes, outlen, i)
