"""Algorithm check of mpyc_b200/csrc/ff_arith.cuh on the HOST (no GPU): the reduction families
(pseudo-Mersenne aligned / shifted, generic Montgomery with guard limb), lazy accumulation bounds and
the pow ladder are compiled with g++ (tests/native/host_check.cpp) and compared with Python integers.
The PTX carry-chain primitives themselves are covered by the -m gpu parity tests."""
import os
import shutil
import subprocess

import pytest


HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'native', 'host_check.cpp')
BIN = os.path.join(HERE, 'native', '_build', 'host_check')

from arith_vectors import PRIMES, commands, expected_kind


@pytest.fixture(scope='module')
def checker():
    if shutil.which('g++') is None:
        pytest.skip('g++ not available')
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    subprocess.run(['g++', '-std=c++17', '-O2', '-fno-strict-aliasing', '-x', 'c++', SRC, '-o', BIN], check=True)

    def run(lines):
        r = subprocess.run([BIN], input='\n'.join(lines) + '\n', capture_output=True, text=True, check=True)
        return r.stdout.split()
    return run


@pytest.mark.parametrize('p', PRIMES, ids=lambda p: f'p{p.bit_length()}_{p & 0xffff:x}')
def test_field_arithmetic_on_host(checker, p):
    lines, want, (L, k) = commands(p)
    out = checker(lines)
    kind, Lr, kr = (int(x) for x in out[:3])
    assert (kind, Lr, kr) == (expected_kind(p), L, k)
    got = out[3:]
    assert len(got) == len(want)
    bad = [(lines[i + 1], got[i], f'{want[i]:x}') for i in range(len(want)) if got[i] != f'{want[i]:x}']
    assert not bad, bad[:5]
