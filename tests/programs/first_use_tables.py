"""Stress program for the table cache of the host-buffer entry points: every recombination below uses a set of
x-coordinates this process has not used before, so each call uploads a fresh Lagrange table and launches the kernel that
stages it right away.  Run several copies concurrently on one GPU (tests/test_gpu_concurrent.py): with the GPU
time-sliced between processes, a table upload that is not complete when the kernel starts shows up as a wrong result
(the bug real 3-party MPyC runs exposed in round 2).  Results are checked against Python-int Lagrange interpolation.
Prints "<tag> done, mismatches: N"."""
import itertools
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import fakefield                      # noqa: E402
from mpyc_b200 import thresha         # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else 'x'
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40


def lagrange(p, xs, rows, x_r=0):
    lam = []
    for i, xi in enumerate(xs):
        num = den = 1
        for j, xj in enumerate(xs):
            if i != j:
                num = num * (x_r - xj) % p
                den = den * (xi - xj) % p
        lam.append(num * pow(den, -1, p) % p)
    return [sum(c * int(r[h]) for c, r in zip(lam, rows)) % p for h in range(len(rows[0]))]


bad = 0
fresh = itertools.permutations(range(1, 12), 3)          # 990 distinct ordered triples: a new table per call
for rep in range(reps):
    for p in (2**64 - 189, 2**128 - 173, 2**80 - 65, 9409569905028393239, 2**256 - 189):
        F = fakefield.make_prime_field(p)
        for n in (48, 1000):
            rng = np.random.default_rng(1000 * rep + n)
            a = np.array([int(v) % p for v in rng.integers(-1000, 1000, size=n)], dtype=object)
            m = 3 + rep % 5                               # split tables for a fresh (t, m) now and then
            sh = thresha.np_random_split(F, a, 1, m)
            if thresha.np_recombine(F, [(m, sh[m - 1]), (1, sh[0])]).value.tolist() != a.tolist():
                bad += 1
                print(tag, 'split/recombine MISMATCH', p.bit_length(), n, m)
            rows = [np.array([int(v) * 3 % p for v in rng.integers(0, 2**62, size=n)], dtype=object) for _ in range(3)]
            xs = list(next(fresh))
            if thresha.np_recombine(F, list(zip(xs, rows))).value.tolist() != lagrange(p, xs, rows):
                bad += 1
                print(tag, 'k=3 MISMATCH', p.bit_length(), n, xs, 'rep', rep)
print(tag, 'done, mismatches:', bad)
