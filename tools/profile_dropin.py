"""cProfile of the drop-in path (np_random_split + np_recombine on object arrays), object rows vs limb wire."""
import cProfile
import os
import pickle
import pstats
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpyc_b200 import thresha   # noqa: E402

p, m, t, k, n = 2**128 - 173, 5, 2, 3, 200_000


class Arr:
    def __init__(self, value, check=True):
        self.value = value


class Field:
    modulus = order = characteristic = p
    ext_deg = 1
    array = Arr


rnd = random.Random(1)
s = np.array([rnd.randrange(p) for _ in range(n)], dtype=object)
thresha.np_random_split(Field, s[:1000], t, m)


def run(limb_wire):
    thresha.limb_wire = limb_wire
    sh = thresha.np_random_split(Field, s, t, m)
    sent = [pickle.dumps(row) for row in sh]
    rows = [pickle.loads(x) for x in sent[:k]]
    return thresha.np_recombine(Field, [(i + 1, rows[i]) for i in range(k)])


for lw in (False, True):
    run(lw)
    pr = cProfile.Profile()
    pr.enable()
    run(lw)
    pr.disable()
    print('==== limb_wire =', lw)
    pstats.Stats(pr).sort_stats('cumulative').print_stats(14)
