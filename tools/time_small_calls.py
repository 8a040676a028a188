"""Per-call latency of the drop-in functions at the demos' small batch shapes (BASELINE configs[0] and configs[3]:
n = 1024 at 61 bits, n = 16/32 over GF(2^8)): mpyc_b200.thresha.np_random_split + np_recombine on object arrays, as
runtime.py calls them.  The reference's own numbers for the same calls are measured in the build container
 -- /root/reference does not exist on the GPU box."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import fakefield                                    # noqa: E402
from mpyc_b200 import thresha                       # noqa: E402

cases = [('gf256', fakefield.make_gf256(283), 16, 3, 1, 3), ('gf256', fakefield.make_gf256(283), 32, 3, 1, 3),
         ('p61', fakefield.make_prime_field(2**61 - 1), 16, 3, 1, 2), ('p61', fakefield.make_prime_field(2**61 - 1), 1024, 3, 1, 2),
         ('p256', fakefield.make_prime_field(2**256 - 189), 3136, 7, 3, 7)]
for name, F, n, m, t, k in cases:
    if name == 'gf256':
        s = np.array([fakefield.Poly(i % 256) for i in range(n)], dtype=object)
    else:
        s = np.array([(i * 0x9E3779B97F4A7C15) % F.modulus for i in range(n)], dtype=object)
    for limb_wire in (False, True):
        thresha.limb_wire = limb_wire
        for _ in range(20):
            sh = thresha.np_random_split(F, s, t, m)
            y = thresha.np_recombine(F, [(i + 1, sh[i]) for i in range(k)])
        reps = 300
        t0 = time.perf_counter()
        for _ in range(reps):
            sh = thresha.np_random_split(F, s, t, m)
        t1 = time.perf_counter()
        for _ in range(reps):
            y = thresha.np_recombine(F, [(i + 1, sh[i]) for i in range(k)])
        t2 = time.perf_counter()
        assert [int(v) for v in y.value] == [int(v) for v in s]
        print(json.dumps({'field': name, 'n': n, 'm': m, 't': t, 'k': k, 'limb_wire': limb_wire,
                          'us_split': (t1 - t0) / reps * 1e6, 'us_recombine': (t2 - t1) / reps * 1e6}), flush=True)
thresha.limb_wire = False
