"""Times K1c (mpyc_b200_ff_matmul, FiniteFieldArray.__matmul__, finfields.py:1126-1146) on device-resident operands at
np_cnnmnist's fully connected layer (1 x 3136 @ 3136 x 1024, demos/np_cnnmnist.py) and on a square batch shape, against
NumPy's object-dtype matmul + % p on one host core (the reference's path) on a bounded sample."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpyc_b200                                   # noqa: E402
from mpyc_b200 import device as dev                # noqa: E402
from mpyc_b200.device import DeviceArray          # noqa: E402

for p in (2**64 - 189, 2**256 - 189, 9409569905028393239):
    ctx = mpyc_b200.context_for(p)
    for r, k, c in ((1, 3136, 1024), (128, 3136, 1024), (1024, 1024, 1024)):
        A = DeviceArray.random(ctx, r * k, seed=1, stream_id=1)
        B = DeviceArray.random(ctx, k * c, seed=2, stream_id=2)
        for _ in range(3):
            C = dev.matmul(ctx, A, B, r, k, c)
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            C = dev.matmul(ctx, A, B, r, k, c)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        line = {'p_bits': p.bit_length(), 'kind': ctx.kind, 'shape': [r, k, c], 'ms': best, 'modmacs_per_s': r * k * c / (best * 1e-3)}
        if r * k * c <= 4_000_000:          # CPU reference path on the same shape (object matmul, then % p)
            a = np.array(A.to_ints(), dtype=object).reshape(r, k)
            b = np.array(B.to_ints(), dtype=object).reshape(k, c)
            t0 = time.perf_counter()
            want = (a @ b) % p
            dt = time.perf_counter() - t0
            assert want.reshape(-1).tolist() == C.to_ints().tolist()
            line['cpu_object_matmul_modmacs_per_s'] = r * k * c / dt
        print(json.dumps(line), flush=True)
