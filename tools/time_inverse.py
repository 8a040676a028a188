"""Times mpyc_b200_ff_inv: Montgomery-trick batches (out of place) vs one Fermat exponentiation per element
(the in-place path) on device-resident arrays.  CUDA events, 3 warm-ups, best of 5."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpyc_b200
from mpyc_b200 import _cabi
from mpyc_b200.device import DeviceArray

lib = _cabi.lib
for p, n in ((2**64 - 189, 1 << 24), (2**128 - 173, 1 << 24), (2**256 - 189, 1 << 23), (9409569905028393239, 1 << 24)):
    ctx = mpyc_b200.context_for(p)
    A = DeviceArray.random(ctx, n, seed=1, stream_id=1)
    out = A._like()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    res = {}
    for name, dst in (('batched', out), ('fermat_in_place', None)):
        best = 1e9
        for it in range(8):
            src = A if dst is not None else DeviceArray(ctx, A.t.clone())
            d = dst if dst is not None else src
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            _cabi.check(lib.mpyc_b200_ff_inv(ctx.handle, src.ptr, d.ptr, n, st))
            e1.record()
            torch.cuda.synchronize()
            if it >= 3:
                best = min(best, e0.elapsed_time(e1))
        res[name] = best
    print('p=%d bits n=%d: batched %.3f ms (%.3e inv/s)  per-element Fermat %.3f ms (%.3e inv/s)  speed-up %.1fx' % (
        p.bit_length(), n, res['batched'], n / res['batched'] * 1e3, res['fermat_in_place'], n / res['fermat_in_place'] * 1e3,
        res['fermat_in_place'] / res['batched']))
