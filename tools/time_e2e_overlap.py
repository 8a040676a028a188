"""Where does the host-buffer (e2e) step spend its time?  C3 shape, pinned host buffers, n = 2^24 per call:
split_host alone, recombine_host alone, both in a steady-state loop from two host threads (separate workspaces in the
library), and the raw link (bench.pcie_probe).  Prints one JSON line."""
import ctypes, json, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import mpyc_b200
from mpyc_b200 import _cabi
from mpyc_b200._cabi import lib, check

bench.bind_to_gpu_numa(0)
p, m, t, k, ne = 2**128 - 173, 5, 2, 3, 1 << 24
ctx = mpyc_b200.context_for(p)
L = ctx.nlimbs
hs = torch.randint(0, 2**62, (ne, L), dtype=torch.int64).pin_memory()
hc = torch.randint(0, 2**62, (t, ne, L), dtype=torch.int64).pin_memory()
hsh = [torch.empty((m, ne, L), dtype=torch.int64).pin_memory() for _ in range(2)]
hout = torch.empty((1, ne, L), dtype=torch.int64).pin_memory()
xs = [1, 2, 3]
rowp = [_cabi.ptr_array([h[x - 1].data_ptr() for x in xs]) for h in hsh]
xs_c, xr_c = _cabi.i64_array(xs), _cabi.i64_array([0])

def split(b): check(lib.mpyc_b200_shamir_split_host(ctx.handle, hs.data_ptr(), hc.data_ptr(), ne, hsh[b].data_ptr(), ne, ne, t, m, 0))
def rec(b): check(lib.mpyc_b200_shamir_recombine_host(ctx.handle, rowp[b], xs_c, k, xr_c, 1, hout.data_ptr(), ne, ne, 0))
def timed(fn, reps=6):
    fn(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    return 1e3 * (time.perf_counter() - t0) / reps
split(0); split(1); rec(0)
out = {'split_alone_ms': timed(lambda: split(0)), 'recombine_alone_ms': timed(lambda: rec(0))}
reps = 8
def loop(fn): 
    for _ in range(reps): fn()
ta, tb = threading.Thread(target=loop, args=(lambda: split(0),)), threading.Thread(target=loop, args=(lambda: rec(1),))
t0 = time.perf_counter(); ta.start(); tb.start(); ta.join(); tb.join()
out['both_concurrent_ms_per_pair_of_calls'] = 1e3 * (time.perf_counter() - t0) / reps
def fused(b):
    check(lib.mpyc_b200_shamir_reshare_step_host(ctx.handle, hs.data_ptr(), hc.data_ptr(), ne, hsh[b].data_ptr(), ne, ne, t, m,
                                                 rowp[1 - b], xs_c, k, xr_c, 1, hout.data_ptr(), ne, ne, 0))
out['fused_reshare_step_ms'] = timed(lambda: fused(0))
out['chunk_mb'] = sys.argv[1] if len(sys.argv) > 1 else '32'
out['pcie_probe'] = bench.pcie_probe() if out['chunk_mb'] == '32' else None
eb = 16
out['bytes'] = {'split_h2d': (1 + t) * eb * ne, 'split_d2h': m * eb * ne, 'rec_h2d': k * eb * ne, 'rec_d2h': eb * ne}
print(json.dumps(out))
