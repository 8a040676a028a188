"""Benchmark of the Shamir hot path: BASELINE.json's metric "GF(p) Shamir share+recombine pairs/sec".

    python bench.py --gpus 1 --steps 10 --warmup 3             (N > 1: launched with torch.distributed.run)
    python bench.py --impl reference --steps 3 --warmup 1       (CPU port of the reference's algorithm)

One step = one pass of the hot path over one batch: share generation of n secrets (degree t, m
parties; coefficient matrix resident in HBM, "parity mode") followed by Lagrange recombination of
t+1 of the resulting share rows.  One *pair* = one secret split + recombined.

Workloads (--workload):
    c3   (default, BASELINE.json configs[2]) p = 2^128-173, m=5, t=2, n = 10^8 per GPU
    ns64 (north_star's 64-bit case)          p = 2^64-189,  m=3, t=1, n = 2*10^8 per GPU
    c5   (configs[4] field/shape)            p = 2^256-189, m=7, t=3, n = 2*10^7 per GPU, recombine 2t+1
    modmul (configs[1])                      p = 2^64-189 elementwise a*b, n = 10^8 (value = elem/s)

The element axis is sharded over the GPUs with no data-path collective (weak scaling: n per GPU fixed).
Printed JSON line: see the driver's contract; extra keys `roofline`, `cpu_baseline`, `e2e`, `clocks`.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    'c3': dict(p=2**128 - 173, m=5, t=2, k=3, n=100_000_000, name='shamir split+recombine p=2^128-173 m=5 t=2 n=1e8/GPU (BASELINE configs[2])'),
    'ns64': dict(p=2**64 - 189, m=3, t=1, k=2, n=200_000_000, name='shamir split+recombine p=2^64-189 m=3 t=1 n=2e8/GPU (north_star 64-bit case)'),
    'c5': dict(p=2**256 - 189, m=7, t=3, k=7, n=20_000_000, name='shamir reshare p=2^256-189 m=7 t=3 recombine 2t+1 n=2e7/GPU (configs[4] shape)'),
    'modmul': dict(p=2**64 - 189, m=0, t=0, k=0, n=100_000_000, name='elementwise modmul p=2^64-189 n=1e8/GPU (BASELINE configs[1])'),
    'modmul_generic': dict(p=9409569905028393239, m=0, t=0, k=0, n=100_000_000, name='elementwise modmul generic 64-bit prime 9409569905028393239 (Montgomery) n=1e8/GPU (BASELINE configs[1])'),
    'c3g': dict(p=0x800000000000000000000000000000fb, m=5, t=2, k=3, n=100_000_000, name='shamir split+recombine GENERIC 128-bit prime (Montgomery path) m=5 t=2 n=1e8/GPU'),
    'prss': dict(p=2**256 - 189, m=7, t=3, k=0, n=1 << 21, prss=True, name='PRSS np_pseudorandom_share p=2^256-189 m=7 t=3 (20 key subsets, 48-byte PRF chunks): combine kernel on n=2^21 resident bytes; e2e at np_cnnmnist call size n=213,248 (configs[4] shape)'),
    'c4': dict(p=283, binary=True, m=3, t=1, k=3, n=1 << 28, name='GF(2^8) reshare (np_aes field, modulus 283) m=3 t=1 recombine 2t+1, batched n=2^28 bytes/GPU + per-call latency at n=16 (BASELINE configs[3] shape)'),
}
METRIC = 'GF(p) Shamir share+recombine pairs/sec'


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


def ncu_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` summary of the same workload (profiles/), or (None, None)."""
    import glob
    import re
    for path in sorted(glob.glob(os.path.join(ROOT, 'profiles', f'r*_ncu_{kernel}_{workload}.txt')), reverse=True):
        total, units = 0.0, {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Tbyte': 1e12}
        for line in open(path):
            mt = re.match(r'\s*dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+(\w+)', line)
            if mt:
                total += float(mt.group(2)) * units.get(mt.group(3), 1.0)
        if total:
            return total, os.path.relpath(path, ROOT)
    return None, None


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every few ms from a
    thread (nvidia_ml_py), falling back to `nvidia-smi -lms` when NVML cannot be loaded."""
    BAD = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown', 0x4: 'sw_power_cap'}

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.stop_flag, self.thread, self.max = index, [], set(), False, None, None
        self.power = []

    def _poll(self):
        import pynvml as nv
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        self.max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        while not self.stop_flag:
            self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            try:
                self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            except Exception:
                mask = 0
            for bit, name in self.BAD.items():
                if mask & bit:
                    self.reasons.add(name)
            time.sleep(0.002)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            time.sleep(0.01)
        except Exception as exc:   # noqa: BLE001
            self.thread = None
            self.error = repr(exc)

    def stop(self):
        if self.thread is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvml unavailable: ' + getattr(self, 'error', '?')]}
        self.stop_flag = True
        self.thread.join(timeout=2)
        sm = sorted(self.samples)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': self.max, 'samples': len(sm),
                'power_w_max': max(self.power) if self.power else None, 'reasons': sorted(self.reasons), 'source': 'nvml'}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's NumPy-object path (np_random_split + np_recombine)
# ---------------------------------------------------------------------------------------------------

def _cpu_pairs(args):
    """One process: split+recombine `n` pairs `reps` times with the reference's own draw (secrets.randbelow)."""
    p, m, t, k, n, reps, with_rng = args
    import secrets
    import numpy as np
    from oracle import shamir_oracle as orc
    s = np.array(orc.synth_elements(p, n, 20260923), dtype=object)
    C = np.array(orc.np_stream_to_C(orc.synth_elements(p, t * n, 7, stream=2), t, n), dtype=object) if t else np.empty((0, n), dtype=object)
    xs = tuple(range(1, k + 1))
    t0 = time.perf_counter()
    for _ in range(reps):
        Cr = orc.np_draw_coefficients(p, t, n, secrets.randbelow) if with_rng else C
        sh = orc.np_split(p, s, Cr, m)
        out = orc.np_recombine(p, xs, sh[:k])
    dt = time.perf_counter() - t0
    if not with_rng:
        assert out.tolist() == s.tolist()
    return n * reps, dt


def cpu_baseline(w, n=20000, target_s=12.0):
    """1 core, bounded sample; returns the cpu_baseline object (kind 'port': the reference is Python and
    cannot travel to the GPU box; oracle/shamir_oracle.py is its restatement, pinned by golden fixtures)."""
    pairs, dt = _cpu_pairs((w['p'], w['m'], w['t'], w['k'], n, 1, True))
    reps = max(1, int(target_s * 0.7 / max(dt, 1e-3)))
    pairs, dt = _cpu_pairs((w['p'], w['m'], w['t'], w['k'], n, reps, True))
    pairs2, dt2 = _cpu_pairs((w['p'], w['m'], w['t'], w['k'], n, max(1, reps // 2), False))
    return {'value': pairs / dt, 'unit': 'pairs/s', 'cores': 1, 'kind': 'port',
            'sample': f'np_split+np_recombine (NumPy object arrays, coefficients via secrets.randbelow as the reference does) '
                      f'on {n} elements x {reps} reps, 1 process', 'value_without_rng': pairs2 / dt2}


_REF = {}


def _ref_init(p, m, t, k, n):
    """Pool initializer: every worker builds its inputs once (outside the timed steps)."""
    import numpy as np
    from oracle import shamir_oracle as orc
    _REF.update(p=p, m=m, t=t, k=k, n=n, s=np.array(orc.synth_elements(p, n, 20260923 + os.getpid() % 1000), dtype=object))


def _ref_step(_):
    """One worker's share of a step: draw coefficients as the reference does, split, recombine t+1 shares."""
    import secrets
    from oracle import shamir_oracle as orc
    r = _REF
    C = orc.np_draw_coefficients(r['p'], r['t'], r['n'], secrets.randbelow)
    sh = orc.np_split(r['p'], r['s'], C, r['m'])
    out = orc.np_recombine(r['p'], tuple(range(1, r['k'] + 1)), sh[:r['k']])
    assert out[0] == r['s'][0]
    return r['n']


def run_reference_arm(a, w):
    """CPU arm: the oracle port of thresha.np_random_split + np_recombine (NumPy object arrays, per-element
    secrets.randbelow -- the reference's own code path) on all host cores, one independent process per core.
    The reference is pure Python and cannot travel to the GPU box, hence kind = 'port'."""
    import multiprocessing as mp
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = 20000
    done = []
    with mp.get_context('fork').Pool(cores, initializer=_ref_init, initargs=(w['p'], w['m'], w['t'], w['k'], n)) as pool:
        pool.map(_ref_step, range(cores))          # make sure every worker is up before timing
        for step in range(a.warmup + a.steps):
            t0 = time.perf_counter()
            res = pool.map(_ref_step, range(cores), chunksize=1)
            dt = time.perf_counter() - t0
            if step >= a.warmup:
                done.append((sum(res), dt))
    pairs = sum(x for x, _ in done)
    dt = sum(y for _, y in done)
    val = pairs / dt
    sample = f'{cores} independent processes x {n} pairs per step (oracle port of thresha.np_random_split+np_recombine incl. secrets.randbelow)'
    line = {'metric': METRIC, 'impl': 'reference', 'value': val, 'unit': 'pairs/s', 'n_gpus': a.gpus, 'steps': a.steps,
            'warmup': a.warmup, 'ms_per_step': 1e3 * dt / max(a.steps, 1), 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'python-int (exact)', 'data': 'synthetic',
            'config': {'workload': w['name'], 'sample_per_step': f'{n} pairs x {cores} processes'},
            'cpu_baseline': {'value': val, 'unit': 'pairs/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': val, 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------

def dropin_rate(w, device, n=200_000):
    """The rate an unmodified MPyC caller sees: mpyc_b200.thresha.np_random_split + np_recombine on NumPy
    object arrays of Python ints (conversion to limbs, CSPRNG draw, PCIe and kernels all inside)."""
    import random
    import numpy as np
    from mpyc_b200 import thresha
    p, m, t, k = w['p'], w['m'], w['t'], w['k']

    class Arr:
        def __init__(self, value, check=True):
            self.value = value

    class Field:
        modulus = order = characteristic = p
        ext_deg = 1
        array = Arr
    thresha.device = device
    rnd = random.Random(20260923)
    s = np.array([rnd.randrange(p) for _ in range(n)], dtype=object)
    thresha.np_random_split(Field, s[:1000], t, m)                     # warm-up (tables, workspace)
    t0 = time.perf_counter()
    sh = thresha.np_random_split(Field, s, t, m)
    out = thresha.np_recombine(Field, [(i + 1, sh[i]) for i in range(k)])
    dt = time.perf_counter() - t0
    assert out.value.tolist() == s.tolist()
    # limb wire (mpyc_b200.wire): rows stay limb buffers; as in runtime.py:655-665 every row is pickled for its
    # peer and the k received rows are unpickled before np_recombine -- pickling is INSIDE this timed region
    import pickle
    thresha.limb_wire = True
    try:
        t0 = time.perf_counter()
        sh = thresha.np_random_split(Field, s, t, m)
        sent = [pickle.dumps(row) for row in sh]
        out = thresha.np_recombine(Field, [(i + 1, pickle.loads(sent[i])) for i in range(k)])
        dt_lw = time.perf_counter() - t0
    finally:
        thresha.limb_wire = False
    assert out.value.tolist() == s.tolist()
    return {'value': n / dt, 'unit': 'pairs/s', 'n': n,
            'path': 'mpyc_b200.thresha.np_random_split + np_recombine on dtype=object arrays (what runtime.py calls)',
            'limb_wire': {'value': n / dt_lw, 'unit': 'pairs/s',
                          'path': 'same calls with install(limb_wire=True): ShareRow rows, pickle.dumps of all m rows and '
                                  'pickle.loads of the k recombined ones inside the timed region',
                          'wire_bytes_per_row': len(sent[0])}}


def run_prss_arm(a, w):
    """PRSS (thresha.np_pseudorandom_share, mpyc/thresha.py:163-173) at the np_cnnmnist shape.  value: elements/s of the
    combine kernel K4 on PRF bytes resident in HBM; e2e: the drop-in call on the host (SHAKE128 sponges on host threads,
    pinned chunks, H2D, K4, D2H, ints) at the largest per-call size of the demo; cpu_baseline: the oracle port."""
    import itertools
    import numpy as np
    import torch
    import mpyc_b200
    from mpyc_b200 import _cabi, thresha
    from mpyc_b200._cabi import lib, check
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (there is no CPU fallback for the product arm)')
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    p, m, t, n = w['p'], w['m'], w['t'], a.n or w['n']
    ctx = mpyc_b200.context_for(p)
    L, eb = ctx.nlimbs, ctx.elem_bytes
    party = 1
    subsets = [S for S in itertools.combinations(range(m), m - t) if party in S]
    nsub, d = len(subsets), 1
    prf = {S: thresha.PRF(bytes([(31 * a_ + 7) % 256 for a_ in S] + [0] * (16 - len(S))), p) for S in subsets}
    chunk = next(iter(prf.values())).byte_length
    stride = (n * d * chunk + 15) // 16 * 16
    peak, peak_src = peaks()
    # device-resident PRF bytes: random bytes stand in for the XOF output (the kernel's work does not depend on them)
    g = torch.Generator(device='cuda')
    g.manual_seed(20260923 + rank)
    d_bytes = torch.randint(0, 256, (nsub, stride), dtype=torch.uint8, device='cuda', generator=g)
    d_out = torch.empty((n, L), dtype=torch.int64, device='cuda')
    nl = L
    coef = []
    for S in subsets:
        coef.extend(_cabi.int_to_limbs(int(thresha._f_S_i(_PrssField(p), m, party, S)), nl))
    wl = _cabi.int_to_limbs(1, nl)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def step():
        check(lib.mpyc_b200_prss_combine(ctx.handle, ctypes.c_void_p(d_bytes.data_ptr()), stride, nsub, d, chunk, 0,
                                         _cabi.u64_array(coef), _cabi.u64_array(wl), ctypes.c_void_p(d_out.data_ptr()), n, st))
    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = mpyc_b200.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    ev[0].record()
    for i in range(a.steps):
        step()
        ev[i + 1].record()
    torch.cuda.synchronize()
    launches = mpyc_b200.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[-1])
    if world > 1:
        tt = torch.tensor([total_ms], device='cuda')
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    ms = total_ms / a.steps
    alg = n * (nsub * d * chunk + eb)
    ach = alg / (ms * 1e-3) / 1e9
    roofline = {'bound': 'hbm', 'kernel': 'k_prss_tiles', 'achieved': ach, 'peak': peak, 'peak_source': peak_src, 'unit': 'GB/s',
                'frac': ach / peak, 'traffic': ncu_traffic('prss', 'prss')[0], 'ms': ms, 'algorithmic_bytes': alg,
                'note': 'includes the per-call table upload (cudaMallocAsync + 1 KB H2D + stream sync) of mpyc_b200_prss_combine'}
    # e2e through the drop-in call, host data in / ints out, at the demo's largest call size
    ne = 213_248
    F = _PrssField(p)
    thresha.device = local
    thresha.np_pseudorandom_share(F, m, party, prf, b'warm', 4096)
    reps, t0 = 3, time.perf_counter()
    for r in range(reps):
        res = thresha.np_pseudorandom_share(F, m, party, prf, b'uci%d' % r, ne)
    dt = (time.perf_counter() - t0) / reps
    out_limbs = np.empty((ne, L), dtype=np.uint64)
    keys = b''.join(f.key for f in prf.values())
    t0 = time.perf_counter()
    for r in range(reps):
        check(lib.mpyc_b200_prss_host(ctx.handle, keys, 16, b'uci%d' % r, 4, nsub, d, chunk, 0, _cabi.u64_array(coef),
                                      _cabi.u64_array(wl), ctypes.c_void_p(out_limbs.ctypes.data), ne, local, 0))
    dt_abi = (time.perf_counter() - t0) / reps
    e2e = {'value': world * ne / dt_abi, 'unit': 'shares/s', 'h2d_bytes_per_step': ne * nsub * d * chunk, 'd2h_bytes_per_step': ne * eb,
           'n_per_step': ne, 'ms_per_step': dt_abi * 1e3,
           'path': 'mpyc_b200_prss_host: SHAKE128 sponges on host threads -> pinned chunks -> H2D -> K4 -> D2H (host buffers in and out)',
           'xof_bytes_per_step': ne * nsub * d * chunk, 'host_threads': min(nsub, os.cpu_count() or 1),
           'dropin': {'value': ne / dt, 'unit': 'shares/s', 'path': 'mpyc_b200.thresha.np_pseudorandom_share (field.array of Python ints out)'}}
    cpu = None
    if rank == 0 and not a.no_cpu:
        from oracle import shamir_oracle as orc
        Fo = orc.field_of(p)
        nc = 10_000
        t0 = time.perf_counter()
        got = orc.prss_share(Fo, m, party, {S: orc.prf_values(f.key, p, b'uci0', nc) for S, f in prf.items()}, nc)
        dtc = time.perf_counter() - t0
        chk = thresha.np_pseudorandom_share(F, m, party, prf, b'uci0', nc)
        assert got == chk.value.tolist(), 'PRSS differs from the oracle'
        cpu = {'value': nc / dtc, 'unit': 'shares/s', 'cores': 1, 'kind': 'port',
               'sample': f'oracle port of np_pseudorandom_share (hashlib SHAKE128 + Python ints) on {nc} elements, 1 process'}
    if rank == 0:
        line = {'metric': 'PRSS pseudorandom shares/sec', 'value': world * n / (ms * 1e-3), 'unit': 'shares/s', 'n_gpus': world,
                'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': f'u64x{L} limbs (exact integer arithmetic mod p)', 'data': 'synthetic',
                'config': {'workload': w['name'], 'p_bits': p.bit_length(), 'm': m, 't': t, 'subsets': nsub, 'chunk_bytes': chunk, 'n_per_gpu': n,
                           'l2_policy': 'PRF bytes (2 GB) far larger than the 126 MB L2; no flush needed'},
                'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': int(launches), 'clocks': clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


class _PrssArr:
    def __init__(self, value, check=True):
        self.value = value


def _PrssField(p):
    class Field:
        modulus = order = characteristic = p
        ext_deg = 1
        array = _PrssArr
    return Field


def run_gpu_arm(a, w):
    if w.get('prss'):
        return run_prss_arm(a, w)
    import torch
    import torch.distributed as dist
    import mpyc_b200
    from mpyc_b200 import _cabi, device as dev
    from mpyc_b200._cabi import lib, check

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (there is no CPU fallback for the product arm)')
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    p, m, t, k, n = w['p'], w['m'], w['t'], w['k'], a.n or w['n']
    ctx = mpyc_b200.context_for(p, binary=bool(w.get('binary')))
    L = max(ctx.nlimbs, 1)
    eb = ctx.elem_bytes
    if w.get('binary'):
        a.no_e2e = True   # host-buffer pipeline is exercised by the prime-field workloads
    is_mul = w['m'] == 0
    peak, peak_src = peaks()

    # ---- device-resident inputs (synthetic, generated on the device; far larger than the 126 MB L2) ----
    S = dev.DeviceArray.random(ctx, n, seed=20260923 + rank, stream_id=1)
    if is_mul:
        B = dev.DeviceArray.random(ctx, n, seed=77 + rank, stream_id=2)
        OUT = dev.DeviceArray.empty(ctx, n)
    else:
        C = dev.DeviceMatrix.empty(ctx, t, n)
        for j in range(t):
            C.t[j].copy_(dev.DeviceArray.random(ctx, n, seed=100 + j + 10 * rank, stream_id=3).t)
        SH = dev.DeviceMatrix.empty(ctx, m, n)
        REC = dev.DeviceMatrix.empty(ctx, 1, n)
        xs = list(range(1, k + 1))
        rows = [SH.row(x - 1) for x in xs]
    torch.cuda.synchronize()

    def step():
        if is_mul:
            check(lib.mpyc_b200_ff_binop(ctx.handle, _cabi.OP_MUL, S.ptr, B.ptr, OUT.ptr, n, dev._stream_ptr()))
        else:
            dev.shamir_split(ctx, S, C, t, m, out=SH)
            dev.shamir_recombine(ctx, xs, rows, 0, out=REC)

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    if not is_mul:   # correctness guard inside the bench: the recombined secrets are the inputs
        assert REC.row(0).count_mismatch(S) == 0, 'recombined secrets differ from inputs'

    sampler = ClockSampler(local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    launches0 = mpyc_b200.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * a.steps + 1)]
    ev[0].record()
    for i in range(a.steps):
        if is_mul:
            step()
            ev[2 * i + 1].record()
        else:
            dev.shamir_split(ctx, S, C, t, m, out=SH)
            ev[2 * i + 1].record()
            dev.shamir_recombine(ctx, xs, rows, 0, out=REC)
        ev[2 * i + 2].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = mpyc_b200.launch_count() - launches0
    if world > 1:   # whole-job count
        tl = torch.tensor([launches], device='cuda', dtype=torch.int64)
        dist.all_reduce(tl)
        launches = int(tl.item())
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[-1])
    split_ms = sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(a.steps)) / a.steps
    rec_ms = 0.0 if is_mul else sum(ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(a.steps)) / a.steps
    if world > 1:
        tt = torch.tensor([total_ms], device='cuda', dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    ms_per_step = total_ms / a.steps
    value = world * n / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (share generation), algorithmic bytes / measured duration ----
    if is_mul:
        dom_name, dom_bytes, dom_ms = 'k_binop<mul>', 3 * eb * n, split_ms
        sec = {}
    else:
        split_bytes = (1 + t + m) * eb * n
        rec_bytes = (k + 1) * eb * n
        dom_name, dom_bytes, dom_ms = 'k_split', split_bytes, split_ms
        sec = {'recombine': {'kernel': 'k_recombine', 'achieved': rec_bytes / (rec_ms * 1e-3) / 1e9, 'unit': 'GB/s',
                             'frac': rec_bytes / (rec_ms * 1e-3) / 1e9 / peak, 'ms': rec_ms, 'algorithmic_bytes': rec_bytes},
               'step_total': {'achieved': (split_bytes + rec_bytes) / (ms_per_step * 1e-3) / 1e9, 'unit': 'GB/s',
                              'frac': (split_bytes + rec_bytes) / (ms_per_step * 1e-3) / 1e9 / peak,
                              'bytes_per_pair': (split_bytes + rec_bytes) / n}}
    ach = dom_bytes / (dom_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(a.workload, 'binop' if is_mul else 'split') if n == w['n'] else (None, None)
    roofline = {'bound': 'hbm', 'kernel': dom_name, 'achieved': ach, 'peak': peak, 'peak_source': peak_src, 'unit': 'GB/s',
                'frac': ach / peak, 'traffic': traffic, 'traffic_source': traffic_src, 'ms': dom_ms,
                'algorithmic_bytes': dom_bytes, **sec}

    # ---- e2e: same path through the C ABI's host-buffer entry points, pinned host memory ----------------
    e2e = None
    if not a.no_e2e:
        ne = a.e2e_n
        hs = torch.empty((ne, L), dtype=torch.int64).pin_memory()
        hs.copy_(S.t[:ne].cpu())
        if is_mul:
            hb = torch.empty((ne, L), dtype=torch.int64).pin_memory()
            hb.copy_(B.t[:ne].cpu())
            ho = torch.empty((ne, L), dtype=torch.int64).pin_memory()

            def e2e_step():
                check(lib.mpyc_b200_ff_binop_host(ctx.handle, _cabi.OP_MUL, hs.data_ptr(), hb.data_ptr(), ho.data_ptr(), ne, local))
            h2d, d2h = 2 * eb * ne, eb * ne
        else:
            hc = torch.empty((t, ne, L), dtype=torch.int64).pin_memory()
            hc.copy_(C.t[:, :ne].cpu())
            hsh = torch.empty((m, ne, L), dtype=torch.int64).pin_memory()
            hout = torch.empty((1, ne, L), dtype=torch.int64).pin_memory()
            rowp = _cabi.ptr_array([hsh[x - 1].data_ptr() for x in xs])
            xs_c, xr_c = _cabi.i64_array(xs), _cabi.i64_array([0])

            def e2e_step():
                check(lib.mpyc_b200_shamir_split_host(ctx.handle, hs.data_ptr(), hc.data_ptr(), ne, hsh.data_ptr(), ne, ne, t, m, local))
                check(lib.mpyc_b200_shamir_recombine_host(ctx.handle, rowp, xs_c, k, xr_c, 1, hout.data_ptr(), ne, ne, local))
            h2d, d2h = (1 + t) * eb * ne + k * eb * ne, m * eb * ne + eb * ne
        for _ in range(max(1, min(a.warmup, 2))):
            e2e_step()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            e2e_step()
        dt = time.perf_counter() - t0
        if world > 1:   # every rank drives its own GPU over its own PCIe link; the job time is the slowest rank
            td = torch.tensor([dt], device='cuda', dtype=torch.float64)
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
            dt = float(td.item())
        if not is_mul:
            assert torch.equal(hout[0], hs), 'e2e: recombined secrets differ from inputs'
        e2e = {'value': world * ne * a.steps / dt, 'unit': 'pairs/s' if not is_mul else 'elem/s', 'h2d_bytes_per_step': h2d,
               'd2h_bytes_per_step': d2h, 'n_per_step': ne, 'ms_per_step': 1e3 * dt / a.steps,
               'path': 'mpyc_b200_shamir_split_host + mpyc_b200_shamir_recombine_host (pinned host buffers, copies inside)',
               'note': 'all ranks concurrently, max over ranks of the host wall clock around the blocking C-ABI calls'}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None if (a.no_cpu or is_mul or world > 1 or w.get('binary')) else cpu_baseline(w)
    small_call = None
    if w.get('binary'):
        # np_aes.py shape: calls on n = 16 bytes; launch latency, not bandwidth, is what counts there
        s16 = dev.DeviceArray.random(ctx, 16, seed=1, stream_id=1)
        c16 = dev.DeviceMatrix.empty(ctx, t, 16)
        c16.t[0].copy_(dev.DeviceArray.random(ctx, 16, seed=2, stream_id=2).t)
        sh16 = dev.DeviceMatrix.empty(ctx, m, 16)
        r16 = dev.DeviceMatrix.empty(ctx, 1, 16)
        rows16 = [sh16.row(i) for i in range(k)]
        for _ in range(20):
            dev.shamir_split(ctx, s16, c16, t, m, out=sh16)
            dev.shamir_recombine(ctx, list(range(1, k + 1)), rows16, 0, out=r16)
        torch.cuda.synchronize()
        reps = 2000
        t0 = time.perf_counter()
        for _ in range(reps):
            dev.shamir_split(ctx, s16, c16, t, m, out=sh16)
            dev.shamir_recombine(ctx, list(range(1, k + 1)), rows16, 0, out=r16)
        torch.cuda.synchronize()
        small_call = {'n': 16, 'us_per_split_plus_recombine': 1e6 * (time.perf_counter() - t0) / reps,
                      'note': 'device-resident, two kernel launches per pair, host-side issue rate included'}
    dropin = None
    if not (a.no_e2e or is_mul or world > 1 or w.get('binary')):
        dropin = dropin_rate(w, local)
    line = {'metric': METRIC if not is_mul else 'GF(p) modmul elem/sec', 'value': value,
            'unit': 'pairs/s' if not is_mul else 'elem/s', 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': ('u8 (GF(2^8) polynomial arithmetic)' if w.get('binary') else f'u64x{L} limbs (exact integer arithmetic mod p)'), 'data': 'synthetic',
            'config': {'workload': w['name'], 'p_bits': p.bit_length(), 'm': m, 't': t, 'recombine_k': k, 'n_per_gpu': n,
                       'parallelism': f'element axis sharded over {world} GPU(s), no data-path collective',
                       'l2_policy': 'inputs (>= 1.6 GB) far larger than the 126 MB L2; no flush needed',
                       'coefficients': 'resident in HBM (parity mode)'},
            'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'e2e_dropin': dropin, 'small_call': small_call,
            'gpu_launches': int(launches), 'clocks': clocks}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c3', choices=sorted(WORKLOADS))
    ap.add_argument('--n', type=int, default=0, help='elements per GPU (default: the workload size)')
    ap.add_argument('--e2e-n', type=int, default=1 << 24)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == 'ours' else a.warmup
    w = WORKLOADS[a.workload]
    if a.impl == 'reference':
        run_reference_arm(a, w)
    else:
        run_gpu_arm(a, w)


if __name__ == '__main__':
    main()
