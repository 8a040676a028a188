"""Two-GPU checks (skipped with fewer than 2 devices): sharded split+recombine equals the single-GPU result,
and the NCCL gather reassembles the element axis."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
    pytest.skip('needs 2 CUDA devices', allow_module_level=True)

import torch.distributed as dist          # noqa: E402
import torch.multiprocessing as mp       # noqa: E402


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        import mpyc_b200
        from mpyc_b200 import device as dev, sharding
        from mpyc_b200.device import DeviceArray, DeviceMatrix
        p, m, t, n = 2**128 - 173, 5, 2, 1_000_003
        ctx = mpyc_b200.context_for(p)
        # identical full inputs on every rank (seeded), each rank processes its slice
        S = DeviceArray.random(ctx, n, seed=11, stream_id=1)
        C = DeviceMatrix.empty(ctx, t, n)
        for j in range(t):
            C.t[j].copy_(DeviceArray.random(ctx, n, seed=20 + j, stream_id=2).t)
        a, b = sharding.shard_bounds(n, world, rank)
        S_loc = DeviceArray(ctx, S.t[a:b].contiguous())
        C_loc = DeviceMatrix.empty(ctx, t, b - a)
        for j in range(t):
            C_loc.t[j].copy_(C.t[j, a:b])
        sh_loc = dev.shamir_split(ctx, S_loc, C_loc, t, m)
        rec_loc = dev.shamir_recombine(ctx, [1, 2, 3], [sh_loc.row(i) for i in range(3)])
        assert rec_loc.count_mismatch(S_loc) == 0
        full_rec = sharding.gather(rec_loc.t.contiguous(), n)           # all-gather over NCCL
        assert torch.equal(full_rec, S.t)
        row3 = sharding.gather(sh_loc.t[3].contiguous(), n, dst=0)      # gather to one rank
        if rank == 0:
            sh_full = dev.shamir_split(ctx, S, C, t, m)
            assert torch.equal(row3, sh_full.t[3])
        q.put((rank, 'ok'))
    except Exception as exc:   # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_sharded_equals_single_gpu_and_gather():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=300) for _ in range(2)]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(results) == [(0, 'ok'), (1, 'ok')], results
