"""N > 1 host logic on CPU: world_size-2 (and 3) gloo process groups exercise the element-axis partition
and the gather collective used when a caller needs the whole vector on one rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mpyc_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, L, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        full = torch.from_numpy(np.arange(n * L, dtype=np.int64).reshape(n, L) * 7 + 3)
        mine = sharding.local_slice(full).clone()
        a, b = sharding.shard_bounds(n, world, rank)
        assert mine.shape[0] == b - a and torch.equal(mine, full[a:b])
        everywhere = sharding.gather(mine, n)
        assert torch.equal(everywhere, full)
        at_root = sharding.gather(mine, n, dst=0)
        assert (at_root is None) == (rank != 0)
        if rank == 0:
            assert torch.equal(at_root, full)
        # "weak-scaling" bookkeeping of bench.py: per-rank work is independent, totals add up
        t = torch.tensor([float(mine.shape[0])])
        dist.all_reduce(t)
        assert int(t.item()) == n
        q.put((rank, 'ok'))
    except Exception as exc:   # noqa: BLE001
        q.put((rank, repr(exc)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,n,L', [(2, 1000, 2), (2, 1001, 1), (3, 10, 4), (2, 1, 2)])
def test_partition_and_gather(world, n, L):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, L, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(r, 'ok') for r in range(world)], results


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 100, 10**8 + 3):
        for world in (1, 2, 4, 8):
            bounds = [sharding.shard_bounds(n, world, r) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == n
            assert all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in bounds]
            assert max(sizes) - min(sizes) <= 1 and sizes == sharding.shard_sizes(n, world)
