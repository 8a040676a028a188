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


def _reshare_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        import mpyc_b200
        from mpyc_b200 import device as dev, exchange
        from mpyc_b200.device import DeviceArray, DeviceMatrix
        for p, m, t, n in ((2**128 - 173, 5, 2, 200_003), (2**64 - 189, 3, 1, 1_000_000), (2**256 - 189, 7, 3, 50_001)):
            ctx = mpyc_b200.context_for(p)
            A = DeviceArray.random(ctx, n, seed=3, stream_id=1)          # same inputs and sharings on every rank
            B = DeviceArray.random(ctx, n, seed=4, stream_id=1)
            CA, CB = DeviceMatrix.empty(ctx, t, n), DeviceMatrix.empty(ctx, t, n)
            for j in range(t):
                CA.t[j].copy_(DeviceArray.random(ctx, n, seed=30 + j, stream_id=2).t)
                CB.t[j].copy_(DeviceArray.random(ctx, n, seed=40 + j, stream_id=2).t)
            sa, sb = dev.shamir_split(ctx, A, CA, t, m), dev.shamir_split(ctx, B, CB, t, m)
            mine = exchange.local_parties(m, world, rank)
            prod = {j: (sa.row(j) * sb.row(j)).t for j in mine}          # degree-2t local products
            if os.environ.get('MPYC_TEST_PEER') == '1':
                # exchange fused into K2: rows stored straight into the recipient GPU (three rounds: both buffer parities)
                peer = exchange.PeerReshare(ctx, m, t, n, first_dealer=1)
                for _ in range(3):
                    new = peer.reshare(prod)
                peer.close()
            else:
                new = exchange.reshare(exchange.DeviceEngine(ctx), prod, t, m, first_dealer=1)
            assert sorted(new) == mine
            # collect every party's new share on every rank and open the product with two different party sets
            full = [None] * m
            for i in range(m):
                buf = new[i].contiguous() if i in new else torch.empty((n, ctx.nlimbs), dtype=torch.int64, device='cuda')
                dist.broadcast(buf, src=exchange.owner(i, world))
                full[i] = DeviceArray(ctx, buf)
            want = A * B
            for xs in (list(range(1, t + 2)), list(range(m - t, m + 1))):
                got = dev.shamir_recombine(ctx, xs, [full[x - 1] for x in xs])
                assert got.count_mismatch(want) == 0, (p.bit_length(), xs)
        q.put((rank, 'ok'))
    except Exception:   # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('mode', ['nccl', 'peer'])
def test_colocated_reshare_over_nccl(mode, monkeypatch):
    """secure multiplication's resharing step with the parties spread over 2 GPUs -- K2 on every
    dealer, limb rows exchanged GPU to GPU (mode nccl: grouped ncclSend/ncclRecv; mode peer: K2 stores each row
    straight into the recipient GPU's memory over NVLink), K3s on every party; any t+1 of the new shares open the
    product."""
    monkeypatch.setenv('MPYC_TEST_PEER', '1' if mode == 'peer' else '0')
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_reshare_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=300) for _ in range(2)]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(results) == [(0, 'ok'), (1, 'ok')], results
