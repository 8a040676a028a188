"""Times the co-located resharing round (mpyc_b200.exchange.reshare) on N GPUs: m parties dealt round-robin to
the ranks, n elements per party.  Launch: python -m torch.distributed.run --nproc-per-node N tools/bench_reshare.py
Reports per-round time (CUDA events, max over ranks) and the bytes that crossed GPUs."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpyc_b200                                   # noqa: E402
from mpyc_b200 import exchange                     # noqa: E402
from mpyc_b200.device import DeviceArray          # noqa: E402

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
for p, m, t, n in ((2**128 - 173, 5, 2, 20_000_000), (2**256 - 189, 7, 3, 213_248), (2**64 - 189, 3, 1, 50_000_000)):
    ctx = mpyc_b200.context_for(p)
    eng = exchange.DeviceEngine(ctx)
    mine = exchange.local_parties(m, world, rank)
    shares = {j: DeviceArray.random(ctx, n, seed=j, stream_id=5).t for j in mine}
    peer = exchange.PeerReshare(ctx, m, t, n)
    results = {}
    for name, fn in (('nccl_send_recv', lambda: exchange.reshare(eng, shares, t, m)), ('peer_stores_in_K2', lambda: peer.reshare(shares))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        steps = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device='cuda')
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        results[name] = float(ms.item())
    ms = torch.tensor([results['nccl_send_recv']], device='cuda')
    dealers = list(range(2 * t + 1))
    crossing = sum(1 for j in dealers for i in range(m) if exchange.owner(j, world) != exchange.owner(i, world))
    if rank == 0:
        eb = ctx.elem_bytes
        print(json.dumps({'workload': f'reshare p={p.bit_length()}b m={m} t={t} n={n}', 'n_gpus': world, 'ms_per_round': float(ms.item()),
                          'elements_reshared_per_s': m * n / (float(ms.item()) * 1e-3),
                          'rows_crossing_gpus': crossing, 'bytes_crossing_gpus': crossing * n * eb,
                          'exchange_GBps_aggregate': crossing * n * eb / (float(ms.item()) * 1e-3) / 1e9,
                          'ms_per_round_peer_stores': results['peer_stores_in_K2'],
                          'elements_reshared_per_s_peer_stores': m * n / (results['peer_stores_in_K2'] * 1e-3)}), flush=True)
    peer.close()
dist.destroy_process_group()
