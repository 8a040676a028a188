"""Host logic of the co-located resharing step (mpyc_b200/exchange.py) on CPU: world_size 2 and 3
gloo groups, the oracle standing in for the kernels.  m parties are dealt round-robin to the ranks; every rank
reshapes the degree-2t products of its parties, the rows cross ranks through batch_isend_irecv, and what the
parties hold afterwards must be a degree-t sharing of the products (runtime.py:603-689 semantics)."""
import os
import random
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import shamir_oracle as orc
from mpyc_b200 import exchange

P = 2**61 - 1


class OracleEngine:
    """split / recombine of mpyc_b200.exchange on int64 (n, 1) CPU tensors via the CPU oracle."""

    def __init__(self, p, seed):
        self.F, self.p, self.rnd = orc.field_of(p), p, random.Random(seed)

    def split(self, x, t, m):
        s = [int(v) for v in x[:, 0].tolist()]
        C = [[self.rnd.randrange(self.p) for _ in s] for _ in range(t)]
        rows = orc.split_np_order(self.F, s, C, m)
        return torch.tensor(rows, dtype=torch.int64).unsqueeze(-1)

    def recombine(self, xs, rows):
        vals = orc.recombine(self.F, xs, [[int(v) for v in r[:, 0].tolist()] for r in rows])
        return torch.tensor([v % self.p for v in vals], dtype=torch.int64).unsqueeze(-1)

    def empty_like_row(self, x):
        return torch.empty_like(x)


def _worker(rank, world, port, m, t, n, first_dealer, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        F = orc.field_of(P)
        rnd = random.Random(7)                     # identical on every rank: the "inputs" a, b and their sharings
        a = [rnd.randrange(P) for _ in range(n)]
        b = [rnd.randrange(P) for _ in range(n)]
        Ca = [[rnd.randrange(P) for _ in range(n)] for _ in range(t)]
        Cb = [[rnd.randrange(P) for _ in range(n)] for _ in range(t)]
        sa, sb = orc.split_np_order(F, a, Ca, m), orc.split_np_order(F, b, Cb, m)
        mine = exchange.local_parties(m, world, rank)
        prod = {j: torch.tensor([x * y % P for x, y in zip(sa[j], sb[j])], dtype=torch.int64).unsqueeze(-1) for j in mine}
        new = exchange.reshare(OracleEngine(P, 100 + rank), prod, t, m, first_dealer=first_dealer)
        assert sorted(new) == mine
        gathered = [None] * world
        dist.all_gather_object(gathered, {i: v[:, 0].tolist() for i, v in new.items()})
        allsh = {i: v for part in gathered for i, v in part.items()}
        assert sorted(allsh) == list(range(m))
        want = [x * y % P for x, y in zip(a, b)]
        for xs in ([1 + i for i in range(t + 1)], [m - i for i in range(t + 1)]):     # any t+1 parties open the product
            got = orc.recombine(F, xs, [allsh[x - 1] for x in xs])
            assert [v % P for v in got] == want
        q.put((rank, 'ok'))
    except Exception:   # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,m,t,n,first_dealer', [(2, 3, 1, 37, 0), (3, 5, 2, 16, 3), (2, 7, 3, 9, 5), (2, 2, 0, 5, 1),
                                                      (3, 2, 0, 4, 0)])   # last: more ranks than parties
def test_reshare_routes_rows_between_ranks(world, m, t, n, first_dealer):
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, m, t, n, first_dealer, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(r, 'ok') for r in range(world)], results
