"""Co-located resharing: the `_reshare` exchange of mpyc/runtime.py:603-689 as one GPU collective step.

In MPyC every one of the 2t+1 dealing parties splits its (degree-2t) share vector into m fresh degree-t
shares, pickles row i for party i and sends it over TCP; every party unpickles the 2t+1 rows it received and
recombines them (runtime.py:660-680; framing asyncoro.py:54-64).  When the m parties of a computation live on
the GPUs of one box (benchmark / simulation mode; party j on rank j % world) the same step is

    K2 on every dealer  ->  exchange of limb rows between GPUs  ->  K3s on every party

with the rows going GPU to GPU as limb buffers: no Python ints, no pickle, no host copies.  The exchange is a
grouped point-to-point batch (`torch.distributed.batch_isend_irecv`: one ncclGroupStart/End of ncclSend/ncclRecv
over NVLink on the nccl backend), posted straight on the rows of the share matrices K2 wrote -- rows that stay on
the same rank are not copied at all.  Both ends enumerate the (dealer, recipient) pairs in the same sorted
order, which is how NCCL matches sends with receives between a pair of ranks.

The arithmetic is injected as an `engine` (split / recombine on tensors) so that the routing can be exercised on
CPU tensors over gloo (tests/test_exchange_gloo.py) with the oracle standing in for the kernels; on the GPU the
engine is `DeviceEngine` below (mpyc_b200.device: K2 in generate mode, K3).
"""
import torch
import torch.distributed as dist


def owner(party, world):
    """Rank that hosts a party."""
    return party % world


def local_parties(m, world, rank):
    return [j for j in range(m) if owner(j, world) == rank]


class DeviceEngine:
    """K2 (coefficients from the in-kernel ChaCha20 stream) and K3 on device-resident limb tensors."""

    def __init__(self, ctx):
        self.ctx = ctx
        self._nonce = 0

    def split(self, x, t, m):
        """x: int64 (n, L) limb tensor -> (m, n, L) tensor of shares (row i for party i)."""
        from mpyc_b200 import device as dev
        self._nonce += 1
        sh = dev.shamir_split_generate(self.ctx, dev.DeviceArray(self.ctx, x), t, m, nonce=self._nonce)
        return sh.t

    def recombine(self, xs, rows):
        from mpyc_b200 import device as dev
        return dev.shamir_recombine(self.ctx, xs, [dev.DeviceArray(self.ctx, r) for r in rows]).t

    def empty_like_row(self, x):
        return torch.empty_like(x)


def reshare(engine, shares, t, m, group=None, first_dealer=0):
    """One resharing round.

    shares: {party j: limb tensor (n, L)} for the parties hosted on this rank (their current shares, e.g. the
    degree-2t local products of a secure multiplication).  Dealers are the 2t+1 parties first_dealer,
    first_dealer+1, ... (mod m) -- runtime.py's `uci` load balancing.  Returns {party i: new degree-t share}
    for the same local parties.  Collective: every rank of the group must call it."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = local_parties(m, world, rank)
    if sorted(shares) != mine:
        raise ValueError(f'rank {rank} hosts parties {mine}, got shares for {sorted(shares)}')
    dealers = [(first_dealer + a) % m for a in range(2 * t + 1)]
    if len(set(dealers)) != len(dealers):
        raise ValueError('resharing needs m >= 2t+1 parties')
    # 1. every local dealer splits its share vector: dealt[j] is (m, n, L), row i goes to party i
    dealt = {j: engine.split(shares[j], t, m) for j in dealers if j in shares}
    some = next(iter(shares.values()))
    # 2. exchange: recv[i][a] = row i of dealer dealers[a]'s matrix
    recv = {i: [None] * len(dealers) for i in mine}
    ops = []
    for a, j in enumerate(dealers):                      # the same (dealer, recipient) order on every rank
        src = owner(j, world)
        for i in range(m):
            dst = owner(i, world)
            if src == rank and dst == rank:
                recv[i][a] = dealt[j][i]                 # stays on this GPU: use the row K2 wrote, no copy
            elif src == rank:
                ops.append(dist.P2POp(dist.isend, dealt[j][i].contiguous(), _global_rank(dst, group), group))
            elif dst == rank:
                buf = engine.empty_like_row(some)
                recv[i][a] = buf
                ops.append(dist.P2POp(dist.irecv, buf, _global_rank(src, group), group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    # 3. every local party recombines the 2t+1 rows it now holds
    xs = [j + 1 for j in dealers]
    return {i: engine.recombine(xs, recv[i]) for i in mine}


def _global_rank(group_rank, group):
    return group_rank if group is None else dist.get_global_rank(group, group_rank)
