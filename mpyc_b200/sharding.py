"""Element-axis sharding over the GPUs of one box.

Every op of the hot path is independent per element, so a vector of n field elements is
split into `world` contiguous slices, one process per GPU; field contexts and the Vandermonde / Lagrange
tables are replicated (bytes).  Share generation, recombination, elementwise ops and the PRSS linear step
need NO communication.  The single collective is `gather`: an all-gather (or gather to one rank) of the
limb rows over NCCL/NVLink, used only when the caller needs the whole vector on one rank -- e.g. to hand
it to MPyC's transport (runtime.py:665) or to return an `output` (runtime.py:586).

Works on CUDA tensors (backend nccl) and CPU tensors (backend gloo: used by the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_bounds(n, world, rank):
    """Contiguous balanced partition: the first n % world ranks get one extra element."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(n, world):
    return [shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0] for r in range(world)]


def local_slice(x, world=None, rank=None, dim=0):
    """This rank's slice of a full tensor / array along the element axis."""
    world = dist.get_world_size() if world is None else world
    rank = dist.get_rank() if rank is None else rank
    a, b = shard_bounds(x.shape[dim], world, rank)
    index = [slice(None)] * x.ndim
    index[dim] = slice(a, b)
    return x[tuple(index)]


def gather(local, n, dst=None, group=None):
    """Reassemble the full (n, ...) limb tensor from per-rank contiguous shards.

    dst=None: all-gather (every rank gets the result); dst=r: only rank r gets it (others return None).
    Shards may differ by one element; they are padded to the largest shard for the collective."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(n, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f'rank {rank}: shard has {local.shape[0]} elements, expected {sizes[rank]}')
    width = max(sizes)
    padded = local
    if local.shape[0] != width:
        padded = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded[:local.shape[0]] = local
    padded = padded.contiguous()
    if dst is None:
        buf = torch.empty((world, width) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        if local.is_cuda:
            dist.all_gather_into_tensor(buf, padded, group=group)
        else:
            parts = [torch.empty_like(padded) for _ in range(world)]
            dist.all_gather(parts, padded, group=group)
            buf = torch.stack(parts)
    else:
        if rank == dst:
            parts = [torch.empty_like(padded) for _ in range(world)]
            dist.gather(padded, parts, dst=dst, group=group)
            buf = torch.stack(parts)
        else:
            dist.gather(padded, None, dst=dst, group=group)
            return None
    if all(s == width for s in sizes):
        return buf.reshape((world * width,) + tuple(local.shape[1:]))[:n]
    return torch.cat([buf[r, :sizes[r]] for r in range(world)], dim=0)
