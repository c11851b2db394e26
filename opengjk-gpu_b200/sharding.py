"""Pair-array sharding across the GPUs of one box.

Pairs are independent, so multi-GPU is plain data parallelism: rank g owns the contiguous slice
[g*n/G, (g+1)*n/G) of both polytope arrays (or of the gkCollisionPair list, with the pool replicated), writes its
slice of distances / simplices / normals, and nothing is exchanged on the data path.  The only collectives are
plumbing: a barrier around the timed region and a MAX-reduce of the per-rank elapsed times.
"""
from __future__ import annotations


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced slice of range(n) owned by `rank` (sizes differ by at most one)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return lo, hi


def max_over_ranks(values, device=None):
    """Element-wise MAX of a list of floats over all ranks (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.cpu()]


def gather_slices(local, n_total: int, rank: int, world: int):
    """all_gather of per-rank result slices into the full array (device-resident consumers; the host-pointer API
    does not need it -- each rank copies straight into its slice of the caller's arrays)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or world == 1:
        return local
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    outs = [torch.empty((hi - lo,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device) for lo, hi in sizes]
    dist.all_gather(outs, local.contiguous()) if len({hi - lo for lo, hi in sizes}) == 1 else _uneven_all_gather(outs, local, dist)
    return torch.cat(outs, 0)


def _uneven_all_gather(outs, local, dist):
    for r, buf in enumerate(outs):
        if r == dist.get_rank():
            buf.copy_(local)
        dist.broadcast(buf, src=r)
