"""Multi-GPU plumbing for the arena-sharded engine.

Arenas never interact, so the path shards with NO data-path collective: rank r simulates the contiguous
block of arena ids ``shard_arenas(total, r, world)`` (arena k is seeded ``seed + k`` no matter which rank
owns it, so a sharded job reproduces the single-GPU job arena for arena).  The only exchange is one
all-reduce per measurement window: SUM of the int64 event counters, MAX of the elapsed device time
(NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations


def shard_arenas(total, rank, world):
    """(first_arena, count) of rank's contiguous block; the first ``total % world`` ranks get one extra."""
    base, rem = divmod(int(total), int(world))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def reduce_window(counters, elapsed_ms, device=None):
    """all-reduce a measurement window: returns (summed counters list, max elapsed ms).

    ``counters``: list of python ints; ``elapsed_ms``: float.  Works un-initialised (single process)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [int(c) for c in counters], float(elapsed_ms)
    c = torch.tensor([int(x) for x in counters], dtype=torch.int64, device=device)
    t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [int(x) for x in c.tolist()], float(t.item())
