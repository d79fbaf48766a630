"""worker for tests/test_sharding_cpu.py: one rank of a world_size-2 gloo job driving the test-only emulation
build of the engine on its shard of arenas; prints the reduced window on rank 0."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch.distributed as dist  # noqa: E402

import magent_b200 as magent  # noqa: E402
import parity_common as pc  # noqa: E402
from magent_b200.sharding import shard_arenas, reduce_window  # noqa: E402


def simulate(first, count, steps, seed):
    env = magent.GridWorld("battle", map_size=30, _lib=pc.EMU_LIB, _num_arenas=count)
    env.set_seed(seed + first)              # arena k is seeded seed + k whichever rank owns it
    env.reset()
    hs = env.get_handles()
    for h in hs:
        env.add_agents(h, method="random", n=80)
    digest = []
    for t in range(steps):
        for h in hs:
            env.set_random_actions(h, 1000 * t + first)      # emulation RNG: per-shard stream
        env.step()
        env.clear_dead()
        digest.append([int(x) for h in hs for x in env.get_arena_nums(h)])
    return env.get_counters(), digest


if __name__ == "__main__":
    total, steps = int(sys.argv[1]), int(sys.argv[2])
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    first, count = shard_arenas(total, rank, world)
    counters, digest = simulate(first, count, steps, seed=7)
    summed, tmax = reduce_window(counters, elapsed_ms=10.0 * (rank + 1))
    gathered = [None] * world
    dist.all_gather_object(gathered, (first, count, counters))
    if rank == 0:
        print(json.dumps({"summed": summed, "tmax": tmax, "shards": gathered}))
    dist.destroy_process_group()
