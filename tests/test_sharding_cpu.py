"""N>1 host logic on CPU (gloo, world_size 2): arena sharding and the counter/elapsed-time reduction."""
import json
import os
import subprocess
import sys

import pytest

import parity_common as pc
from magent_b200.sharding import shard_arenas


def test_shard_arenas_partitions_exactly():
    for total in (1, 7, 8, 512, 4096, 4099):
        for world in (1, 2, 3, 8):
            blocks = [shard_arenas(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _f, c in blocks) == total
            for (f0, c0), (f1, _c1) in zip(blocks, blocks[1:]):
                assert f0 + c0 == f1
            assert max(c for _f, c in blocks) - min(c for _f, c in blocks) <= 1


def test_world2_gloo_reduction():
    if not os.path.exists(pc.EMU_LIB):
        subprocess.run([os.path.join(pc.REPO, "tests", "emu", "build.sh")], check=True, capture_output=True)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29617",
           os.path.join(pc.REPO, "tests", "dist_worker.py"), "5", "6"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=pc.REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    res = json.loads(line)
    shards = res["shards"]
    assert [s[:2] for s in shards] == [[0, 3], [3, 2]]
    # SUM over ranks of every counter, MAX of the elapsed time
    for k in range(len(res["summed"])):
        assert res["summed"][k] == shards[0][2][k] + shards[1][2][k]
    assert res["tmax"] == 20.0
    # agent-steps counter == arenas x agents x steps at the start of a sparse battle (few deaths): sanity bound
    assert 0 < res["summed"][0] <= 5 * 160 * 6
