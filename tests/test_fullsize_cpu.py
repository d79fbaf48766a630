"""Pins the full-size checkers (tests/fullsize_common.py) on the CPU: the PyTorch restatement of the battle
observation and the state invariants must accept what the compiled reference itself produces, and must
reject corrupted observations; the batched driver is exercised on the test-only host emulation."""
import os

import numpy as np
import pytest
import torch

import fullsize_common as fs
import parity_common as pc


def checker_lib():
    for p in (pc.REF_LIB, pc.PORT_LIB):
        if os.path.exists(p):
            return p
    pytest.skip("no oracle library available")


@pytest.mark.parametrize("size,n,seed", [(40, 300, 0), (27, 120, 3)])
def test_restatement_accepts_the_reference(size, n, seed):
    env = pc.make_battle(checker_lib(), size, n, seed)
    steps = fs.play_battle_and_check(env, size, size, 25, seed)
    assert steps == 25


def test_restatement_rejects_corruption():
    env = pc.make_battle(checker_lib(), 30, 150, 1)
    hs = env.get_handles()
    pos = [env.get_pos(h).copy() for h in hs]
    ids = [env.get_agent_id(h).copy() for h in hs]
    nums = [np.array([env.get_num(h)]) for h in hs]
    la = [np.full(n[0], 21, dtype=np.int64) for n in nums]
    lr = [np.zeros(n[0], dtype=np.float32) for n in nums]
    obs = [env.get_observation(h) for h in hs]
    views = [torch.from_numpy(o[0].copy()) for o in obs]
    feats = [torch.from_numpy(o[1].copy()) for o in obs]
    fs.check_battle_observation(views, feats, pos, ids, nums, la, lr, 30, 30)
    for what in ("wall", "enemy", "hp", "minimap", "feature"):
        v2 = [v.clone() for v in views]
        f2 = [f.clone() for f in feats]
        if what == "wall":
            i = int(np.argmin(pos[0][:, 0]))          # the agent nearest to the west wall sees it
            idx = (v2[0][i, :, :, 0] == 1).nonzero()[0]
            v2[0][i, idx[0], idx[1], 0] = 0
        elif what == "enemy":
            seen = (v2[1][..., 4] == 1).nonzero()[0]
            v2[1][seen[0], seen[1], seen[2], 4] = 0
        elif what == "hp":
            seen = (v2[0][..., 4] == 1).nonzero()[0]
            v2[0][seen[0], seen[1], seen[2], 5] = 0.5
        elif what == "minimap":
            v2[1][7, 0, 0, 3] += 0.25
        else:
            f2[0][3, 0] = 1 - f2[0][3, 0]
        with pytest.raises(AssertionError):
            fs.check_battle_observation(v2, f2, pos, ids, nums, la, lr, 30, 30)
    # state invariants: a duplicated cell and an out-of-order id are caught
    p2 = [p.copy() for p in pos]
    p2[1][0] = p2[0][0]
    with pytest.raises(AssertionError):
        fs.check_state(p2, ids, nums, 30, 30)
    i2 = [i.copy() for i in ids]
    i2[0][[0, 1]] = i2[0][[1, 0]]
    with pytest.raises(AssertionError):
        fs.check_state(pos, i2, nums, 30, 30)


def test_batched_driver_on_the_host_emulation():
    """3 arenas behind one handle (test-only emulation of the engine), arenas 0 and 2 replayed by the checker"""
    if not os.path.exists(pc.EMU_LIB):
        pytest.skip("tests/_emu not built")
    import magent_b200 as magent
    A, size, n = 3, 32, 140
    env = magent.GridWorld("battle", map_size=size, _lib=pc.EMU_LIB, _num_arenas=A)
    env.set_seed(11)
    env.reset()
    for h in env.get_handles():
        env.add_agents(h, method="random", n=n)
    samples = {a: pc.make_battle(checker_lib(), size, n, 11 + a) for a in (0, 2)}
    fs.play_battle_and_check(env, size, size, 20, 5, samples=samples)


def test_soak_driver_on_the_host_emulation():
    """device-drawn actions, whole-batch checks only (the loop tests/test_zz_fullsize_gpu.py runs at full size)"""
    if not os.path.exists(pc.EMU_LIB):
        pytest.skip("tests/_emu not built")
    import magent_b200 as magent
    A, size, n = 3, 30, 200
    env = magent.GridWorld("battle", map_size=size, _lib=pc.EMU_LIB, _num_arenas=A)
    env.set_seed(4)
    env.reset()
    for h in env.get_handles():
        env.add_agents(h, method="random", n=n)
    deaths = fs.soak_battle_and_check(env, size, size, 40, 9, obs_every=5)
    assert deaths > 0
