"""The two places where the engine deliberately does NOT do what the reference does, pinned:
  * set_action called twice for one group in one step: the second call's actions are the ones executed (the reference
    queues both calls' moves / attacks: GridWorld.cc:403-454 pushes into the buffers on every call);
  * action ids outside [0, n_action): ignored -- the agent neither moves nor attacks (the reference indexes its
    action tables out of bounds: GridWorld.cc:412-453).
Both are checked against the reference driven with the equivalent well-formed input."""
import numpy as np

import parity_common as pc


def second_set_action_wins(engine_lib, checker_lib):
    env, ref = pc.make_battle(engine_lib, 30, 200, 4), pc.make_battle(checker_lib, 30, 200, 4)
    rs = np.random.RandomState(4)
    for t in range(12):
        hs, hr = env.get_handles(), ref.get_handles()
        first = [rs.randint(0, 21, size=env.get_num(h)).astype(np.int32) for h in hs]
        second = [rs.randint(0, 21, size=env.get_num(h)).astype(np.int32) for h in hs]
        for h, a in zip(hs, first):
            env.set_action(h, a)
        for h, a in zip(hs, second):                       # overwrites; the call order of the FIRST calls stands
            env.set_action(h, a)
        for h, a in zip(hr, second):
            ref.set_action(h, a)
        assert env.step() == ref.step()
        for h, k in zip(hs, hr):
            np.testing.assert_array_equal(env.get_pos(h), ref.get_pos(k))
            np.testing.assert_array_equal(env.get_alive(h), ref.get_alive(k))
            np.testing.assert_allclose(env.get_reward(h), ref.get_reward(k), rtol=0, atol=pc.REWARD_TOL)
        env.clear_dead()
        ref.clear_dead()


def invalid_actions_are_ignored(engine_lib, checker_lib):
    """an out-of-range id behaves like the reference's 'stay' move (action 6 of battle's 13 moves = offset (0, 0))"""
    env, ref = pc.make_battle(engine_lib, 30, 200, 9), pc.make_battle(checker_lib, 30, 200, 9)
    stay = None
    rs = np.random.RandomState(9)
    for t in range(12):
        hs, hr = env.get_handles(), ref.get_handles()
        for h, k in zip(hs, hr):
            n = env.get_num(h)
            a = rs.randint(0, 21, size=n).astype(np.int32)
            bad = rs.rand(n) < 0.3
            junk = np.where(rs.rand(n) < 0.5, 21 + rs.randint(0, 1000, size=n), -1 - rs.randint(0, 1000, size=n)).astype(np.int32)
            if stay is None:                                # the move action whose offset is (0, 0): find it once by trying
                stay = 6
            env.set_action(h, np.where(bad, junk, a).astype(np.int32))
            ref.set_action(k, np.where(bad, stay, a).astype(np.int32))
        assert env.step() == ref.step()
        for h, k in zip(hs, hr):
            np.testing.assert_array_equal(env.get_pos(h), ref.get_pos(k))
            np.testing.assert_array_equal(env.get_alive(h), ref.get_alive(k))
            np.testing.assert_allclose(env.get_reward(h), ref.get_reward(k), rtol=0, atol=pc.REWARD_TOL)
        env.clear_dead()
        ref.clear_dead()
