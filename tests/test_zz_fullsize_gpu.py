"""Parity at BASELINE.json's full sizes (configs[1..4] as bench.py runs them), on a real GPU, through the C ABI.

The sequential reference cannot replay 512 arenas (or 64) in test time, so these tests check
  * SAMPLED arenas exactly against independent checker environments seeded seed+arena, and
  * the WHOLE batch through size-independent properties (tests/fullsize_common.py): unique in-board cells,
    ordered ids, bounded moves, and every observation record against a plain PyTorch restatement of
    get_observation that is itself pinned to the compiled reference on the CPU (tests/test_fullsize_cpu.py).
The single 1000x1000 arena (2x400k agents) is small enough for the reference: it is compared record by record.
Observations are taken through device pointers (`get_observation_torch`), the path bench.py's `value` times.
"""
import os

import numpy as np
import pytest

import fullsize_common as fs
import parity_common as pc

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]      # (pytest-timeout: a stuck test fails instead of hanging the suite)

ENGINE = os.environ.get("MAGENT_FULLSIZE_ENGINE", pc.CUDA_LIB)      # (tests/_emu library for a dry run on the CPU)
ON_GPU = ENGINE == pc.CUDA_LIB


def checker_lib():
    for p in (pc.REF_LIB, pc.PORT_LIB):
        if os.path.exists(p):
            return p
    pytest.skip("no oracle library available (oracle/_ref or oracle/_build)")


def batched_battle(arenas, size, n, seed):
    import magent_b200 as magent
    kw = {"_num_arenas": arenas} if arenas != 1 else {}
    env = magent.GridWorld("battle", map_size=size, _lib=ENGINE, **kw)
    env.set_seed(seed)
    env.reset()
    for h in env.get_handles():
        env.add_agents(h, method="random", n=n)
    return env


# ---- randomised differential games on maps in the reference's large_map_mode (8 / 16 move bands), CUDA engine vs
# the checker (tests/fuzz_common.py; the same seeds run against the host emulation in tests/test_fuzz_cpu.py)
@pytest.mark.parametrize("seed", list(range(100000, 100012)) + [200000, 200001, 200002])
def test_random_games_on_banded_maps(seed):
    import fuzz_common as fz
    fz.play(seed, checker_lib(), ENGINE, steps=20)


@pytest.mark.parametrize("seed", [101000, 101001, 101002, 101003, 201000])
def test_random_games_on_banded_maps_with_an_irregular_caller(seed):
    import fuzz_common as fz
    fz.play_irregular(seed, checker_lib(), ENGINE, steps=20)


@pytest.mark.parametrize("seed", [102000, 102001, 102002])
def test_random_arena_batches_on_banded_maps(seed):
    import fuzz_common as fz
    fz.play_batch(seed, checker_lib(), ENGINE, n_arenas=2 + seed % 3, steps=10)


def test_plain_c_caller_prints_the_same_trace_as_on_the_checker():
    """tests/c/battle_caller.c: the same binary, host buffers, reference ABI only --
    once on the checker library and once on the CUDA library"""
    import c_caller_common as cc
    want = cc.run(checker_lib(), steps=60, size=40, n=250)
    got = cc.run(ENGINE, steps=60, size=40, n=250)
    assert want.returncode == 0, want.stderr
    assert got.returncode == 0, got.stderr
    assert got.stdout == want.stdout


def test_arena_whose_step_scratch_does_not_fit_shared_memory():
    """2x4000 agents on 120x120: one CTA per arena, but 8000 agents x 41 B of step scratch exceed the 200 KB the
    launch keeps in shared memory, so the scratch arrays stay in HBM (backend_cuda.cu launch_step)"""
    want = pc.run_trace(pc.make_battle(checker_lib(), 120, 4000, 2), 8, 3, keep_obs=True)
    got = pc.run_trace(pc.make_battle(ENGINE, 120, 4000, 2), 8, 3, keep_obs=True)
    pc.compare_traces(want, got)


@pytest.mark.parametrize("which", ["battle", "pursuit", "mixed", "arrange"])
def test_cold_info_getters_match_the_reference(which):
    """view2attack / attack_base / groups_info / walls_info / global_minimap / mean_info (GridWorld.cc:717-894),
    served from a host snapshot of the device state"""
    make = {"battle": lambda lib: pc.make_battle(lib, 30, 120, 1), "pursuit": lambda lib: pc.make_pursuit(lib, 40, 2),
            "mixed": lambda lib: pc.make_mixed(lib), "arrange": lambda lib: pc.make_arrange(lib)}[which]
    pc.play_and_compare_info(make, checker_lib(), ENGINE)


def test_select_arena_and_event_counters():
    """per-arena setup through magent_b200_select_arena (own seed, walls, extra agents per arena) against independent
    checkers; the device event counters against host-side counts"""
    pc.play_selected_arenas(ENGINE, checker_lib())


@pytest.mark.parametrize("seed", list(range(40000, 40024)) + [40029, 47002, 47012, 47086, 47170, 110000, 110001, 110002, 113018])
def test_random_games_with_a_chaotic_caller(seed):
    """reads at every point of the loop, acting subset changes every step, reset in mid-run (fuzz_common.trace_chaotic)"""
    import fuzz_common as fz
    fz.play_chaotic(seed, checker_lib(), ENGINE)


def test_uncollected_group_reward_survives_reset():
    want = pc.group_reward_across_reset(checker_lib())
    got = pc.group_reward_across_reset(ENGINE)
    for a, b in zip(want, got):
        np.testing.assert_allclose(a, b, rtol=0, atol=pc.REWARD_TOL)


def test_self_kill_feeds_the_corpse(tmp_path):
    """hp of an un-culled corpse in the replay dump after a self-aimed in-group attack (Map.cc:265-273)"""
    act = pc.self_kill_frames(checker_lib(), None)
    want = pc.self_kill_frames(checker_lib(), str(tmp_path / "ref"), act)
    got = pc.self_kill_frames(ENGINE, str(tmp_path / "b200"), act)
    np.testing.assert_allclose(want[0], got[0], rtol=0, atol=pc.REWARD_TOL)
    assert want[1] == got[1]


@pytest.mark.parametrize("seed", list(range(60000, 60014)) + [115000, 115001])
def test_random_arena_batches_with_a_chaotic_caller(seed):
    import fuzz_common as fz
    fz.play_batch_chaotic(seed, checker_lib(), ENGINE, n_arenas=1 + seed % 4)


# 5-8 groups: up to 25 observation channels per view cell (fuzz_common.MANY_GROUPS_SEED)
@pytest.mark.parametrize("seed", list(range(70000, 70016)))
def test_random_games_with_many_groups(seed):
    import fuzz_common as fz
    fz.play(seed, checker_lib(), ENGINE, steps=20)


@pytest.mark.parametrize("seed", list(range(72000, 72008)) + [73000, 73001])
def test_random_games_with_many_groups_and_a_chaotic_caller(seed):
    import fuzz_common as fz
    if seed >= 73000:
        fz.play_batch_chaotic(seed, checker_lib(), ENGINE, n_arenas=1 + seed % 4)
    else:
        fz.play_chaotic(seed, checker_lib(), ENGINE)


def test_every_step_loop_buffer_as_a_device_pointer():
    """include/magent_runtime_api.h: observation, action, reward and id / pos / alive buffers may be CUDA device pointers
    (read / written in place).  One engine is driven through device tensors only, the checker through host arrays."""
    import ctypes
    import torch
    dev = torch.device("cuda") if ON_GPU else torch.device("cpu")           # (CPU tensors in the dry run: host pointers)
    env = pc.make_battle(ENGINE, 40, 180, 4)
    ref = pc.make_battle(checker_lib(), 40, 180, 4)
    L = env._lib
    rs = np.random.RandomState(6)
    for t in range(25):
        for g, h in enumerate(env.get_handles()):
            rh = ref.get_handles()[g]
            n = env.get_num(h)
            assert n == ref.get_num(rh)
            if n == 0:
                continue
            v, f = env.get_observation_torch(h) if ON_GPU else [torch.from_numpy(x.copy()) for x in env.get_observation(h)]
            rv, rf = ref.get_observation(rh)
            np.testing.assert_array_equal(v.cpu().numpy().view(np.uint32), rv.view(np.uint32))
            np.testing.assert_array_equal(f.cpu().numpy().view(np.uint32), rf.view(np.uint32))
            act = rs.randint(0, 21, size=n).astype(np.int32)
            d_act = torch.from_numpy(act).to(dev)
            L.env_set_action(env.game, env._hv(h), ctypes.c_void_p(d_act.data_ptr()))
            ref.set_action(rh, act)
        if ON_GPU:
            torch.cuda.synchronize()
        env.step()
        ref.step()
        for g, h in enumerate(env.get_handles()):
            rh = ref.get_handles()[g]
            n = env.get_num(h)
            d_rew = torch.full((n,), -7.0, dtype=torch.float32, device=dev)
            d_pos = torch.full((n, 2), -7, dtype=torch.int32, device=dev)
            d_id = torch.full((n,), -7, dtype=torch.int32, device=dev)
            d_alive = torch.full((n,), 7, dtype=torch.uint8, device=dev)
            L.env_get_reward(env.game, env._hv(h), ctypes.c_void_p(d_rew.data_ptr()))
            L.env_get_info(env.game, env._hv(h), b"pos", ctypes.c_void_p(d_pos.data_ptr()))
            L.env_get_info(env.game, env._hv(h), b"id", ctypes.c_void_p(d_id.data_ptr()))
            L.env_get_info(env.game, env._hv(h), b"alive", ctypes.c_void_p(d_alive.data_ptr()))
            env.sync()
            np.testing.assert_allclose(d_rew.cpu().numpy(), ref.get_reward(rh), rtol=0, atol=pc.REWARD_TOL)
            np.testing.assert_array_equal(d_pos.cpu().numpy(), ref.get_pos(rh))
            np.testing.assert_array_equal(d_id.cpu().numpy(), ref.get_agent_id(rh))
            np.testing.assert_array_equal(d_alive.cpu().numpy().astype(bool), ref.get_alive(rh).astype(bool))
        env.clear_dead()
        ref.clear_dead()


def test_observation_into_one_host_and_one_device_buffer():
    """env_get_observation with bufs[0] on the device and bufs[1] on the host, and the other way round"""
    import ctypes
    import torch
    dev = torch.device("cuda") if ON_GPU else torch.device("cpu")
    env = pc.make_battle(ENGINE, 36, 140, 9)
    h = env.get_handles()[1]
    want_v, want_f = [x.copy() for x in env.get_observation(h)]
    n = env.get_num(h)
    for view_on_device in (True, False):
        v = torch.full((n,) + env.get_view_space(h), -3.0, dtype=torch.float32, device=dev if view_on_device else "cpu")
        f = torch.full((n,) + env.get_feature_space(h), -3.0, dtype=torch.float32, device="cpu" if view_on_device else dev)
        bufs = (ctypes.c_void_p * 2)(v.data_ptr(), f.data_ptr())
        env._lib.env_get_observation(env.game, env._hv(h), bufs)
        env.sync()
        np.testing.assert_array_equal(v.cpu().numpy().view(np.uint32), want_v.view(np.uint32))
        np.testing.assert_array_equal(f.cpu().numpy().view(np.uint32), want_f.view(np.uint32))


@pytest.mark.parametrize("seed", list(range(9000, 9012)))
def test_three_engines_interleaved_with_chaotic_callers(seed):
    """three different random games alive in one process, their callers switched between any two API calls: the
    backend's process-global scratch (per-agent headers, padded minimap rows, cached launch configurations, staging
    buffers) must not leak from one engine into another"""
    import fuzz_common as fz
    fz.play_interleaved_engines(seed, checker_lib(), ENGINE)


def test_golden_edge_cases(tmp_path):
    """tests/golden/edge_cases.npz (recorded from the compiled reference): group reward across reset, replay frames
    after a self-kill -- needs no checker library"""
    import golden_common as gc
    gc.check_edge_cases(ENGINE, str(tmp_path / "frames"))


# ---- bench.py's workloads at their full sizes
def test_battle_512_arenas_of_2x1000():
    """BASELINE configs[4] per-GPU share (= bench.py's default workload): 1.024 M agents per step"""
    A, size, n, seed = 512, 200, 1000, 100
    env = batched_battle(A, size, n, seed)
    samples = {a: pc.make_battle(checker_lib(), size, n, seed + a) for a in (0, 255, 511)}
    fs.play_battle_and_check(env, size, size, 3, 17, samples=samples, use_torch_obs=ON_GPU)


def test_gather_64_arenas():
    """BASELINE configs[2]: gather 200x200 (495 agents + 1847 food per arena), 64 arenas, only the agent group
    observes and acts; arenas 0, 31 and 63 against independent checkers"""
    import bench
    import magent_b200 as magent
    A, size, seed = 64, 200, 40
    wl = bench.WORKLOADS["gather64"]
    env, act = bench.build_env(wl, ENGINE, A, seed0=seed)
    hs = env.get_handles()
    gi = [list(hs).index(h) for h in act]
    refs = {}
    for a in (0, 31, 63):
        r, _ = bench.build_env(wl, checker_lib(), 1, seed0=seed + a)
        refs[a] = r
    rs = np.random.RandomState(7)
    for t in range(6):
        nums = [env.get_arena_nums(h).astype(np.int64) for h in hs]
        off = [np.concatenate([[0], np.cumsum(k)]) for k in nums]
        assert all(int(k.sum()) == env.get_num(h) for k, h in zip(nums, hs))
        for g in gi:
            if ON_GPU:
                v, f = env.get_observation_torch(hs[g])
                v, f = v.cpu().numpy(), f.cpu().numpy()
            else:
                v, f = env.get_observation(hs[g])
            for a, r in refs.items():
                rv, rf = r.get_observation(r.get_handles()[g])
                sl = slice(int(off[g][a]), int(off[g][a + 1]))
                np.testing.assert_array_equal(v[sl].view(np.uint32), rv.view(np.uint32), err_msg="view t%d arena %d" % (t, a))
                np.testing.assert_array_equal(f[sl].view(np.uint32), rf.view(np.uint32), err_msg="feature t%d arena %d" % (t, a))
        acts = {g: rs.randint(0, env.get_action_space(hs[g])[0], size=int(nums[g].sum())).astype(np.int32) for g in gi}
        for g in gi:
            env.set_action(hs[g], acts[g])
            for a, r in refs.items():
                r.set_action(r.get_handles()[g], np.ascontiguousarray(acts[g][off[g][a]:off[g][a + 1]]))
        env.step()
        done = env.get_arena_done() != 0
        for a, r in refs.items():
            assert bool(done[a]) == bool(r.step())
        for g, h in enumerate(hs):
            rew, pos, alive, ids = env.get_reward(h), env.get_pos(h), env.get_alive(h), env.get_agent_id(h)
            for a, r in refs.items():
                rh = r.get_handles()[g]
                sl = slice(int(off[g][a]), int(off[g][a + 1]))
                np.testing.assert_allclose(rew[sl], r.get_reward(rh), atol=pc.REWARD_TOL, rtol=0)
                np.testing.assert_array_equal(pos[sl], r.get_pos(rh))
                np.testing.assert_array_equal(alive[sl], r.get_alive(rh))
                np.testing.assert_array_equal(ids[sl], r.get_agent_id(rh))
        # every arena starts from the same layout and differs only by its seed and its actions: arenas must
        # not leak into each other -- the food group of an arena only ever shrinks
        env.clear_dead()
        for r in refs.values():
            r.clear_dead()
        after = env.get_arena_nums(hs[0]).astype(np.int64)
        assert (after <= nums[0]).all()


# ---- the heaviest tests run last (a problem in one of them must not hide the rest under -x)
def test_battle_one_arena_of_2x400k():
    """BASELINE configs[3] as placeable (2x400k on 1000x1000; the whole-grid cooperative kernels): every record
    of every step against the reference"""
    size, n, seed = 1000, 400000, 3
    env = batched_battle(1, size, n, seed)
    ref = pc.make_battle(checker_lib(), size, n, seed)
    fs.play_battle_and_check(env, size, size, 2, 23, samples={0: ref}, use_torch_obs=ON_GPU)


def test_battle_one_arena_in_the_reference_1m_geometry():
    """BASELINE configs[3] in the reference's own 1 M-agent geometry (scripts/test/test_1m.py:66-74: map = sqrt(20 N) =
    4472x4472, 2x500k agents, 16 move bands, boundary buffer): every record of the first steps against the reference"""
    size, n, seed = 4472, 500000, 5
    env = batched_battle(1, size, n, seed)
    ref = pc.make_battle(checker_lib(), size, n, seed)
    fs.play_battle_and_check(env, size, size, 2, 31, samples={0: ref}, use_torch_obs=ON_GPU)


def test_two_huge_arenas_behind_one_handle():
    """arena batch x whole-grid kernels: 2 arenas of 2x20000 agents (more than 32768 per arena, so every arena is
    stepped by the cooperative grid one after the other); both arenas exactly against independent checkers and the
    whole batch against the PyTorch restatement"""
    A, size, n, seed = 2, 320, 20000, 60
    env = batched_battle(A, size, n, seed)
    samples = {a: pc.make_battle(checker_lib(), size, n, seed + a) for a in range(A)}
    fs.play_battle_and_check(env, size, size, 4, 29, samples=samples, use_torch_obs=ON_GPU)


SOAK_ARENAS = int(os.environ.get("MAGENT_FULLSIZE_SOAK_ARENAS", "512"))   # (smaller for a dry run on the CPU)


@pytest.mark.parametrize("workload", ["battle512", "battle512_blocks"])
def test_soak_of_the_throughput_loop(workload):
    """40 steps of bench.py's own loop (actions drawn on the device) at its own size -- 512 arenas of 2x1000 randomly
    placed agents, and of 2x1600 agents packed in two facing blocks (every move contended) -- with the whole-batch
    checks: one agent per cell, ordered ids, bounded moves and exact survivor counts after every step, every
    observation record against the PyTorch restatement every 10 steps"""
    import bench
    wl = bench.WORKLOADS[workload]
    env, _ = bench.build_env(wl, ENGINE, SOAK_ARENAS, seed0=77)
    size = wl["map_size"]
    fs.soak_battle_and_check(env, size, size, 40, 13, obs_every=10, use_torch_obs=ON_GPU)
