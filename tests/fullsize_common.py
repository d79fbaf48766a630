"""Size-independent checks for the battle game at BASELINE.json's full sizes (test infrastructure).

At 512 arenas x 2x1000 agents or 2x400k agents in one arena the sequential reference cannot replay every arena
in test time, so the full-size tests combine

  * exact parity of SAMPLED arenas against independent checker environments (tests/test_zz_fullsize_gpu.py), with
  * properties of the WHOLE batch that need no second engine:
      - `check_state`: every live agent stands on its own in-board, non-wall cell (cell keys of an arena are
        unique), ids of a group are strictly increasing inside an arena (clear_dead is a stable compaction of
        ids handed out in order), nobody moved further than its speed;
      - `battle_observation`: a plain PyTorch restatement of get_observation for the battle config
        (GridWorld.cc:292-401, Map.cc:129-207: 1x1 bodies, heading north, CircleRange view, minimap_mode) that
        rebuilds EVERY agent's view and feature row from the engine's own positions / ids / last actions / last
        rewards and is compared bit for bit.  The hp an observer sees on a cell is taken from the occupant's own
        centre cell (`hp / max_hp` is the same float whoever looks at it), the in-range mask of the view is
        CircleRange's `dx^2 + dy^2 <= R^2` (Range.h:149-190; 113 cells for radius 6).

The restatement itself is pinned on the CPU against the compiled reference (tests/test_fullsize_cpu.py).
"""
import numpy as np
import torch


def device():
    return torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")


def arena_index(nums, dev):
    """arena id of every record of a group's concatenation (nums = per-arena counts)"""
    nums_t = torch.as_tensor(np.asarray(nums, dtype=np.int64), device=dev)
    return torch.repeat_interleave(torch.arange(len(nums), device=dev), nums_t)


def check_state(pos, ids, nums, width, height, prev=None, speed=None):
    """pos: list (per group) of int32 [n_g, 2] (x, y); ids: list of int32 [n_g]; nums: list of per-arena counts.
    prev = (prev_pos_by_id dict arrays) is optional: (ids, pos) lists of the previous step for the speed check."""
    keys = []
    for g, (p, i, n) in enumerate(zip(pos, ids, nums)):
        assert p.shape == (int(np.sum(n)), 2) and i.shape == (int(np.sum(n)),)
        x, y = p[:, 0].astype(np.int64), p[:, 1].astype(np.int64)
        assert (x >= 1).all() and (x <= width - 2).all() and (y >= 1).all() and (y <= height - 2).all(), \
            "group %d: an agent stands on the border wall or off the board" % g
        a = np.repeat(np.arange(len(n), dtype=np.int64), n)
        keys.append((a * height + y) * width + x)
        # ids inside one arena strictly increase (stable compaction of ids handed out in order)
        d = np.diff(i.astype(np.int64))
        same_arena = np.diff(a) == 0
        assert (d[same_arena] > 0).all(), "group %d: ids out of order inside an arena" % g
    allk = np.concatenate(keys)
    assert np.unique(allk).size == allk.size, "two live agents share a cell"
    if prev is not None:
        for g in range(len(pos)):
            pid, ppos, pn = prev[0][g], prev[1][g], prev[2][g]
            # match survivors by (arena, id)
            a_now = np.repeat(np.arange(len(nums[g]), dtype=np.int64), nums[g])
            a_prev = np.repeat(np.arange(len(pn), dtype=np.int64), pn)
            k_now = a_now * (1 << 32) + ids[g].astype(np.int64)
            k_prev = a_prev * (1 << 32) + pid.astype(np.int64)
            idx = np.searchsorted(k_prev, k_now)
            assert (idx < k_prev.size).all() and (k_prev[idx] == k_now).all(), "an agent appeared from nowhere"
            d = np.abs(pos[g].astype(np.int64) - ppos[idx].astype(np.int64))
            assert speed is None or (d[:, 0] ** 2 + d[:, 1] ** 2 <= speed ** 2).all(), \
                "group %d: an agent moved further than its speed" % g


def learned_mask(views, own_channels=(0, 1, 4)):
    """union over all observers of the cells where a wall / agent was marked -> bool [H_v, W_v]"""
    m = None
    for v in views:
        if v.shape[0] == 0:
            continue
        mm = (v[..., list(own_channels)] != 0).any(dim=3).any(dim=0)
        m = mm if m is None else (m | mm)
    return m


def battle_observation(pos, ids, nums, last_action, last_reward, centre_hp, xy_feature, mask, width, height,
                       view_hw=13, n_action=21, embedding=10, chunk=65536):
    """Yield (group, start, stop, view, feature) blocks of the expected battle observation (float32 tensors).

    pos / ids / last_action / last_reward / centre_hp: per-group tensors on one device ([n,2] int, [n] int,
    [n] int, [n] float32, [n] float32); xy_feature: per-group float32 [n, 2] = (x / width, y / height) divided on
    the host in IEEE float32 (a CUDA tensor divided by a Python scalar is multiplied by the rounded reciprocal, which
    is not the reference's division); nums: per-group per-arena counts (numpy); mask: bool [13,13]."""
    dev = pos[0].device
    G = len(pos)
    A = len(nums[0])
    R = view_hw // 2
    P = R
    Hp, Wp = height + 2 * P, width + 2 * P
    kind = torch.zeros((A, Hp, Wp), dtype=torch.int8, device=dev)
    kind[:, P, P:P + width] = 1
    kind[:, P + height - 1, P:P + width] = 1
    kind[:, P:P + height, P] = 1
    kind[:, P:P + height, P + width - 1] = 1
    hpn = torch.zeros((A, Hp, Wp), dtype=torch.float32, device=dev)
    sw = (width + view_hw - 1) // view_hw
    sh = (height + view_hw - 1) // view_hw
    mm = torch.zeros((A, G, view_hw, view_hw), dtype=torch.float32, device=dev)
    arena = [arena_index(n, dev) for n in nums]
    for g in range(G):
        x, y = pos[g][:, 0].long(), pos[g][:, 1].long()
        kind[arena[g], y + P, x + P] = 2 + g
        hpn[arena[g], y + P, x + P] = centre_hp[g]
        flat = (arena[g] * G + g) * (view_hw * view_hw) + (y // sh) * view_hw + (x // sw)
        cnt = torch.bincount(flat, minlength=A * G * view_hw * view_hw).view(A, G, view_hw, view_hw)
        # counts / total in float32 (GridWorld.cc:331-360: `minimap /= total_ct`), divided on the host (IEEE)
        tot = np.asarray(nums[g]).astype(np.float32).reshape(A, 1, 1)
        with np.errstate(invalid="ignore", divide="ignore"):
            mm[:, g] = torch.from_numpy(cnt[:, g].cpu().numpy().astype(np.float32) / tot).to(dev)
    dy, dx = torch.meshgrid(torch.arange(-R, R + 1, device=dev), torch.arange(-R, R + 1, device=dev), indexing="ij")
    bits = torch.arange(embedding, device=dev)
    for g in range(G):
        n = pos[g].shape[0]
        for s in range(0, n, chunk):
            e = min(n, s + chunk)
            x, y = pos[g][s:e, 0].long(), pos[g][s:e, 1].long()
            ar = arena[g][s:e]
            yy = (y + P).view(-1, 1, 1) + dy
            xx = (x + P).view(-1, 1, 1) + dx
            k = kind[ar.view(-1, 1, 1), yy, xx]
            hv = hpn[ar.view(-1, 1, 1), yy, xx]
            view = torch.zeros((e - s, view_hw, view_hw, 1 + 3 * G), dtype=torch.float32, device=dev)
            view[..., 0] = ((k == 1) & mask).float()
            order = [g] + [o for o in range(G) if o != g]          # own group first (GridWorld.cc:897-913)
            for slot, og in enumerate(order):
                here = (k == 2 + og) & mask
                view[..., 1 + 3 * slot] = here.float()
                view[..., 2 + 3 * slot] = torch.where(here, hv, torch.zeros_like(hv))
                ch = mm[ar, og].clone()
                ch[torch.arange(e - s, device=dev), y // sh, x // sw] += 1.0
                view[..., 3 + 3 * slot] = ch
            feat = torch.zeros((e - s, embedding + n_action + 3), dtype=torch.float32, device=dev)
            feat[:, :embedding] = ((ids[g][s:e].long().view(-1, 1) >> bits) & 1).float()
            feat[torch.arange(e - s, device=dev), embedding + last_action[g][s:e].long()] = 1.0
            feat[:, embedding + n_action] = last_reward[g][s:e]
            feat[:, embedding + n_action + 1:embedding + n_action + 3] = xy_feature[g][s:e]
            yield g, s, e, view, feat


def assert_bits_equal(got, want, what):
    a = got.contiguous().view(torch.int32)
    b = want.contiguous().view(torch.int32)
    if not torch.equal(a, b):
        bad = (a != b).nonzero()
        first = bad[0].tolist()
        raise AssertionError("%s: %d elements differ, first at %s: got %r want %r" % (
            what, bad.shape[0], first, got[tuple(first)].item(), want[tuple(first)].item()))


def check_battle_observation(views, feats, pos, ids, nums, last_action, last_reward, width, height,
                             expect_mask_cells=113):
    """views / feats: per-group float32 tensors as the engine returned them (any device).  Everything else numpy."""
    dev = views[0].device
    R = views[0].shape[1] // 2
    # CircleRange(R) (Range.h:149-190): cells whose distance from the centre is < R + 1e-8, i.e. dx^2 + dy^2 <= R^2
    d = torch.arange(-R, R + 1, device=dev)
    mask = (d.view(-1, 1) ** 2 + d.view(1, -1) ** 2) <= R * R
    if expect_mask_cells is not None:
        assert int(mask.sum()) == expect_mask_cells
    seen = learned_mask(views)
    assert seen is not None and bool((seen & ~mask).sum() == 0), "something is marked outside the view range"
    centre = [v[:, R, R, 2].contiguous() for v in views]
    for g, c in enumerate(centre):
        assert bool(((c > 0) & (c <= 1)).all()), "group %d: a live observer's own hp/max_hp is outside (0, 1]" % g
        assert bool((views[g][:, R, R, 1] == 1).all()), "group %d: observer missing from its own centre cell" % g
    t = lambda arrs, dt: [torch.as_tensor(np.ascontiguousarray(a), device=dev).to(dt) for a in arrs]
    xy = [np.stack([p[:, 0].astype(np.float32) / np.float32(width), p[:, 1].astype(np.float32) / np.float32(height)],
                   axis=1) for p in pos]
    exp = battle_observation(t(pos, torch.int64), t(ids, torch.int64), nums, t(last_action, torch.int64),
                             t(last_reward, torch.float32), centre, t(xy, torch.float32), mask, width, height,
                             view_hw=views[0].shape[1], n_action=feats[0].shape[1] - 13)
    for g, s, e, v, f in exp:
        assert_bits_equal(views[g][s:e], v, "view of group %d records %d..%d" % (g, s, e))
        assert_bits_equal(feats[g][s:e], f, "feature of group %d records %d..%d" % (g, s, e))
    return mask


def play_battle_and_check(env, width, height, steps, seed, samples=None, use_torch_obs=False, speed=2, expect_mask_cells=113):
    """Drive the standard loop on a (possibly batched) battle environment with host-generated uniform actions and
    check every step: whole-batch state invariants, the whole-batch observation against the PyTorch restatement,
    and -- `samples` = {arena: independent single-arena checker environment} -- exact parity of those arenas."""
    hs = env.get_handles()
    G = len(hs)
    A = getattr(env, "num_arenas", 1)
    W, H = width, height
    rs = np.random.RandomState(seed)
    n_action = [env.get_action_space(h)[0] for h in hs]
    nums = [env.get_arena_nums(h).astype(np.int64) if A > 1 else np.array([env.get_num(h)], dtype=np.int64) for h in hs]
    last_action = [np.full(int(n.sum()), n_action[g], dtype=np.int64) for g, n in enumerate(nums)]
    last_reward = [np.zeros(int(n.sum()), dtype=np.float32) for n in nums]
    prev = None
    samples = samples or {}
    for t in range(steps):
        pos = [env.get_pos(h).copy() for h in hs]
        ids = [env.get_agent_id(h).copy() for h in hs]
        check_state(pos, ids, nums, W, H, prev=prev, speed=speed)
        if use_torch_obs:
            obs = [env.get_observation_torch(h) for h in hs]
            views, feats = [o[0] for o in obs], [o[1] for o in obs]
        else:
            obs = [env.get_observation(h) for h in hs]
            views = [torch.from_numpy(o[0].copy()).to(device()) for o in obs]
            feats = [torch.from_numpy(o[1].copy()).to(device()) for o in obs]
        check_battle_observation(views, feats, pos, ids, nums, last_action, last_reward, W, H,
                                 expect_mask_cells=expect_mask_cells)
        off = [np.concatenate([[0], np.cumsum(n)]) for n in nums]
        for a, ref in samples.items():
            for g, rh in enumerate(ref.get_handles()):
                rv, rf = ref.get_observation(rh)
                sl = slice(int(off[g][a]), int(off[g][a + 1]))
                assert rv.shape[0] == sl.stop - sl.start, "arena %d group %d: %d agents, checker has %d" % (
                    a, g, sl.stop - sl.start, rv.shape[0])
                np.testing.assert_array_equal(views[g][sl].cpu().numpy().view(np.uint32), rv.view(np.uint32),
                                              err_msg="view t%d arena %d g%d" % (t, a, g))
                np.testing.assert_array_equal(feats[g][sl].cpu().numpy().view(np.uint32), rf.view(np.uint32),
                                              err_msg="feature t%d arena %d g%d" % (t, a, g))
                np.testing.assert_array_equal(pos[g][sl], ref.get_pos(rh))
                np.testing.assert_array_equal(ids[g][sl], ref.get_agent_id(rh))
        acts = [rs.randint(0, n_action[g], size=int(nums[g].sum())).astype(np.int32) for g in range(G)]
        for g, h in enumerate(hs):
            env.set_action(h, acts[g])
            for a, ref in samples.items():
                ref.set_action(ref.get_handles()[g], np.ascontiguousarray(acts[g][off[g][a]:off[g][a + 1]]))
        env.step()
        done = env.get_arena_done() != 0 if A > 1 else None
        rew = [env.get_reward(h) for h in hs]
        alive = [env.get_alive(h) for h in hs]
        for a, ref in samples.items():
            d = ref.step()
            if done is not None:
                assert bool(done[a]) == bool(d), "done flag of arena %d" % a
            for g, rh in enumerate(ref.get_handles()):
                sl = slice(int(off[g][a]), int(off[g][a + 1]))
                np.testing.assert_allclose(rew[g][sl], ref.get_reward(rh), atol=1e-6, rtol=0)
                np.testing.assert_array_equal(alive[g][sl], ref.get_alive(rh))
                np.testing.assert_array_equal(env.get_pos(hs[g])[sl], ref.get_pos(rh))
            ref.clear_dead()
        env.clear_dead()
        # what the next observation must show: survivors keep their action and the reward they just got
        prev = (ids, pos, nums)
        keep = [al.astype(bool) for al in alive]
        new_nums = [env.get_arena_nums(h).astype(np.int64) if A > 1 else np.array([env.get_num(h)], dtype=np.int64) for h in hs]
        for g in range(G):
            ar = np.repeat(np.arange(len(nums[g])), nums[g])
            np.testing.assert_array_equal(np.bincount(ar[keep[g]], minlength=len(nums[g])), new_nums[g],
                                          err_msg="clear_dead kept a different number of agents than were alive")
        last_action = [acts[g][keep[g]].astype(np.int64) for g in range(G)]
        last_reward = [rew[g][keep[g]].astype(np.float32) for g in range(G)]
        nums = new_nums
    return t + 1


def soak_battle_and_check(env, width, height, steps, seed, obs_every=8, use_torch_obs=False, speed=2,
                          expect_mask_cells=113):
    """Many steps of the throughput loop (uniform actions drawn on the device, `set_random_actions`) on a batched
    battle environment with the whole-batch checks only: state invariants after every step, the PyTorch restatement
    of every observation record every `obs_every` steps.  The actions are not known to the host, so the one-hot of
    the last action is read from the feature row itself and only required to BE a one-hot (or empty before the first
    action); everything else is rebuilt and compared bit for bit."""
    hs = env.get_handles()
    G = len(hs)
    A = getattr(env, "num_arenas", 1)
    n_action = env.get_action_space(hs[0])[0]
    emb = env.get_feature_space(hs[0])[0] - n_action - 3
    arena_nums = lambda: [env.get_arena_nums(h).astype(np.int64) if A > 1 else np.array([env.get_num(h)], dtype=np.int64)
                          for h in hs]
    nums = arena_nums()
    last_reward = [np.zeros(int(n.sum()), dtype=np.float32) for n in nums]
    prev = None
    deaths = 0
    for t in range(steps):
        pos = [env.get_pos(h).copy() for h in hs]
        ids = [env.get_agent_id(h).copy() for h in hs]
        check_state(pos, ids, nums, width, height, prev=prev, speed=speed)
        if t % obs_every == 0 or t == steps - 1:
            if use_torch_obs:
                obs = [env.get_observation_torch(h) for h in hs]
                views, feats = [o[0] for o in obs], [o[1] for o in obs]
            else:
                obs = [env.get_observation(h) for h in hs]
                views = [torch.from_numpy(o[0].copy()).to(device()) for o in obs]
                feats = [torch.from_numpy(o[1].copy()).to(device()) for o in obs]
            last_action = []
            for g in range(G):
                hot = feats[g][:, emb:emb + n_action]
                assert bool(((hot == 0) | (hot == 1)).all()), "last-action slots hold something else than 0 / 1"
                cnt = hot.sum(dim=1)
                assert bool((cnt == (0 if t == 0 else 1)).all()), "last-action slots are not a one-hot"
                la = hot.argmax(dim=1) if t > 0 else torch.full((hot.shape[0],), n_action, device=hot.device)
                last_action.append(la.cpu().numpy().astype(np.int64))
            check_battle_observation(views, feats, pos, ids, nums, last_action, last_reward, width, height,
                                     expect_mask_cells=expect_mask_cells)
        for h in hs:
            env.set_random_actions(h, seed * 1000 + t)
        env.step()
        rew = [env.get_reward(h) for h in hs]
        alive = [env.get_alive(h).astype(bool) for h in hs]
        env.clear_dead()
        prev = (ids, pos, nums)
        new_nums = arena_nums()
        for g in range(G):
            ar = np.repeat(np.arange(len(nums[g])), nums[g])
            np.testing.assert_array_equal(np.bincount(ar[alive[g]], minlength=len(nums[g])), new_nums[g],
                                          err_msg="clear_dead kept a different number of agents than were alive")
            deaths += int((~alive[g]).sum())
        last_reward = [rew[g][alive[g]].astype(np.float32) for g in range(G)]
        nums = new_nums
    return deaths
