"""More than one device in one process: every engine owns a device context
(magent_b200/csrc/backend.h be::Ctx) -- device id, streams, events, scratch -- so engines on different GPUs can be
stepped alternately from one thread, whatever device the caller (or torch) made current in between.  Needs two GPUs:
`gpurun --gpus 2 -- python -m pytest tests/test_multigpu_gpu.py -m gpu`; skipped on a one-GPU box."""
import numpy as np
import pytest

import parity_common as pc

pytestmark = pytest.mark.gpu


def _n_gpus():
    from magent_b200.c_lib import load_library
    return load_library(pc.CUDA_LIB).magent_b200_device_count()


def test_two_engines_on_two_devices_in_one_process():
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    torch = pytest.importorskip("torch")
    from test_parity_gpu import checker_lib
    e0 = pc.make_battle(pc.CUDA_LIB, 40, 150, 0, _device=0)
    e1 = pc.make_pursuit(pc.CUDA_LIB, 40, 0, _device=1)
    r0, r1 = pc.make_battle(checker_lib(), 40, 150, 0), pc.make_pursuit(checker_lib(), 40, 0)
    rs = np.random.RandomState(0)
    for t in range(20):
        for k, (env, ref) in enumerate(((e0, r0), (e1, r1))):
            torch.cuda.set_device((t + k) % 2)                    # the caller's current device is not the engine's business
            for h, hr in zip(env.get_handles(), ref.get_handles()):
                v, f = env.get_observation(h)
                rv, rf = ref.get_observation(hr)
                np.testing.assert_array_equal(v.view(np.uint32), rv.view(np.uint32), err_msg="view t%d dev%d" % (t, k))
                np.testing.assert_array_equal(f.view(np.uint32), rf.view(np.uint32), err_msg="feat t%d dev%d" % (t, k))
        for k, (env, ref) in enumerate(((e0, r0), (e1, r1))):
            for h, hr in zip(env.get_handles(), ref.get_handles()):
                a = rs.randint(0, env.get_action_space(h)[0], size=env.get_num(h)).astype(np.int32)
                env.set_action(h, a)
                ref.set_action(hr, a)
            assert env.step() == ref.step()
            for h, hr in zip(env.get_handles(), ref.get_handles()):
                np.testing.assert_allclose(env.get_reward(h), ref.get_reward(hr), rtol=0, atol=pc.REWARD_TOL)
                np.testing.assert_array_equal(env.get_pos(h), ref.get_pos(hr))
            env.clear_dead()
            ref.clear_dead()


def test_device_buffers_on_the_engines_device():
    """device-pointer observations on device 1 while device 0 is current"""
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    torch = pytest.importorskip("torch")
    import ctypes
    from magent_b200.c_lib import load_library
    from test_parity_gpu import checker_lib
    L = load_library(pc.CUDA_LIB)
    env = pc.make_battle(pc.CUDA_LIB, 40, 150, 3, _device=1)
    ref = pc.make_battle(checker_lib(), 40, 150, 3)
    torch.cuda.set_device(0)
    for h, hr in zip(env.get_handles(), ref.get_handles()):
        n = env.get_num(h)
        v = torch.empty((n,) + env.get_view_space(h), dtype=torch.float32, device="cuda:1")
        f = torch.empty((n,) + env.get_feature_space(h), dtype=torch.float32, device="cuda:1")
        L.env_get_observation(env.game, env._hv(h), (ctypes.c_void_p * 2)(v.data_ptr(), f.data_ptr()))
        env.sync()
        rv, rf = ref.get_observation(hr)
        np.testing.assert_array_equal(v.cpu().numpy().view(np.uint32), rv.view(np.uint32))
        np.testing.assert_array_equal(f.cpu().numpy().view(np.uint32), rf.view(np.uint32))
