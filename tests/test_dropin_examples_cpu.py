"""Drop-in acceptance on CPU: the reference's own example scripts, imported UNCHANGED from
/root/reference/examples, run their generate_map / play_a_round against (i) the compiled reference and
(ii) this repository's engine code (the test-only emulation build of the same engine sources the CUDA
library is built from), through `import magent` -> magent_b200.  Outputs must be identical.
Skipped where /root/reference does not exist (the GPU box); the -m gpu parity suite covers the same flows there."""
import argparse
import importlib.util
import os
import sys
import types

import numpy as np
import pytest

import parity_common as pc

EXAMPLES = "/root/reference/examples"
pytestmark = pytest.mark.skipif(not (os.path.isdir(EXAMPLES) and os.path.exists(pc.REF_LIB)),
                                reason="needs /root/reference and oracle/_ref")


class StubModel:
    """stands in for the TF/MXNet DQN processes: uniform random actions from a private stream"""

    def __init__(self, env, handle, seed):
        self.n_action = env.get_action_space(handle)[0]
        self.rs = np.random.RandomState(seed)
        self._pending = None

    def infer_action(self, obs, ids, policy=None, eps=None, block=True):
        assert obs[0].shape[0] == len(ids) == obs[1].shape[0]
        acts = self.rs.randint(0, self.n_action, size=len(ids)).astype(np.int32)
        self._pending = acts
        return acts

    def fetch_action(self):
        return self._pending

    def sample_step(self, rewards, alives, block=True):
        assert len(rewards) == len(alives)

    def check_done(self):
        pass


def load_example(name):
    import magent  # noqa: F401  alias package -> magent_b200
    for backend in ("tf_model", "mx_model"):
        mod = types.ModuleType("magent.builtin." + backend)
        mod.DeepQNetwork = object
        sys.modules["magent.builtin." + backend] = mod
    spec = importlib.util.spec_from_file_location("ref_example_" + name, os.path.join(EXAMPLES, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def emu():
    import subprocess
    if not os.path.exists(pc.EMU_LIB):
        subprocess.run([os.path.join(pc.REPO, "tests", "emu", "build.sh")], check=True, capture_output=True)
    return pc.EMU_LIB


def test_train_battle_play_a_round(emu, capsys):
    import magent
    ex = load_example("train_battle")
    ex.args = argparse.Namespace(train=False)

    def run(lib):
        ex.leftID, ex.rightID = 0, 1
        env = magent.GridWorld("battle", map_size=50, _lib=lib)
        handles = env.get_handles()
        models = [StubModel(env, h, 10 + i) for i, h in enumerate(handles)]
        out = ex.play_a_round(env, 50, handles, models, print_every=50, train=False, render=False, eps=0.5)
        return out, [env.get_num(h) for h in handles]
    assert run(pc.REF_LIB) == run(emu)


def test_train_pursuit_play_a_round(emu, capsys):
    import magent
    ex = load_example("train_pursuit")

    def run(lib):
        env = magent.GridWorld("pursuit", map_size=40, _lib=lib)
        handles = env.get_handles()
        models = [StubModel(env, h, 20 + i) for i, h in enumerate(handles)]
        return ex.play_a_round(env, 40, handles, models, print_every=100, train=False, render=False, eps=0.3)
    assert run(pc.REF_LIB) == run(emu)


def test_train_gather_generate_map_and_round(emu, capsys):
    import magent
    ex = load_example("train_gather")

    def run(lib):
        env = magent.GridWorld(ex.load_config(size=80), _lib=lib)
        handles = env.get_handles()
        food_handle, player_handles = handles[0], handles[1:]
        models = [StubModel(env, h, 30 + i) for i, h in enumerate(player_handles)]
        out = ex.play_a_round(env, 80, food_handle, player_handles, models, train_id=-1, print_every=100, eps=0.2)
        return out, env.get_num(food_handle), [env.get_num(h) for h in player_handles], env.get_pos(player_handles[0]).tolist()
    a, b = run(pc.REF_LIB), run(emu)
    assert a == b
