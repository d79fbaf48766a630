"""Shared pieces of the literal drop-in tests: the reference's OWN Python package (python/magent, untouched, bound with
ctypes and no argtypes, python/magent/c_lib.py:10-22) and its own example scripts on top of an engine library.

Where the reference comes from: /root/reference when it exists (the build container); otherwise the scratch copy that
`__graft_entry__.build()` places under oracle/_ref/py/ (git-ignored like oracle/_ref/libmagent.so, shipped to the GPU
box with it).  Nothing of it ever enters the repository's history."""
import os
import shutil
import subprocess
import sys

import parity_common as pc

_CANDIDATES = ["/root/reference", os.path.join(pc.REPO, "oracle", "_ref", "py")]
REF_ROOT = next((r for r in _CANDIDATES if os.path.isdir(os.path.join(r, "python", "magent"))), None)
REF_PKG = os.path.join(REF_ROOT, "python", "magent") if REF_ROOT else None
EXAMPLES = os.path.join(REF_ROOT, "examples") if REF_ROOT else None
AVAILABLE = REF_PKG is not None and os.path.exists(pc.REF_LIB)

# plays the built-in games through the reference wrapper and prints per-step digests
DRIVER = r'''
import hashlib, sys
import numpy as np
import magent
assert "/magent_b200" not in (magent.__file__ or "") and "/root/repo/magent/" not in (magent.__file__ or ""), magent.__file__
game, size, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
env = magent.GridWorld(game, map_size=size)
env.set_seed(3)
env.reset()
hs = env.get_handles()
env.add_walls(method="random", n=size)
env.add_walls(method="fill", pos=(size // 2, 2), size=(2, 3))
for i, h in enumerate(hs):
    env.add_agents(h, method="random", n=size * size // (25 if game != "pursuit" else 60))
    env.add_agents(h, method="fill", pos=(3 + 9 * i, size - 8), size=(4, 4))
    env.add_agents(h, method="custom", pos=[[5 + i, 5], [6 + i, 7]])
print("spaces", [(env.get_view_space(h), env.get_feature_space(h), env.get_action_space(h)) for h in hs])
print("view2attack", [hashlib.sha256(env.get_view2attack(h)[1].tobytes()).hexdigest()[:16] for h in hs])
rs = np.random.RandomState(11)
d = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]
for t in range(steps):
    line = ["t%d" % t]
    for h in hs:
        n = env.get_num(h)
        if n:
            v, f = env.get_observation(h)
            line += [d(v), d(f)]
        env.set_action(h, rs.randint(0, env.get_action_space(h)[0], size=n).astype(np.int32))
    done = env.step()
    for h in hs:
        line += [d(env.get_pos(h)), d(env.get_agent_id(h)), d(env.get_alive(h)),
                 "%.5f" % float(np.asarray(env.get_reward(h), dtype=np.float64).sum())]
    env.clear_dead()
    line += [str(done), str([env.get_num(h) for h in hs])]
    print(" ".join(line))
print("minimap", d(env.get_global_minimap(10, 10)))
'''

# imports examples/train_{battle,pursuit,gather}.py UNCHANGED (the RL model packages stubbed out: TF / MXNet are not
# part of the engine), runs their own generate_map / play_a_round loop with stub models that draw uniform random
# actions and digests everything the loop hands to the models
EXAMPLES_DRIVER = r'''
import argparse, hashlib, importlib.util, os, sys, types
import collections, collections.abc
collections.Iterable = collections.abc.Iterable     # the reference predates Python 3.10 (python/magent/utility.py:205)
import numpy as np
import magent
assert "/magent_b200" not in (magent.__file__ or "") and "/root/repo/magent/" not in (magent.__file__ or ""), magent.__file__
name, examples = sys.argv[1], sys.argv[2]
for backend in ("tf_model", "mx_model"):
    mod = types.ModuleType("magent.builtin." + backend)
    mod.DeepQNetwork = object
    sys.modules["magent.builtin." + backend] = mod
spec = importlib.util.spec_from_file_location("ref_example_" + name, os.path.join(examples, name + ".py"))
ex = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ex)
H = hashlib.sha256()

class StubModel:
    def __init__(self, env, handle, seed):
        self.n_action = env.get_action_space(handle)[0]
        self.rs = np.random.RandomState(seed)
        self._pending = None
    def infer_action(self, obs, ids, policy=None, eps=None, block=True):
        assert obs[0].shape[0] == len(ids) == obs[1].shape[0]
        H.update(np.ascontiguousarray(obs[0]).tobytes()); H.update(np.ascontiguousarray(obs[1]).tobytes())
        H.update(np.ascontiguousarray(ids).tobytes())
        self._pending = self.rs.randint(0, self.n_action, size=len(ids)).astype(np.int32)
        return self._pending
    def fetch_action(self):
        return self._pending
    def sample_step(self, rewards, alives, block=True):
        assert len(rewards) == len(alives)
        H.update(np.ascontiguousarray(alives).tobytes())
        H.update(np.round(np.asarray(rewards, dtype=np.float64), 4).tobytes())
    def check_done(self):
        pass
    def train(self, *a, **k):
        return 0.0, 0.0
    def fetch_train(self):
        return 0.0, 0.0

ex.args = argparse.Namespace(train=True)
if name == "train_battle":
    ex.leftID, ex.rightID = 0, 1
    env = magent.GridWorld("battle", map_size=50)
    handles = env.get_handles()
    models = [StubModel(env, h, 10 + i) for i, h in enumerate(handles)]
    out = ex.play_a_round(env, 50, handles, models, print_every=50, train=True, render=False, eps=0.5)
elif name == "train_pursuit":
    env = magent.GridWorld("pursuit", map_size=40)
    handles = env.get_handles()
    models = [StubModel(env, h, 20 + i) for i, h in enumerate(handles)]
    out = ex.play_a_round(env, 40, handles, models, print_every=100, train=True, render=False, eps=0.3)
else:
    env = magent.GridWorld(ex.load_config(size=80))
    handles = env.get_handles()
    food_handle, player_handles = handles[0], handles[1:]
    models = [StubModel(env, h, 30 + i) for i, h in enumerate(player_handles)]
    out = ex.play_a_round(env, 80, food_handle, player_handles, models, train_id=-1, print_every=100, eps=0.2)
    handles = player_handles
print("RESULT", repr(out))
print("NUMS", [env.get_num(h) for h in handles], "POS", hashlib.sha256(env.get_pos(handles[0]).tobytes()).hexdigest()[:16])
print("STREAM", H.hexdigest())
'''


def scratch_tree(tmp_path, lib, tag):
    """<root>/python/magent = scratch copy of the reference package; <root>/build/libmagent.so -> lib (the reference
    loader looks for <package>/../../build/libmagent.so)"""
    root = tmp_path / tag
    (root / "python").mkdir(parents=True)
    (root / "build").mkdir()
    shutil.copytree(REF_PKG, str(root / "python" / "magent"))
    os.symlink(lib, str(root / "build" / "libmagent.so"))
    return root


def run_driver(tmp_path, lib, tag, source, argv, timeout=900):
    root = scratch_tree(tmp_path, lib, tag)
    script = root / "driver.py"
    script.write_text(source)
    env = dict(os.environ, PYTHONPATH=str(root / "python"), OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, str(script)] + [str(a) for a in argv], capture_output=True, text=True,
                       env=env, cwd=str(root), timeout=timeout)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-2500:])
    return r.stdout


def run_with(tmp_path, lib, tag, game, size, steps):
    return run_driver(tmp_path, lib, tag, DRIVER, [game, size, steps])


def run_example(tmp_path, lib, tag, name):
    out = run_driver(tmp_path, lib, tag, EXAMPLES_DRIVER, [name, EXAMPLES])
    # wall-clock figures the example prints differ from run to run: keep the deterministic lines
    return "\n".join(l for l in out.splitlines() if l.startswith(("RESULT", "NUMS", "STREAM", "step ", "eps ")))
