"""The literal drop-in: the reference's OWN Python package (python/magent, untouched, bound with ctypes and no
argtypes) on top of this repository's engine.

A scratch directory gets a COPY of /root/reference/python/magent (the reference loader looks for
`<package>/../../build/libmagent.so`, python/magent/c_lib.py:12-22) next to a `build/libmagent.so` that is a symlink
to (i) the compiled reference and (ii) the test-only host build of this repository's engine sources (the same
engine.cc / shim.cc / phase functions the CUDA library is built from).  One driver script, run in a subprocess per
library, plays the reference's built-in games through the reference wrapper and prints per-step digests; the two
outputs must be identical.  This exercises the calling convention the reference really uses: Python ints as C ints,
handles as c_int32 objects, 6 of 7 arguments to gridworld_add_reward_rule, 8 arguments for "fill".
Nothing is copied into the repository; skipped where /root/reference does not exist (the GPU box)."""
import os
import shutil
import subprocess
import sys

import pytest

import parity_common as pc
from test_emu_parity_cpu import emu  # noqa: F401  (fixture: builds tests/_emu on demand)

REF_PKG = "/root/reference/python/magent"
pytestmark = pytest.mark.skipif(not (os.path.isdir(REF_PKG) and os.path.exists(pc.REF_LIB)),
                                reason="needs /root/reference and oracle/_ref")

DRIVER = r'''
import hashlib, sys
import numpy as np
import magent
assert "/root/repo" not in (magent.__file__ or ""), magent.__file__      # the reference package, not the mirror
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


def run_with(tmp_path, lib, tag, game, size, steps):
    root = tmp_path / tag
    (root / "python").mkdir(parents=True)
    (root / "build").mkdir()
    shutil.copytree(REF_PKG, str(root / "python" / "magent"))          # scratch copy, never enters the repository
    os.symlink(lib, str(root / "build" / "libmagent.so"))
    script = root / "driver.py"
    script.write_text(DRIVER)
    env = dict(os.environ, PYTHONPATH=str(root / "python"), OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, str(script), game, str(size), str(steps)], capture_output=True, text=True,
                       env=env, cwd=str(root), timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


@pytest.mark.parametrize("game,size,steps", [("battle", 40, 30), ("pursuit", 40, 30), ("double_attack", 30, 25),
                                             ("forest", 30, 25)])
def test_reference_python_package_runs_unchanged_on_this_engine(emu, tmp_path, game, size, steps):
    want = run_with(tmp_path, pc.REF_LIB, "ref", game, size, steps)
    got = run_with(tmp_path, emu, "b200", game, size, steps)
    assert want.count("\n") == steps + 3
    assert got == want
