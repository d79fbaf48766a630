"""The literal drop-in ON THE GPU (BASELINE.json north_star; VERDICT r1 row g): the reference's own Python package
(no argtypes, 6-of-7-argument add_reward_rule, plain numpy buffers) and its own examples/train_{battle,pursuit,gather}.py
play_a_round loops, unchanged, with build/libmagent.so = the CUDA engine of this repository, against the same scripts
on the compiled reference.  Outputs (per-step digests of observations, ids, positions, rewards, alive flags; the
examples' own printed rounds) must be identical.  The reference tree travels to the GPU box as the git-ignored scratch
copy oracle/_ref/py (see dropin_common.py)."""
import pytest

import dropin_common as dc
import parity_common as pc

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not dc.AVAILABLE, reason="needs the reference tree (oracle/_ref/py) and oracle/_ref")]


@pytest.mark.parametrize("game,size,steps", [("battle", 40, 30), ("pursuit", 40, 30), ("double_attack", 30, 25),
                                             ("forest", 30, 25)])
def test_reference_python_package_on_the_cuda_library(tmp_path, game, size, steps):
    want = dc.run_with(tmp_path, pc.REF_LIB, "ref", game, size, steps)
    got = dc.run_with(tmp_path, pc.CUDA_LIB, "b200", game, size, steps)
    assert want.count("\n") == steps + 3
    assert got == want


@pytest.mark.parametrize("name", ["train_battle", "train_pursuit", "train_gather"])
def test_reference_examples_on_the_cuda_library(tmp_path, name):
    want = dc.run_example(tmp_path, pc.REF_LIB, "ref", name)
    got = dc.run_example(tmp_path, pc.CUDA_LIB, "b200", name)
    assert "STREAM" in want and got == want
