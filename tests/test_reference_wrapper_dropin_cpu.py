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

import dropin_common as dc
from dropin_common import run_with

pytestmark = pytest.mark.skipif(not dc.AVAILABLE, reason="needs the reference tree (/root/reference or oracle/_ref/py) and oracle/_ref")


@pytest.mark.parametrize("game,size,steps", [("battle", 40, 30), ("pursuit", 40, 30), ("double_attack", 30, 25),
                                             ("forest", 30, 25)])
def test_reference_python_package_runs_unchanged_on_this_engine(emu, tmp_path, game, size, steps):
    want = run_with(tmp_path, pc.REF_LIB, "ref", game, size, steps)
    got = run_with(tmp_path, emu, "b200", game, size, steps)
    assert want.count("\n") == steps + 3
    assert got == want


@pytest.mark.parametrize("name", ["train_battle", "train_pursuit", "train_gather"])
def test_reference_examples_run_unchanged_through_the_reference_package(emu, tmp_path, name):
    """examples/train_*.py::play_a_round, imported unchanged, on the reference's own wrapper: same printed rounds,
    same stream of observations / ids / rewards / alive flags handed to the (stub) models"""
    want = dc.run_example(tmp_path, pc.REF_LIB, "ref", name)
    got = dc.run_example(tmp_path, emu, "b200", name)
    assert "STREAM" in want and got == want
