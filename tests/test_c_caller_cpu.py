"""The drop-in boundary from plain C: tests/c/battle_caller.c dlopen()s an engine library,
sets up the battle game through the reference ABI only and prints per-step checksums.  On the CPU the same binary
must print the same lines for the compiled reference, the C restatement and the test-only host emulation of the
engine; the CUDA library must refuse to start without a GPU (no CPU fallback)."""
import os

import pytest

import c_caller_common as cc
import parity_common as pc
from test_emu_parity_cpu import emu  # noqa: F401  (fixture: builds tests/_emu on demand)


def test_headers_are_plain_c_and_the_caller_builds():
    assert os.path.exists(cc.build())


def test_c_caller_prints_the_same_trace_on_every_cpu_engine(emu):
    libs = [p for p in (pc.REF_LIB, pc.PORT_LIB) if os.path.exists(p)] + [emu]
    assert len(libs) >= 2
    outs = []
    for lib in libs:
        r = cc.run(lib)
        assert r.returncode == 0, r.stderr
        assert r.stdout.count("\n") == 41 and "spaces view 13x13x7 feature 34 actions 21" in r.stdout
        outs.append(r.stdout)
    for lib, out in zip(libs[1:], outs[1:]):
        assert out == outs[0], "C caller trace differs between %s and %s" % (libs[0], lib)
    last = outs[0].strip().splitlines()[-1].split()
    assert (int(last[3]), int(last[4])) != (250, 250), "the scenario is supposed to see kills"


def test_c_caller_is_refused_by_the_cuda_library_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    if not os.path.exists(pc.CUDA_LIB):
        pytest.skip("CUDA library not built")
    r = cc.run(pc.CUDA_LIB, steps=2)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr
