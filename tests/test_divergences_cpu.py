"""Pinned behaviour where the engine is deliberately safer than the reference (tests/divergence_common.py); CPU leg on the
host emulation of the engine sources, GPU leg in tests/test_parity_gpu.py."""
import os

import pytest

import divergence_common as dv
import parity_common as pc
from test_emu_parity_cpu import emu  # noqa: F401


def _checker():
    return pc.REF_LIB if os.path.exists(pc.REF_LIB) else pc.PORT_LIB


def test_second_set_action_of_a_step_wins(emu):
    dv.second_set_action_wins(emu, _checker())


def test_out_of_range_actions_are_ignored(emu):
    dv.invalid_actions_are_ignored(emu, _checker())
