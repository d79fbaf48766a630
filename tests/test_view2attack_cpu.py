"""get_info("view2attack") when the attack range is wider than the view: the reference writes
`ret.at(dy - y1, dx - x1) = i` through a linear index without a bounds check (GridWorld.cc:864-870, utility.h NDPointer::at),
so cells whose column is outside the view rectangle land wrapped on a neighbouring row, and indices outside the buffer
corrupt the heap.  The engine and the oracle port reproduce the in-buffer part; the test hands every library a pointer
into the middle of a large guard area so that the reference's stray writes stay inside memory we own."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import parity_common as pc


def wide_attack_config(size, view_r, attack_r):
    import magent_b200 as magent
    gw = magent.gridworld
    cfg = gw.Config()
    cfg.set({"map_width": size, "map_height": size})
    t = cfg.register_agent_type(name="t", attr={'width': 1, 'length': 1, 'hp': 3, 'speed': 1,
                                                'view_range': gw.CircleRange(view_r), 'attack_range': gw.CircleRange(attack_r),
                                                'damage': 1, 'step_recover': 0})
    cfg.add_group(t)
    return cfg


def view2attack_window(lib, view_r, attack_r):
    import magent_b200 as magent
    env = magent.GridWorld(wide_attack_config(20, view_r, attack_r), _lib=lib)
    env.reset()
    h = env.get_handles()[0]
    vh, vw, _ = env.get_view_space(h)
    guard = 4096
    buf = np.full(2 * guard + vh * vw, -7, dtype=np.int32)
    env._lib.env_get_info(env.game, env._hv(h), b"view2attack", buf.ctypes.data + 4 * guard)
    return buf[guard:guard + vh * vw].reshape(vh, vw).copy(), buf, guard, vh * vw


@pytest.mark.parametrize("view_r,attack_r", [(2, 3), (1, 3), (3, 2), (2, 2)])
def test_view2attack_with_an_attack_range_wider_than_the_view(view_r, attack_r):
    if not os.path.exists(pc.EMU_LIB):
        subprocess.run([os.path.join(pc.REPO, "tests", "emu", "build.sh")], check=True, capture_output=True)
    libs = [pc.EMU_LIB, pc.PORT_LIB]
    got = [view2attack_window(lib, view_r, attack_r) for lib in libs]
    np.testing.assert_array_equal(got[0][0], got[1][0])
    for win, buf, guard, n in got:                       # engine and port never write outside the window
        assert (buf[:guard] == -7).all() and (buf[guard + n:] == -7).all()
    if os.path.exists(pc.REF_LIB):
        want, _buf, _g, _n = view2attack_window(pc.REF_LIB, view_r, attack_r)
        np.testing.assert_array_equal(got[0][0], want)
        if attack_r > view_r:
            assert (want >= 0).sum() > 0
