"""CPU-side checks of the drop-in boundary: the built library loads, exports every symbol declared in
include/*.h, and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

import parity_common as pc

INCLUDE = os.path.join(pc.REPO, "include")


def declared_symbols(header):
    text = open(os.path.join(INCLUDE, header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.findall(r"\b((?:env|gridworld|discrete_snake|magent_b200)_\w+)\s*\(", text)


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(pc.CUDA_LIB):
        import __graft_entry__ as g
        g.build()
    return ctypes.CDLL(pc.CUDA_LIB, mode=ctypes.RTLD_LOCAL)


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols("magent_runtime_api.h") + declared_symbols("magent_b200_ext.h")
    assert len(names) >= 21 + 9
    for name in names:
        assert hasattr(lib, name), "missing export " + name


def test_reference_symbol_set_is_covered(lib):
    """the 21 unmangled entry points of the reference header (src/runtime_api.h:20-61)"""
    ref = ["env_new_game", "env_delete_game", "env_config_game", "env_reset", "env_get_observation",
           "env_set_action", "env_step", "env_get_reward", "env_get_info", "env_render", "env_render_next_file",
           "gridworld_register_agent_type", "gridworld_new_group", "gridworld_add_agents", "gridworld_clear_dead",
           "gridworld_set_goal", "gridworld_define_agent_symbol", "gridworld_define_event_node",
           "gridworld_add_reward_rule", "discrete_snake_clear_dead", "discrete_snake_add_object"]
    assert sorted(ref) == sorted(declared_symbols("magent_runtime_api.h"))
    for name in ref:
        assert hasattr(lib, name)
    if os.path.exists(pc.REF_LIB):
        r = ctypes.CDLL(pc.REF_LIB, mode=ctypes.RTLD_LOCAL)
        for name in ref:
            assert hasattr(r, name)


def test_python_abi_table_matches_header():
    from magent_b200.c_lib import ABI_SIGNATURES, EXT_SIGNATURES
    assert sorted(ABI_SIGNATURES) == sorted(declared_symbols("magent_runtime_api.h"))
    assert sorted(EXT_SIGNATURES) == sorted(declared_symbols("magent_b200_ext.h"))


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib.magent_b200_device_count.restype = ctypes.c_int
    assert lib.magent_b200_device_count() == 0
    import magent_b200 as magent
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        magent.GridWorld("battle", map_size=30, _lib=pc.CUDA_LIB)


def test_product_library_does_not_link_the_oracle():
    out = subprocess.run(["ldd", pc.CUDA_LIB], capture_output=True, text=True).stdout
    assert "oracle" not in out and "magent_emu" not in out and "gomp" not in out
    # and the package never mentions the oracle or the emulation
    pkg = os.path.join(pc.REPO, "magent_b200")
    for root, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".h", ".cc", ".cu")):
                text = open(os.path.join(root, f)).read()
                assert "oracle/" not in text.replace("oracle/_ref", "").replace("oracle/Makefile", "") or f in ("gridworld.py", "c_lib.py"), f
                assert "libmagent_emu" not in text and "libmagent_oracle" not in text, f


def test_alias_package_resolves():
    code = "import magent, magent.gridworld as gw, magent.builtin.config.battle as b; print(magent.GridWorld.__name__, gw.Config.__name__)"
    out = subprocess.run([sys.executable, "-c", code], cwd=pc.REPO, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert out.stdout.split() == ["GridWorld", "Config"]
