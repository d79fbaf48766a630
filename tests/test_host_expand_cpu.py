"""The host half of env_get_observation (magent_b200/csrc/host_expand.cc): wire records -> dense float32 records.

Runs on CPU against the test-only host emulation, which produces the same wire records (WireHdr / WireMark,
backend.h) the CUDA kernels do, so that the expansion threads, the chunk / wave protocol, the persistent-record undo
logic and the unaligned-buffer handling are checked bit for bit against the compiled reference (oracle/_ref) and
against the dense path of the same engine.  The `-m gpu` tests repeat this with the CUDA producer."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import parity_common as pc


@pytest.fixture(scope="module")
def emu():
    src_dir = os.path.join(pc.REPO, "magent_b200", "csrc")
    newest = max(os.path.getmtime(os.path.join(src_dir, f)) for f in os.listdir(src_dir))
    newest = max(newest, os.path.getmtime(os.path.join(pc.REPO, "tests", "emu", "backend_emu.cc")))
    if not os.path.exists(pc.EMU_LIB) or os.path.getmtime(pc.EMU_LIB) < newest:
        subprocess.run([os.path.join(pc.REPO, "tests", "emu", "build.sh")], check=True, capture_output=True)
    return pc.EMU_LIB


def checker():
    return pc.REF_LIB if os.path.exists(pc.REF_LIB) else pc.PORT_LIB


MAKERS = {
    "battle": lambda lib, **kw: pc.make_battle(lib, 40, 150, 0, **kw),                 # minimap channels, self marker
    "battle_blocks": lambda lib, **kw: pc.make_battle_blocks(lib, 40, **kw),           # dense marks
    "pursuit": lambda lib, **kw: pc.make_pursuit(lib, 40, 0, **kw),                    # no minimap, 2x2 bodies, walls
    "gather": lambda lib, **kw: pc.make_gather(lib, 40, 0, **kw),
    "turn": lambda lib, **kw: pc.make_turn(lib, **kw),                                 # per-heading view LUTs
    "food": lambda lib, **kw: pc.make_food(lib, **kw),                                 # food channel
    "sector": lambda lib, **kw: pc.make_sector(lib, **kw),
    "multi4": lambda lib, **kw: pc.make_multi4(lib, **kw),                             # 4 groups: 13 channels
}


@pytest.mark.parametrize("threads", [1, 3, 7])
@pytest.mark.parametrize("game", sorted(MAKERS))
def test_wire_expansion_matches_the_reference(emu, game, threads):
    from magent_b200.c_lib import load_library
    L = load_library(emu)
    L.magent_b200_set_host_threads(threads)
    try:
        want = pc.run_trace(MAKERS[game](checker()), 12, 5, keep_obs=True)
        got = pc.run_trace(MAKERS[game](emu, _host_path="wire"), 12, 5, keep_obs=True)
        pc.compare_traces(want, got, "%s wire x%d threads" % (game, threads))
        dense = pc.run_trace(MAKERS[game](emu, _host_path="dense"), 12, 5, keep_obs=True)
        pc.compare_traces(dense, got, "%s wire vs dense" % game)
    finally:
        L.magent_b200_set_host_threads(0)


def test_many_chunks_many_arenas(emu):
    """more observers than one wire chunk (1024) per thread, arena changes inside chunks, several waves"""
    import magent_b200 as magent
    from magent_b200.c_lib import load_library
    L = load_library(emu)
    L.magent_b200_set_host_threads(4)
    try:
        envs = []
        for path in ("wire", "dense"):
            env = magent.GridWorld("battle", map_size=30, _lib=emu, _num_arenas=37, _host_path=path)
            env.set_seed(11)
            env.reset()
            for h in env.get_handles():
                env.add_agents(h, method="random", n=90)
            envs.append(env)
        rs = np.random.RandomState(3)
        for t in range(6):
            acts = None
            for env in envs:
                hs = env.get_handles()
                obs = [tuple(x.copy() for x in env.get_observation(h)) for h in hs]
                if acts is None:
                    acts = [rs.randint(0, 21, size=env.get_num(h)).astype(np.int32) for h in hs]
                    first = obs
                else:
                    for (v0, f0), (v1, f1) in zip(first, obs):
                        np.testing.assert_array_equal(v0.view(np.uint32), v1.view(np.uint32))
                        np.testing.assert_array_equal(f0.view(np.uint32), f1.view(np.uint32))
                for h, a in zip(hs, acts):
                    env.set_action(h, a)
                env.step()
                env.clear_dead()
            assert first[0][0].shape[0] > 3 * 1024
    finally:
        L.magent_b200_set_host_threads(0)


@pytest.mark.parametrize("shift", [4, 20, 36, 60])
def test_caller_buffers_of_any_alignment(emu, shift):
    """the ABI takes any float*: records are streamed in 64-byte lines, the ragged ends with ordinary stores; nothing
    outside [buffer, buffer + n * record) may be touched"""
    from magent_b200.c_lib import load_library
    L = load_library(emu)
    L.magent_b200_set_host_threads(3)
    try:
        env = pc.make_battle(emu, 40, 150, 0, _host_path="wire")
        ref = pc.make_battle(emu, 40, 150, 0, _host_path="dense")
        h = env.get_handles()[0]
        g = env._hv(h)
        n = env.get_num(h)
        vs, fs = env.get_view_space(h), env.get_feature_space(h)
        nv, nf = n * int(np.prod(vs)), n * fs[0]
        raw_v = np.full(nv * 4 + 256, 0xAB, dtype=np.uint8)
        raw_f = np.full(nf * 4 + 256, 0xCD, dtype=np.uint8)
        base_v = (-raw_v.ctypes.data) % 64 + shift
        base_f = (-raw_f.ctypes.data) % 64 + shift
        ptrs = (ctypes.c_void_p * 2)(raw_v.ctypes.data + base_v, raw_f.ctypes.data + base_f)
        L.env_get_observation(env.game, g, ptrs)
        v, f = ref.get_observation(ref.get_handles()[0])
        np.testing.assert_array_equal(raw_v[base_v:base_v + nv * 4].view(np.uint32), v.reshape(-1).view(np.uint32))
        np.testing.assert_array_equal(raw_f[base_f:base_f + nf * 4].view(np.uint32), f.reshape(-1).view(np.uint32))
        assert (raw_v[:base_v] == 0xAB).all() and (raw_v[base_v + nv * 4:] == 0xAB).all()
        assert (raw_f[:base_f] == 0xCD).all() and (raw_f[base_f + nf * 4:] == 0xCD).all()
    finally:
        L.magent_b200_set_host_threads(0)


def test_numa_split_buffers_with_pretend_nodes(emu):
    """MAGENT_B200_NUMA_FAKE=2 (a subprocess: the topology is read once): the wrapper's big receive buffers come from
    numa_split_alloc, chunks are dealt per part, threads help the other part when theirs is done -- same bytes"""
    import subprocess
    import sys
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import magent_b200 as magent
from magent_b200.c_lib import load_library
L = load_library(%r)
assert L.magent_b200_numa_nodes() == 2
L.magent_b200_set_host_threads(5)
p = L.magent_b200_host_alloc(80 << 20)
assert p
L.magent_b200_host_free(p)
envs = []
for path in ("wire", "dense"):
    env = magent.GridWorld("battle", map_size=40, _lib=L.path, _num_arenas=100, _host_path=path)
    env.set_seed(2); env.reset()
    for h in env.get_handles(): env.add_agents(h, method="random", n=150)
    envs.append(env)
rs = np.random.RandomState(1)
for t in range(3):
    acts = [rs.randint(0, 21, size=envs[0].get_num(h)).astype(np.int32) for h in envs[0].get_handles()]
    obs = []
    for env in envs:
        hs = env.get_handles()
        obs.append([tuple(x.copy() for x in env.get_observation(h)) for h in hs])
        for h, a in zip(hs, acts): env.set_action(h, a)
        env.step(); env.clear_dead()
    for (v0, f0), (v1, f1) in zip(obs[0], obs[1]):
        assert v0.nbytes >= (64 << 20) and np.array_equal(v0.view(np.uint32), v1.view(np.uint32)) and np.array_equal(f0.view(np.uint32), f1.view(np.uint32))
# somebody else's buffers (plain numpy memory): chunks dealt to the nodes round-robin
import ctypes
wire, dense = envs
h = wire.get_handles()[0]; n = wire.get_num(h)
v = np.empty((n,) + wire.get_view_space(h), np.float32); f = np.empty((n,) + wire.get_feature_space(h), np.float32)
L.env_get_observation(wire.game, wire._hv(h), (ctypes.c_void_p * 2)(v.ctypes.data, f.ctypes.data))
dv, df = dense.get_observation(dense.get_handles()[0])
assert np.array_equal(v.view(np.uint32), dv.view(np.uint32)) and np.array_equal(f.view(np.uint32), df.view(np.uint32))
print("OK", obs[0][0][0].nbytes)
''' % (pc.REPO, os.path.join(pc.REPO, "tests"), emu)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, MAGENT_B200_NUMA_FAKE="2", OMP_NUM_THREADS="1"))
    assert r.returncode == 0 and "OK" in r.stdout, (r.stdout[-500:], r.stderr[-2000:])


@pytest.mark.parametrize("isa", ["avx2", "sse2"])
def test_streaming_stores_without_avx512(emu, isa):
    """MAGENT_B200_HOST_ISA picks the 32-byte / 16-byte non-temporal store loops (hosts without AVX-512): same bytes"""
    import subprocess
    import sys
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import parity_common as pc
from magent_b200.c_lib import load_library
L = load_library(%r)
L.magent_b200_set_host_threads(3)
w = pc.run_trace(pc.make_battle(L.path, 40, 150, 0, _host_path="wire"), 6, 5, keep_obs=True)
d = pc.run_trace(pc.make_battle(L.path, 40, 150, 0, _host_path="dense"), 6, 5, keep_obs=True)
pc.compare_traces(d, w, "isa")
w = pc.run_trace(pc.make_pursuit(L.path, 40, 0, _host_path="wire"), 6, 5, keep_obs=True)
d = pc.run_trace(pc.make_pursuit(L.path, 40, 0, _host_path="dense"), 6, 5, keep_obs=True)
pc.compare_traces(d, w, "isa")
print("OK")
''' % (pc.REPO, os.path.join(pc.REPO, "tests"), emu)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, MAGENT_B200_HOST_ISA=isa, OMP_NUM_THREADS="1"))
    assert r.returncode == 0 and "OK" in r.stdout, (r.stdout[-500:], r.stderr[-2000:])
