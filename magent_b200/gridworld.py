"""Host-side mirror of the reference ``magent.gridworld`` surface for the B200 engine.

Same public names, argument meaning and return types as the reference wrapper
(reference: python/magent/gridworld.py:14-800) so that ``examples/train_{battle,pursuit,gather}.py``
run unchanged, but written against the typed ABI table in :mod:`magent_b200.c_lib`:

* the library is per-instance (``_lib=`` keyword or ``$MAGENT_B200_LIB``), which lets the parity
  tests drive the compiled reference engine and the CUDA engine with one and the same host code;
* observation / reward / info buffers are cached per group and, when the library is the B200
  engine, live in page-locked host memory handed out by ``magent_b200_host_alloc`` so the
  device->host copy behind ``env_get_observation`` runs at PCIe speed;
* B200 extensions: ``_num_arenas`` (independent arenas batched behind one handle; every group is
  presented as the concatenation over arenas), device-pointer observation (``get_observation_torch``),
  on-device random actions for throughput runs, event counters.

Reserved constructor keywords start with an underscore so they can never collide with the keyword
arguments forwarded to a built-in config's ``get_config``.
"""
from __future__ import annotations

import ctypes
import importlib
import os

import numpy as np

from .c_lib import load_library, as_float_c_array, as_int32_c_array
from .environment import Environment


def _cint(v):
    return ctypes.byref(ctypes.c_int(int(v)))


class _HostBlock:
    """A growable host buffer; page-locked when the engine offers an allocator."""

    def __init__(self, lib):
        self._lib = lib
        self._ptr = None
        self._raw = None
        self._cap = 0

    def view(self, shape, dtype):
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        if nbytes > self._cap or self._raw is None:
            self._grow(max(nbytes, 64))
        return self._raw[:nbytes].view(dtype).reshape(shape)

    def _grow(self, nbytes):
        self.release()
        # populations only shrink between resets: little slack for the big observation buffers (which a multi-socket host
        # splits evenly over its NUMA nodes), room to grow for the small ones
        cap = int(nbytes * (1.02 if nbytes >= (64 << 20) else 1.25)) + 64
        if self._lib.is_b200:
            ptr = self._lib.magent_b200_host_alloc(cap)
            if ptr:
                self._ptr = ptr
                self._raw = np.ctypeslib.as_array(
                    ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint8)), shape=(cap,))
                self._cap = cap
                return
        self._raw = np.empty((cap,), dtype=np.uint8)
        self._cap = cap

    def release(self):
        if self._ptr is not None:
            self._lib.magent_b200_host_free(self._ptr)
        self._ptr = None
        self._raw = None
        self._cap = 0


class GridWorld(Environment):
    """The grid-world environment (reference: python/magent/gridworld.py:14-566)."""

    OBS_INDEX_VIEW = 0
    OBS_INDEX_HP = 1

    # global keys accepted by env_config_game and the C type behind the void*
    # (reference: python/magent/gridworld.py:46-52, src/gridworld/GridWorld.cc:120-149)
    _CONFIG_TYPES = {
        "map_width": int, "map_height": int,
        "food_mode": bool, "turn_mode": bool, "minimap_mode": bool,
        "revive_mode": bool, "goal_mode": bool,
        "embedding_size": int,
        "render_dir": str,
    }

    def __init__(self, config, **kwargs):
        Environment.__init__(self)
        self._lib = load_library(kwargs.pop("_lib", None))
        self.num_arenas = int(kwargs.pop("_num_arenas", 1))
        device = kwargs.pop("_device", None)
        host_path = kwargs.pop("_host_path", None)      # None: engine default (wire); "dense": dense records over PCIe
        self.game = None

        if isinstance(config, str):
            try:
                module = importlib.import_module("magent.builtin.config." + config)
            except ImportError:
                module = importlib.import_module(__package__ + ".builtin.config." + config)
            try:
                config = getattr(module, "get_config")(**kwargs)
            except AttributeError:
                raise BaseException('unknown built-in game "' + config + '"')
        self.config = config

        L = self._lib
        game = ctypes.c_void_p()
        rc = L.env_new_game(ctypes.byref(game), b"GridWorld")
        if rc != 0 or not game.value:
            why = L.magent_b200_last_error().decode() if L.is_b200 else "env_new_game failed"
            raise RuntimeError("cannot create the GridWorld engine: " + why)
        self.game = game

        # extension keys go first: they size the arena batch before anything else is configured
        if self.num_arenas != 1 or device is not None:
            if not L.is_b200:
                raise ValueError("_num_arenas/_device need the B200 engine library")
            if device is not None:
                L.env_config_game(game, b"device_id", _cint(device))
            L.env_config_game(game, b"num_arenas", _cint(self.num_arenas))
        if host_path is not None:
            if not L.is_b200:
                raise ValueError("_host_path needs the B200 engine library")
            # "dense": dense records over PCIe; "wire": wire records + host expansion whatever the size; default: wire for
            # observations of 16 MB and more, dense below
            L.env_config_game(game, b"host_path", _cint({"dense": 0, "wire": 2, "auto": 1}[host_path]))

        for key, value in config.config_dict.items():
            kind = self._CONFIG_TYPES[key]
            name = key.encode("ascii")
            if kind is int:
                L.env_config_game(game, name, _cint(value))
            elif kind is bool:
                L.env_config_game(game, name, ctypes.byref(ctypes.c_bool(bool(value))))
            elif kind is str:
                text = value if isinstance(value, bytes) else str(value).encode("ascii")
                L.env_config_game(game, name, ctypes.cast(ctypes.c_char_p(text), ctypes.c_void_p))

        for name, attrs in config.agent_type_dict.items():
            flat = self._flatten_type(attrs)
            n = len(flat)
            keys = (ctypes.c_char_p * n)(*[k.encode("ascii") for k in flat])
            vals = (ctypes.c_float * n)(*[float(v) for v in flat.values()])
            L.gridworld_register_agent_type(game, name.encode("ascii"), n, keys, vals)

        self._serialize_event_exp(config)

        self.group_handles = []
        for type_name in config.groups:
            h = ctypes.c_int32()
            L.gridworld_new_group(game, type_name.encode("ascii"),
                                  ctypes.cast(ctypes.byref(h), ctypes.POINTER(ctypes.c_int)))
            self.group_handles.append(h)

        self._init_obs_buf()

        self.view_space, self.feature_space, self.action_space = {}, {}, {}
        tmp = np.empty((3,), dtype=np.int32)
        for h in self.group_handles:
            L.env_get_info(game, h.value, b"view_space", tmp.ctypes.data)
            self.view_space[h.value] = (int(tmp[0]), int(tmp[1]), int(tmp[2]))
            L.env_get_info(game, h.value, b"feature_space", tmp.ctypes.data)
            self.feature_space[h.value] = (int(tmp[0]),)
            L.env_get_info(game, h.value, b"action_space", tmp.ctypes.data)
            self.action_space[h.value] = (int(tmp[0]),)

    @staticmethod
    def _flatten_type(attrs):
        """Expand Range objects into radius/angle scalars
        (reference: python/magent/gridworld.py:69-80)."""
        flat = {}
        for key, val in attrs.items():
            if key == "view_range":
                flat["view_radius"], flat["view_angle"] = val.radius, val.angle
            elif key == "attack_range":
                flat["attack_radius"], flat["attack_angle"] = val.radius, val.angle
            else:
                flat[key] = val
        return flat

    @staticmethod
    def _hv(handle):
        return handle.value if hasattr(handle, "value") else int(handle)

    # ------------------------------------------------------------------ episode setup
    def reset(self):
        """reset environment (reference: gridworld.py:117-119)"""
        self._lib.env_reset(self.game)

    def add_walls(self, method, **kwargs):
        """add walls; ``method`` in random/custom/fill (reference: gridworld.py:121-141)"""
        kwargs["dir"] = 0
        self.add_agents(-1, method, **kwargs)

    def new_group(self, name):
        """register a new group (reference: gridworld.py:144-148)"""
        h = ctypes.c_int32()
        self._lib.gridworld_new_group(self.game, name.encode("ascii"),
                                      ctypes.cast(ctypes.byref(h), ctypes.POINTER(ctypes.c_int)))
        return h

    def add_agents(self, handle, method, **kwargs):
        """add agents (or walls when handle == -1) (reference: gridworld.py:150-200)

        method="random": kwargs["n"]; "custom": kwargs["pos"] = [(x, y[, dir]), ...];
        "fill": kwargs["pos"]=(x, y), kwargs["size"]=(w, h)[, kwargs["dir"]].
        """
        L, g = self._lib, self._hv(handle)
        if method == "random":
            L.gridworld_add_agents(self.game, g, int(kwargs["n"]), b"random", None, None, None)
        elif method == "custom":
            pos = np.asarray(kwargs["pos"], dtype=np.int32)
            if pos.size == 0:
                return
            n = pos.shape[0]
            xs = np.ascontiguousarray(pos[:, 0])
            ys = np.ascontiguousarray(pos[:, 1])
            dirs = (np.ascontiguousarray(pos[:, 2]) if pos.shape[1] == 3
                    else np.zeros((n,), dtype=np.int32))
            L.gridworld_add_agents(self.game, g, n, b"custom",
                                   xs.ctypes.data, ys.ctypes.data, dirs.ctypes.data)
        elif method == "fill":
            x, y = kwargs["pos"][0], kwargs["pos"][1]
            w, h = kwargs["size"][0], kwargs["size"][1]
            d = kwargs.get("dir", 0)
            bind = np.array([x, y, w, h, d], dtype=np.int32)
            L.gridworld_add_agents(self.game, g, 0, b"fill", bind.ctypes.data, None, None)
        elif method == "maze":
            raise NotImplementedError("maze placement is not implemented by the engine "
                                      "(reference: GridWorld.cc:215-217 raises FATAL as well)")
        else:
            print("Unknown type of position")
            exit(-1)

    # ------------------------------------------------------------------ run
    def _init_obs_buf(self):
        self.obs_bufs = [{}, {}, {}, {}]       # view, feature (reference); f16 view, f16 feature (extension)
        self._blocks = {}

    def _get_obs_buf(self, group, key, shape, dtype):
        """cached receive buffer, resized in place like the reference (gridworld.py:203-213)"""
        block = self._blocks.get((group, key))
        if block is None:
            block = self._blocks[(group, key)] = _HostBlock(self._lib)
        buf = block.view(shape, dtype)
        self.obs_bufs[key][group] = buf
        return buf

    def get_observation(self, handle):
        """(views, features) of a whole group (reference: gridworld.py:221-248)

        views: float32 (n, view_h, view_w, n_channel); features: float32 (n, feature_size).
        The arrays are the engine's cached receive buffers: valid until the next call.
        """
        g = self._hv(handle)
        n = self.get_num(handle)
        view = self._get_obs_buf(g, self.OBS_INDEX_VIEW, (n,) + self.view_space[g], np.float32)
        feat = self._get_obs_buf(g, self.OBS_INDEX_HP, (n,) + self.feature_space[g], np.float32)
        bufs = (ctypes.c_void_p * 2)(view.ctypes.data, feat.ctypes.data)
        self._lib.env_get_observation(self.game, g, bufs)
        return view, feat

    def set_action(self, handle, actions):
        """actions: int32 numpy array of length get_num(handle) (reference: gridworld.py:250-261)"""
        assert isinstance(actions, np.ndarray)
        assert actions.dtype == np.int32
        actions = np.ascontiguousarray(actions)
        self._lib.env_set_action(self.game, self._hv(handle), actions.ctypes.data)

    def step(self):
        """simulate one step; returns done (reference: gridworld.py:263-273)"""
        done = ctypes.c_int(0)
        self._lib.env_step(self.game, ctypes.byref(done))
        return bool(done.value)

    def get_reward(self, handle):
        """float32 rewards of a group (reference: gridworld.py:275-287)"""
        n = self.get_num(handle)
        buf = np.empty((n,), dtype=np.float32)
        self._lib.env_get_reward(self.game, self._hv(handle), buf.ctypes.data)
        return buf

    def clear_dead(self):
        """remove dead agents; call after step() (reference: gridworld.py:289-293)"""
        self._lib.gridworld_clear_dead(self.game)

    # ------------------------------------------------------------------ info
    def get_handles(self):
        return self.group_handles

    def _info(self, handle, name, buf):
        self._lib.env_get_info(self.game, self._hv(handle), name, buf.ctypes.data)
        return buf

    def get_num(self, handle):
        num = ctypes.c_int(0)
        self._lib.env_get_info(self.game, self._hv(handle), b"num",
                               ctypes.cast(ctypes.byref(num), ctypes.c_void_p))
        return num.value

    def get_action_space(self, handle):
        return self.action_space[self._hv(handle)]

    def get_view_space(self, handle):
        return self.view_space[self._hv(handle)]

    def get_feature_space(self, handle):
        return self.feature_space[self._hv(handle)]

    def get_agent_id(self, handle):
        """int32 ids (reference: gridworld.py:333-345)"""
        return self._info(handle, b"id", np.empty((self.get_num(handle),), dtype=np.int32))

    def get_alive(self, handle):
        """bool alive flags (reference: gridworld.py:347-359)"""
        return self._info(handle, b"alive", np.empty((self.get_num(handle),), dtype=np.bool_))

    def get_pos(self, handle):
        """int32 (n, 2) positions (reference: gridworld.py:361-373)"""
        return self._info(handle, b"pos", np.empty((self.get_num(handle), 2), dtype=np.int32))

    def get_mean_info(self, handle):
        """deprecated (reference: gridworld.py:375-380)"""
        n_act = self.action_space[self._hv(handle)][0]
        return self._info(handle, b"mean_info", np.empty(2 + n_act, dtype=np.float32))

    def get_view2attack(self, handle):
        """(attack_base, view-shaped int32 map of attack action numbers, -1 elsewhere)
        (reference: gridworld.py:382-399)"""
        size = self.get_view_space(handle)[0:2]
        buf = self._info(handle, b"view2attack", np.empty(size, dtype=np.int32))
        base = ctypes.c_int(0)
        self._lib.env_get_info(self.game, self._hv(handle), b"attack_base",
                               ctypes.cast(ctypes.byref(base), ctypes.c_void_p))
        return base.value, buf

    def get_global_minimap(self, height, width):
        """(height, width, n_group) float32 density maps (reference: gridworld.py:401-420)"""
        buf = np.empty((height, width, len(self.group_handles)), dtype=np.float32)
        buf[0, 0, 0] = height
        buf[0, 0, 1] = width
        self._lib.env_get_info(self.game, -1, b"global_minimap", buf.ctypes.data)
        return buf

    def set_seed(self, seed):
        """seed the engine RNG (reference: gridworld.py:422-424)"""
        self._lib.env_config_game(self.game, b"seed", _cint(seed))

    # ------------------------------------------------------------------ render
    def set_render_dir(self, name):
        if not os.path.exists(name):
            os.mkdir(name)
        text = name.encode("ascii")
        self._lib.env_config_game(self.game, b"render_dir",
                                  ctypes.cast(ctypes.c_char_p(text), ctypes.c_void_p))

    def render(self):
        self._lib.env_render(self.game)

    def _get_groups_info(self):
        buf = np.empty((len(self.group_handles), 5), dtype=np.int32)
        self._lib.env_get_info(self.game, -1, b"groups_info", buf.ctypes.data)
        return buf

    def _get_walls_info(self):
        buf = np.empty((100 * 100, 2), dtype=np.int32)
        self._lib.env_get_info(self.game, -1, b"walls_info", buf.ctypes.data)
        n = buf[0, 0]
        return buf[1:1 + n]

    def _get_render_info(self, x_range, y_range):
        n = sum(self.get_num(h) for h in self.group_handles)
        buf = np.empty((n + 1, 4), dtype=np.int32)
        buf[0] = x_range[0], y_range[0], x_range[1], y_range[1]
        self._lib.env_get_info(self.game, -1, b"render_window_info", buf.ctypes.data)
        agent_ct, attack_ct = buf[0][0], buf[0][1]
        agent_info = {row[0]: [row[1], row[2], row[3]] for row in buf[1:1 + agent_ct]}
        events = np.empty((attack_ct, 3), dtype=np.int32)
        self._lib.env_get_info(self.game, -1, b"attack_event", events.ctypes.data)
        return agent_info, events

    def __del__(self):
        game, self.game = getattr(self, "game", None), None
        if game is not None:
            try:
                for block in getattr(self, "_blocks", {}).values():
                    block.release()
                self._lib.env_delete_game(game)
            except Exception:       # interpreter shutdown
                pass

    # ------------------------------------------------------------------ special rule
    def set_goal(self, handle, method, *args, **kwargs):
        """deprecated (reference: gridworld.py:485-490)"""
        if method == "random":
            self._lib.gridworld_set_goal(self.game, self._hv(handle), b"random", None)
        else:
            raise NotImplementedError

    # ------------------------------------------------------------------ B200 extensions
    def get_arena_nums(self, handle):
        """per-arena agent counts of a group (extension; int32 [num_arenas])"""
        buf = np.empty((self.num_arenas,), dtype=np.int32)
        return self._info(handle, b"arena_num", buf)

    def get_arena_done(self):
        """per-arena done flags of the last step (extension; int32 [num_arenas])"""
        buf = np.empty((self.num_arenas,), dtype=np.int32)
        self._lib.env_get_info(self.game, -1, b"arena_done", buf.ctypes.data)
        return buf

    def select_arena(self, arena):
        """route subsequent setup calls (add_agents/add_walls/set_seed) to one arena; -1 = all"""
        self._lib.magent_b200_select_arena(self.game, int(arena))

    def get_observation_torch(self, handle, out=None, dtype=None):
        """observation written straight into CUDA tensors through the same ABI call
        (device pointers are detected by the engine; no PCIe traffic).

        dtype: torch.float32 (default, the reference layout) or torch.float16 (compact hand-off: every element
        is the float32 value rounded to nearest-even; half the bytes)."""
        import torch
        g = self._hv(handle)
        n = self.get_num(handle)
        if out is None:
            dtype = torch.float32 if dtype is None else dtype
            view = torch.empty((n,) + self.view_space[g], dtype=dtype, device="cuda")
            feat = torch.empty((n,) + self.feature_space[g], dtype=dtype, device="cuda")
        else:
            view, feat = out
            dtype = view.dtype
            assert feat.dtype == dtype and view.is_contiguous() and feat.is_contiguous()
        bufs = (ctypes.c_void_p * 2)(view.data_ptr(), feat.data_ptr())
        if dtype == torch.float16:
            self._lib.magent_b200_get_observation_f16(self.game, g, bufs)
        elif dtype == torch.float32:
            self._lib.env_get_observation(self.game, g, bufs)
        else:
            raise ValueError("observation dtype must be torch.float32 or torch.float16")
        return view, feat

    def get_observation_f16(self, handle):
        """(views, features) as float16 numpy arrays (compact hand-off; see get_observation for the layout)"""
        g = self._hv(handle)
        n = self.get_num(handle)
        view = self._get_obs_buf(g, 2, (n,) + self.view_space[g], np.float16)
        feat = self._get_obs_buf(g, 3, (n,) + self.feature_space[g], np.float16)
        bufs = (ctypes.c_void_p * 2)(view.ctypes.data, feat.ctypes.data)
        self._lib.magent_b200_get_observation_f16(self.game, g, bufs)
        return view, feat

    def set_random_actions(self, handle, seed):
        """uniform random actions generated on the device (throughput runs only)"""
        self._lib.magent_b200_random_actions(self.game, self._hv(handle), None, int(seed))

    def get_counters(self):
        """int64 event counters accumulated on the device since construction"""
        buf = (ctypes.c_longlong * 16)()
        n = self._lib.magent_b200_get_counters(self.game, buf, 16)
        return [int(buf[i]) for i in range(n)]

    def sync(self):
        self._lib.magent_b200_sync(self.game)

    def get_io_stats(self):
        """step-loop traffic so far: dict(d2h=PCIe device->host bytes, h2d=..., host_written=bytes the engine's host
        threads wrote into caller buffers)"""
        buf = (ctypes.c_longlong * 6)()
        self._lib.magent_b200_get_io_stats(self.game, buf, 6)
        return {"d2h": int(buf[0]), "h2d": int(buf[1]), "host_written": int(buf[2]),
                "us_wire": int(buf[3]), "us_expand": int(buf[4]), "us_feature": int(buf[5])}

    def capture_graph(self, fn):
        """record the step-loop calls fn() makes (CUDA device pointers only, an even number of clear_dead) into a CUDA
        graph; returns its id for launch_graph"""
        self._lib.magent_b200_graph_begin(self.game)
        fn()
        return int(self._lib.magent_b200_graph_end(self.game))

    def launch_graph(self, graph_id, times=1):
        self._lib.magent_b200_graph_launch(self.game, int(graph_id), int(times))

    def step_device_done(self, done_ptr):
        """env_step with a CUDA device int for `done`: returns without waiting for the step (device-resident loops)"""
        self._lib.env_step(self.game, ctypes.cast(ctypes.c_void_p(int(done_ptr)), ctypes.POINTER(ctypes.c_int)))

    def set_profiling(self, on):
        """bracket every obs-render kernel launch with CUDA events (adds a sync per launch)"""
        self._lib.magent_b200_set_profiling(self.game, 1 if on else 0)

    def get_profile(self):
        """(total obs-render kernel milliseconds, launches) since profiling was enabled"""
        ms = ctypes.c_double(0.0)
        n = ctypes.c_longlong(0)
        self._lib.magent_b200_get_profile(self.game, ctypes.byref(ms), ctypes.byref(n))
        return ms.value, n.value

    def launch_count(self):
        """number of kernels this library has launched in this process"""
        return int(self._lib.magent_b200_launch_count())

    # ------------------------------------------------------------------ reward DSL -> ABI
    def _serialize_event_exp(self, config):
        """Number symbols and event nodes and send them to the engine.

        The numbering must match the reference exactly (gridworld.py:493-565): symbol numbers fix
        the binding order of the rule evaluation (RewardEngine.cc:156-189 iterates a std::set of
        symbol pointers), so: per rule, receivers first, then a pre-order walk of the trigger.
        """
        L, game = self._lib, self.game
        sym_no, node_no = {}, {}

        def number_symbols(node):
            for item in node.inputs:
                if isinstance(item, EventNode):
                    number_symbols(item)
                elif isinstance(item, AgentSymbol) and item not in sym_no:
                    sym_no[item] = len(sym_no)

        def number_nodes(node):
            if node not in node_no:
                node_no[node] = len(node_no)
            for item in node.inputs:
                if isinstance(item, EventNode):
                    number_nodes(item)

        for on, receivers, _values, _terminal in config.reward_rules:
            for sym in receivers:
                if sym not in sym_no:
                    sym_no[sym] = len(sym_no)
            number_symbols(on)
        for rule in config.reward_rules:
            number_nodes(rule[0])
        config.symbol_ct, config.node_ct = len(sym_no), len(node_no)

        for sym, no in sym_no.items():
            L.gridworld_define_agent_symbol(game, no, sym.group, sym.index)

        for node, no in node_no.items():
            raw = np.zeros((len(node.inputs),), dtype=np.int32)
            for i, item in enumerate(node.inputs):
                if isinstance(item, EventNode):
                    raw[i] = node_no[item]
                elif isinstance(item, AgentSymbol):
                    raw[i] = sym_no[item]
                else:
                    raw[i] = item
            L.gridworld_define_event_node(game, no, node.op, as_int32_c_array(raw), len(raw))

        for on, receivers, values, terminal in config.reward_rules:
            recv = np.array([sym_no[s] for s in receivers], dtype=np.int32)
            auto = len(values) == 1 and values[0] == "auto"
            val = (np.zeros((len(recv),), dtype=np.float32) if auto
                   else np.array(values, dtype=np.float32))
            L.gridworld_add_reward_rule(game, node_no[on], as_int32_c_array(recv),
                                        as_float_c_array(val), len(recv), bool(terminal), bool(auto))


# ---------------------------------------------------------------------- reward description DSL
class EventNode:
    """AST node of an event expression (reference: gridworld.py:571-650).

    Leaves are built by calling the module-level ``Event`` object:
    ``Event(subject, 'attack', target)``; nodes combine with ``&``, ``|`` and ``~``.
    Operator numbers follow src/gridworld/grid_def.h:17-23.
    """
    OP_AND, OP_OR, OP_NOT = 0, 1, 2
    OP_KILL, OP_AT, OP_IN, OP_COLLIDE, OP_ATTACK, OP_DIE = 3, 4, 5, 6, 7, 8
    OP_IN_A_LINE, OP_ALIGN = 9, 10

    _BINARY_AGENT = {"kill": OP_KILL, "attack": OP_ATTACK, "collide": OP_COLLIDE}
    _UNARY_AGENT = {"die": OP_DIE, "in_a_line": OP_IN_A_LINE, "align": OP_ALIGN}

    def __init__(self, op=None, inputs=(), predicate=None):
        self.op = op
        self.predicate = predicate
        self.inputs = list(inputs)

    def __call__(self, subject, predicate, *args):
        if predicate in self._BINARY_AGENT:
            return EventNode(self._BINARY_AGENT[predicate], [subject, args[0]], predicate)
        if predicate in self._UNARY_AGENT:
            return EventNode(self._UNARY_AGENT[predicate], [subject], predicate)
        if predicate == "at":
            x, y = args[0][0], args[0][1]
            return EventNode(self.OP_AT, [subject, x, y], predicate)
        if predicate == "in":
            (ax, ay), (bx, by) = args[0][0], args[0][1]
            return EventNode(self.OP_IN,
                             [subject, min(ax, bx), min(ay, by), max(ax, bx), max(ay, by)], predicate)
        raise Exception("invalid predicate of event " + predicate)

    def __and__(self, other):
        return EventNode(self.OP_AND, [self, other])

    def __or__(self, other):
        return EventNode(self.OP_OR, [self, other])

    def __invert__(self):
        return EventNode(self.OP_NOT, [self])


Event = EventNode()


class AgentSymbol:
    """symbol standing for some agent(s) of a group (reference: gridworld.py:654-675)

    index: 'any' (-1), 'all' (-2) or a deterministic int index into the group.
    """

    def __init__(self, group, index):
        self.group = group if group is not None else -1
        if index == "any":
            self.index = -1
        elif index == "all":
            self.index = -2
        else:
            assert isinstance(index, (int, np.integer)), "index must be a deterministic int"
            self.index = int(index)

    def __str__(self):
        return "agent(%d,%d)" % (self.group, self.index)


class Config:
    """game configuration (reference: gridworld.py:678-766)"""

    def __init__(self):
        self.config_dict = {}
        self.agent_type_dict = {}
        self.groups = []
        self.reward_rules = []

    def set(self, args):
        """global key/value pairs, e.g. {"map_width": 100, "minimap_mode": True}"""
        self.config_dict.update(args)

    def register_agent_type(self, name, attr):
        """register an agent type; attr keys as in src/gridworld/AgentType.cc:52-78
        (width, length, speed, hp, view_range, attack_range, damage, step_recover, kill_supply,
        step_reward, kill_reward, dead_penalty, attack_penalty, attack_in_group, ...)"""
        if name in self.agent_type_dict:
            raise Exception("type name %s already exists" % name)
        self.agent_type_dict[name] = attr
        return name

    def add_group(self, agent_type):
        """returns the handle (an int) of the new group"""
        self.groups.append(agent_type)
        return len(self.groups) - 1

    def add_reward_rule(self, on, receiver, value, terminal=False):
        """when event ``on`` holds, give ``value`` to ``receiver`` (lists allowed)"""
        if not isinstance(receiver, (tuple, list)):
            assert not isinstance(value, (tuple, list))
            receiver, value = [receiver], [value]
        if len(receiver) != len(value):
            raise Exception("the length of receiver and value should be equal")
        self.reward_rules.append([on, list(receiver), list(value), terminal])


class CircleRange:
    """circular view/attack range (reference: gridworld.py:769-781)"""

    def __init__(self, radius):
        self.radius = radius
        self.angle = 360

    def __str__(self):
        return "circle(%g)" % self.radius


class SectorRange:
    """sector view/attack range, angle < 180 (reference: gridworld.py:784-800)"""

    def __init__(self, radius, angle):
        self.radius = radius
        self.angle = angle
        if self.angle >= 180:
            raise Exception("the angle of a sector should be smaller than 180 degree")

    def __str__(self):
        return "sector(%g, %g)" % (self.radius, self.angle)
