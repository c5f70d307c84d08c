"""`magent.GridWorld` -- the Python face of the engine C ABI.

Mirrors the public surface of the reference binding
(examples/battle_model/python/magent/gridworld.py:15-633 for GridWorld, :731-1018 for the config DSL)
so that senario_battle.play()/battle() and the algo classes run unchanged:

    GridWorld('battle', map_size=40)      reset()             add_agents(handle, method=, pos=/n=)
    get_handles()  get_num()              get_observation()   set_action()   step() -> bool
    get_reward()   get_alive()            clear_dead()        get_agent_id() get_pos()
    get_action_space/view_space/feature_space               get_mean_info() get_view2attack()
    get_global_minimap()  set_seed()      set_render_dir()    render()

Buffers are caller-owned numpy arrays, sized from get_num() at call time, exactly as in the
reference (gridworld.py:282-342,377-389); the engine fills them.  The engine itself is whichever
library `lib` points to: by default the CUDA build of this package (c_lib._LIB); tests pass the
reference build to run both side by side.
"""
import ctypes
import importlib
import os

import numpy as np

from . import c_lib
from .c_lib import as_float_c_array, as_int32_c_array
from .environment import Environment

_GAME_KEYS = {
    # key -> ctypes scalar used to pass the value (reference gridworld.py:58-88)
    "map_width": ctypes.c_int, "map_height": ctypes.c_int, "embedding_size": ctypes.c_int,
    "food_mode": ctypes.c_bool, "turn_mode": ctypes.c_bool, "minimap_mode": ctypes.c_bool,
    "revive_mode": ctypes.c_bool, "goal_mode": ctypes.c_bool,
    "render_dir": str,
}


class GridWorld(Environment):
    OBS_INDEX_VIEW = 0
    OBS_INDEX_HP = 1

    def __init__(self, config, lib=None, **kwargs):
        self._lib = lib if lib is not None else c_lib._LIB
        if isinstance(config, str):
            module = importlib.import_module("magent.builtin.config." + config)
            if not hasattr(module, "get_config"):
                raise BaseException('unknown built-in game "' + config + '"')
            config = module.get_config(**kwargs)

        self.game = ctypes.c_void_p()
        self._lib.env_new_game(ctypes.byref(self.game), b"GridWorld")

        for key, value in config.config_dict.items():
            kind = _GAME_KEYS[key]
            if kind is str:
                self._lib.env_config_game(self.game, key.encode("ascii"),
                                          ctypes.c_char_p(value.encode("ascii")))
            else:
                self._lib.env_config_game(self.game, key.encode("ascii"), ctypes.byref(kind(value)))

        for name, attrs in config.agent_type_dict.items():
            flat = {}
            for key, value in attrs.items():
                if key in ("view_range", "attack_range"):
                    stem = key.split("_")[0]
                    flat[stem + "_radius"] = value.radius
                    flat[stem + "_angle"] = value.angle
                else:
                    flat[key] = value
            keys = (ctypes.c_char_p * len(flat))(*[k.encode("ascii") for k in flat])
            vals = (ctypes.c_float * len(flat))(*flat.values())
            self._lib.gridworld_register_agent_type(self.game, name.encode("ascii"), len(flat),
                                                    keys, vals)

        self._serialize_event_exp(config)

        self.group_handles = []
        for type_name in config.groups:
            handle = ctypes.c_int32()
            self._lib.gridworld_new_group(self.game, type_name.encode("ascii"), ctypes.byref(handle))
            self.group_handles.append(handle)

        self._init_obs_buf()

        self.view_space, self.feature_space, self.action_space = {}, {}, {}
        buf = np.empty((3,), dtype=np.int32)
        for handle in self.group_handles:
            self._lib.env_get_info(self.game, handle, b"view_space", as_int32_c_array(buf))
            self.view_space[handle.value] = (buf[0], buf[1], buf[2])      # numpy int32 scalars, as the reference returns them
            self._lib.env_get_info(self.game, handle, b"feature_space", as_int32_c_array(buf))
            self.feature_space[handle.value] = (buf[0],)
            self._lib.env_get_info(self.game, handle, b"action_space", as_int32_c_array(buf))
            self.action_space[handle.value] = (buf[0],)

    # ------------------------------------------------------------------ episode set-up
    def reset(self):
        self._lib.env_reset(self.game)

    def add_walls(self, method, **kwargs):
        kwargs["dir"] = 0
        self.add_agents(-1, method, **kwargs)

    def new_group(self, name):
        handle = ctypes.c_int32()
        self._lib.gridworld_new_group(self.game, name.encode("ascii"), ctypes.byref(handle))
        return handle

    def add_agents(self, handle, method, **kwargs):
        """method 'random' (n=), 'custom' (pos=[[x, y(, dir)], ...]) or 'fill' (pos=, size=)."""
        if method == "random":
            self._lib.gridworld_add_agents(self.game, handle, int(kwargs["n"]), b"random", 0, 0, 0)
        elif method == "custom":
            pos = np.asarray(kwargs["pos"], dtype=np.int32)
            if len(pos) == 0:
                return
            n = len(pos)
            xs = np.ascontiguousarray(pos[:, 0])
            ys = np.ascontiguousarray(pos[:, 1])
            dirs = (np.ascontiguousarray(pos[:, 2]) if pos.shape[1] == 3
                    else np.zeros((n,), dtype=np.int32))
            self._lib.gridworld_add_agents(self.game, handle, n, b"custom", as_int32_c_array(xs),
                                           as_int32_c_array(ys), as_int32_c_array(dirs))
        elif method == "fill":
            x, y = kwargs["pos"][0], kwargs["pos"][1]
            w, h = kwargs["size"][0], kwargs["size"][1]
            bind = np.array([x, y, w, h, kwargs.get("dir", 0)], dtype=np.int32)
            self._lib.gridworld_add_agents(self.game, handle, 0, b"fill", as_int32_c_array(bind),
                                           0, 0)
        else:
            raise ValueError("unknown add_agents method: %r" % (method,))

    # ------------------------------------------------------------------ step loop
    def _init_obs_buf(self):
        self.obs_bufs = [{}, {}]

    def _get_obs_buf(self, group, key, shape, dtype):
        cache = self.obs_bufs[key]
        buf = cache.get(group)
        if buf is None:
            buf = cache[group] = np.empty(shape=shape, dtype=dtype)
        elif buf.shape != shape:
            buf.resize(shape, refcheck=False)
        return buf

    def get_observation(self, handle):
        """-> (views float32[n, 13, 13, 7], features float32[n, 34]) for the whole group."""
        no = handle.value
        n = self.get_num(handle)
        view = self._get_obs_buf(no, self.OBS_INDEX_VIEW, (n,) + self.view_space[no], np.float32)
        feat = self._get_obs_buf(no, self.OBS_INDEX_HP, (n,) + self.feature_space[no], np.float32)
        bufs = (ctypes.POINTER(ctypes.c_float) * 2)(as_float_c_array(view), as_float_c_array(feat))
        self._lib.env_get_observation(self.game, handle, bufs)
        return view, feat

    def set_action(self, handle, actions):
        assert isinstance(actions, np.ndarray)
        assert actions.dtype == np.int32
        self._lib.env_set_action(self.game, handle, as_int32_c_array(actions))

    def step(self):
        done = ctypes.c_int32()
        self._lib.env_step(self.game, ctypes.byref(done))
        return bool(done)

    def get_reward(self, handle):
        buf = np.empty((self.get_num(handle),), dtype=np.float32)
        self._lib.env_get_reward(self.game, handle, as_float_c_array(buf))
        return buf

    def clear_dead(self):
        self._lib.gridworld_clear_dead(self.game)

    # ------------------------------------------------------------------ info
    def get_handles(self):
        return self.group_handles

    def get_num(self, handle):
        num = ctypes.c_int32()
        self._lib.env_get_info(self.game, handle, b"num", ctypes.byref(num))
        return num.value

    def get_action_space(self, handle):
        return self.action_space[handle.value]

    def get_view_space(self, handle):
        return self.view_space[handle.value]

    def get_feature_space(self, handle):
        return self.feature_space[handle.value]

    def _info_array(self, handle, key, shape, dtype):
        buf = np.empty(shape, dtype=dtype)
        self._lib.env_get_info(self.game, handle, key, buf.ctypes.data_as(ctypes.c_void_p))
        return buf

    def get_agent_id(self, handle):
        return self._info_array(handle, b"id", (self.get_num(handle),), np.int32)

    def get_alive(self, handle):
        return self._info_array(handle, b"alive", (self.get_num(handle),), np.bool_)

    def get_pos(self, handle):
        return self._info_array(handle, b"pos", (self.get_num(handle), 2), np.int32)

    def get_mean_info(self, handle):
        """[mean_x, mean_y, action histogram / n] of the agents currently in the group."""
        return self._info_array(handle, b"mean_info", (2 + self.action_space[handle.value][0],),
                                np.float32)

    def get_view2attack(self, handle):
        size = self.get_view_space(handle)[0:2]
        buf = self._info_array(handle, b"view2attack", size, np.int32)
        base = ctypes.c_int32()
        self._lib.env_get_info(self.game, handle, b"attack_base", ctypes.byref(base))
        return base.value, buf

    def get_global_minimap(self, height, width):
        buf = np.empty((height, width, len(self.group_handles)), dtype=np.float32)
        buf[0, 0, 0] = height
        buf[0, 0, 1] = width
        self._lib.env_get_info(self.game, -1, b"global_minimap", buf.ctypes.data_as(ctypes.c_void_p))
        return buf

    def set_seed(self, seed):
        self._lib.env_config_game(self.game, b"seed", ctypes.byref(ctypes.c_int(seed)))

    # ------------------------------------------------------------------ render (no-op in the CUDA engine)
    def set_render_dir(self, name):
        if not os.path.exists(name):
            os.mkdir(name)
        self._lib.env_config_game(self.game, b"render_dir", ctypes.c_char_p(name.encode("ascii")))

    def render(self):
        self._lib.env_render(self.game)

    def __del__(self):
        game = getattr(self, "game", None)
        if game is not None and game.value:
            self._lib.env_delete_game(game)
            self.game = None

    # ------------------------------------------------------------------ deprecated in the reference
    def set_goal(self, handle, method, *args, **kwargs):
        if method != "random":
            raise NotImplementedError
        self._lib.gridworld_set_goal(self.game, handle, b"random", 0)

    # ------------------------------------------------------------------ reward DSL -> engine
    def _serialize_event_exp(self, config):
        """Number the symbols and event nodes of every rule and ship them to the engine
        (wire format of reference gridworld.py:646-722: symbols first, then nodes, then rules)."""
        symbols, nodes = {}, {}

        def see_symbol(sym):
            symbols.setdefault(sym, len(symbols))

        def walk(node):
            nodes.setdefault(node, len(nodes))
            for item in node.inputs:
                if isinstance(item, EventNode):
                    walk(item)

        def walk_symbols(node):
            for item in node.inputs:
                if isinstance(item, EventNode):
                    walk_symbols(item)
                elif isinstance(item, AgentSymbol):
                    see_symbol(item)

        for on, receivers, _values, _terminal in config.reward_rules:
            for sym in receivers:
                see_symbol(sym)
            walk_symbols(on)
        for on, _r, _v, _t in config.reward_rules:
            walk(on)
        config.symbol_ct, config.node_ct = len(symbols), len(nodes)

        for sym, no in symbols.items():
            self._lib.gridworld_define_agent_symbol(self.game, no, sym.group, sym.index)

        for node, no in nodes.items():
            args = np.zeros((len(node.inputs),), dtype=np.int32)
            for i, item in enumerate(node.inputs):
                if isinstance(item, EventNode):
                    args[i] = nodes[item]
                elif isinstance(item, AgentSymbol):
                    args[i] = symbols[item]
                else:
                    args[i] = item
            self._lib.gridworld_define_event_node(self.game, no, node.op, as_int32_c_array(args),
                                                  len(args))

        for on, receivers, values, terminal in config.reward_rules:
            recv = np.array([symbols[s] for s in receivers], dtype=np.int32)
            auto = len(values) == 1 and values[0] == "auto"
            vals = (np.zeros((len(recv),), dtype=np.float32) if auto
                    else np.array(values, dtype=np.float32))
            self._lib.gridworld_add_reward_rule(self.game, nodes[on], as_int32_c_array(recv),
                                                as_float_c_array(vals), len(recv), bool(terminal),
                                                bool(auto))


# ---------------------------------------------------------------------- reward description DSL
class EventNode:
    """AST node of an event expression; op codes are the engine's EventOp enum
    (reference grid_def.h:18-24, gridworld.py:734-746)."""
    OP_AND, OP_OR, OP_NOT = 0, 1, 2
    OP_KILL, OP_AT, OP_IN, OP_COLLIDE, OP_ATTACK, OP_DIE, OP_IN_A_LINE, OP_ALIGN = range(3, 11)

    _BINARY = {"kill": OP_KILL, "attack": OP_ATTACK, "collide": OP_COLLIDE}
    _UNARY = {"die": OP_DIE, "in_a_line": OP_IN_A_LINE, "align": OP_ALIGN}

    def __init__(self, op=None, inputs=(), predicate=None):
        self.op = op
        self.predicate = predicate
        self.inputs = list(inputs)

    def __call__(self, subject, predicate, *args):
        if predicate in self._BINARY:
            return EventNode(self._BINARY[predicate], [subject, args[0]], predicate)
        if predicate in self._UNARY:
            return EventNode(self._UNARY[predicate], [subject], predicate)
        if predicate == "at":
            return EventNode(self.OP_AT, [subject, args[0][0], args[0][1]], predicate)
        if predicate == "in":
            (xa, ya), (xb, yb) = args[0]
            return EventNode(self.OP_IN, [subject, min(xa, xb), min(ya, yb), max(xa, xb), max(ya, yb)],
                             predicate)
        raise Exception("invalid predicate of event " + predicate)

    def __and__(self, other):
        return EventNode(self.OP_AND, [self, other])

    def __or__(self, other):
        return EventNode(self.OP_OR, [self, other])

    def __invert__(self):
        return EventNode(self.OP_NOT, [self])


Event = EventNode()


class AgentSymbol:
    """A (group, index) placeholder; index 'any' -> -1, 'all' -> -2, or a fixed int."""

    def __init__(self, group, index):
        self.group = group if group is not None else -1
        if index == "any":
            self.index = -1
        elif index == "all":
            self.index = -2
        else:
            assert isinstance(index, int), "index must be a deterministic int"
            self.index = index

    def __str__(self):
        return "agent(%d,%d)" % (self.group, self.index)


class Config:
    """Game description: global keys, agent types, groups, reward rules."""

    def __init__(self):
        self.config_dict = {}
        self.agent_type_dict = {}
        self.groups = []
        self.reward_rules = []

    def set(self, args):
        self.config_dict.update(args)

    def register_agent_type(self, name, attr):
        if name in self.agent_type_dict:
            raise Exception("type name %s already exists" % name)
        self.agent_type_dict[name] = attr
        return name

    def add_group(self, agent_type):
        self.groups.append(agent_type)
        return len(self.groups) - 1

    def add_reward_rule(self, on, receiver, value, terminal=False):
        if not isinstance(receiver, (tuple, list)):
            assert not isinstance(value, (tuple, list))
            receiver, value = [receiver], [value]
        if len(receiver) != len(value):
            raise Exception("the length of receiver and value should be equal")
        self.reward_rules.append([on, list(receiver), list(value), terminal])


class CircleRange:
    def __init__(self, radius):
        self.radius = radius
        self.angle = 360

    def __str__(self):
        return "circle(%g)" % self.radius


class SectorRange:
    def __init__(self, radius, angle):
        if angle >= 180:
            raise Exception("the angle of a sector should be smaller than 180 degree")
        self.radius = radius
        self.angle = angle

    def __str__(self):
        return "sector(%g, %g)" % (self.radius, self.angle)
