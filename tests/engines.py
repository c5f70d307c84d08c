"""Engine adapters used by the differential tests (test infrastructure).

Three engines expose one duck-typed interface (group = 0/1 ints):
  RefEngine    the UNMODIFIED reference C++ engine, oracle/_ref/libmagent_ref.so, through this
               package's `magent` mirror (OMP_NUM_THREADS=1)
  OracleEngine the plain-C restatement, oracle/_build/liboracle.so
  CudaEngine   the product: build/libmagent.so through the same `magent` mirror (C ABI, host buffers)
"""
import ctypes
import os
import subprocess

import numpy as np

from conftest import PKG, REPO

REF_SO = os.path.join(REPO, "oracle", "_ref", "libmagent_ref.so")
ORACLE_SO = os.path.join(REPO, "oracle", "_build", "liboracle.so")
CUDA_SO = os.path.join(PKG, "build", "libmagent.so")


def have_ref():
    return os.path.exists(REF_SO)


def build_oracle():
    if not os.path.exists(ORACLE_SO):
        subprocess.check_call(["make", "-C", os.path.join(REPO, "oracle"), "oracle"],
                              stdout=subprocess.DEVNULL)
    return ORACLE_SO


class _MagentEngine:
    """magent.GridWorld('battle') over a given library."""

    def __init__(self, so_path, map_size):
        import magent
        from magent import c_lib
        self.lib = c_lib.load(so_path)
        self.env = magent.GridWorld("battle", lib=self.lib, map_size=map_size)
        self.h = self.env.get_handles()
        self.map_size = map_size

    def set_seed(self, seed): self.env.set_seed(seed)
    def reset(self): self.env.reset()
    def add_agents(self, g, pos): self.env.add_agents(self.h[g], method="custom", pos=pos)
    def add_walls(self, pos): self.env.add_walls(method="custom", pos=pos)
    def get_num(self, g): return self.env.get_num(self.h[g])
    def get_observation(self, g):
        v, f = self.env.get_observation(self.h[g])
        return v.copy(), f.copy()
    def set_action(self, g, acts): self.env.set_action(self.h[g], np.ascontiguousarray(acts, dtype=np.int32))
    def step(self): return self.env.step()
    def get_reward(self, g): return self.env.get_reward(self.h[g])
    def get_alive(self, g): return self.env.get_alive(self.h[g])
    def get_agent_id(self, g): return self.env.get_agent_id(self.h[g])
    def get_pos(self, g): return self.env.get_pos(self.h[g])
    def get_mean_info(self, g): return self.env.get_mean_info(self.h[g])
    def clear_dead(self): self.env.clear_dead()


class RefEngine(_MagentEngine):
    def __init__(self, map_size):
        super().__init__(REF_SO, map_size)


class CudaEngine(_MagentEngine):
    def __init__(self, map_size):
        super().__init__(CUDA_SO, map_size)


class OracleEngine:
    def __init__(self, map_size, embedding_size=10):
        lib = ctypes.CDLL(build_oracle())
        vp, ci = ctypes.c_void_p, ctypes.c_int
        lib.mo_new.restype = vp
        lib.mo_new.argtypes = [ci, ci, ci, vp]
        for name in ("mo_free", "mo_reset", "mo_clear_dead"):
            getattr(lib, name).argtypes = [vp]
            getattr(lib, name).restype = None
        lib.mo_set_seed.argtypes = [vp, ctypes.c_ulong]
        lib.mo_add_agents.argtypes = [vp, ci, ci, vp, vp]
        lib.mo_add_agents.restype = ci
        for name in ("mo_get_num",):
            getattr(lib, name).argtypes = [vp, ci]
            getattr(lib, name).restype = ci
        for name in ("mo_view_size", "mo_n_channel", "mo_feature_size", "mo_n_action", "mo_view_count",
                     "mo_attack_count", "mo_step"):
            getattr(lib, name).argtypes = [vp]
            getattr(lib, name).restype = ci
        lib.mo_get_observation.argtypes = [vp, ci, vp, vp]
        for name in ("mo_set_action", "mo_get_reward", "mo_get_alive", "mo_get_id", "mo_get_pos",
                     "mo_get_hp", "mo_get_mean_info"):
            getattr(lib, name).argtypes = [vp, ci, vp]
            getattr(lib, name).restype = None
        lib.mo_inject_attack_order.argtypes = [vp, vp, ci]
        lib.mo_action_table.argtypes = [vp, vp, vp]
        lib.mo_rng_next.argtypes = [vp]
        lib.mo_rng_next.restype = ctypes.c_ulong
        lib.mo_mean_action.argtypes = [vp, ci, ci, vp]
        self.lib = lib
        self.e = lib.mo_new(map_size, map_size, embedding_size, None)
        self.map_size = map_size
        self.vs = lib.mo_view_size(self.e)
        self.nc = lib.mo_n_channel(self.e)
        self.fs = lib.mo_feature_size(self.e)
        self.n_action = lib.mo_n_action(self.e)

    def __del__(self):
        if getattr(self, "e", None):
            self.lib.mo_free(self.e)
            self.e = None

    @staticmethod
    def _p(a): return a.ctypes.data_as(ctypes.c_void_p)

    def set_seed(self, seed): self.lib.mo_set_seed(self.e, seed)
    def reset(self): self.lib.mo_reset(self.e)
    def add_agents(self, g, pos):
        pos = np.asarray(pos, dtype=np.int32)
        xs, ys = np.ascontiguousarray(pos[:, 0]), np.ascontiguousarray(pos[:, 1])
        return self.lib.mo_add_agents(self.e, g, len(pos), self._p(xs), self._p(ys))
    def add_walls(self, pos): return self.add_agents(-1, pos)
    def get_num(self, g): return self.lib.mo_get_num(self.e, g)
    def get_observation(self, g):
        n = self.get_num(g)
        v = np.empty((n, self.vs, self.vs, self.nc), np.float32)
        f = np.empty((n, self.fs), np.float32)
        self.lib.mo_get_observation(self.e, g, self._p(v), self._p(f))
        return v, f
    def set_action(self, g, acts):
        acts = np.ascontiguousarray(acts, dtype=np.int32)
        self.lib.mo_set_action(self.e, g, self._p(acts))
    def inject_attack_order(self, perm):
        perm = np.ascontiguousarray(perm, dtype=np.int32)
        self.lib.mo_inject_attack_order(self.e, self._p(perm), len(perm))
    def attack_count(self): return self.lib.mo_attack_count(self.e)
    def step(self): return bool(self.lib.mo_step(self.e))
    def _get(self, fn, g, shape, dtype):
        buf = np.empty(shape, dtype)
        fn(self.e, g, self._p(buf))
        return buf
    def get_reward(self, g): return self._get(self.lib.mo_get_reward, g, (self.get_num(g),), np.float32)
    def get_alive(self, g): return self._get(self.lib.mo_get_alive, g, (self.get_num(g),), np.uint8).astype(bool)
    def get_agent_id(self, g): return self._get(self.lib.mo_get_id, g, (self.get_num(g),), np.int32)
    def get_pos(self, g): return self._get(self.lib.mo_get_pos, g, (self.get_num(g), 2), np.int32)
    def get_hp(self, g): return self._get(self.lib.mo_get_hp, g, (self.get_num(g),), np.float32)
    def get_mean_info(self, g): return self._get(self.lib.mo_get_mean_info, g, (2 + self.n_action,), np.float32)
    def clear_dead(self): self.lib.mo_clear_dead(self.e)
    def rng_next(self): return self.lib.mo_rng_next(self.e)
    def action_table(self):
        t = np.zeros((self.n_action, 2), np.int32)
        base = ctypes.c_int()
        self.lib.mo_action_table(self.e, self._p(t), ctypes.byref(base))
        return t, base.value
    def mean_action(self, acts):
        acts = np.ascontiguousarray(acts, dtype=np.int32)
        out = np.empty((self.n_action,), np.float64)
        self.lib.mo_mean_action(self._p(acts), len(acts), self.n_action, self._p(out))
        return out
