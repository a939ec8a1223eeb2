"""ORACLE -- test infrastructure only.

ctypes front-end of oracle/gm_env.c presenting the reference's single-env API
(/root/reference/src/envs/group_matching/group_matching.py:5-127) so parity tests read like the
reference's own usage: ``env.reset(); rew, done, info = env.step(actions); env.get_entities()``.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.build()
        L = C.CDLL(path)
        L.gm_create.restype = C.c_void_p
        L.gm_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_uint32]
        L.gm_destroy.argtypes = [C.c_void_p]
        L.gm_reset.argtypes = [C.c_void_p, C.c_int]
        L.gm_step.restype = C.c_int
        L.gm_step.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.gm_get_entities.argtypes = [C.c_void_p, C.c_void_p]
        L.gm_get_masks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.gm_get_locs.argtypes = [C.c_void_p, C.c_void_p]
        L.gm_get_t.restype = C.c_int
        L.gm_get_t.argtypes = [C.c_void_p]
        L.mt_create.restype = C.c_void_p
        L.mt_create.argtypes = [C.c_uint32]
        L.mt_destroy.argtypes = [C.c_void_p]
        L.mt_next_u32.restype = C.c_uint32
        L.mt_next_u32.argtypes = [C.c_void_p]
        L.mt_next_double.restype = C.c_double
        L.mt_next_double.argtypes = [C.c_void_p]
        L.mt_next_bounded.restype = C.c_uint32
        L.mt_next_bounded.argtypes = [C.c_void_p, C.c_uint32]
        L.gm_bench_loop.restype = C.c_long
        L.gm_bench_loop.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class MT19937Oracle:
    """numpy legacy RandomState draw model (u32 / uniform / bounded)."""

    def __init__(self, seed):
        self._h = lib().mt_create(int(seed) & 0xFFFFFFFF)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().mt_destroy(self._h)
            self._h = None

    def u32(self):
        return int(lib().mt_next_u32(self._h))

    def uniform(self):
        return float(lib().mt_next_double(self._h))

    def bounded(self, mx):
        return int(lib().mt_next_bounded(self._h, int(mx)))


class GroupMatchingOracle:
    def __init__(self, entity_scheme=True, n_agents=4, n_states=10, n_groups=2, rand_trans=0.1,
                 episode_limit=50, fixed_scen=False, seed=None):
        assert entity_scheme
        assert seed is not None, "the oracle is deterministic; pass a seed"
        self.n_agents, self.n_states, self.n_groups = n_agents, n_states, n_groups
        self.rand_trans, self.episode_limit, self.fixed_scen = rand_trans, episode_limit, fixed_scen
        self.n_actions = 3
        self._h = lib().gm_create(n_agents, n_states, n_groups, float(rand_trans), episode_limit,
                                  int(seed) & 0xFFFFFFFF)
        assert self._h, "n_agents/n_groups beyond oracle limits"

    def __del__(self):
        if getattr(self, "_h", None):
            lib().gm_destroy(self._h)
            self._h = None

    def reset(self, **kwargs):
        lib().gm_reset(self._h, int(self.fixed_scen))
        return self.get_entities(), self.get_masks()

    def step(self, actions):
        a = np.ascontiguousarray(np.asarray(actions, dtype=np.int64)[: self.n_agents])
        rew = C.c_double()
        flags = lib().gm_step(self._h, a.ctypes.data, C.byref(rew))
        info = {"solved": bool(flags & 2)}
        if flags & 4:
            info["episode_limit"] = True
        return rew.value, bool(flags & 1), info

    def get_entity_size(self):
        return self.n_states + self.n_groups + self.n_agents

    def get_entities(self):
        out = np.empty((self.n_agents, self.get_entity_size()), dtype=np.float32)
        lib().gm_get_entities(self._h, out.ctypes.data)
        return [out[i] for i in range(self.n_agents)]

    def get_masks(self):
        na = self.n_agents
        obs = np.empty((na, na), np.uint8)
        ent = np.empty((na,), np.uint8)
        gt = np.empty((na, na), np.uint8)
        lib().gm_get_masks(self._h, obs.ctypes.data, ent.ctypes.data, gt.ctypes.data)
        return obs, ent, gt

    def get_locs(self):
        out = np.empty((self.n_agents,), np.int32)
        lib().gm_get_locs(self._h, out.ctypes.data)
        return out

    def get_avail_actions(self):
        return [[1] * self.n_actions for _ in range(self.n_agents)]

    def get_total_actions(self):
        return self.n_actions

    def get_env_info(self, args=None):
        return {"entity_shape": self.get_entity_size(), "n_actions": self.n_actions,
                "n_agents": self.n_agents, "n_entities": self.n_agents, "gt_mask_avail": True,
                "episode_limit": self.episode_limit}

    def bench_loop(self, n_steps):
        na = self.n_agents
        ent = np.empty((na, self.get_entity_size()), np.float32)
        msk = np.empty((2 * na * na + na,), np.uint8)
        return lib().gm_bench_loop(self._h, int(n_steps), ent.ctypes.data, msk.ctypes.data)
