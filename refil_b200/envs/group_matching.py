"""Group Matching on device: thousands of instances stepped by one sm_100a kernel launch.

`GroupMatchingBatch` is the batched environment the device runner uses; `GroupMatching` presents the reference's
single-environment MultiAgentEnv API (/root/reference/src/envs/group_matching/group_matching.py:5-127: reset / step /
get_entities / get_masks / get_avail_actions / get_env_info ...) on top of a 1-instance batch, so code written against
the reference class keeps working (and is bit-exact for the same seed: same numpy-legacy MT19937 stream).
"""
import numpy as np
import torch

from .. import _lib, ops

F_ALIVE, F_INLIST, F_SOLVED, F_LIMIT = 1, 2, 4, 8


class GroupMatchingBatch:
    def __init__(self, n_envs, n_agents=4, n_states=10, n_groups=2, rand_trans=0.1, episode_limit=50,
                 fixed_scen=False, seed=0, n_entities=None, device="cuda", first_env_index=0, entity_scheme=True):
        assert entity_scheme, "This environment only supports entity scheme"
        self.E, self.n_agents, self.n_states, self.n_groups = int(n_envs), int(n_agents), int(n_states), int(n_groups)
        self.rand_trans, self.episode_limit, self.fixed_scen = float(rand_trans), int(episode_limit), bool(fixed_scen)
        self.n_entities = int(n_entities) if n_entities is not None else self.n_agents
        self.n_actions = 3
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.RefilError("GroupMatchingBatch runs on a CUDA device only (got %s); no CPU fallback" % device)
        E, dev = self.E, self.device
        self.mt_key = torch.empty(624, E, dtype=torch.int32, device=dev)
        self.mt_pos = torch.zeros(E, dtype=torch.int32, device=dev)
        self.loc = torch.zeros(self.n_agents, E, dtype=torch.int32, device=dev)
        self.grp = torch.zeros(self.n_groups, E, dtype=torch.int32, device=dev)
        self.est = torch.zeros(4, E, dtype=torch.int32, device=dev)   # prev_matches, t, flags, ep_len
        self.ep_ret = torch.zeros(E, dtype=torch.float64, device=dev)
        self.step_counter = torch.zeros(1, dtype=torch.int64, device=dev)
        self.seed(seed, first_env_index)

    # ---- RNG ------------------------------------------------------------------------------------------------------
    def seed(self, seed, first_env_index=0):
        """env i gets numpy RandomState(seed + first_env_index + i)  (parallel_runner.py:23-26: base_seed + rank)."""
        seeds = (int(seed) + int(first_env_index) + np.arange(self.E, dtype=np.int64)) % (2 ** 32)
        s = torch.from_numpy(seeds.astype(np.uint32).view(np.int32)).to(self.device)
        _lib.call("gm_env_seed", self.mt_key.data_ptr(), self.mt_pos.data_ptr(), s.data_ptr(), self.E,
                  torch.cuda.current_stream().cuda_stream)
        ops._launches += 1

    # ---- dynamics -------------------------------------------------------------------------------------------------
    def get_entity_size(self):
        return self.n_states + self.n_groups + self.n_agents

    def get_env_info(self, args=None):
        return {"entity_shape": self.get_entity_size(), "n_actions": self.n_actions, "n_agents": self.n_agents,
                "n_entities": self.n_entities, "gt_mask_avail": True, "episode_limit": self.episode_limit}

    @staticmethod
    def _ptr(batch, key):
        try:
            t = batch[key]
        except (KeyError, ValueError):
            return None
        if not t.is_cuda or not t.is_contiguous():
            raise _lib.RefilError("rollout tensor %r must be a contiguous CUDA tensor" % key)
        return t.data_ptr()

    def reset(self, batch, env_offset=0):
        """Reset every instance and write timestep 0 of `batch` (EpisodeBatch layout [B, T, ...])."""
        T = batch["entities"].shape[1]
        _lib.call("gm_env_reset", self.mt_key.data_ptr(), self.mt_pos.data_ptr(), self.loc.data_ptr(),
                  self.grp.data_ptr(), self.est.data_ptr(), self.ep_ret.data_ptr(), self._ptr(batch, "entities"),
                  self._ptr(batch, "obs_mask"), self._ptr(batch, "entity_mask"), self._ptr(batch, "gt_mask"),
                  self._ptr(batch, "avail_actions"), self._ptr(batch, "filled"), self.E, self.n_agents, self.n_entities,
                  self.n_states, self.n_groups, int(self.fixed_scen), self.episode_limit, T, int(env_offset),
                  torch.cuda.current_stream().cuda_stream)
        ops._launches += 1

    def step(self, batch, ts, env_offset=0, write_gt=False):
        """Advance every live instance with batch['actions'][:, ts]; writes reward / terminated at ts and the
        observation rows at ts + 1 (plus gt_mask when write_gt: the EpisodeRunner's per-step pre-transition data)."""
        T = batch["entities"].shape[1]
        _lib.call("gm_env_step", self.mt_key.data_ptr(), self.mt_pos.data_ptr(), self.loc.data_ptr(),
                  self.grp.data_ptr(), self.est.data_ptr(), self.ep_ret.data_ptr(), self._ptr(batch, "actions"),
                  self._ptr(batch, "entities"), self._ptr(batch, "obs_mask"), self._ptr(batch, "entity_mask"),
                  self._ptr(batch, "gt_mask") if write_gt else None, self._ptr(batch, "avail_actions"),
                  self._ptr(batch, "reward"), self._ptr(batch, "terminated"),
                  self._ptr(batch, "filled"), self.step_counter.data_ptr(), self.E, self.n_agents, self.n_entities,
                  self.n_states, self.n_groups, self.rand_trans, self.episode_limit, T, int(ts), int(env_offset),
                  torch.cuda.current_stream().cuda_stream)
        ops._launches += 1

    # ---- views ----------------------------------------------------------------------------------------------------
    @property
    def flags(self):
        return self.est[2]

    def alive(self):
        return (self.est[2] & F_ALIVE) != 0


class _MiniBatch(dict):
    """the rollout tensors of a 1-env batch"""


class GroupMatching:
    """Single-environment API of the reference class, backed by a 1-instance `GroupMatchingBatch` on the GPU."""

    def __init__(self, entity_scheme=True, n_agents=4, n_states=10, n_groups=2, rand_trans=0.1, episode_limit=50,
                 fixed_scen=False, seed=None, device="cuda"):
        assert entity_scheme, "This environment only supports entity scheme"
        if seed is None:
            seed = int(np.random.SeedSequence().entropy % (2 ** 32))
        self.n_agents, self.n_states, self.n_groups = n_agents, n_states, n_groups
        self.rand_trans, self.episode_limit, self.fixed_scen = rand_trans, episode_limit, fixed_scen
        self.n_actions = 3
        self._env = GroupMatchingBatch(1, n_agents, n_states, n_groups, rand_trans, episode_limit, fixed_scen, seed,
                                       device=device)
        dev, na, T = self._env.device, n_agents, episode_limit + 1
        ed = self.get_entity_size()
        self._b = _MiniBatch(
            entities=torch.zeros(1, T, na, ed, device=dev), obs_mask=torch.zeros(1, T, na, na, dtype=torch.uint8, device=dev),
            entity_mask=torch.zeros(1, T, na, dtype=torch.uint8, device=dev),
            gt_mask=torch.zeros(1, T, na, na, dtype=torch.uint8, device=dev),
            avail_actions=torch.zeros(1, T, na, 3, dtype=torch.int32, device=dev),
            actions=torch.zeros(1, T, na, 1, dtype=torch.int64, device=dev), reward=torch.zeros(1, T, 1, device=dev),
            terminated=torch.zeros(1, T, 1, dtype=torch.uint8, device=dev),
            filled=torch.zeros(1, T, 1, dtype=torch.int64, device=dev))
        self.t = 0
        self._gt0 = None

    def seed(self, seed):
        self._env.seed(seed)

    def reset(self, **kwargs):
        for v in self._b.values():
            v.zero_()
        self._env.reset(self._b)
        self.t = 0
        est = self._env.est.cpu()
        self.prev_matches = int(est[0, 0])
        self._gt0 = self._b["gt_mask"][0, 0].cpu().numpy().copy()
        return self.get_entities(), self.get_masks()

    def step(self, actions):
        a = torch.as_tensor(np.asarray([int(x) for x in actions[:self.n_agents]], dtype=np.int64))
        self._b["actions"][0, self.t, :, 0] = a.to(self._env.device)
        self._env.step(self._b, self.t)
        est = self._env.est.cpu()
        m, flags = int(est[0, 0]), int(est[2, 0])
        rew = -0.1
        rew += 2.5 * (m - self.prev_matches)
        self.prev_matches = m
        self.t += 1
        info = {"solved": bool(flags & F_SOLVED)}
        if flags & F_LIMIT:
            info["episode_limit"] = True
        return rew, not (flags & F_ALIVE), info

    def get_entities(self):
        e = self._b["entities"][0, self.t].cpu().numpy()
        return [e[i] for i in range(self.n_agents)]

    def get_entity_size(self):
        return self.n_states + self.n_groups + self.n_agents

    def get_masks(self):
        na = self.n_agents
        return np.zeros((na, na), np.uint8), np.zeros(na, np.uint8), self._gt0.copy()

    def get_locs(self):
        return self._env.loc[:, 0].cpu().numpy()

    def get_avail_actions(self):
        return [[1] * self.n_actions for _ in range(self.n_agents)]

    def get_avail_agent_actions(self, agent_id):
        return [1] * self.n_actions

    def get_total_actions(self):
        return self.n_actions

    def get_stats(self):
        return {}

    def get_agg_stats(self, stats):
        return {}

    def close(self):
        return

    def get_env_info(self, args=None):
        return {"entity_shape": self.get_entity_size(), "n_actions": self.n_actions, "n_agents": self.n_agents,
                "n_entities": self.n_agents, "gt_mask_avail": True, "episode_limit": self.episode_limit}
