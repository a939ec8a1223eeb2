"""On-device rollout driver with the reference ParallelRunner interface.

Interface mirror of /root/reference/src/runners/parallel_runner.py:11-245 (setup / get_env_info / reset / run / close_env,
`t_env`, `batch_size`, stat keys).  The reference keeps one OS process per environment and exchanges pickled messages over
pipes every step; here all `batch_size_run` instances live in one `GroupMatchingBatch` on the GPU, the agent forward, the
epsilon-greedy selection and the env step are three kernel sequences per timestep, and observations / rewards are
written by the env kernel straight into the EpisodeBatch tensors (no host round trip inside an episode).

Reference quirks kept on purpose: `terminated` excludes time-limit endings; `t_env` counts one per
live env per step in train mode only; an env that finished at step t-1 still has an action selected (and stored) at step
t because the live-list is refreshed after selection (the env kernel keeps that list in bit 1 of its flags); `gt_mask`
is written at reset only."""
from functools import partial

import numpy as np
import torch

from .. import parallel
from ..components.episode_buffer import EpisodeBatch
from ..envs import BATCHED_REGISTRY
from ..envs.group_matching import F_LIMIT, F_SOLVED


class ParallelRunner:
    write_gt_every_step = False          # ParallelRunner: gt_mask at reset only (SURVEY.md section 3.2 quirk iv)

    def __init__(self, args, logger):
        self.args = args
        self.logger = logger
        self.batch_size = self.args.batch_size_run
        if args.env not in BATCHED_REGISTRY:
            raise KeyError("environment %r has no device-batched implementation (available: %s)"
                           % (args.env, sorted(BATCHED_REGISTRY)))
        env_args = dict(args.env_args)
        seed = env_args.pop("seed", 0)
        rank = int(getattr(args, "rank", 0))
        # env i of rank r is seeded base_seed + r * batch_size + i (parallel_runner.py:23-26: base_seed + worker rank)
        self.env = BATCHED_REGISTRY[args.env](self.batch_size, seed=seed, device=args.device,
                                              first_env_index=rank * self.batch_size, **env_args)
        self.env_info = self.env.get_env_info()
        self.episode_limit = self.env_info["episode_limit"]
        self.t = 0
        self.t_env = 0
        self.train_returns, self.test_returns = [], []
        self.train_stats, self.test_stats = {}, {}
        self.log_train_stats_t = -100000

    def setup(self, scheme, groups, preprocess, mac):
        self.new_batch = partial(EpisodeBatch, scheme, groups, self.batch_size, self.episode_limit + 1,
                                 preprocess=preprocess, device=self.args.device)
        self.mac = mac
        self.scheme, self.groups, self.preprocess = scheme, groups, preprocess

    def get_env_info(self):
        return self.env_info

    def save_replay(self):
        pass

    def close_env(self):
        pass

    def reset(self, **kwargs):
        self.batch = self.new_batch()
        self.env.reset(self.batch)          # writes entities / masks / avail_actions / filled at ts = 0
        self.t = 0
        self.env_steps_this_run = 0

    def run(self, test_mode=False, test_scen=None, index=None, vid_writer=None):
        assert vid_writer is None, "Writing videos not supported for ParallelRunner"
        self.reset()
        self.mac.init_hidden(batch_size=self.batch_size)
        self.mac.eval()
        batch, env = self.batch, self.env
        actions = torch.zeros(self.batch_size, self.args.n_agents, dtype=torch.int64, device=self.args.device)
        for t in range(self.episode_limit):
            q = self.mac.forward(batch, t, test_mode=test_mode)
            actions.zero_()
            self.mac.action_selector.select_action(q, batch["avail_actions"][:, t], self.t_env, test_mode=test_mode,
                                                   est_flags=env.flags, out=actions)
            batch.update({"actions": actions.unsqueeze(1)}, ts=t, mark_filled=False)
            env.step(batch, t, write_gt=self.write_gt_every_step)   # reward / terminated at t, observations + filled at t + 1
            self.t = t + 1
            if (t & 7) == 7 and not bool(env.alive().any()):
                break
        est = env.est.cpu().numpy()
        returns = env.ep_ret.cpu().numpy()
        lengths, flags = est[3], est[2]
        if not test_mode:
            # t_env counts the env steps of ALL ranks' instances, so every rank sees the same value (loop exit, target
            # updates, epsilon schedule and logging decisions stay in lock-step under torchrun)
            self.env_steps_this_run = parallel.all_reduce_sum_int(int(lengths.sum()), self.args.device)
            self.t_env += self.env_steps_this_run

        cur_stats = self.test_stats if test_mode else self.train_stats
        cur_returns = self.test_returns if test_mode else self.train_returns
        log_prefix = "test_" if test_mode else ""
        cur_stats["solved"] = cur_stats.get("solved", 0) + int(((flags & F_SOLVED) != 0).sum())
        cur_stats["episode_limit"] = cur_stats.get("episode_limit", 0) + int(((flags & F_LIMIT) != 0).sum())
        cur_stats["n_episodes"] = self.batch_size + cur_stats.get("n_episodes", 0)
        cur_stats["ep_length"] = int(lengths.sum()) + cur_stats.get("ep_length", 0)
        cur_returns.extend(float(r) for r in returns)

        n_test_runs = max(1, self.args.test_nepisode // self.batch_size) * self.batch_size
        if test_mode and len(self.test_returns) == n_test_runs:
            self._log(cur_returns, cur_stats, log_prefix)
        elif not test_mode and self.t_env - self.log_train_stats_t >= self.args.runner_log_interval:
            self._log(cur_returns, cur_stats, log_prefix)
            if hasattr(self.mac.action_selector, "epsilon"):
                self.logger.log_stat("epsilon", self.mac.action_selector.epsilon, self.t_env)
            self.log_train_stats_t = self.t_env
        return self.batch

    def _log(self, returns, stats, prefix):
        self.logger.log_stat(prefix + "return_mean", np.mean(returns), self.t_env)
        self.logger.log_stat(prefix + "return_std", np.std(returns), self.t_env)
        returns.clear()
        for k, v in stats.items():
            if k != "n_episodes":
                self.logger.log_stat(prefix + k + "_mean", v / stats["n_episodes"], self.t_env)
        stats.clear()


class EpisodeRunner(ParallelRunner):
    """`runner: episode` (runners/episode_runner.py:8-139): the same device loop, with the one data difference between the two
    reference runners kept -- the episode runner stores gt_mask in its pre-transition data at EVERY step (:52-67), so
    `train_gt_factors` sees the ground-truth factorisation on all timesteps."""
    write_gt_every_step = True
