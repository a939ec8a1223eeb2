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
        self._graph_state = None

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

    # ---- the rollout body: everything between reset and the statistics, enqueued without a host sync -------------------
    def _rollout_steps(self, batch, eps_dev=None, test_mode=False, early_exit=False):
        """mac.forward -> epsilon-greedy selection (written straight into batch["actions"][:, t]) -> env step, for every
        timestep.  With eps_dev (device scalar) the sequence contains no host-dependent value and can be captured once."""
        env = self.env
        self.mac.init_hidden(batch_size=self.batch_size)
        avail, acts = batch["avail_actions"], batch["actions"]
        # the exploration uniforms of the whole rollout in ONE draw (one launch instead of one per timestep)
        explore = eps_dev is not None or (not test_mode and self.mac.action_selector.schedule.eval(self.t_env) > 0.0)
        u_all = torch.rand(self.episode_limit, 2, self.batch_size, self.args.n_agents, device=self.args.device) if explore else None
        fused = (hasattr(self.mac, "act_and_select") and self.mac._fused_acting_ok(batch) and avail.is_contiguous()
                 and avail.dtype == torch.int32 and acts.is_contiguous())
        for t in range(self.episode_limit):
            u_t = u_all[t] if u_all is not None else None
            if fused:       # small FF agent: forward + selection are one launch (csrc/ffact.cu)
                self.mac.act_and_select(batch, t, self.t_env, test_mode=test_mode, est_flags=env.flags, eps_dev=eps_dev, uniforms=u_t)
            else:
                q = self.mac.forward(batch, t, test_mode=test_mode)
                self.mac.action_selector.select_action(q, avail[:, t], self.t_env, test_mode=test_mode, est_flags=env.flags,
                                                       out=acts[:, t, :, 0], eps_dev=eps_dev, uniforms=u_t)
            env.step(batch, t, write_gt=self.write_gt_every_step)   # reward / terminated at t, observations + filled at t + 1
            self.t = t + 1
            if early_exit and (t & 7) == 7 and not bool(env.alive().any()):
                break
        # actions_onehot (the scheme's preprocess of `actions`, episode_buffer.py:88-95) once for the whole rollout
        batch.update({"actions": acts}, mark_filled=False)

    def _run_graph(self, test_mode):
        """args.rollout_graph: the whole rollout -- reset, `episode_limit` x (agent forward, selection, env step) -- is ONE CUDA
        graph on a static EpisodeBatch, replayed per run (the uniforms come from torch's graph-safe philox state, epsilon from a
        device scalar); the result is copied into a fresh EpisodeBatch, as `run` must return one the caller may keep.  The first
        run is eager (sizes the workspaces), the second captures."""
        st = self._graph_state
        if st is None:
            st = self._graph_state = {"runs": 0, "graph": None, "failed": False}
        eps = 0.0 if test_mode else self.mac.action_selector.schedule.eval(self.t_env)
        self.mac.action_selector.epsilon = eps
        if st["failed"] or st["runs"] == 0:
            st["runs"] += 1
            self.reset()
            self._rollout_steps(self.batch, test_mode=test_mode, early_exit=True)
            return
        if st["graph"] is None:
            st["static"] = self.new_batch()
            st["eps"] = torch.zeros(1, dtype=torch.float32, device=self.args.device)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            from .. import ops
            l0 = ops.launch_count()
            try:
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    for v in st["static"].data.transition_data.values():
                        v.zero_()
                    self.env.reset(st["static"])
                    self._rollout_steps(st["static"], eps_dev=st["eps"])
            except Exception as e:      # capture only changes how the work is submitted: never lose the rollout over it
                st["failed"] = True
                torch.cuda.synchronize()
                self.logger.console_logger.info("rollout graph capture failed (%s: %s); running eagerly" % (type(e).__name__, e))
                return self._run_graph(test_mode)
            st["graph"], st["launches"] = g, ops.launch_count() - l0
            ops.add_launches(-st["launches"])
        st["eps"].fill_(eps)
        st["graph"].replay()
        from .. import ops
        ops.add_launches(st["launches"])
        self.batch = self.new_batch(zero_init=False)       # every tensor is overwritten by the copy below
        for k, v in st["static"].data.transition_data.items():
            self.batch.data.transition_data[k].copy_(v)
        self.t = self.episode_limit

    def run(self, test_mode=False, test_scen=None, index=None, vid_writer=None):
        assert vid_writer is None, "Writing videos not supported for ParallelRunner"
        self.mac.eval()
        if getattr(self.args, "rollout_graph", False):
            self._run_graph(test_mode)
        else:
            self.reset()
            self._rollout_steps(self.batch, test_mode=test_mode, early_exit=True)
        batch, env = self.batch, self.env
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
