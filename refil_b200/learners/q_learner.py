"""Double-Q TD learner with the REFIL auxiliary loss, as one explicit forward/backward schedule of sm_100a kernels.

API mirror of /root/reference/src/learners/q_learner.py:10-229 (QLearner): ctor (mac, scheme, logger, args);
train(batch, t_env, episode_num); cuda(); save_models(path); load_models(path, evaluate); _update_targets().
Arithmetic follows q_learner.py:66-182; the schedule differs from the reference on purpose:
  * fc1 / in_trans of the agent and of hyper_w_1 run ONCE for the three imagine copies (the reference repeats them 3x);
  * hyper_w_final / hyper_b_1 / V are evaluated once for the plain and the imagine mix (identical inputs);
  * losses are kept as sums; the division by sum(mask), the global-norm clip and RMSprop run in one kernel after the
    single all-reduce of [flat grads | loss statistics] (multi-GPU: replay sharded over episodes).
"""
import os

import torch
from .. import ops, parallel
from ..modules.nets import Mixer, hypernets_backward_group, hypernets_forward_group
from ..modules.params import Workspace

N_STATS = 8


class _StaticBatch:
    """Fixed-address copy of the EpisodeBatch tensors the learner reads: the input side of a captured CUDA graph."""

    def __init__(self, batch, keys):
        self.batch_size, self.max_seq_length = batch.batch_size, batch.max_seq_length
        self.data = {}
        for k in keys:
            try:
                t = batch[k]
            except (KeyError, ValueError, AttributeError):
                continue
            self.data[k] = t.contiguous().clone()

    def __getitem__(self, k):
        return self.data[k]

    def load(self, batch):
        for k, dst in self.data.items():
            src = batch[k]
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)


class QLearner:
    def __init__(self, mac, scheme, logger, args):
        self.args = args
        self.mac = mac
        self.logger = logger
        self.device = mac.agent.store.flat.device
        self.last_target_update_episode = 0
        self.imagine = "imagine" in args.agent
        self.ein = mac.agent.ein

        self.mixer = None
        if args.mixer is not None:
            if args.mixer not in Mixer.HYPERS:
                raise ValueError("Mixer %s not recognised (flat-state `qmix` is out of scope)." % args.mixer)
            self.mixer = Mixer(args, self.ein, self.device, tag="mixer")
            self.target_mixer = Mixer(args, self.ein, self.device, tag="target_mixer")
            self.target_mixer.load_state_dict(self.mixer.state_dict())
            names = list(self.mixer.nets.keys())
            self.mixer.set_scratch_groups([[h] for h in names])        # every hypernetwork's backward runs on its own stream
        self.target_mac = type(mac)(mac.scheme, mac.groups, args)
        self.target_mac.load_state(mac)
        self._bind_flat()
        self.ws = Workspace(self.device)
        self.log_stats_t = -self.args.learner_log_interval - 1
        self._graphs = {}

    # ---- flat parameter / gradient / optimiser state -----------------------------------------------------------
    def _bind_flat(self):
        stores = [self.mac.agent.store] + ([self.mixer.store] if self.mixer is not None else [])
        self.n_params = sum(s.padded_size for s in stores)
        dev = self.device
        self.flat = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
        self.gradbuf = torch.zeros(self.n_params + N_STATS, dtype=torch.float32, device=dev)   # [grads | loss stats]
        self.square_avg = torch.zeros(self.n_params, dtype=torch.float32, device=dev)          # RMSprop state
        off = 0
        for s in stores:
            s.rebind(self.flat[off:off + s.padded_size], self.gradbuf[off:off + s.padded_size])
            off += s.padded_size
        self.params = list(self.mac.parameters()) + (list(self.mixer.parameters()) if self.mixer is not None else [])
        self.stats64 = torch.zeros(N_STATS, dtype=torch.float64, device=dev)
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)

    def sync_replicas(self, src=0):
        """Multi-GPU: broadcast rank `src`'s parameters, optimiser state and target networks (no-op on one rank)."""
        bufs = [self.flat, self.square_avg, self.target_mac.agent.store.flat]
        if self.mixer is not None:
            bufs.append(self.target_mixer.store.flat)
        for b in bufs:
            parallel.broadcast_(b, src=src)

    N_SIDE = 14      # target agent | 4 online hypernetworks | 4 target hypernetworks | weight-gradient companions (agent + 4)

    def _side_stream(self, i):
        if getattr(self, "_side", None) is None:
            self._side = [torch.cuda.Stream(device=self.device) for _ in range(self.N_SIDE)]
        return self._side[i]

    def _crit_stream(self):
        """High-priority stream for the agent chain (the step's critical path): when SMs free up, its CTAs are placed first."""
        if getattr(self, "_crit", None) is None:
            self._crit = torch.cuda.Stream(device=self.device, priority=-1)
        return self._crit

    def cuda(self):
        self.mac.cuda()
        self.target_mac.cuda()
        if self.mixer is not None:
            self.mixer.cuda()
            self.target_mixer.cuda()
        if self.mac.agent.store.flat.device != self.device:
            self.device = self.mac.agent.store.flat.device
            self._bind_flat()
            self.ws = Workspace(self.device)

    # ---- the training step ---------------------------------------------------------------------------------------
    def train(self, batch, t_env, episode_num, group_bits=None):
        args = self.args
        na = args.n_agents
        will_log = t_env - self.log_stats_t >= args.learner_log_interval
        log_gt = will_log and self.imagine and getattr(args, "test_gt_factors", False) and args.mixer == "lin_flex_qmix"
        # the device part of the step: eagerly, or (args.cuda_graph) as ONE replayed CUDA graph per batch shape
        if getattr(args, "cuda_graph", False) and group_bits is None and not log_gt:
            ingroup = gt_ingroup = None
            self._graph_step(batch)
        else:
            ingroup, gt_ingroup = self._device_step(batch, group_bits, log_gt)

        if (episode_num - self.last_target_update_episode) / args.target_update_interval >= 1.0:
            self._update_targets()
            self.last_target_update_episode = episode_num

        if will_log:
            st = self.gradbuf[self.n_params:].double().cpu()      # the only host sync of the step
            m = float(st[0])
            td, td_im = float(st[1]) / m, float(st[2]) / m
            loss = (1 - args.lmbda) * td + args.lmbda * td_im if self.imagine else td
            self.logger.log_stat("loss", loss, t_env)
            if self.imagine:
                self.logger.log_stat("im_loss", td_im, t_env)
            if log_gt:
                self.logger.log_stat("ingroup_prop", float(ingroup.item()), t_env)
                self.logger.log_stat("gt_ingroup_prop", float(gt_ingroup.item()), t_env)
            self.logger.log_stat("grad_norm", float(self.grad_norm.item()), t_env)
            self.logger.log_stat("td_error_abs", float(st[3]) / m, t_env)
            self.logger.log_stat("q_taken_mean", float(st[4]) / (m * na), t_env)
            self.logger.log_stat("target_mean", float(st[5]) / (m * na), t_env)
            self.log_stats_t = t_env

    # ---- CUDA-graph replay of the device step -------------------------------------------------------------------------
    _GRAPH_KEYS = ("entities", "obs_mask", "entity_mask", "gt_mask", "actions", "avail_actions", "reward", "terminated",
                   "filled")

    def _graph_step(self, batch):
        """First call for a (B, T) shape runs eagerly (sizes every workspace), the second captures the whole step -- all
        three streams, ~200 kernel launches -- into two CUDA graphs on static input tensors (forward + backward | normalise +
        clip + RMSprop; the NCCL gradient all-reduce between them stays an eager call), later calls copy the batch into those
        tensors (device to device) and replay.  Host cost per step: a dozen copies + 2 graph launches."""
        key = (batch.batch_size, batch.max_seq_length)
        gen = self._ws_generation()
        if gen != getattr(self, "_graphs_gen", gen):
            self._graphs.clear()               # a workspace grew and moved: every captured address is stale
        self._graphs_gen = gen
        ent = self._graphs.pop(key, None)
        if ent is not None:
            self._graphs[key] = ent            # most recently used last
        if ent is None:
            while len(self._graphs) >= self.MAX_GRAPHS:      # LRU bound: each entry pins a static batch + two graphs
                self._graphs.pop(next(iter(self._graphs)))
            self._device_step(batch, None, False)
            if self._ws_generation() != gen:   # this shape grew a workspace: graphs of the other shapes are stale too
                self._graphs.clear()
                self._graphs_gen = self._ws_generation()
            self._graphs[key] = {"static": None, "graph": None, "launches": 0}
            return
        if ent.get("failed"):
            self._device_step(batch, None, False)
            return
        if ent["graph"] is None:
            static = _StaticBatch(batch, self._GRAPH_KEYS)
            torch.cuda.synchronize()
            g, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            l0 = ops.launch_count()
            try:
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    self._device_step(static, None, False, reduce_and_update=False)
                with torch.cuda.graph(g2, capture_error_mode="thread_local"):
                    self._update_step()
            except Exception as e:           # capture is an optimisation of the submission path only: never lose the step
                ops.add_launches(l0 - ops.launch_count())
                ent["failed"] = True
                torch.cuda.synchronize()
                if self.logger is not None:
                    self.logger.console_logger.info("CUDA graph capture failed (%s: %s); training eagerly" % (type(e).__name__, e))
                self._device_step(batch, None, False)
                return
            ent["static"], ent["graph"], ent["update"], ent["launches"] = static, g, g2, ops.launch_count() - l0
            ops.add_launches(-ent["launches"])                     # capturing launched nothing
        ent["static"].load(batch)
        ent["graph"].replay()
        parallel.all_reduce_sum_(self.gradbuf)                     # eager: NCCL stays outside the captured graphs
        ent["update"].replay()
        ops.add_launches(ent["launches"])

    MAX_GRAPHS = 8
    GROUP_MAX_ENTITY_ROWS = 65536

    def _ws_generation(self):
        wss = [self.ws, self.mac.agent.ws, self.target_mac.agent.ws]
        if self.mixer is not None:
            wss += [self.mixer.ws, self.target_mixer.ws]
        return sum(w.generation for w in wss)

    def _device_step(self, batch, group_bits, log_gt, reduce_and_update=True):
        args, ws = self.args, self.ws
        B, T = batch.batch_size, batch.max_seq_length
        na, A = args.n_agents, args.n_actions
        N = B * T
        actions = batch["actions"].contiguous().view(N, na)
        avail = batch["avail_actions"].contiguous().view(N, na, A)
        reward = batch["reward"].contiguous().view(N)
        terminated = batch["terminated"].contiguous().view(N)
        filled = batch["filled"].contiguous().view(N)
        gt_ingroup = ingroup = None
        if log_gt:
            # logging-only pass with the ground-truth factorisation (q_learner.py:98-105,138-147), same weights, no grads
            self.mac.init_hidden(B)
            q_gt, _, mix_gt, inp_gt = self.mac.forward(batch, None, imagine=True, use_gt_factors=True, ret_plan=True)
            ch_gt = ops.gather_chosen(q_gt, actions, ws.get("chosen_gt", (3, N, na)), 3, N * na, A)
            self.mixer.forward(ch_gt[0], ch_gt[1], ch_gt[2], inp_gt["entities"], inp_gt["last_action"],
                               inp_gt["entity_mask"], T, imagine_masks=mix_gt, xin=inp_gt.get("xin"), ret_ingroup=True)
            gt_ingroup = self.mixer.ingroup.view(B, T)[:, :-1].mean()
        # One CUDA stream per network (args.concurrent_streams, default on): the online agent on the main stream, the target agent,
        # the online hypernetworks and the target hypernetworks each on their own.  The dense kernels size their grids to the
        # problem (a small shard does not occupy every SM), so independent networks really run side by side; for large
        # batches, where every kernel fills the GPU, what overlaps is each launch's fixed cost (tile prologue, pipeline fill
        # and drain, tail rounds).  Under CUDA-graph capture the forks and joins become the graph's parallel branches.
        two = bool(getattr(args, "concurrent_streams", True)) and self.mixer is not None
        main = torch.cuda.current_stream()
        n_h = len(self.mixer.nets) if self.mixer is not None else 0
        s_tgt = self._side_stream(0) if two else main
        s_hyp = [self._side_stream(1 + i) for i in range(n_h)] if two else None
        s_thyp = [self._side_stream(1 + n_h + i) for i in range(n_h)] if two else None
        side_all = ([s_tgt] + s_hyp + s_thyp) if two else []
        # shared inputs first (entities / last-action index / masks / packed fc1 input) and the mask plan, then fork
        T_all = batch["avail_actions"].shape[1]
        inp = self.mac._build_inputs(batch, slice(0, T_all))
        use_gt = getattr(args, "train_gt_factors", False)
        use_rgt = getattr(args, "train_rand_gt_factors", False)
        if self.imagine and group_bits is None and (self.mac.agent.rnn or not use_gt):   # the partition is drawn ONCE, here
            group_bits = self.mac.draw_groups(inp["bs"], inp["ne"], inp["entity_mask"].device)
        self.last_group_bits = group_bits        # under graph replay: the static tensor the captured draw writes (tests read it)
        _, mix, _ = self.mac.mask_plan(inp, self.imagine, use_gt, use_rgt, group_bits)
        for st in side_all:
            st.wait_stream(main)
        ents, la, em = inp["entities"], inp["last_action"], inp["entity_mask"]
        # ---- target agent + target hypernetworks (q_learner.py:111-118,154) ------------------------------------------
        with torch.cuda.stream(s_tgt):
            self.target_mac.init_hidden(B)
            q_tgt, _, _, _ = self.target_mac.forward(batch, None, ret_plan=True, inputs=inp)
        # ---- hypernetworks of the target and of the online mixer (they only need the entities and the partition) --------
        # args.group_hypernets: True / False, default "auto" = grouped launches while the shard is small (where a launch's fixed
        # cost dominates: 55 instead of 114 launches per step, 1.10 vs 1.12 ms at 16 episodes), per-network streams for a batch that
        # fills the GPU (6.25 vs 6.38 ms at 128 episodes: more independent chains to overlap)
        gh = getattr(args, "group_hypernets", "auto")
        grouped = n_h > 0 and (bool(gh) if gh != "auto" else N * inp["ne"] <= self.GROUP_MAX_ENTITY_ROWS)
        if grouped:
            # all of them in lock-step on ONE stream: every dense layer of the 2 x n_h networks is a single grouped launch
            with torch.cuda.stream(s_hyp[0] if two else main):
                jobs = self.target_mixer.hyper_jobs(em) + self.mixer.hyper_jobs(em, mix if self.imagine else None)
                outs = hypernets_forward_group(jobs, ents, la, T, inp.get("xin"))
                self.target_mixer.set_hyper_outputs(outs[:n_h], N, False)
                self.mixer.set_hyper_outputs(outs[n_h:], N, self.imagine)
        else:
            self.target_mixer.hyper_forward(ents, la, em, T, xin=inp.get("xin"), streams=s_thyp)
            self.mixer.hyper_forward(ents, la, em, T, imagine_masks=mix if self.imagine else None, xin=inp.get("xin"), streams=s_hyp)
        # ---- main: online agent on all T steps, 3 mask copies when imagining (q_learner.py:79-109) -------------------
        # the agent chain is the step's critical path: short kernels (few m-tiles per CTA) on a high-priority stream
        # (measured, profiles/r2_tuning.md: priority 1.107 -> 1.089 ms at 16 episodes, 4 m-tiles per CTA on the chain -> 1.080; at
        # 128 episodes everything within noise, so the tile hint is only given to small shards)
        small = N * inp["ne"] <= self.GROUP_MAX_ENTITY_ROWS
        crit_tiles = int(getattr(args, "critical_min_tiles", 4 if small else 0)) if two else 0
        crit = self._crit_stream() if (two and getattr(args, "critical_priority", True)) else None
        if crit is not None:
            crit.wait_stream(main)
        with torch.cuda.stream(crit if crit is not None else main), ops.min_tiles(crit_tiles):
            self.mac.init_hidden(B)
            q_all, spec, _, _ = self.mac.forward(batch, None, imagine=self.imagine, use_gt_factors=use_gt,
                                                 use_rand_gt_factors=use_rgt, group_bits=group_bits, train=True,
                                                 ret_plan=True, inputs=inp)
            C = spec.C
            chosen = ops.gather_chosen(q_all, actions, ws.get("chosen", (3, N, na)), C, N * na, A)
        if crit is not None:
            main.wait_stream(crit)
        for st in (s_hyp or []):
            main.wait_stream(st)
        qtot, qtot_im = self.mixer.mix(chosen[0], chosen[1] if self.imagine else None,
                                       chosen[2] if self.imagine else None, ret_ingroup=log_gt)
        ingroup = self.mixer.ingroup.view(B, T)[:, :-1].mean() if log_gt else None
        for st in ([s_tgt] + s_thyp if two else []):
            main.wait_stream(st)
        # double-Q selection and the target mix (q_learner.py:121-126,154)
        tgt_max = ops.target_max(q_all[0], q_tgt[0], avail, ws.get("tgt_max", (N, na)), None, N * na, A, args.double_q)
        tgt_tot, _ = self.target_mixer.mix(tgt_max, None, None)
        # targets, TD errors, masked losses as sums (q_learner.py:157-172)
        self.stats64.zero_()
        g_plain = ws.get("g_plain", (N,))
        g_im = ws.get("g_im", (N,)) if self.imagine else None
        ops.td_loss(qtot, qtot_im, tgt_tot, reward, terminated, filled, g_plain, g_im, None, self.stats64, B, T,
                    args.gamma, args.lmbda if self.imagine else 0.0)
        # backward (q_learner.py:175-176): hypernetworks on the side streams, agent on the main stream
        self.gradbuf.zero_()
        dq, dhyper = self.mixer.backward_mix(g_plain, g_im)
        for st in (s_hyp or []):
            st.wait_stream(main)
        # weight-gradient kernels leave the data-gradient chains: one companion stream per network
        s_w = [self._side_stream(1 + 2 * n_h + i) for i in range(n_h + 1)] if two else None
        if grouped:
            with torch.cuda.stream(s_hyp[0] if two else main):
                hypernets_backward_group(list(self.mixer.nets.values()), [dhyper[h] for h in self.mixer.nets],
                                         wstream=s_w[1] if two else None)
        else:
            self.mixer.backward_hyper(dhyper, streams=s_hyp, wstreams=s_w[1:] if two else None)
        Ap = self.mac.agent.dq_width(C * N * na)        # one-hot scatter of d(chosen) into (padded) action columns
        dQ = ops.scatter_dq(dq, actions, ws.get("dQ", (C * N * na, Ap)), C, N * na, Ap, T, na)
        if crit is not None:
            crit.wait_stream(main)
        with torch.cuda.stream(crit if crit is not None else main), ops.min_tiles(crit_tiles):
            self.mac.agent.backward(dQ, wstream=s_w[0] if two else None)
        if crit is not None:
            main.wait_stream(crit)
        for st in (s_hyp or []):
            main.wait_stream(st)
        # one all-reduce of [grads | stats], then normalise + clip + RMSprop (q_learner.py:177-178)
        ops.pack_stats(self.stats64, self.gradbuf[self.n_params:], N_STATS)
        if not reduce_and_update:
            return ingroup, gt_ingroup
        parallel.all_reduce_sum_(self.gradbuf)
        self._update_step()
        return ingroup, gt_ingroup

    def _update_step(self):
        args = self.args
        self.sumsq.zero_()
        ops.grad_sumsq(self.gradbuf, self.n_params, self.sumsq)
        ops.clip_rmsprop_step(self.flat, self.gradbuf, self.square_avg, self.n_params,
                              self.gradbuf[self.n_params:self.n_params + 1], self.sumsq, self.grad_norm,
                              args.grad_norm_clip, args.lr, args.optim_alpha, args.optim_eps,
                              getattr(args, "weight_decay", 0))

    def _update_targets(self):
        self.target_mac.load_state(self.mac)
        if self.mixer is not None:
            self.target_mixer.load_state_dict(self.mixer.state_dict())
        if self.logger is not None:
            self.logger.console_logger.info("Updated target network")

    # ---- checkpoints (file names and state_dict keys of q_learner.py:216-229) ---------------------------------------
    def _opt_state_dict(self):
        state, off = {}, 0
        stores = [self.mac.agent.store] + ([self.mixer.store] if self.mixer is not None else [])
        idx = 0
        for s in stores:
            o = off
            for k, shp in s.specs.items():
                n = 1
                for d in shp:
                    n *= d
                state[idx] = {"step": 0, "square_avg": self.square_avg[o:o + n].view(shp).detach().cpu().clone()}
                o += n
                idx += 1
            off += s.padded_size
        a = self.args
        group = {"lr": a.lr, "momentum": 0, "alpha": a.optim_alpha, "eps": a.optim_eps, "centered": False,
                 "weight_decay": getattr(a, "weight_decay", 0), "params": list(range(idx))}
        return {"state": state, "param_groups": [group]}

    def _load_opt_state_dict(self, sd):
        off, idx = 0, 0
        stores = [self.mac.agent.store] + ([self.mixer.store] if self.mixer is not None else [])
        for s in stores:
            o = off
            for k, shp in s.specs.items():
                n = 1
                for d in shp:
                    n *= d
                st = sd["state"].get(idx)
                if st is not None and "square_avg" in st:
                    self.square_avg[o:o + n].copy_(st["square_avg"].reshape(-1).to(self.device))
                o += n
                idx += 1
            off += s.padded_size

    def save_models(self, path):
        self.mac.save_models(path)
        if self.mixer is not None:
            torch.save({k: v.detach().cpu() for k, v in self.mixer.state_dict().items()}, os.path.join(path, "mixer.th"))
        torch.save(self._opt_state_dict(), os.path.join(path, "opt.th"))

    def load_models(self, path, evaluate=False):
        self.mac.load_models(path)
        self.target_mac.load_models(path)      # target nets are re-loaded from the online weights (q_learner.py:224-225)
        if not evaluate:
            if self.mixer is not None:
                sd = torch.load(os.path.join(path, "mixer.th"), map_location="cpu")
                self.mixer.load_state_dict(sd)
                self.target_mixer.load_state_dict(sd)
            self._load_opt_state_dict(torch.load(os.path.join(path, "opt.th"), map_location="cpu"))
