"""Entity-scheme controller: agents consume (entities, obs_mask, entity_mask[, gt_mask]).

API mirror of /root/reference/src/controllers/entity_controller.py:7-36.  The concat of entities with the one-hot of the
previous action (:14-27) is never materialised: `_build_inputs` produces an int32 last-action index per entity row and the
embed kernel gathers the matching column of fc1."""
import torch

from .. import ops
from ..modules.nets import MaskSpec
from .basic_controller import BasicMAC


def _logical_or(a, b):
    return ((a.to(torch.int32) + b.to(torch.int32)) > 0).to(torch.uint8)


class EntityMAC(BasicMAC):
    def __init__(self, scheme, groups, args):
        super().__init__(scheme, groups, args)

    def _get_input_shape(self, scheme):
        shape = scheme["entities"]["vshape"]
        shape = shape if isinstance(shape, int) else shape[0]
        if self.args.entity_last_action:
            a = scheme["actions_onehot"]["vshape"]
            shape += a if isinstance(a, int) else a[0]
        return shape

    # ---- inputs ---------------------------------------------------------------------------------------------
    def _build_inputs(self, batch, t):
        """t: slice over time.  -> dict of contiguous device tensors with rows (b, t)."""
        args = self.args
        ents = batch["entities"][:, t]
        bs, ts, ne = ents.shape[0], ents.shape[1], ents.shape[2]
        inp = {"bs": bs, "ts": ts, "ne": ne}
        inp["entities"] = ents.contiguous().view(bs * ts, ne, -1)
        la = None
        if args.entity_last_action:
            acts = batch["actions"]
            if t.start == 0 and t.stop == acts.shape[1] and acts.is_contiguous():
                # whole sequence (the learner): one kernel -- la[b, t, e] = actions[b, t-1, e] for agents at t > 0, else -1
                la = ops.last_action_index(acts, torch.empty((bs, ts, ne), dtype=torch.int32, device=ents.device), ne)
            else:                          # an acting slice of a few rows: index plumbing with torch ops
                la = torch.full((bs, ts, ne), -1, dtype=torch.int32, device=ents.device)
                if t.start == 0:
                    if t.stop > 1:
                        la[:, 1:, :args.n_agents] = acts[:, 0:t.stop - 1, :, 0].to(torch.int32)
                else:
                    la[:, :, :args.n_agents] = acts[:, t.start - 1:t.stop - 1, :, 0].to(torch.int32)
            la = la.view(bs * ts, ne)
        inp["last_action"] = la
        inp["obs_mask"] = batch["obs_mask"][:, t].contiguous().view(bs * ts, ne, ne)
        inp["entity_mask"] = batch["entity_mask"][:, t].contiguous().view(bs * ts, ne)
        if getattr(args, "gt_mask_avail", False):
            inp["gt_mask"] = batch["gt_mask"][:, t].contiguous().view(bs * ts, args.n_agents, ne)
        inp["xin"] = None
        if ops.USE_TENSOR_CORES and bs * ts * ne >= ops.TC_MIN_ROWS:
            # [entities | onehot(last action) | 0] padded to a multiple of 32 columns, shared by fc1 of every network
            kp = (self.agent.ein + 31) // 32 * 32
            xin = torch.empty(bs * ts * ne, kp, dtype=torch.float32, device=ents.device)
            inp["xin"] = ops.pack_inputs(inp["entities"], la, args.n_actions, xin)
        return inp

    def _fused_acting_ok(self, batch):
        """args.fused_acting (default on): FF agent with the Group Matching widths (d = 64, 4 heads, <= 8 entities, <= 32 input
        features) on contiguous device tensors -> the fused acting kernel; anything else takes the layer-by-layer path."""
        a = self.agent
        if a.rnn or a.pool is not None or not getattr(self.args, "fused_acting", True):
            return False
        ok = getattr(self, "_fused_ok", None)
        if ok is None:
            ok = self._fused_ok = ops.ff_agent_act_supported(int(self.args.n_entities), self.n_agents, a.ein, a.d, a.H, a.A)
        if not ok:
            return False
        keys = ["entities", "obs_mask", "entity_mask"] + (["gt_mask"] if getattr(self.args, "gt_obs_mask", False) else []) + \
               (["actions"] if a.one_hot_la else [])
        try:
            return all(batch[k].is_cuda and batch[k].is_contiguous() for k in keys)
        except (KeyError, ValueError):
            return False

    def act_and_select(self, batch, t, t_env, test_mode=False, est_flags=None, eps_dev=None, uniforms=None):
        """Acting step t in ONE launch: fused FF-agent forward + epsilon-greedy selection written to batch["actions"][:, t]
        (callers check _fused_acting_ok first).  Same rules as select_actions + EpsilonGreedyActionSelector.select_action."""
        a, sel = self.agent, self.action_selector
        ents = batch["entities"]
        if eps_dev is None:
            sel.epsilon = 0.0 if test_mode else sel.schedule.eval(t_env)
        eps = 0.0 if eps_dev is not None else sel.epsilon
        explore = eps_dev is not None or eps > 0.0
        if explore and uniforms is None:
            uniforms = torch.rand(2, ents.shape[0], self.n_agents, device=ents.device)
        gt_obs = bool(getattr(self.args, "gt_obs_mask", False))
        q = a.ws.get(a.tag + ".q_act", (ents.shape[0], self.n_agents, a.A))
        ops.ff_agent_act(ents, batch["actions"] if a.one_hot_la else None, a.A, batch["gt_mask"] if gt_obs else batch["obs_mask"],
                         batch["entity_mask"], a.store.p, q, t,
                         select=dict(avail=batch["avail_actions"], actions_out=batch["actions"], est_flags=est_flags, epsilon=eps,
                                     eps_dev=eps_dev, u_pick=uniforms[0] if explore else None,
                                     u_act=uniforms[1] if explore else None))
        return q

    def draw_groups(self, bs, ne, device):
        """Random 2-partition of the entities, one Bernoulli parameter per episode
        (entity_rnn_agent.py:94-96 / entity_ff_agent.py:87,96)."""
        probs = torch.rand(bs, 1, 1, device=device).repeat(1, 1, ne)
        return torch.bernoulli(probs).to(torch.uint8).view(bs, ne)

    def mask_plan(self, inp, imagine=False, use_gt_factors=False, use_rand_gt_factors=False, group_bits=None):
        """-> (agent MaskSpec, mixer imagine copies or None, explicit (W, I) 'noobs' masks or None)."""
        args = self.args
        ne, na, ts = inp["ne"], args.n_agents, inp["ts"]
        is_rnn = self.agent.rnn
        obs = inp["obs_mask"]
        if (not is_rnn) and getattr(args, "gt_obs_mask", False):      # entity_ff_agent.py:34-35
            obs, obs_stride = inp["gt_mask"], na * ne
        else:
            obs_stride = ne * ne
        em = inp["entity_mask"]
        plain = (obs, obs_stride, 0)
        if not imagine:
            return MaskSpec([plain], None, em), None, None
        # the RNN agent ignores use_gt_factors / use_rand_gt_factors (entity_rnn_agent.py:83 takes **kwargs) and always
        # draws a random partition; the FF agent skips the draw only for the pure ground-truth factorisation
        if group_bits is None and (is_rnn or not use_gt_factors):
            group_bits = self.draw_groups(inp["bs"], ne, em.device)
        if is_rnn or not (use_gt_factors or use_rand_gt_factors):
            spec = MaskSpec([plain, (obs, obs_stride, ops.ATTN_PART_WITHIN), (obs, obs_stride, ops.ATTN_PART_INTERACT)],
                            group_bits, em)
            mix = ([(None, 0, ops.ATTN_PART_WITHIN | ops.ATTN_ACTIVE0), (None, 0, ops.ATTN_PART_INTERACT | ops.ATTN_ACTIVE0)],
                   group_bits)
            return spec, mix, None
        # ground-truth factor modes of the FF agent (entity_ff_agent.py:93-114): masks become explicit tensors
        gt = inp["gt_mask"].view(inp["bs"], ts, na, ne)
        ina0 = em.view(inp["bs"], ts, ne)[:, 0].bool()
        active = (ina0[:, :na].unsqueeze(2) | ina0.unsqueeze(1)).to(torch.uint8).unsqueeze(1)      # (bs,1,na,ne)
        if use_gt_factors:
            within = gt
        else:
            g = group_bits.bool()
            in_a, in_b = (~g) & (~ina0), g & (~ina0)
            same = (in_a[:, :na].unsqueeze(2) & in_a.unsqueeze(1)) | (in_b[:, :na].unsqueeze(2) & in_b.unsqueeze(1))
            within = _logical_or((~same).to(torch.uint8).unsqueeze(1), gt)
        interact = 1 - within
        w_noobs = _logical_or(within, active).expand(inp["bs"], ts, na, ne).contiguous().view(-1, na, ne)
        i_noobs = _logical_or(interact, active).expand(inp["bs"], ts, na, ne).contiguous().view(-1, na, ne)
        obs4 = obs.view(inp["bs"], ts, -1, ne)[:, :, :na]
        w_obs = _logical_or(within, obs4).contiguous().view(-1, na, ne)
        i_obs = _logical_or(interact, obs4).contiguous().view(-1, na, ne)
        spec = MaskSpec([plain, (w_obs, na * ne, 0), (i_obs, na * ne, 0)], None, em)
        mix = ([(w_noobs, na * ne, 0), (i_noobs, na * ne, 0)], None)
        return spec, mix, (w_noobs, i_noobs)

    # ---- forward --------------------------------------------------------------------------------------------
    def forward(self, ep_batch, t, test_mode=False, imagine=False, use_gt_factors=False, use_rand_gt_factors=False,
                group_bits=None, train=False, ret_plan=False, inputs=None):
        """t=int: one acting step -> (bs, na, A), hidden state carried.  t=None: whole sequence -> (bs, T, na, A), or with
        imagine=True ((3 bs, T, na, A), (Wmask, Imask)) exactly as basic_controller.py:28-67."""
        int_t = isinstance(t, int)
        if int_t and not imagine and inputs is None and not ret_plan and self._fused_acting_ok(ep_batch):
            # acting step of the small FF agent: ONE fused kernel reading the rollout tensors in place (csrc/ffact.cu)
            a = self.agent
            ents = ep_batch["entities"]
            gt_obs = bool(getattr(self.args, "gt_obs_mask", False))
            q = a.ws.get(a.tag + ".q_act", (ents.shape[0], self.n_agents, a.A))
            ops.ff_agent_act(ents, ep_batch["actions"] if a.one_hot_la else None, a.A,
                             ep_batch["gt_mask"] if gt_obs else ep_batch["obs_mask"], ep_batch["entity_mask"], a.store.p, q, t)
            return q
        T_all = ep_batch["avail_actions"].shape[1]
        sl = slice(t, t + 1) if int_t else (slice(0, T_all) if t is None else t)
        inp = inputs if inputs is not None else self._build_inputs(ep_batch, sl)
        bs, ts, ne, na = inp["bs"], inp["ts"], inp["ne"], self.n_agents
        spec, mix, explicit = self.mask_plan(inp, imagine, use_gt_factors, use_rand_gt_factors, group_bits)
        C = spec.C
        h0 = None
        if self.agent.rnn and self.hidden_states is not None:
            h0 = self.hidden_states.reshape(-1, self.agent.r)
            if h0.shape[0] == bs * na and C > 1:
                h0 = h0.repeat(C, 1)
            h0 = h0.contiguous()
        q, hs = self.agent.forward(inp["entities"], inp["last_action"], spec, bs, ts, h0=h0, train=train,
                                   xin=inp.get("xin"))
        if self.agent.rnn:
            # entity_rnn_agent.py:64 returns the whole stack; only the last step is ever carried forward
            self.hidden_states = hs.view(C * bs, ts, na, self.agent.r)[:, -1].clone()
        if ret_plan:
            return q, spec, mix, inp
        out = q.view(C * bs, ts, na, -1)
        if int_t:
            return out[:, 0]
        if imagine:
            return out, self._imagine_groups(inp, spec, explicit)
        return out

    def _imagine_groups(self, inp, spec, explicit):
        """Explicit (Wattnmask_noobs, Iattnmask_noobs) tensors for API parity with the reference return value."""
        bs, ts, ne, na = inp["bs"], inp["ts"], inp["ne"], self.n_agents
        if explicit is not None:
            return explicit[0].view(bs, ts, na, ne), explicit[1].view(bs, ts, na, ne)
        rows = ne if self.agent.rnn else na
        g = spec.group_bits.bool()
        ina0 = inp["entity_mask"].view(bs, ts, ne)[:, 0].bool()
        in_a, in_b = (~g) & (~ina0), g & (~ina0)
        same = (in_a[:, :rows].unsqueeze(2) & in_a.unsqueeze(1)) | (in_b[:, :rows].unsqueeze(2) & in_b.unsqueeze(1))
        active = ina0[:, :rows].unsqueeze(2) | ina0.unsqueeze(1)
        w = ((~same) | active).to(torch.uint8).unsqueeze(1).repeat(1, ts, 1, 1)
        i = (same | active).to(torch.uint8).unsqueeze(1).repeat(1, ts, 1, 1)
        return w, i
