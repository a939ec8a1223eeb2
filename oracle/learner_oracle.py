"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not shipped, not on the product path.

Functional torch-CPU (fp32) restatement of the reference learner hot path.  Each function cites the
reference file:line it follows (paths relative to /root/reference/src).  Parameters are plain dicts keyed
by the reference ``state_dict`` names (``fc1.weight``, ``attn.in_trans.weight`` ...), batches are plain
dicts of tensors with the EpisodeBatch layout of run.py:178-196.

Parity is PINNED: tests/test_oracle_learner.py checks every function here against fixtures produced by
the reference's own modules (EntityMAC + QLearner.train) run in the build container
(tests/golden/make_golden.py -> tests/golden/learner_*.npz).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import math
from types import SimpleNamespace

import torch
import torch.nn.functional as F

NEG_UNAVAIL = -9999999.0  # q_learner.py:118,124


# --------------------------------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------------------------------
def entity_attention(x, w_in, w_out, b_out, pre_mask, post_mask, n_heads):
    """modules/layers/attention.py:24-79.

    x (N, ne, d) ; pre_mask (N, >=nq, ne) 1=masked ; post_mask (N, nq) 1=masked.  Returns (N, nq, d_out).
    QKV = x W_in^T chunked Q|K|V; heads are contiguous hd slices; logits / sqrt(hd); -inf fill;
    softmax over entities with NaN rows (all masked) -> 0; concat heads; out_trans; zero post-masked rows.
    """
    N, ne, d_in = x.shape
    d = w_in.shape[0] // 3
    hd = d // n_heads
    nq = post_mask.shape[1]
    qkv = x @ w_in.t()
    q, k, v = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
    q = q[:, :nq].reshape(N, nq, n_heads, hd).permute(0, 2, 1, 3)
    k = k.reshape(N, ne, n_heads, hd).permute(0, 2, 1, 3)
    v = v.reshape(N, ne, n_heads, hd).permute(0, 2, 1, 3)
    logits = (q @ k.transpose(-1, -2)) / torch.tensor(float(hd)).sqrt()
    m = pre_mask[:, :nq].bool().unsqueeze(1)
    logits = logits.masked_fill(m, -float("inf"))
    w = torch.softmax(logits, dim=-1)
    w = torch.where(torch.isnan(w), torch.zeros_like(w), w)
    o = (w @ v).permute(0, 2, 1, 3).reshape(N, nq, d)
    o = o @ w_out.t() + b_out
    return o.masked_fill(post_mask.bool().unsqueeze(-1), 0.0)


def entity_pooling(x, w_in, b_in, w_out, b_out, pre_mask, post_mask, pooling_type):
    """EntityPoolingLayer.forward, modules/layers/attention.py:95-132: in_trans (with bias), masked entities set to ZERO (they
    still count in the mean's divisor and take part in the max), pool over entities per agent, out_trans, post-mask."""
    nq = post_mask.shape[1]
    e = x @ w_in.t() + b_in
    rep = e.unsqueeze(1).repeat(1, nq, 1, 1).masked_fill(pre_mask[:, :nq].bool().unsqueeze(3), 0.0)
    pooled = rep.max(dim=2)[0] if pooling_type == "max" else rep.mean(dim=2)
    o = pooled @ w_out.t() + b_out
    return o.masked_fill(post_mask.bool().unsqueeze(-1), 0.0)


def entity_layer(p, args, x1, pre_mask, post_mask):
    """`self.attn` of the agents / hypernetworks: attention unless args.pooling_type is set (entity_rnn_agent.py:13-22)."""
    if getattr(args, "pooling_type", None) is None:
        return entity_attention(x1, p["attn.in_trans.weight"], p["attn.out_trans.weight"], p["attn.out_trans.bias"],
                                pre_mask, post_mask, args.attn_n_heads)
    return entity_pooling(x1, p["attn.in_trans.weight"], p["attn.in_trans.bias"], p["attn.out_trans.weight"],
                          p["attn.out_trans.bias"], pre_mask, post_mask, args.pooling_type)


def _sub(p, prefix):
    return {k[len(prefix):]: v for k, v in p.items() if k.startswith(prefix)}


# --------------------------------------------------------------------------------------------------
# agents
# --------------------------------------------------------------------------------------------------
def gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    """torch.nn.GRUCell semantics (gate order r, z, n) used at modules/agents/entity_rnn_agent.py:24,53."""
    gi = x @ w_ih.t() + b_ih
    gh = h @ w_hh.t() + b_hh
    i_r, i_z, i_n = gi.chunk(3, dim=1)
    h_r, h_z, h_n = gh.chunk(3, dim=1)
    r = torch.sigmoid(i_r + h_r)
    z = torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    return (1.0 - z) * n + z * h


def rnn_agent_forward(p, args, entities, obs_mask, entity_mask, h0):
    """EntityAttentionRNNAgent.forward, modules/agents/entity_rnn_agent.py:31-64.

    entities (bs,ts,ne,ein) ; obs_mask (bs,ts,ne,ne) ; entity_mask (bs,ts,ne) ; h0 (bs,na,r).
    returns q (bs,ts,na,A), hs (bs,ts,na,r)
    """
    bs, ts, ne, ed = entities.shape
    na, r = args.n_agents, args.rnn_hidden_dim
    e = entities.reshape(bs * ts, ne, ed)
    om = obs_mask.reshape(bs * ts, ne, ne)
    em = entity_mask.reshape(bs * ts, ne)
    am = em[:, :na]
    x1 = F.relu(e @ p["fc1.weight"].t() + p["fc1.bias"])
    x2 = entity_layer(p, args, x1, om, am)
    x3 = F.relu(x2 @ p["fc2.weight"].t() + p["fc2.bias"]).reshape(bs, ts, na, r)
    h = h0.reshape(-1, r)
    hs = []
    for t in range(ts):
        h = gru_cell(x3[:, t].reshape(-1, r), h, p["rnn.weight_ih"], p["rnn.weight_hh"],
                     p["rnn.bias_ih"], p["rnn.bias_hh"])
        hs.append(h.reshape(bs, na, r))
    hs = torch.stack(hs, dim=1)
    q = hs @ p["fc3.weight"].t() + p["fc3.bias"]
    q = q.masked_fill(am.reshape(bs, ts, na, 1).bool(), 0.0)
    return q, hs


def ff_agent_forward(p, args, entities, obs_mask, entity_mask, gt_mask=None):
    """EntityAttentionFFAgent.forward, modules/agents/entity_ff_agent.py:30-57 (relu AFTER the attention
    layer's out_trans + post-mask; ``gt_obs_mask`` swaps obs_mask for gt_mask)."""
    if getattr(args, "gt_obs_mask", False):
        obs_mask = gt_mask
    bs, ts, ne, ed = entities.shape
    na = args.n_agents
    e = entities.reshape(bs * ts, ne, ed)
    om = obs_mask.reshape(bs * ts, obs_mask.shape[2], ne)
    em = entity_mask.reshape(bs * ts, ne)
    am = em[:, :na]
    x1 = F.relu(e @ p["fc1.weight"].t() + p["fc1.bias"])
    x2 = F.relu(entity_layer(p, args, x1, om, am))
    q = x2 @ p["fc2.weight"].t() + p["fc2.bias"]
    q = q.reshape(bs, ts, na, -1).masked_fill(am.reshape(bs, ts, na, 1).bool(), 0.0)
    return q, x2


def imagine_masks(group_a, inactive0, n_rows):
    """Closed form of the random-partition masks (entity_rnn_agent.py:79-114, entity_ff_agent.py:72-118).

    group_a (bs, ne) 0/1 Bernoulli draw BEFORE OR-ing with the inactive mask; inactive0 (bs, ne) =
    entity_mask[:, 0]; n_rows = ne (RNN agent) or na (FF agent).
    Returns dict of (bs, n_rows, ne) uint8 masks, 1 = masked:
      within        = not(same group and both active)              (agent 'W' copy before OR obs_mask)
      interact      = same group and both active                   (agent 'I' copy before OR obs_mask)
      within_noobs  = within   OR inact_i OR inact_j  (mixer W)    == within
      interact_noobs= interact OR inact_i OR inact_j  (mixer I)
    """
    g = group_a.bool()
    ina = inactive0.bool()
    in_a = (~g) & (~ina)          # member of group A and present
    in_b = g & (~ina)             # member of group B and present
    ra, rb, rin = in_a[:, :n_rows], in_b[:, :n_rows], ina[:, :n_rows]
    interact = (ra.unsqueeze(2) & in_a.unsqueeze(1)) | (rb.unsqueeze(2) & in_b.unsqueeze(1))
    within = ~interact
    active = rin.unsqueeze(2) | ina.unsqueeze(1)
    return {
        "within": within.to(torch.uint8),
        "interact": interact.to(torch.uint8),
        "within_noobs": (within | active).to(torch.uint8),
        "interact_noobs": (interact | active).to(torch.uint8),
    }


def build_agent_inputs(batch, args):
    """EntityMAC._build_inputs with t=None, controllers/entity_controller.py:11-30."""
    ents = batch["entities"]
    bs, T, ne, _ = ents.shape
    if args.entity_last_action:
        la = torch.zeros(bs, T, ne, args.n_actions, dtype=ents.dtype)
        la[:, 1:, :args.n_agents] = batch["actions_onehot"][:, :-1]
        ents = torch.cat([ents, la], dim=3)
    return ents, batch["obs_mask"], batch["entity_mask"]


def agent_forward(p, args, batch, imagine=False, group_a=None, use_gt_factors=False, use_rand_gt_factors=False):
    """BasicMAC.forward(t=None[, imagine=True]) -> agent (controllers/basic_controller.py:28-67) for the four
    entity agents of modules/agents/__init__.py.  Hidden state starts at zero (basic_controller.py:69-70).
    Returns q (bs or 3bs, T, na, A) and, when imagine, the mixer masks (W_noobs, I_noobs)."""
    ents, obs_mask, entity_mask = build_agent_inputs(batch, args)
    bs, T, ne, _ = ents.shape
    na = args.n_agents
    is_rnn = "rnn" in args.agent
    h0 = torch.zeros(bs, na, args.rnn_hidden_dim)
    gt_mask = batch.get("gt_mask", None)
    if not imagine:
        if is_rnn:
            return rnn_agent_forward(p, args, ents, obs_mask, entity_mask, h0)[0]
        return ff_agent_forward(p, args, ents, obs_mask, entity_mask, gt_mask)[0]
    n_rows = ne if is_rnn else na
    ina0 = entity_mask[:, 0]
    rep_t = T
    if (not is_rnn) and use_gt_factors:          # entity_ff_agent.py:93-95
        within = gt_mask
        interact = 1 - within
        active = (ina0[:, :na].bool().unsqueeze(2) | ina0.bool().unsqueeze(1)).to(torch.uint8).unsqueeze(1)
        rep_t = 1
    else:
        m = imagine_masks(group_a, ina0, n_rows)
        within, interact = m["within"].unsqueeze(1), m["interact"].unsqueeze(1)
        active = (ina0[:, :n_rows].bool().unsqueeze(2) | ina0.bool().unsqueeze(1)).to(torch.uint8).unsqueeze(1)
        if (not is_rnn) and use_rand_gt_factors:  # entity_ff_agent.py:111-114
            within = ((within + gt_mask) > 0).to(torch.uint8)
            interact = 1 - within
            rep_t = 1
    w_noobs = ((within + active) > 0).to(torch.uint8)
    i_noobs = ((interact + active) > 0).to(torch.uint8)
    om_rows = obs_mask[:, :, :n_rows] if is_rnn else obs_mask
    # FF agent: obs_mask is (bs,T,ne,ne) with ne == na (entity_ff_agent.py:121 only broadcasts when na == ne)
    w_obs = ((within + om_rows) > 0).to(torch.uint8)
    i_obs = ((interact + om_rows) > 0).to(torch.uint8)
    ents3 = ents.repeat(3, 1, 1, 1)
    om3 = torch.cat([om_rows, w_obs, i_obs], dim=0)
    em3 = entity_mask.repeat(3, 1, 1)
    if is_rnn:
        q = rnn_agent_forward(p, args, ents3, om3, em3, h0.repeat(3, 1, 1))[0]
    else:
        q = ff_agent_forward(p, SimpleNamespace(**{**vars(args), "gt_obs_mask": False}), ents3, om3, em3)[0]
    wm = w_noobs.expand(bs, w_noobs.shape[1], n_rows, ne).repeat(1, rep_t if w_noobs.shape[1] == 1 else 1, 1, 1)
    im = i_noobs.expand(bs, i_noobs.shape[1], n_rows, ne).repeat(1, rep_t if i_noobs.shape[1] == 1 else 1, 1, 1)
    return q, (wm, im)


# --------------------------------------------------------------------------------------------------
# mixers
# --------------------------------------------------------------------------------------------------
def hypernet(p, args, entities, entity_mask, attn_mask, mode):
    """AttentionHyperNet.forward, modules/mixers/flex_qmix.py:40-57.  entities (N,ne,ein), entity_mask (N,ne)."""
    na = args.n_agents
    x1 = F.relu(entities @ p["fc1.weight"].t() + p["fc1.bias"])
    am = entity_mask[:, :na]
    if attn_mask is None:
        attn_mask = (am.bool().unsqueeze(2) | entity_mask.bool().unsqueeze(1)).to(torch.uint8)
    x2 = entity_layer(p, args, x1, attn_mask, am)
    x3 = x2 @ p["fc2.weight"].t() + p["fc2.bias"]
    x3 = x3.masked_fill(am.bool().unsqueeze(2), 0.0)
    if mode == "vector":
        return x3.mean(dim=1)
    if mode == "alt_vector":
        return x3.mean(dim=2)
    if mode == "scalar":
        return x3.mean(dim=(1, 2))
    return x3


def mixer_inputs(batch, args):
    """QLearner._get_mixer_ins (entity scheme), learners/q_learner.py:45-64.  Returns ((ents[:-1], em[:-1]), (ents[1:], em[1:]))."""
    ents = batch["entities"]
    bs, T, ne, _ = ents.shape
    if args.entity_last_action:
        la = torch.zeros(bs, T, ne, args.n_actions, dtype=ents.dtype)
        la[:, 1:, :args.n_agents] = batch["actions_onehot"][:, :-1]
        ents = torch.cat([ents, la], dim=3)
    em = batch["entity_mask"]
    return (ents[:, :-1], em[:, :-1]), (ents[:, 1:], em[:, 1:])


def flex_mixer(p, args, agent_qs, entities, entity_mask, imagine_groups=None):
    """FlexQMixer.forward, modules/mixers/flex_qmix.py:79-121.  agent_qs (bs,t,na) or (bs,t,2na) when imagine."""
    bs, t, ne, ed = entities.shape
    na, me = args.n_agents, args.mixing_embed_dim
    e = entities.reshape(bs * t, ne, ed)
    em = entity_mask.reshape(bs * t, ne)
    hw1, hwf, hb1, hv = _sub(p, "hyper_w_1."), _sub(p, "hyper_w_final."), _sub(p, "hyper_b_1."), _sub(p, "V.")
    if imagine_groups is not None:
        qs = agent_qs.reshape(-1, 1, 2 * na)
        wm, im = imagine_groups
        w1 = torch.cat([hypernet(hw1, args, e, em, wm.reshape(bs * t, -1, ne), "matrix"),
                        hypernet(hw1, args, e, em, im.reshape(bs * t, -1, ne), "matrix")], dim=1)
    else:
        qs = agent_qs.reshape(-1, 1, na)
        w1 = hypernet(hw1, args, e, em, None, "matrix")
    b1 = hypernet(hb1, args, e, em, None, "vector").reshape(-1, 1, me)
    w1 = w1.reshape(bs * t, -1, me)
    w1 = torch.softmax(w1, dim=-1) if args.softmax_mixing_weights else w1.abs()
    pre = torch.bmm(qs, w1) + b1
    hidden = torch.tanh(pre) if getattr(args, "mixer_non_lin", "elu") == "tanh" else F.elu(pre)
    wf = hypernet(hwf, args, e, em, None, "vector")
    wf = torch.softmax(wf, dim=-1) if args.softmax_mixing_weights else wf.abs()
    v = hypernet(hv, args, e, em, None, "scalar").reshape(-1, 1, 1)
    y = torch.bmm(hidden, wf.reshape(-1, me, 1)) + v
    return y.reshape(bs, -1, 1)


def lin_flex_mixer(p, args, agent_qs, entities, entity_mask, imagine_groups=None, ret_ingroup_prop=False):
    """LinearFlexQMixer.forward, modules/mixers/flex_qmix.py:136-172."""
    bs, t, ne, ed = entities.shape
    na = args.n_agents
    e = entities.reshape(bs * t, ne, ed)
    em = entity_mask.reshape(bs * t, ne)
    hw1, hv = _sub(p, "hyper_w_1."), _sub(p, "V.")
    if imagine_groups is not None:
        qs = agent_qs.reshape(-1, 2 * na)
        wm, im = imagine_groups
        w1 = torch.cat([hypernet(hw1, args, e, em, wm.reshape(bs * t, na, ne), "alt_vector"),
                        hypernet(hw1, args, e, em, im.reshape(bs * t, na, ne), "alt_vector")], dim=1)
    else:
        qs = agent_qs.reshape(-1, na)
        w1 = hypernet(hw1, args, e, em, None, "alt_vector")
    w1 = w1.reshape(bs * t, -1)
    w1 = torch.softmax(w1, dim=1) if args.softmax_mixing_weights else w1.abs()
    v = hypernet(hv, args, e, em, None, "scalar")
    q_tot = ((qs * w1).sum(dim=1) + v).reshape(bs, -1, 1)
    if ret_ingroup_prop:
        return q_tot, w1[:, :na].sum(dim=1).mean()
    return q_tot


def mix(p, args, agent_qs, entities, entity_mask, imagine_groups=None):
    if args.mixer == "vdn":                      # modules/mixers/vdn.py:9-10
        return agent_qs.sum(dim=2, keepdim=True)
    if args.mixer == "flex_qmix":
        return flex_mixer(p, args, agent_qs, entities, entity_mask, imagine_groups)
    if args.mixer == "lin_flex_qmix":
        return lin_flex_mixer(p, args, agent_qs, entities, entity_mask, imagine_groups)
    raise ValueError(args.mixer)


# --------------------------------------------------------------------------------------------------
# learner
# --------------------------------------------------------------------------------------------------
def td_loss(agent_p, mixer_p, tgt_agent_p, tgt_mixer_p, batch, args, group_a=None):
    """QLearner.train up to the loss, learners/q_learner.py:66-172.  Returns (loss, aux dict)."""
    rewards = batch["reward"][:, :-1]
    actions = batch["actions"][:, :-1]
    terminated = batch["terminated"][:, :-1].float()
    mask = batch["filled"][:, :-1].float().clone()
    mask[:, 1:] = mask[:, 1:] * (1 - terminated[:, :-1])
    avail = batch["avail_actions"]
    imagine = "imagine" in args.agent
    aux = {}
    if imagine:
        all_out, groups = agent_forward(agent_p, args, batch, imagine=True, group_a=group_a,
                                        use_gt_factors=getattr(args, "train_gt_factors", False),
                                        use_rand_gt_factors=getattr(args, "train_rand_gt_factors", False))
        chosen_all = torch.gather(all_out[:, :-1], 3, actions.repeat(3, 1, 1, 1)).squeeze(3)
        mac_out = all_out.chunk(3, dim=0)[0]
        chosen, caq_w, caq_i = chosen_all.chunk(3, dim=0)
        caq_imagine = torch.cat([caq_w, caq_i], dim=2)
        aux["all_mac_out"] = all_out
    else:
        mac_out = agent_forward(agent_p, args, batch)
        chosen = torch.gather(mac_out[:, :-1], 3, actions).squeeze(3)
    with torch.no_grad():
        tgt_out = agent_forward(tgt_agent_p, args, batch)[:, 1:].clone()
        tgt_out[avail[:, 1:] == 0] = NEG_UNAVAIL
        if args.double_q:
            det = mac_out.clone().detach()
            det[avail == 0] = NEG_UNAVAIL
            cur_max = det[:, 1:].max(dim=3, keepdim=True)[1]
            tgt_max = torch.gather(tgt_out, 3, cur_max).squeeze(3)
            aux["cur_max_actions"] = cur_max.squeeze(3)
        else:
            tgt_max = tgt_out.max(dim=3)[0]
    (m_e, m_m), (t_e, t_m) = mixer_inputs(batch, args)
    q_tot = mix(mixer_p, args, chosen, m_e, m_m)
    if imagine:
        groups = [g[:, :-1] for g in groups]
        q_tot_im = mix(mixer_p, args, caq_imagine, m_e, m_m, imagine_groups=groups)
        if getattr(args, "test_gt_factors", False) and args.mixer == "lin_flex_qmix":
            # logging-only pass of q_learner.py:98-105,138-147: the share of the mixing weight that stays inside each agent's
            # own group, for the partition used in training and for the ground-truth factorisation
            with torch.no_grad():
                aux["ingroup_prop"] = lin_flex_mixer(mixer_p, args, caq_imagine, m_e, m_m, imagine_groups=groups,
                                                     ret_ingroup_prop=True)[1]
                gt_out, gt_groups = agent_forward(agent_p, args, batch, imagine=True, group_a=group_a, use_gt_factors=True)
                gt_chosen = torch.gather(gt_out[:, :-1], 3, actions.repeat(3, 1, 1, 1)).squeeze(3)
                _, gt_w, gt_i = gt_chosen.chunk(3, dim=0)
                gt_groups = [g[:, :-1] for g in gt_groups]
                aux["gt_ingroup_prop"] = lin_flex_mixer(mixer_p, args, torch.cat([gt_w, gt_i], dim=2), m_e, m_m,
                                                        imagine_groups=gt_groups, ret_ingroup_prop=True)[1]
    with torch.no_grad():
        tgt_tot = mix(tgt_mixer_p, args, tgt_max, t_e, t_m)
    targets = rewards + args.gamma * (1 - terminated) * tgt_tot
    td = q_tot - targets.detach()
    mask = mask.expand_as(td)
    masked_td = td * mask
    loss = (masked_td ** 2).sum() / mask.sum()
    aux.update(mac_out=mac_out, chosen=chosen, q_tot=q_tot, targets=targets, mask=mask, td_loss=loss)
    if imagine:
        im_td = (q_tot_im - targets.detach()) * mask
        im_loss = (im_td ** 2).sum() / mask.sum()
        aux.update(q_tot_im=q_tot_im, im_loss=im_loss)
        loss = (1 - args.lmbda) * loss + args.lmbda * im_loss
    aux["td_error_abs"] = masked_td.abs().sum() / mask.sum()
    aux["q_taken_mean"] = (q_tot * mask).sum() / (mask.sum() * args.n_agents)
    aux["target_mean"] = (targets * mask).sum() / (mask.sum() * args.n_agents)
    return loss, aux


def param_order(agent_p, mixer_p):
    """QLearner.params = list(mac.parameters()) + list(mixer.parameters()) (q_learner.py:16,35)."""
    return [("agent", k) for k in agent_p] + [("mixer", k) for k in mixer_p]


def train_step(agent_p, mixer_p, tgt_agent_p, tgt_mixer_p, batch, args, group_a=None, opt_state=None):
    """One QLearner.train optimisation step (q_learner.py:174-178): backward, clip_grad_norm_(params, clip),
    RMSprop(lr, alpha, eps) as torch.optim.RMSprop (centered=False, momentum=0, weight_decay=args.weight_decay):
        v <- alpha v + (1-alpha) g^2 ;  p <- p - lr g / (sqrt(v) + eps)
    Returns dict(loss, aux, grads, grad_norm, new_agent, new_mixer, opt_state)."""
    ap = {k: v.detach().clone().requires_grad_(True) for k, v in agent_p.items()}
    mp = {k: v.detach().clone().requires_grad_(True) for k, v in mixer_p.items()}
    loss, aux = td_loss(ap, mp, tgt_agent_p, tgt_mixer_p, batch, args, group_a)
    leaves = list(ap.values()) + list(mp.values())
    grads = torch.autograd.grad(loss, leaves, allow_unused=True)
    grads = [torch.zeros_like(l) if g is None else g for g, l in zip(grads, leaves)]
    total = torch.norm(torch.stack([torch.norm(g, 2.0) for g in grads]), 2.0)   # clip_grad_norm_ order of ops
    coef = torch.clamp(args.grad_norm_clip / (total + 1e-6), max=1.0)
    grads = [g * coef for g in grads]
    if opt_state is None:
        opt_state = [torch.zeros_like(l) for l in leaves]
    new_state, new_leaves = [], []
    for l, g, v in zip(leaves, grads, opt_state):
        if getattr(args, "weight_decay", 0) != 0:
            g = g + args.weight_decay * l.detach()
        v = args.optim_alpha * v + (1 - args.optim_alpha) * g * g
        new_leaves.append(l.detach() - args.lr * g / (v.sqrt() + args.optim_eps))
        new_state.append(v)
    keys_a, keys_m = list(ap.keys()), list(mp.keys())
    return dict(loss=loss.detach(), aux={k: (v.detach() if torch.is_tensor(v) else v) for k, v in aux.items()},
                grads_agent=dict(zip(keys_a, grads[:len(keys_a)])),
                grads_mixer=dict(zip(keys_m, grads[len(keys_a):])),
                grad_norm=total,
                new_agent=dict(zip(keys_a, new_leaves[:len(keys_a)])),
                new_mixer=dict(zip(keys_m, new_leaves[len(keys_a):])),
                opt_state=new_state)


def greedy_actions(q, avail):
    """EpsilonGreedyActionSelector.select_action with test_mode=True (components/action_selectors.py:45-63):
    unavailable -> -inf, first max index."""
    mq = q.clone()
    mq[avail == 0] = -float("inf")
    return mq.max(dim=2)[1]


def epsilon(args, t_env):
    """DecayThenFlatSchedule(linear).eval, components/epsilon_schedules.py:20-22."""
    delta = (args.epsilon_start - args.epsilon_finish) / args.epsilon_anneal_time
    return max(args.epsilon_finish, args.epsilon_start - delta * t_env)


def default_args(**over):
    """Hyper-parameters of config/default.yaml + config/algs/refil.yaml restated as a namespace."""
    d = dict(gamma=0.99, lr=0.0005, optim_alpha=0.99, optim_eps=0.00001, grad_norm_clip=10, weight_decay=0,
             double_q=True, lmbda=0.5, attn_n_heads=4, attn_embed_dim=128, hypernet_embed=128,
             mixing_embed_dim=32, rnn_hidden_dim=64, softmax_mixing_weights=True, entity_last_action=True,
             agent="imagine_entity_attend_rnn", mixer="flex_qmix", gt_obs_mask=False, train_gt_factors=False,
             train_rand_gt_factors=False, test_gt_factors=False, pooling_type=None,
             epsilon_start=1.0, epsilon_finish=0.05, epsilon_anneal_time=500000)
    d.update(over)
    return SimpleNamespace(**d)


# --------------------------------------------------------------------------------------------------
# synthetic data (SURVEY.md §8d config-3 recipe) and default-style init, shared by tests and bench
# --------------------------------------------------------------------------------------------------
def init_linear(gen, out_f, in_f, bias=True):
    """nn.Linear default init bounds (kaiming_uniform a=sqrt(5) -> U(-1/sqrt(in), 1/sqrt(in)))."""
    b = 1.0 / math.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=gen) * 2 - 1) * b
    if not bias:
        return w, None
    return w, (torch.rand(out_f, generator=gen) * 2 - 1) * b


def init_agent_params(gen, args, ein):
    d, r, A = args.attn_embed_dim, args.rnn_hidden_dim, args.n_actions
    p = {}
    p["fc1.weight"], p["fc1.bias"] = init_linear(gen, d, ein)
    if getattr(args, "pooling_type", None) is None:
        p["attn.in_trans.weight"], _ = init_linear(gen, 3 * d, d, bias=False)
    else:
        p["attn.in_trans.weight"], p["attn.in_trans.bias"] = init_linear(gen, d, d)
    p["attn.out_trans.weight"], p["attn.out_trans.bias"] = init_linear(gen, d, d)
    if "rnn" in args.agent:
        p["fc2.weight"], p["fc2.bias"] = init_linear(gen, r, d)
        b = 1.0 / math.sqrt(r)
        p["rnn.weight_ih"] = (torch.rand(3 * r, r, generator=gen) * 2 - 1) * b
        p["rnn.weight_hh"] = (torch.rand(3 * r, r, generator=gen) * 2 - 1) * b
        p["rnn.bias_ih"] = (torch.rand(3 * r, generator=gen) * 2 - 1) * b
        p["rnn.bias_hh"] = (torch.rand(3 * r, generator=gen) * 2 - 1) * b
        p["fc3.weight"], p["fc3.bias"] = init_linear(gen, A, r)
    else:
        p["fc2.weight"], p["fc2.bias"] = init_linear(gen, A, d)
    return p


def init_hypernet_params(gen, args, ein, prefix):
    he, me = args.hypernet_embed, args.mixing_embed_dim
    p = {}
    p[prefix + "fc1.weight"], p[prefix + "fc1.bias"] = init_linear(gen, he, ein)
    if getattr(args, "pooling_type", None) is None:
        p[prefix + "attn.in_trans.weight"], _ = init_linear(gen, 3 * he, he, bias=False)
    else:
        p[prefix + "attn.in_trans.weight"], p[prefix + "attn.in_trans.bias"] = init_linear(gen, he, he)
    p[prefix + "attn.out_trans.weight"], p[prefix + "attn.out_trans.bias"] = init_linear(gen, he, he)
    p[prefix + "fc2.weight"], p[prefix + "fc2.bias"] = init_linear(gen, me, he)
    return p


def init_mixer_params(gen, args, ein):
    p = {}
    if args.mixer == "flex_qmix":
        names = ["hyper_w_1.", "hyper_w_final.", "hyper_b_1.", "V."]
    elif args.mixer == "lin_flex_qmix":
        names = ["hyper_w_1.", "V."]
    else:
        names = []
    for n in names:
        p.update(init_hypernet_params(gen, args, ein, n))
    return p


def synthetic_batch(gen, B, T, na, ne, ed, A, gt_mask=False, pad=True):
    """Synthetic replay of the EpisodeBatch layout."""
    ents = torch.rand(B, T, ne, ed, generator=gen)
    obs = (torch.rand(B, T, ne, ne, generator=gen) < 0.2).to(torch.uint8)
    idx = torch.arange(ne)
    obs[:, :, idx, idx] = 0
    em = torch.zeros(B, T, ne, dtype=torch.uint8)
    if pad:
        for b in range(B):
            ka = int(torch.randint(max(1, na - 2), na + 1, (1,), generator=gen))
            em[b, :, ka:na] = 1
            if ne > na:
                ke = int(torch.randint(max(1, ne - na - 3), ne - na + 1, (1,), generator=gen))
                em[b, :, na + ke:] = 1
    # padded entities are never observable (SC2 semantics): mask their rows/cols
    obs = ((obs + em.unsqueeze(2) + em.unsqueeze(3)) > 0).to(torch.uint8)
    ents = ents * (1 - em).unsqueeze(-1).float()
    avail = (torch.rand(B, T, na, A, generator=gen) < 0.7).to(torch.int32)
    avail[..., min(1, A - 1)] = 1
    prob = avail.float() + 1e-9
    actions = torch.multinomial(prob.reshape(-1, A), 1, generator=gen).reshape(B, T, na, 1)
    onehot = torch.zeros(B, T, na, A).scatter_(3, actions, 1.0)
    reward = torch.randn(B, T, 1, generator=gen)
    term = torch.zeros(B, T, 1, dtype=torch.uint8)
    filled = torch.zeros(B, T, 1, dtype=torch.int64)
    for b in range(B):
        L = int(torch.randint(max(1, T // 2), T, (1,), generator=gen))   # env steps in the episode, <= T-1
        filled[b, :L + 1] = 1
        if bool(torch.rand(1, generator=gen) < 0.7):                     # else ended by the time limit
            term[b, L - 1] = 1
    batch = dict(entities=ents, obs_mask=obs, entity_mask=em, actions=actions, actions_onehot=onehot,
                 avail_actions=avail, reward=reward, terminated=term, filled=filled)
    if gt_mask:
        batch["gt_mask"] = (torch.rand(B, T, na, ne, generator=gen) < 0.5).to(torch.uint8)
    return batch
