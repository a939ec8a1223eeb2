"""Entity-attention networks of the learner hot path, as explicit forward / hand-written backward schedules over the
sm_100a kernels (no autograd on the product path).

Mirrors (state_dict keys included):
  /root/reference/src/modules/layers/attention.py:6-79          EntityAttentionLayer  -> AttnTrunk (fc1 + in_trans + MHA + out_trans)
  /root/reference/src/modules/agents/entity_rnn_agent.py:7-126  (Imagine)EntityAttentionRNNAgent -> EntityAttnAgent(rnn=True)
  /root/reference/src/modules/agents/entity_ff_agent.py:7-135   (Imagine)EntityAttentionFFAgent  -> EntityAttnAgent(rnn=False)
  /root/reference/src/modules/mixers/flex_qmix.py:7-57          AttentionHyperNet -> AttnHyperNet

Row conventions: N = B*T rows of (b, t); entity rows [N*ne, .]; agent rows of a copy stack [C*N*na, .] ordered
(copy, b, t, agent).  The reference's 3x `repeat` of the inputs for the imagine copies never happens: fc1 and in_trans
run once, the attention kernel resolves the three masks against one QKV tile.
"""
from collections import OrderedDict

import torch

from .. import ops
from .params import ParamStore, Workspace, init_linear_, init_uniform_


class MaskSpec:
    """Attention masks of one forward: up to 3 copies of (explicit u8 mask or None, stride per (b,t) row, mode bits),
    plus the per-episode random group bits and the entity mask the closed-form modes need."""

    def __init__(self, copies, group_bits=None, entity_mask=None):
        self.copies = list(copies)
        self.group_bits = group_bits
        self.entity_mask = entity_mask

    @property
    def C(self):
        return len(self.copies)


def _on_side(wstream, fn):
    """Enqueue fn() on `wstream` behind everything already enqueued on the current stream (None: run it inline).  Used for the
    weight-gradient kernels of a backward pass: they read what the data-gradient chain has produced so far but nothing
    downstream waits for them, so they leave the chain's critical path; the caller joins `wstream` when the pass is done."""
    if wstream is None:
        fn()
        return
    wstream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(wstream):
        fn()


class AttnTrunk:
    """x1 = relu(fc1([ents | onehot(last action)])); QKV = in_trans(x1); masked MHA for C mask copies;
    x2 = [relu] out_trans(.) with inactive-agent rows zeroed."""

    def __init__(self, store, prefix, ws, tag, ein, d, n_heads, n_agents, n_actions, pooling_type=None):
        self.s, self.pre, self.ws, self.tag = store, prefix, ws, tag
        self.ein, self.d, self.H, self.na, self.A = ein, d, n_heads, n_agents, n_actions
        self.pool = pooling_type      # None: EntityAttentionLayer; "mean" / "max": EntityPoolingLayer (attention.py:82-132)
        self.scratch = "scratch"      # tag prefix of the backward scratch buffers: nets that run concurrently use distinct ones
        if self.pool is None:
            # attention.py:18-19 registers sqrt(head_dim) as a buffer: keep the key so checkpoints interchange
            store.buffers[prefix + "attn.scale_factor"] = torch.tensor(float(d // n_heads)).sqrt()

    @staticmethod
    def specs(prefix, ein, d, pooling_type=None):
        if pooling_type is not None:      # EntityPoolingLayer: in_trans is d -> d WITH a bias (attention.py:92)
            if pooling_type not in ops.POOL_TYPES:
                raise ValueError("pooling_type %r not in %s" % (pooling_type, sorted(ops.POOL_TYPES)))
            mid = [(prefix + "attn.in_trans.weight", (d, d)), (prefix + "attn.in_trans.bias", (d,))]
        else:
            mid = [(prefix + "attn.in_trans.weight", (3 * d, d))]
        return OrderedDict([(prefix + "fc1.weight", (d, ein)), (prefix + "fc1.bias", (d,))] + mid +
                           [(prefix + "attn.out_trans.weight", (d, d)), (prefix + "attn.out_trans.bias", (d,))])

    def init(self, gen):
        p, pre = self.s.p, self.pre
        init_linear_(gen, p[pre + "fc1.weight"], p[pre + "fc1.bias"])
        init_linear_(gen, p[pre + "attn.in_trans.weight"], p.get(pre + "attn.in_trans.bias"))
        init_linear_(gen, p[pre + "attn.out_trans.weight"], p[pre + "attn.out_trans.bias"])

    def forward(self, ents, la, masks, T, relu_out=False, xin=None):
        """ents [N, ne, ed] f32, la [N, ne] i32 or None -> x2 [C*N*na, d].
        xin: optional packed input [N*ne, Kp] (ops.pack_inputs) -> fc1 runs on the tensor cores with a zero-padded weight."""
        p, pre, ws, tag = self.s.p, self.pre, self.ws, self.tag
        N, ne = ents.shape[0], ents.shape[1]
        C, d, na = masks.C, self.d, self.na
        x1 = ws.get(tag + ".x1", (N * ne, d))
        if xin is not None:          # packed [entities | onehot | 0] input: the (d, ein) weight is read in place, K zero-padded
            ops.linear_fwd(xin, p[pre + "fc1.weight"], p[pre + "fc1.bias"], x1, relu=True)
        else:
            ops.embed_fwd(ents, la, self.A, p[pre + "fc1.weight"], p[pre + "fc1.bias"], x1, relu=True)
        att = ws.get(tag + ".att", (C * N * na, d))
        if self.pool is None:
            qkv = ws.get(tag + ".qkv", (N * ne, 3 * d))
            if ops.qkv_split_ok(N, ne, na, d):    # queries for the agent rows only (the reference slices them after in_trans)
                ops.in_trans_fwd_split(x1, p[pre + "attn.in_trans.weight"], qkv, N, ne, na)
            else:
                ops.linear_fwd(x1, p[pre + "attn.in_trans.weight"], None, qkv)
            ops.masked_attn_fwd(qkv, att, masks.copies, masks.group_bits, masks.entity_mask, N, T, ne, na, d, self.H)
        else:                               # pooling ablation: qkv holds E = in_trans(x1) [N*ne, d]
            qkv = ws.get(tag + ".qkv", (N * ne, d))
            ops.linear_fwd(x1, p[pre + "attn.in_trans.weight"], p[pre + "attn.in_trans.bias"], qkv)
            ops.entity_pool_fwd(qkv, att, masks.copies, masks.group_bits, masks.entity_mask, N, T, ne, na, d, self.pool)
        x2 = ws.get(tag + ".x2", (C * N * na, d))
        self.row_mask = (masks.entity_mask, na, N * na) if masks.entity_mask is not None else None
        ops.linear_fwd(att, p[pre + "attn.out_trans.weight"], p[pre + "attn.out_trans.bias"], x2, relu=relu_out,
                       row_mask=self.row_mask)
        self.saved = (ents, la, masks, T, relu_out, x1, qkv, att, x2, N, ne, xin)
        return x2

    def backward(self, dx2, wstream=None):
        p, g, pre, ws = self.s.p, self.s.g, self.pre, self.ws
        ents, la, masks, T, relu_out, x1, qkv, att, x2, N, ne, xin = self.saved
        C, d, na = masks.C, self.d, self.na
        relu_y = x2 if relu_out else None
        _on_side(wstream, lambda: ops.linear_bwd_weight(dx2, att, g[pre + "attn.out_trans.weight"], g[pre + "attn.out_trans.bias"],
                                                        relu_y=relu_y, row_mask=self.row_mask))
        datt = ws.get(self.scratch + ".datt", (C * N * na, d))
        ops.linear_bwd_data(dx2, p[pre + "attn.out_trans.weight"], datt, relu_y=relu_y, row_mask=self.row_mask)
        if self.pool is None:
            dqkv = ws.get(self.scratch + ".dqkv", (N * ne, 3 * d))
            ops.masked_attn_bwd(qkv, datt, dqkv, masks.copies, masks.group_bits, masks.entity_mask, N, T, ne, na, d, self.H)
            if ops.qkv_split_ok(N, ne, na, d) and ops.tc_wgrad_y_tma():
                _on_side(wstream, lambda: ops.in_trans_bwd_weight_split(dqkv, x1, g[pre + "attn.in_trans.weight"], N, ne, na))
            else:
                _on_side(wstream, lambda: ops.linear_bwd_weight(dqkv, x1, g[pre + "attn.in_trans.weight"], None))
        else:
            dqkv = ws.get(self.scratch + ".dqkv", (N * ne, d))
            ops.entity_pool_bwd(qkv, datt, dqkv, masks.copies, masks.group_bits, masks.entity_mask, N, T, ne, na, d, self.pool)
            _on_side(wstream, lambda: ops.linear_bwd_weight(dqkv, x1, g[pre + "attn.in_trans.weight"],
                                                            g[pre + "attn.in_trans.bias"]))
        dx1 = ws.get(self.scratch + ".dx1", (N * ne, d))
        if self.pool is None and ops.qkv_split_ok(N, ne, na, d):
            ops.in_trans_bwd_data_split(dqkv, p[pre + "attn.in_trans.weight"], dx1, N, ne, na)
        else:
            ops.linear_bwd_data(dqkv, p[pre + "attn.in_trans.weight"], dx1)
        if xin is not None:
            ops.linear_bwd_weight(dx1, xin, g[pre + "fc1.weight"], g[pre + "fc1.bias"], relu_y=x1)
        else:
            ops.embed_bwd_weight(dx1, x1, ents, la, self.A, g[pre + "fc1.weight"], g[pre + "fc1.bias"])


class EntityAttnAgent:
    """Per-agent utility network (shared parameters).  rnn=True: fc1-attn-fc2-GRU-fc3 (entity_rnn_agent.py:31-64);
    rnn=False: fc1-attn-relu-fc2 (entity_ff_agent.py:30-57)."""

    def __init__(self, input_shape, args, device, ws=None, tag="agent", seed=None):
        self.args = args
        self.rnn = "rnn" in args.agent
        self.ein, self.d, self.H = int(input_shape), int(args.attn_embed_dim), int(args.attn_n_heads)
        self.na, self.A, self.r = int(args.n_agents), int(args.n_actions), int(args.rnn_hidden_dim)
        self.one_hot_la = bool(args.entity_last_action)
        self.pool = getattr(args, "pooling_type", None)
        specs = AttnTrunk.specs("", self.ein, self.d, self.pool)
        if self.rnn:
            specs.update(OrderedDict([("fc2.weight", (self.r, self.d)), ("fc2.bias", (self.r,)),
                                      ("rnn.weight_ih", (3 * self.r, self.r)), ("rnn.weight_hh", (3 * self.r, self.r)),
                                      ("rnn.bias_ih", (3 * self.r,)), ("rnn.bias_hh", (3 * self.r,)),
                                      ("fc3.weight", (self.A, self.r)), ("fc3.bias", (self.A,))]))
        else:
            specs.update(OrderedDict([("fc2.weight", (self.A, self.d)), ("fc2.bias", (self.A,))]))
        self.store = ParamStore(specs, device)
        self.ws = ws if ws is not None else Workspace(device)
        self.tag = tag
        self.trunk = AttnTrunk(self.store, "", self.ws, tag, self.ein, self.d, self.H, self.na, self.A, self.pool)
        # like nn.Linear, draw from the GLOBAL torch RNG (a private generator seeded with torch.initial_seed() made the
        # agent's and the mixer's first layers bit-identical whenever their shapes matched)
        gen = torch.Generator().manual_seed(int(seed)) if seed is not None else None
        self.trunk.init(gen)
        p = self.store.p
        if self.rnn:
            init_linear_(gen, p["fc2.weight"], p["fc2.bias"])
            bound = 1.0 / (self.r ** 0.5)
            for k in ("rnn.weight_ih", "rnn.weight_hh", "rnn.bias_ih", "rnn.bias_hh"):
                init_uniform_(gen, p[k], bound)
            init_linear_(gen, p["fc3.weight"], p["fc3.bias"])
        else:
            init_linear_(gen, p["fc2.weight"], p["fc2.bias"])

    # reference nn.Module surface used by BasicMAC ---------------------------------------------------------
    def init_hidden(self):
        return torch.zeros(1, self.r, device=self.store.flat.device)  # entity_rnn_agent.py:27-29

    def parameters(self):
        return self.store.parameters()

    def named_parameters(self):
        return self.store.named_parameters()

    def state_dict(self):
        return self.store.state_dict()

    def load_state_dict(self, sd, strict=True):
        self.store.load_state_dict(sd, strict)

    def cuda(self):
        self.store.to("cuda")
        self.ws.device = self.store.flat.device
        return self

    def train(self, mode=True):
        return self

    def eval(self):
        return self

    # ---------------------------------------------------------------------------------------------------------
    def forward(self, ents, la, masks, B, T, h0=None, train=False, xin=None):
        """ents [B*T, ne, ed]; masks: MaskSpec (C copies); h0 [C*B*na, r] or None (zeros).
        Returns q [C, B*T, na, A] and the hidden-state stack hs [C*B*T*na, r] (rnn) / x2 (ff)."""
        p, ws, tag = self.store.p, self.ws, self.tag
        N, C, na = B * T, masks.C, self.na
        R = C * N * na
        x2 = self.trunk.forward(ents, la if self.one_hot_la else None, masks, T, relu_out=not self.rnn, xin=xin)
        rm = self.trunk.row_mask
        q = ws.get(tag + ".q", (R, self.A))
        if self.rnn:
            x3 = ws.get(tag + ".x3", (R, self.r))
            ops.linear_fwd(x2, p["fc2.weight"], p["fc2.bias"], x3, relu=True)
            gi = ws.get(tag + ".gi", (R, 3 * self.r))
            ops.linear_fwd(x3, p["rnn.weight_ih"], p["rnn.bias_ih"], gi)
            hs = ws.get(tag + ".hs", (R, self.r))
            gates = ws.get(tag + ".gates", (R, 4 * self.r)) if train else None
            ops.gru_scan_fwd(gi, p["rnn.weight_hh"], p["rnn.bias_hh"], h0, hs, gates, C * B * na, T, na)
            ops.linear_fwd(hs, p["fc3.weight"], p["fc3.bias"], q, row_mask=rm)
            self.saved = (x2, x3, hs, gates, h0, B, T, C)
            return q.view(C, N, na, self.A), hs
        ops.linear_fwd(x2, p["fc2.weight"], p["fc2.bias"], q, row_mask=rm)
        self.saved = (x2, None, None, None, None, B, T, C)
        return q.view(C, N, na, self.A), x2

    def dq_width(self, rows):
        """Row width of the dQ buffer the learner scatters into: padded to 32 columns when the head's backward GEMMs run on the
        tensor cores (>= TC_MIN_ROWS rows; aligned reduction, the padding columns are zero), else the plain A columns."""
        width = self.r if self.rnn else self.d
        return 32 if (self.A <= 32 and ops.tc_head_ok(rows, self.A, width)) else self.A

    def backward(self, dq, wstream=None):
        """dq [C*N*na, dq_width()] -> accumulates into the gradient views of the store.  wstream: optional companion stream for
        the weight-gradient kernels (joined before returning)."""
        p, g, ws = self.store.p, self.store.g, self.ws
        x2, x3, hs, gates, h0, B, T, C = self.saved
        na, rm = self.na, self.trunk.row_mask
        R, Ap = dq.shape
        dx2 = ws.get("scratch.dx2", (R, self.d))
        wl, key = ("fc3.weight", "fc3.bias") if self.rnn else ("fc2.weight", "fc2.bias")
        inp = hs if self.rnn else x2
        dhead = ws.get("scratch.dhs", (R, self.r)) if self.rnn else dx2
        # dq may be padded to 32 columns (dq_width): the (A, width) head weight and its gradient are used in place, the
        # kernels take the missing rows / columns as zero
        _on_side(wstream, lambda: ops.linear_bwd_weight(dq, inp, g[wl], g[key], row_mask=rm))
        ops.linear_bwd_data(dq, p[wl], dhead, row_mask=rm)
        if self.rnn:
            dhs = dhead
            dgi = ws.get("scratch.dgi", (R, 3 * self.r))
            dgh = ws.get("scratch.dgh", (R, 3 * self.r))
            ops.gru_scan_bwd(dhs, gates, hs, h0, p["rnn.weight_hh"], dgi, dgh, C * B * na, T, na)

            def rnn_wgrads():
                ops.linear_bwd_weight(dgi, x3, g["rnn.weight_ih"], g["rnn.bias_ih"])
                ops.gru_bwd_weight_hh(dgh, hs, na, T, g["rnn.weight_hh"], g["rnn.bias_hh"])
            _on_side(wstream, rnn_wgrads)
            dx3 = ws.get("scratch.dx3", (R, self.r))
            ops.linear_bwd_data(dgi, p["rnn.weight_ih"], dx3)
            _on_side(wstream, lambda: ops.linear_bwd_weight(dx3, x2, g["fc2.weight"], g["fc2.bias"], relu_y=x3))
            ops.linear_bwd_data(dx3, p["fc2.weight"], dx2, relu_y=x3)
        self.trunk.backward(dx2, wstream=wstream)
        if wstream is not None:
            torch.cuda.current_stream().wait_stream(wstream)


class AttnHyperNet:
    """fc1 - masked attention - fc2 hypernetwork (flex_qmix.py:40-50); the mode reduction (matrix / vector /
    alt_vector / scalar, :51-57) is fused into the mixer kernel."""

    def __init__(self, store, prefix, ws, tag, args, ein):
        self.s, self.pre, self.ws, self.tag = store, prefix, ws, tag
        self.he, self.me, self.na = int(args.hypernet_embed), int(args.mixing_embed_dim), int(args.n_agents)
        self.trunk = AttnTrunk(store, prefix, ws, tag, ein, self.he, int(args.attn_n_heads), self.na, int(args.n_actions),
                               getattr(args, "pooling_type", None))

    @staticmethod
    def specs(prefix, ein, he, me, pooling_type=None):
        s = AttnTrunk.specs(prefix, ein, he, pooling_type)
        s.update(OrderedDict([(prefix + "fc2.weight", (me, he)), (prefix + "fc2.bias", (me,))]))
        return s

    def init(self, gen):
        self.trunk.init(gen)
        init_linear_(gen, self.s.p[self.pre + "fc2.weight"], self.s.p[self.pre + "fc2.bias"])

    def forward(self, ents, la, masks, T, xin=None):
        p, pre = self.s.p, self.pre
        x2 = self.trunk.forward(ents, la, masks, T, xin=xin)
        x3 = self.ws.get(self.tag + ".x3", (x2.shape[0], self.me))
        ops.linear_fwd(x2, p[pre + "fc2.weight"], p[pre + "fc2.bias"], x3, row_mask=self.trunk.row_mask)
        self.x2 = x2
        return x3

    def backward(self, dx3, wstream=None):
        p, g, pre = self.s.p, self.s.g, self.pre
        rm = self.trunk.row_mask
        _on_side(wstream, lambda: ops.linear_bwd_weight(dx3, self.x2, g[pre + "fc2.weight"], g[pre + "fc2.bias"], row_mask=rm))
        dx2 = self.ws.get(self.trunk.scratch + ".dx2h", (dx3.shape[0], self.he))
        ops.linear_bwd_data(dx3, p[pre + "fc2.weight"], dx2, row_mask=rm)
        self.trunk.backward(dx2, wstream=wstream)
        if wstream is not None:
            torch.cuda.current_stream().wait_stream(wstream)


# ---- lock-step execution of several hypernetworks: every dense layer of the group is ONE launch -------------------------------
def hypernets_forward_group(jobs, ents, la, T, xin):
    """jobs: list of (AttnHyperNet, MaskSpec) over the SAME entity rows (hypernetworks of the online and of the target mixer).
    The nets advance layer by layer; the same layer of all nets is one grouped tensor-core launch (ops.linear_fwd_group), the
    attention of each net its own launch.  Falls back to net-by-net execution when the packed input is absent (small batches run
    on the FFMA kernels) or a pooling layer replaces the attention.  -> list of outputs [rows_i, me]."""
    if xin is None or len(jobs) < 2 or any(net.trunk.pool is not None for net, _ in jobs):
        return [net.forward(ents, la, m, T, xin=xin) for net, m in jobs]
    N, ne = ents.shape[0], ents.shape[1]
    trunks = [net.trunk for net, _ in jobs]
    x1 = [t.ws.get(t.tag + ".x1", (N * ne, t.d)) for t in trunks]
    ops.linear_fwd_group([(xin, t.s.p[t.pre + "fc1.weight"], t.s.p[t.pre + "fc1.bias"], x, True, None) for t, x in zip(trunks, x1)])
    qkv = [t.ws.get(t.tag + ".qkv", (N * ne, 3 * t.d)) for t in trunks]
    ops.linear_fwd_group([(x, t.s.p[t.pre + "attn.in_trans.weight"], None, q, False, None) for t, x, q in zip(trunks, x1, qkv)])
    att, x2, x3 = [], [], []
    t0 = trunks[0]
    for (net, m), t, q in zip(jobs, trunks, qkv):
        a = t.ws.get(t.tag + ".att", (m.C * N * t.na, t.d))
        att.append(a)
        t.row_mask = (m.entity_mask, t.na, N * t.na) if m.entity_mask is not None else None
        x2.append(t.ws.get(t.tag + ".x2", (m.C * N * t.na, t.d)))
        x3.append(net.ws.get(net.tag + ".x3", (m.C * N * t.na, net.me)))
    ops.masked_attn_fwd_group([(q, a, m.copies, m.group_bits, m.entity_mask) for (_, m), q, a in zip(jobs, qkv, att)],
                              N, T, ne, t0.na, t0.d, t0.H)
    ops.linear_fwd_group([(a, t.s.p[t.pre + "attn.out_trans.weight"], t.s.p[t.pre + "attn.out_trans.bias"], y, False, t.row_mask)
                          for t, a, y in zip(trunks, att, x2)])
    ops.linear_fwd_group([(y, net.s.p[net.pre + "fc2.weight"], net.s.p[net.pre + "fc2.bias"], z, False, net.trunk.row_mask)
                          for (net, _), y, z in zip(jobs, x2, x3)])
    for (net, m), t, a, b, c, y in zip(jobs, trunks, x1, qkv, att, x2):
        t.saved = (ents, la, m, T, False, a, b, c, y, N, ne, xin)
        net.x2 = y
    return x3


def hypernets_backward_group(nets, douts, wstream=None):
    """Backward of several hypernetworks (forwarded by hypernets_forward_group on the packed input) in lock-step: the data-gradient
    chain fc2 -> out_trans -> attention -> in_trans is 3 grouped launches + one attention launch per net; the weight gradients
    (5 grouped launches) go to the companion stream `wstream`, off the chain."""
    trunks = [net.trunk for net in nets]
    if len(nets) < 2 or any(t.saved[11] is None or t.pool is not None for t in trunks):
        for net, d in zip(nets, douts):
            net.backward(d, wstream=wstream)
        return
    g = [net.s.g for net in nets]
    p = [net.s.p for net in nets]
    _on_side(wstream, lambda: ops.linear_bwd_weight_group(
        [(d, net.x2, gg[net.pre + "fc2.weight"], gg[net.pre + "fc2.bias"], None, net.trunk.row_mask) for net, d, gg in zip(nets, douts, g)]))
    dx2 = [net.ws.get(net.trunk.scratch + ".dx2h", (d.shape[0], net.he)) for net, d in zip(nets, douts)]
    ops.linear_bwd_data_group([(d, pp[net.pre + "fc2.weight"], y, None, net.trunk.row_mask) for net, d, y, pp in zip(nets, douts, dx2, p)])
    sv = [t.saved for t in trunks]          # (ents, la, masks, T, relu_out, x1, qkv, att, x2, N, ne, xin)
    _on_side(wstream, lambda: ops.linear_bwd_weight_group(
        [(y, s_[7], gg[t.pre + "attn.out_trans.weight"], gg[t.pre + "attn.out_trans.bias"], None, t.row_mask)
         for t, y, s_, gg in zip(trunks, dx2, sv, g)]))
    datt = [t.ws.get(t.scratch + ".datt", tuple(s_[7].shape)) for t, s_ in zip(trunks, sv)]
    ops.linear_bwd_data_group([(y, pp[t.pre + "attn.out_trans.weight"], a, None, t.row_mask) for t, y, a, pp in zip(trunks, dx2, datt, p)])
    T, N, ne, t0 = sv[0][3], sv[0][9], sv[0][10], trunks[0]
    dqkv = [t.ws.get(t.scratch + ".dqkv", (N * ne, 3 * t.d)) for t in trunks]
    ops.masked_attn_bwd_group([(s_[6], a, q, s_[2].copies, s_[2].group_bits, s_[2].entity_mask) for s_, a, q in zip(sv, datt, dqkv)],
                              N, T, ne, t0.na, t0.d, t0.H)
    _on_side(wstream, lambda: ops.linear_bwd_weight_group(
        [(q, s_[5], gg[t.pre + "attn.in_trans.weight"], None, None, None) for t, q, s_, gg in zip(trunks, dqkv, sv, g)]))
    dx1 = [t.ws.get(t.scratch + ".dx1", tuple(s_[5].shape)) for t, s_ in zip(trunks, sv)]
    ops.linear_bwd_data_group([(q, pp[t.pre + "attn.in_trans.weight"], x, None, None) for t, q, x, pp in zip(trunks, dqkv, dx1, p)])
    ops.linear_bwd_weight_group([(x, s_[11], gg[t.pre + "fc1.weight"], gg[t.pre + "fc1.bias"], s_[5], None)
                                 for t, x, s_, gg in zip(trunks, dx1, sv, g)])
    if wstream is not None:
        torch.cuda.current_stream().wait_stream(wstream)


class Mixer:
    """FlexQMixer / LinearFlexQMixer / VDNMixer (modules/mixers/flex_qmix.py:60-172, vdn.py:5-10) over [N] rows."""

    HYPERS = {"flex_qmix": ["hyper_w_1.", "hyper_w_final.", "hyper_b_1.", "V."], "lin_flex_qmix": ["hyper_w_1.", "V."],
              "vdn": []}

    def __init__(self, args, ein, device, ws=None, tag="mixer", seed=None):
        self.args, self.kind_name = args, args.mixer
        self.kind = ops.MIX_KIND[args.mixer]
        self.na, self.me = int(args.n_agents), int(getattr(args, "mixing_embed_dim", 1) or 1)
        self.softmax_w = bool(getattr(args, "softmax_mixing_weights", True))
        self.tanh_nl = getattr(args, "mixer_non_lin", "elu") == "tanh"
        self.ws = ws if ws is not None else Workspace(device)
        self.tag = tag
        specs = OrderedDict()
        for h in self.HYPERS[args.mixer]:
            specs.update(AttnHyperNet.specs(h, ein, int(args.hypernet_embed), self.me,
                                            getattr(args, "pooling_type", None)))  # vdn: no hypernets
        self.store = ParamStore(specs, device)
        self.nets = OrderedDict((h, AttnHyperNet(self.store, h, self.ws, tag + "." + h, args, ein))
                                for h in self.HYPERS[args.mixer])
        gen = torch.Generator().manual_seed(int(seed)) if seed is not None else None      # None: the global torch RNG
        for n in self.nets.values():
            n.init(gen)

    def parameters(self):
        return self.store.parameters()

    def named_parameters(self):
        return self.store.named_parameters()

    def state_dict(self):
        return self.store.state_dict()

    def load_state_dict(self, sd, strict=True):
        self.store.load_state_dict(sd, strict)

    def cuda(self):
        self.store.to("cuda")
        return self

    def train(self, mode=True):
        return self

    def eval(self):
        return self

    def set_scratch_groups(self, groups):
        """groups: list of lists of hypernet names that may run concurrently with each other's group -> distinct scratch."""
        for gi, names in enumerate(groups):
            for h in names:
                self.nets[h].trunk.scratch = "scratch%d" % gi

    def hyper_forward(self, ents, la, entity_mask, T, imagine_masks=None, xin=None, streams=None):
        """The heavy part of the mixer: every hypernetwork evaluated on the entity rows (independent of the agent utilities,
        so the learner may run it concurrently with the agent forward).  streams: optional list of CUDA streams, hypernetwork i
        is enqueued on streams[i % len] (the nets are independent of each other; every net has its own activations).
        -> {name: [rows, me]}"""
        imagine = imagine_masks is not None
        outs = {}
        if self.kind != 2:
            default = (None, 0, ops.ATTN_DEFAULT)
            for i, (h, net) in enumerate(self.nets.items()):
                if h == "hyper_w_1." and imagine:
                    copies, gbits = imagine_masks
                    m = MaskSpec([default] + list(copies), gbits, entity_mask)
                else:
                    m = MaskSpec([default], None, entity_mask)
                if streams:
                    with torch.cuda.stream(streams[i % len(streams)]):
                        outs[h] = net.forward(ents, la, m, T, xin=xin)
                else:
                    outs[h] = net.forward(ents, la, m, T, xin=xin)
        self._hyper = (outs, ents.shape[0], imagine)
        return outs

    def hyper_jobs(self, entity_mask, imagine_masks=None):
        """(net, MaskSpec) pairs of this mixer's hypernetworks, for hypernets_forward_group."""
        imagine = imagine_masks is not None
        jobs = []
        if self.kind != 2:
            default = (None, 0, ops.ATTN_DEFAULT)
            for h, net in self.nets.items():
                if h == "hyper_w_1." and imagine:
                    copies, gbits = imagine_masks
                    jobs.append((net, MaskSpec([default] + list(copies), gbits, entity_mask)))
                else:
                    jobs.append((net, MaskSpec([default], None, entity_mask)))
        return jobs

    def set_hyper_outputs(self, outs_list, n_rows, imagine):
        self._hyper = (dict(zip(self.nets.keys(), outs_list)), n_rows, imagine)

    def mix(self, q, qW, qI, ret_ingroup=False):
        """q [N, na] (+ qW, qI when imagine) combined with the hypernetwork outputs of the last hyper_forward.
        Returns (q_tot [N], q_tot_im [N] or None)."""
        ws, tag = self.ws, self.tag
        outs, N, imagine = self._hyper
        qtot = ws.get(tag + ".qtot", (N,))
        qtot_im = ws.get(tag + ".qtot_im", (N,)) if imagine else None
        w1 = outs.get("hyper_w_1.")
        self.saved = (q, qW, qI, outs, N, imagine)
        self.ingroup = ws.get(tag + ".ingroup", (N,)) if (ret_ingroup and imagine and self.kind == 1) else None
        ops.mixer_fwd(self.kind, w1, outs.get("hyper_b_1."), outs.get("hyper_w_final."), outs.get("V."), q, qW, qI,
                      qtot, qtot_im, N, self.na, self.me, 3 if imagine else 1, imagine, self.softmax_w, self.tanh_nl,
                      ingroup=self.ingroup)
        return qtot, qtot_im

    def forward(self, q, qW, qI, ents, la, entity_mask, T, imagine_masks=None, xin=None, ret_ingroup=False):
        """q [N, na] (+ qW, qI when imagine); ents [N, ne, ed]; entity_mask [N, ne].
        imagine_masks: (copy_W, copy_I) MaskSpec-style copy tuples + group bits, or None.
        Returns (q_tot [N], q_tot_im [N] or None)."""
        self.hyper_forward(ents, la, entity_mask, T, imagine_masks=imagine_masks, xin=xin)
        return self.mix(q, qW, qI, ret_ingroup=ret_ingroup)

    def backward_mix(self, g_plain, g_im):
        """Mixer combine backward: -> (dq [3, N, na], {name: d(hypernet output)}); cheap, no parameter gradients."""
        ws, tag = self.ws, self.tag
        q, qW, qI, outs, N, imagine = self.saved
        na = self.na
        dq = ws.get(tag + ".dq", (3, N, na))
        d = {h: ws.get(tag + ".d" + h, tuple(o.shape)) for h, o in outs.items()}
        ops.mixer_bwd(self.kind, outs.get("hyper_w_1."), outs.get("hyper_b_1."), outs.get("hyper_w_final."),
                      outs.get("V."), q, qW, qI, g_plain, g_im, d.get("hyper_w_1."), d.get("hyper_b_1."),
                      d.get("hyper_w_final."), d.get("V."), dq[0], dq[1] if imagine else None,
                      dq[2] if imagine else None, N, na, self.me, 3 if imagine else 1, imagine, self.softmax_w,
                      self.tanh_nl)
        return dq, d

    def backward_hyper(self, d, names=None, streams=None, wstreams=None):
        """Hypernetwork backward passes (parameter gradients are accumulated); independent of the agent backward and, given
        distinct scratch groups (set_scratch_groups), of each other: net i goes to streams[i % len] when streams are given, its
        weight-gradient kernels to the companion wstreams[i % len]."""
        for i, (h, net) in enumerate(self.nets.items()):
            if names is None or h in names:
                if streams:
                    with torch.cuda.stream(streams[i % len(streams)]):
                        net.backward(d[h], wstream=wstreams[i % len(wstreams)] if wstreams else None)
                else:
                    net.backward(d[h])

    def backward(self, g_plain, g_im):
        """-> dq, dqW, dqI [N, na]; hypernet parameter gradients are accumulated."""
        dq, d = self.backward_mix(g_plain, g_im)
        self.backward_hyper(d)
        return dq
