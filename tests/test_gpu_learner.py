"""GPU parity of the learner hot path against the fixtures recorded from the reference's EntityMAC + QLearner
(tests/golden/learner_*.npz) and against the CPU oracle at a larger size.  Tolerances: 1e-4 relative on utilities
and losses (north_star); greedy action indices bit-exact."""
import pytest
import torch

from golden_util import learner_cases, load_learner_case
from gpu_util import build_product

pytestmark = pytest.mark.gpu
CASES = learner_cases()
DEV = "cuda:0"


def _close(a, b, rtol=1e-4, atol=2e-6, what=""):
    a, b = a.detach().cpu().float(), b.detach().cpu().float()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert torch.allclose(a, b, rtol=rtol, atol=atol), (what, (a - b).abs().max().item(), b.abs().max().item())


def _setup(name):
    c = load_learner_case(name)
    batch, mac, learner, logger = build_product(c.args, c.dims, c.batch, DEV)
    mac.agent.load_state_dict(c.agent)
    learner.target_mac.agent.load_state_dict(c.tagent)
    if c.mixer:
        learner.mixer.load_state_dict(c.mixer)
        learner.target_mixer.load_state_dict(c.tmixer)
    return c, batch, mac, learner, logger


@pytest.mark.parametrize("name", CASES)
def test_forward_and_greedy_actions(name):
    c, batch, mac, learner, _ = _setup(name)
    B, T, na, ne, ed, A = c.dims
    mac.init_hidden(B)
    _close(mac.forward(batch, t=None), c.fwd["mac_out"], what="mac_out")
    if "imagine" in c.args.agent:
        mac.init_hidden(B)
        q3, (wm, im) = mac.forward(batch, t=None, imagine=True, use_gt_factors=c.args.train_gt_factors,
                                   use_rand_gt_factors=c.args.train_rand_gt_factors, group_bits=c.group_a.to(DEV))
        _close(q3, c.fwd["imagine_out"], what="imagine_out")
        assert torch.equal(wm.cpu(), c.fwd["wmask"]) and torch.equal(im.cpu(), c.fwd["imask"])
    mac.init_hidden(B)
    acts, qs = [], []
    for t in range(T):
        a, q = mac.select_actions(batch, t_ep=t, t_env=0, test_mode=True, ret_agent_outs=True)
        acts.append(a.cpu())
        qs.append(q.cpu())
    _close(torch.stack(qs, 1), c.greedy_q, what="greedy_q")
    assert torch.equal(torch.stack(acts, 1), c.greedy_actions)


@pytest.mark.parametrize("name", CASES)
def test_train_step_matches_reference(name):
    c, batch, mac, learner, logger = _setup(name)
    learner.train(batch, t_env=10, episode_num=0, group_bits=c.group_a.to(DEV))
    torch.cuda.synchronize()
    s = c.stats
    for k in ("loss", "im_loss", "grad_norm", "td_error_abs", "q_taken_mean", "target_mean", "ingroup_prop",
              "gt_ingroup_prop"):
        if k in s:
            got = logger.stats[k][0]
            assert abs(got - s[k]) <= 1e-4 * max(1.0, abs(s[k])), (k, got, s[k])
    # real-width case (refil_ns: reductions over 768 / 256 rows on the 3xTF32 tensor-core path with fp32 atomics): the absolute
    # floor scales with the largest gradient entry, as in the mid-size oracle test below
    atol = 2e-6
    if c.compact:
        atol = 2e-5 * max(float(g.abs().max()) for g in list(c.grad_agent.values()) + list(c.grad_mixer.values()))
    for k, g in c.grad_agent.items():
        _close(mac.agent.store.g[k], g, rtol=2e-4, atol=atol, what="grad agent " + k)
    for k, g in c.grad_mixer.items():
        _close(learner.mixer.store.g[k], g, rtol=2e-4, atol=atol, what="grad mixer " + k)
    for k, p in c.new_agent.items():
        assert float((mac.agent.store.p[k].cpu() - p).abs().max()) <= 1e-4, k
    for k, p in c.new_mixer.items():
        assert float((learner.mixer.store.p[k].cpu() - p).abs().max()) <= 1e-4, k


@pytest.mark.parametrize("alg", ["refil", "qmix_atten", "refil_gm", "refil_pool_mean", "refil_pool_max", "refil_cfg5", "refil_qsplit"])
def test_train_step_matches_oracle_mid_size(alg, monkeypatch):
    """Same seeded inputs through the CUDA path and the CPU oracle at a size the oracle finishes in seconds.  `refil_cfg5` is
    BASELINE config 5's shape (sc2 3-8csz: 16 entities, T = 120) on 2 episodes; `refil_pool_*` the EntityPoolingLayer ablation at
    the real layer widths; `refil_qsplit` forces the split in_trans (queries for the agent rows only, forward and backward) that
    large shards take by default."""
    from oracle import learner_oracle as lo
    gen = torch.Generator().manual_seed(123)
    if alg == "refil_qsplit":
        monkeypatch.setenv("REFIL_QKV_SPLIT", "1")
        alg = "refil"
    pool = alg.split("_pool_")[1] if "_pool_" in alg else None
    if alg == "refil_cfg5":
        B, T, na, ne, ed, A = 2, 120, 8, 16, 39, 14
        args = lo.default_args(agent="imagine_entity_attend_rnn")
        alg = "refil"
    elif pool:
        B, T, na, ne, ed, A = 5, 12, 8, 24, 39, 14
        args = lo.default_args(agent="imagine_entity_attend_rnn", pooling_type=pool)
        alg = "refil"
    elif alg == "refil_gm":
        B, T, na, ne, ed, A = 6, 9, 4, 4, 12, 3
        args = lo.default_args(agent="imagine_entity_attend_ff", mixer="lin_flex_qmix", attn_embed_dim=64,
                               hypernet_embed=64, entity_last_action=False)
    else:
        B, T, na, ne, ed, A = 5, 12, 8, 24 if alg == "refil" else 16, 39, 14
        args = lo.default_args(agent="imagine_entity_attend_rnn" if alg == "refil" else "entity_attend_rnn")
    args.n_agents, args.n_actions, args.n_entities, args.entity_shape = na, A, ne, ed
    args.mac, args.learner, args.agent_output_type, args.action_selector = "entity_mac", "q_learner", "q", "epsilon_greedy"
    args.target_update_interval, args.learner_log_interval, args.gt_mask_avail = 200, 1, False
    ein = ed + (A if args.entity_last_action else 0)
    syn = lo.synthetic_batch(gen, B, T, na, ne, ed, A, pad=alg != "refil_gm")
    agent_p, mixer_p = lo.init_agent_params(gen, args, ein), lo.init_mixer_params(gen, args, ein)
    tagent_p = {k: v + 0.05 * torch.randn(v.shape, generator=gen) for k, v in agent_p.items()}
    tmixer_p = {k: v + 0.05 * torch.randn(v.shape, generator=gen) for k, v in mixer_p.items()}
    group_a = (torch.rand(B, ne, generator=gen) < 0.5).to(torch.uint8)
    ref = lo.train_step(agent_p, mixer_p, tagent_p, tmixer_p, syn, args, group_a=group_a)
    batch, mac, learner, logger = build_product(args, (B, T, na, ne, ed, A), syn, DEV)
    mac.agent.load_state_dict(agent_p)
    learner.target_mac.agent.load_state_dict(tagent_p)
    learner.mixer.load_state_dict(mixer_p)
    learner.target_mixer.load_state_dict(tmixer_p)
    learner.train(batch, t_env=10, episode_num=0, group_bits=group_a.to(DEV))
    torch.cuda.synchronize()
    loss = float(ref["loss"])
    assert abs(logger.stats["loss"][0] - loss) <= 1e-4 * max(1.0, abs(loss))
    gn = float(ref["grad_norm"])
    assert abs(logger.stats["grad_norm"][0] - gn) <= 1e-4 * max(1.0, gn)
    scale = max(float(g.abs().max()) for g in ref["grads_agent"].values())
    for k, g in ref["grads_agent"].items():
        _close(mac.agent.store.g[k], g, rtol=2e-4, atol=2e-5 * scale, what="grad agent " + k)
    for k, g in ref["grads_mixer"].items():
        _close(learner.mixer.store.g[k], g, rtol=2e-4, atol=2e-5 * scale, what="grad mixer " + k)


@pytest.mark.gpu
def test_concurrent_streams_match_single_stream():
    """The three-stream schedule of QLearner.train is a pure re-ordering: same loss statistics and gradients as the
    single-stream schedule (up to the summation order of the fp32 atomics in the weight-gradient kernels)."""
    import copy
    import torch
    from gpu_util import build_product
    from oracle import learner_oracle as lo
    gen = torch.Generator().manual_seed(5)
    B, T, na, ne, ed, A = 6, 9, 8, 24, 39, 14
    args = lo.default_args()
    args.n_agents, args.n_actions, args.n_entities, args.entity_shape = na, A, ne, ed
    args.mac, args.learner, args.agent_output_type, args.action_selector = "entity_mac", "q_learner", "q", "epsilon_greedy"
    args.target_update_interval, args.learner_log_interval, args.gt_mask_avail = 200, 1, False
    syn = lo.synthetic_batch(gen, B, T, na, ne, ed, A)
    ap, mp = lo.init_agent_params(gen, args, ed + A), lo.init_mixer_params(gen, args, ed + A)
    gb = (torch.rand(B, ne, generator=gen) < 0.5).to(torch.uint8)
    outs = []
    for conc in (False, True):
        a2 = copy.copy(args)
        a2.concurrent_streams = conc
        batch, mac, learner, logger = build_product(a2, (B, T, na, ne, ed, A), syn, "cuda:0")
        mac.agent.load_state_dict(ap)
        learner.target_mac.agent.load_state_dict(ap)
        learner.mixer.load_state_dict(mp)
        learner.target_mixer.load_state_dict(mp)
        learner.train(batch, t_env=10, episode_num=0, group_bits=gb.to("cuda:0"))
        torch.cuda.synchronize()
        outs.append((logger.stats["loss"][0], logger.stats["grad_norm"][0], learner.flat.clone()))
    assert abs(outs[0][0] - outs[1][0]) <= 1e-6 * max(1.0, abs(outs[0][0]))
    assert abs(outs[0][1] - outs[1][1]) <= 1e-5 * max(1.0, abs(outs[0][1]))
    assert torch.allclose(outs[0][2], outs[1][2], rtol=0, atol=1e-6)


@pytest.mark.gpu
def test_cuda_graph_step_matches_eager():
    """args.cuda_graph: the captured + replayed step trains exactly like the eager one (same injected partition is not possible
    under capture, so a non-imagine algorithm is used: qmix_atten has no random partition)."""
    import copy
    import torch
    from gpu_util import build_product
    from oracle import learner_oracle as lo
    gen = torch.Generator().manual_seed(7)
    B, T, na, ne, ed, A = 5, 8, 8, 16, 39, 14
    args = lo.default_args()
    args.agent = "entity_attend_rnn"
    args.n_agents, args.n_actions, args.n_entities, args.entity_shape = na, A, ne, ed
    args.mac, args.learner, args.agent_output_type, args.action_selector = "entity_mac", "q_learner", "q", "epsilon_greedy"
    args.target_update_interval, args.learner_log_interval, args.gt_mask_avail = 200, 10 ** 9, False
    syn = lo.synthetic_batch(gen, B, T, na, ne, ed, A)
    ap, mp = lo.init_agent_params(gen, args, ed + A), lo.init_mixer_params(gen, args, ed + A)
    finals = []
    for graph in (False, True):
        a2 = copy.copy(args)
        a2.cuda_graph = graph
        batch, mac, learner, logger = build_product(a2, (B, T, na, ne, ed, A), syn, "cuda:0")
        mac.agent.load_state_dict(ap)
        learner.target_mac.agent.load_state_dict(ap)
        learner.mixer.load_state_dict(mp)
        learner.target_mixer.load_state_dict(mp)
        for step in range(4):                      # eager, capture + replay, replay, replay
            learner.train(batch, t_env=step, episode_num=step)
        torch.cuda.synchronize()
        finals.append(learner.flat.clone())
    assert torch.allclose(finals[0], finals[1], rtol=0, atol=2e-6), (finals[0] - finals[1]).abs().max().item()


def _mid_args(lo, agent, mixer="flex_qmix", d=128, last_action=True, dims=None):
    args = lo.default_args(agent=agent, mixer=mixer, attn_embed_dim=d, hypernet_embed=d, entity_last_action=last_action)
    B, T, na, ne, ed, A = dims
    args.n_agents, args.n_actions, args.n_entities, args.entity_shape = na, A, ne, ed
    args.mac, args.learner, args.agent_output_type, args.action_selector = "entity_mac", "q_learner", "q", "epsilon_greedy"
    args.target_update_interval, args.learner_log_interval, args.gt_mask_avail = 200, 1, False
    return args


@pytest.mark.parametrize("kind", ["refil_rnn_64env_8ag_24ent", "gm_ff_4096env_4ag"])
def test_greedy_indices_on_tensor_core_path(kind):
    """Acting steps large enough (>= 256 rows) to run on the tcgen05 3xTF32 GEMMs: BASELINE config 2's 4096 envs x 4 agents
    (FF agent, d=64) and a 64-env step of the north-star network.  Greedy indices must equal the oracle's first-max wherever
    the oracle's top-2 margin exceeds 1e-5 (a near-tie inside fp32 rounding is not an index error); utilities within 1e-4."""
    from oracle import learner_oracle as lo
    from refil_b200 import ops
    gen = torch.Generator().manual_seed(31)
    if kind.startswith("gm"):
        dims = (4096, 3, 4, 4, 12, 3)
        args = _mid_args(lo, "imagine_entity_attend_ff", "lin_flex_qmix", 64, False, dims)
    else:
        dims = (64, 3, 8, 24, 39, 14)
        args = _mid_args(lo, "imagine_entity_attend_rnn", dims=dims)
    B, T, na, ne, ed, A = dims
    assert B * ne >= ops.TC_MIN_ROWS and B * na >= ops.TC_MIN_ROWS
    args.fused_acting = False                  # this test is about the layer-by-layer path on the tensor cores
    syn = lo.synthetic_batch(gen, B, T, na, ne, ed, A, pad=ne > na)
    ein = ed + (A if args.entity_last_action else 0)
    ap = lo.init_agent_params(gen, args, ein)
    with torch.no_grad():
        q_ref = lo.agent_forward(ap, args, syn)
    batch, mac, learner, _ = build_product(args, dims, syn, DEV)
    mac.agent.load_state_dict(ap)
    mac.init_hidden(B)
    l0 = ops.launch_count()
    acts, qs = [], []
    for t in range(T):
        a, q = mac.select_actions(batch, t_ep=t, t_env=0, test_mode=True, ret_agent_outs=True)
        acts.append(a.cpu())
        qs.append(q.cpu())
    assert ops.launch_count() > l0
    acts, qs = torch.stack(acts, 1), torch.stack(qs, 1)
    _close(qs, q_ref, what="acting utilities")
    avail = syn["avail_actions"]
    ref_acts = torch.stack([lo.greedy_actions(q_ref[:, t], avail[:, t]) for t in range(T)], 1)
    masked = q_ref.clone()
    masked[avail == 0] = -float("inf")
    top2 = masked.topk(2, dim=-1).values
    margin = top2[..., 0] - top2[..., 1]
    decided = margin > 1e-5
    n_bad = int((acts != ref_acts)[decided].sum())
    print("%s: %d decisions, min top-2 margin %.3e, %d near-ties (<=1e-5), mismatches among near-ties %d"
          % (kind, acts.numel(), float(margin.min()), int((~decided).sum()), int((acts != ref_acts)[~decided].sum())))
    assert n_bad == 0


def test_cuda_graph_refil_step_redraws_partition_and_matches_eager():
    """The benchmarked configuration: REFIL under CUDA-graph replay.  The random partition is drawn INSIDE the captured graph
    (torch's philox offset advances per replay): two replays must see different partitions, and a replayed step must train
    exactly like an eager step fed the bits that replay drew."""
    import copy
    from oracle import learner_oracle as lo
    gen = torch.Generator().manual_seed(9)
    dims = (8, 6, 8, 24, 39, 14)
    B, T, na, ne, ed, A = dims
    args = _mid_args(lo, "imagine_entity_attend_rnn", dims=dims)
    args.learner_log_interval = 10 ** 9
    syn = lo.synthetic_batch(gen, B, T, na, ne, ed, A)
    ap, mp = lo.init_agent_params(gen, args, ed + A), lo.init_mixer_params(gen, args, ed + A)
    learners = []
    for graph in (True, False):
        a2 = copy.copy(args)
        a2.cuda_graph = graph
        batch, mac, learner, _ = build_product(a2, dims, syn, DEV)
        mac.agent.load_state_dict(ap)
        learner.target_mac.agent.load_state_dict(ap)
        learner.mixer.load_state_dict(mp)
        learner.target_mixer.load_state_dict(mp)
        learners.append((learner, batch))
    (lg, bg), (le, be) = learners
    torch.manual_seed(77)
    bits = []
    for step in range(4):                          # eager, capture + replay, replay, replay
        lg.train(bg, t_env=step, episode_num=step)
        torch.cuda.synchronize()
        b = lg.last_group_bits.clone()
        bits.append(b.cpu())
        le.train(be, t_env=step, episode_num=step, group_bits=b)
        torch.cuda.synchronize()
        assert torch.allclose(lg.flat, le.flat, rtol=0, atol=2e-6), (step, (lg.flat - le.flat).abs().max().item())
    assert lg._graphs[(B, T)]["graph"] is not None, "the step was not captured"
    assert not torch.equal(bits[1], bits[2]) and not torch.equal(bits[2], bits[3]), "replays re-used one partition"
    assert all(set(b.unique().tolist()) <= {0, 1} for b in bits)


@pytest.mark.parametrize("case", ["gm4", "gm8_pad", "gm8_lastaction", "gt_obs"])
def test_fused_ff_acting_kernel_matches_oracle(case):
    """csrc/ffact.cu: the acting step of the FF entity-attention agent as ONE kernel reading the rollout tensors in place.  Utilities
    within 1e-5 of the oracle (fp32 FFMA), greedy indices equal wherever the oracle's top-2 margin exceeds 1e-5; the launch count shows
    the fused path ran.  Cases: BASELINE config 2's shape; 8 agents with masked / padded slots and partial observability; the one-hot
    last-action input; gt_obs_mask (gt_mask [na, ne] as the observation mask, entity_ff_agent.py:34-35)."""
    from oracle import learner_oracle as lo
    from refil_b200 import ops
    gen = torch.Generator().manual_seed(41)
    la = case == "gm8_lastaction"
    gt = case == "gt_obs"
    dims = (300, 4, 4, 4, 12, 3) if case == "gm4" else (77, 5, 8, 8, 16, 3) if case != "gm8_lastaction" else (64, 5, 6, 8, 14, 5)
    B, T, na, ne, ed, A = dims
    args = _mid_args(lo, "entity_attend_ff" if gt else "imagine_entity_attend_ff", "lin_flex_qmix", 64, la, dims)
    args.gt_obs_mask = gt
    args.gt_mask_avail = gt
    syn = lo.synthetic_batch(gen, B, T, na, ne, ed, A, gt_mask=gt, pad=case != "gm4")
    ein = ed + (A if la else 0)
    ap = lo.init_agent_params(gen, args, ein)
    with torch.no_grad():
        q_ref = lo.agent_forward(ap, args, syn)
    batch, mac, learner, _ = build_product(args, dims, syn, DEV)
    mac.agent.load_state_dict(ap)
    assert mac._fused_acting_ok(batch)
    mac.init_hidden(B)
    acts, qs = [], []
    for t in range(T):
        l0 = ops.launch_count()
        a, q = mac.select_actions(batch, t_ep=t, t_env=0, test_mode=True, ret_agent_outs=True)
        assert ops.launch_count() - l0 == 2          # the fused forward + the selection kernel
        acts.append(a.cpu())
        qs.append(q.cpu().clone())
    acts, qs = torch.stack(acts, 1), torch.stack(qs, 1)
    _close(qs, q_ref, rtol=1e-5, atol=2e-6, what="fused acting utilities")
    avail = syn["avail_actions"]
    ref_acts = torch.stack([lo.greedy_actions(q_ref[:, t], avail[:, t]) for t in range(T)], 1)
    masked = q_ref.clone()
    masked[avail == 0] = -float("inf")
    top2 = masked.topk(2, dim=-1).values
    decided = (top2[..., 0] - top2[..., 1]) > 1e-5
    assert int((acts != ref_acts)[decided].sum()) == 0
    # and the layer-by-layer path gives the same utilities
    args.fused_acting = False
    mac.init_hidden(B)
    q_slow = torch.stack([mac.forward(batch, t).cpu().clone() for t in range(T)], 1)
    _close(q_slow, qs, rtol=1e-5, atol=5e-6, what="fused vs layer-by-layer")
