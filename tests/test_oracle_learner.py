"""Pins oracle/learner_oracle.py against fixtures produced by the reference's EntityMAC + QLearner.train
(tests/golden/learner_*.npz, generated from /root/reference by tests/golden/make_golden.py)."""
import pytest
import torch

from golden_util import learner_cases, load_learner_case
from oracle import learner_oracle as lo

CASES = learner_cases()


def _close(a, b, rtol=1e-5, atol=1e-6):
    assert a.shape == b.shape, (a.shape, b.shape)
    assert torch.allclose(a, b, rtol=rtol, atol=atol), float((a - b).abs().max())


def test_have_cases():
    assert {"refil", "qmix_atten", "refil_gm", "qmix_atten_gm", "refil_vdn", "vdn_atten", "refil_ns", "refil_gm_gt",
            "qmix_atten_gm_gtobs", "refil_gtflag"} <= set(CASES)
    assert "ingroup_prop" in load_learner_case("refil_gm").stats and "gt_ingroup_prop" in load_learner_case("refil_gm_gt").stats


@pytest.mark.parametrize("name", CASES)
def test_forward_matches_reference(name):
    c = load_learner_case(name)
    with torch.no_grad():
        q = lo.agent_forward(c.agent, c.args, c.batch)
        _close(q, c.fwd["mac_out"])
        if "imagine" in c.args.agent:
            q3, (wm, im) = lo.agent_forward(c.agent, c.args, c.batch, imagine=True, group_a=c.group_a,
                                            use_gt_factors=c.args.train_gt_factors,
                                            use_rand_gt_factors=c.args.train_rand_gt_factors)
            _close(q3, c.fwd["imagine_out"])
            assert torch.equal(wm, c.fwd["wmask"]) and torch.equal(im, c.fwd["imask"])


@pytest.mark.parametrize("name", CASES)
def test_greedy_actions_match_reference(name):
    c = load_learner_case(name)
    with torch.no_grad():
        q = lo.agent_forward(c.agent, c.args, c.batch)          # full-sequence forward == stepwise with carried h
    B, T, na, A = q.shape
    _close(q, c.greedy_q)
    acts = torch.stack([lo.greedy_actions(q[:, t], c.batch["avail_actions"][:, t]) for t in range(T)], 1)
    assert torch.equal(acts, c.greedy_actions)


@pytest.mark.parametrize("name", CASES)
def test_train_step_matches_reference(name):
    c = load_learner_case(name)
    out = lo.train_step(c.agent, c.mixer, c.tagent, c.tmixer, c.batch, c.args, group_a=c.group_a)
    s = c.stats
    assert abs(float(out["loss"]) - s["loss"]) <= 1e-5 * max(1.0, abs(s["loss"]))
    if "imagine" in c.args.agent:
        assert abs(float(out["aux"]["im_loss"]) - s["im_loss"]) <= 1e-5 * max(1.0, abs(s["im_loss"]))
    assert abs(float(out["grad_norm"]) - s["grad_norm"]) <= 1e-4 * max(1.0, s["grad_norm"])
    for k in ("td_error_abs", "q_taken_mean", "target_mean"):
        assert abs(float(out["aux"][k]) - s[k]) <= 1e-5 * max(1.0, abs(s[k])), k
    for k in ("ingroup_prop", "gt_ingroup_prop"):          # test_gt_factors logging pass (q_learner.py:98-105,138-147)
        if k in s:
            assert abs(float(out["aux"][k]) - s[k]) <= 1e-6, k
    for k, g in c.grad_agent.items():
        _close(out["grads_agent"][k], g, rtol=1e-4, atol=1e-6)
    for k, g in c.grad_mixer.items():
        _close(out["grads_mixer"][k], g, rtol=1e-4, atol=1e-6)
    # RMSprop's first step is ~ lr*10*sign(g): compare with an absolute tolerance of 2% of that step
    for k, p in c.new_agent.items():
        assert float((out["new_agent"][k] - p).abs().max()) <= 1e-4
    for k, p in c.new_mixer.items():
        assert float((out["new_mixer"][k] - p).abs().max()) <= 1e-4


def test_imagine_mask_closed_form_properties():
    g = torch.tensor([[0, 1, 1, 0, 1]], dtype=torch.uint8)
    ina = torch.tensor([[0, 0, 1, 0, 1]], dtype=torch.uint8)
    m = lo.imagine_masks(g, ina, 5)
    w, i = m["within_noobs"][0], m["interact_noobs"][0]
    assert int(i[0, 0]) == 1 and int(w[0, 0]) == 0          # an entity interacts with itself only in 'within'
    assert int(w[0, 3]) == 0 and int(i[0, 3]) == 1          # same group
    assert int(w[0, 1]) == 1 and int(i[0, 1]) == 0          # different groups
    assert bool((w[2] == 1).all()) and bool((i[:, 4] == 1).all())   # inactive rows/cols masked in both


def test_epsilon_schedule():
    a = lo.default_args()
    assert lo.epsilon(a, 0) == 1.0 and lo.epsilon(a, 10 ** 9) == 0.05
    assert abs(lo.epsilon(a, 250000) - 0.525) < 1e-12
