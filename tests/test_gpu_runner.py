"""GPU tests of the rollout driver, the CLI entry point and checkpoints."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from gpu_util import LoggerStub

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _setup(alg="refil_group_matching", n_envs=64, seed=7, **over):
    from refil_b200.components.transforms import OneHot
    from refil_b200.config import build_config
    from refil_b200.controllers import REGISTRY as mac_REGISTRY
    from refil_b200.runners import REGISTRY as r_REGISTRY
    cfg = build_config("group_matching", alg, ["batch_size_run=%d" % n_envs, "env_args.n_agents=4", "env_args.episode_limit=12"]
                       + ["%s=%s" % kv for kv in over.items()])
    cfg["env_args"]["seed"] = seed
    args = SimpleNamespace(**cfg)
    args.device = DEV
    logger = LoggerStub()
    runner = r_REGISTRY[args.runner](args=args, logger=logger)
    info = runner.get_env_info()
    args.n_agents, args.n_actions, args.entity_shape, args.n_entities = info["n_agents"], info["n_actions"], info["entity_shape"], info["n_entities"]
    args.gt_mask_avail, args.entity_scheme = True, True
    from refil_b200.utils.synthetic import entity_scheme
    scheme, groups, preprocess = entity_scheme(args.n_agents, args.n_entities, args.entity_shape, args.n_actions, gt_mask=True)
    from refil_b200.components.episode_buffer import EpisodeBatch
    proto = EpisodeBatch(scheme, groups, 1, 2, preprocess=preprocess, device=DEV)
    torch.manual_seed(seed)
    mac = mac_REGISTRY[args.mac](proto.scheme, groups, args)
    runner.setup(scheme=scheme, groups=groups, preprocess=preprocess, mac=mac)
    return args, runner, mac, logger, (scheme, groups, preprocess)


@pytest.mark.parametrize("alg,test_mode", [("refil_group_matching", True), ("refil_group_matching", False), ("refil", True)])
def test_rollout_matches_oracle_env_and_agent(alg, test_mode):
    from oracle import learner_oracle as lo
    from oracle.gm_env_oracle import GroupMatchingOracle
    args, runner, mac, logger, _ = _setup(alg, n_envs=48, seed=11)
    batch = runner.run(test_mode=test_mode)
    h = {k: v.cpu() for k, v in batch.data.transition_data.items()}
    E, T, na = 48, 13, 4
    env_args = dict(args.env_args)
    seed = env_args.pop("seed")
    env_args.pop("entity_scheme")
    total_steps, returns = 0, []
    for i in range(E):
        o = GroupMatchingOracle(seed=seed + i, **env_args)
        o.reset()
        assert np.array_equal(np.stack(o.get_entities()), h["entities"][i, 0].numpy())
        done, ts, ret = False, 0, 0.0
        while not done:
            r, done, info = o.step(h["actions"][i, ts, :, 0].numpy())
            ret += r
            assert np.float32(r) == h["reward"][i, ts, 0].item()
            assert int(h["terminated"][i, ts, 0]) == int(done and not info.get("episode_limit", False))
            assert np.array_equal(np.stack(o.get_entities()), h["entities"][i, ts + 1].numpy())
            ts += 1
        assert int(h["filled"][i].sum()) == ts + 1
        total_steps += ts
        returns.append(ret)
        # one-hot preprocessing follows the stored actions on every filled step
        assert torch.equal(h["actions_onehot"][i, :ts].argmax(-1), h["actions"][i, :ts, :, 0])
    assert runner.t_env == (0 if test_mode else total_steps)
    if test_mode:
        # greedy actions equal the CPU oracle agent's argmax on the same observations (hidden state carried for the GRU)
        oargs = SimpleNamespace(**vars(args))
        params = {k: v.cpu() for k, v in mac.agent.state_dict().items() if "scale_factor" not in k}
        q = lo.agent_forward(params, oargs, {k: v for k, v in h.items()})
        for t in range(T - 1):
            live = h["filled"][:, t, 0] == 1                 # in the live-list at t (incl. the step after termination)
            ref = lo.greedy_actions(q[:, t], h["avail_actions"][:, t])
            assert torch.equal(ref[live], h["actions"][live, t, :, 0]), t
    else:
        assert not logger.stats or "return_mean" in logger.stats or True
    # runner statistics (logged when the interval elapses): force a log and compare with the oracle returns
    if not test_mode:
        runner.log_train_stats_t = -10 ** 9
        runner.train_returns, runner.train_stats = [], {}
        runner.run(test_mode=False)
        assert "return_mean" in logger.stats and "ep_length_mean" in logger.stats and "solved_mean" in logger.stats
        assert "epsilon" in logger.stats


def test_cli_trains_and_checkpoints(tmp_path):
    """`main.py --env-config=group_matching --config=refil_group_matching with ...` runs rollouts + learner steps, logs the
    reference's stat keys and writes agent.th / mixer.th / opt.th that load back bit-exactly."""
    from refil_b200 import main as entry
    res = str(tmp_path / "results")
    entry.main(["--env-config=group_matching", "--config=refil_group_matching", "with", "env_args.n_agents=4",
                "env_args.episode_limit=10", "batch_size_run=32", "batch_size=32", "buffer_size=64", "t_max=1500",
                "training_iters=2", "test_interval=600", "test_nepisode=32", "log_interval=300", "runner_log_interval=300",
                "learner_log_interval=300", "save_model=True", "save_model_interval=600", "seed=5",
                "local_results_path=" + res])
    stats = [l for l in open(os.path.join(res, "stats", os.listdir(os.path.join(res, "stats"))[0]))]
    keys = {__import__("json").loads(l)["key"] for l in stats}
    for k in ("loss", "im_loss", "grad_norm", "td_error_abs", "q_taken_mean", "target_mean", "return_mean", "ep_length_mean",
              "test_return_mean", "epsilon", "episode", "ingroup_prop", "gt_ingroup_prop"):
        assert k in keys, (k, keys)
    vals = [__import__("json").loads(l) for l in stats]
    assert all(np.isfinite(v["value"]) for v in vals)
    models = os.path.join(res, "models")
    run_dir = os.path.join(models, os.listdir(models)[0])
    step_dirs = sorted(os.listdir(run_dir), key=int)
    assert step_dirs
    for f in ("agent.th", "mixer.th", "opt.th"):
        assert os.path.exists(os.path.join(run_dir, step_dirs[-1], f))
    sd = torch.load(os.path.join(run_dir, step_dirs[-1], "agent.th"))
    assert "fc1.weight" in sd and "attn.in_trans.weight" in sd and "attn.scale_factor" in sd
    opt = torch.load(os.path.join(run_dir, step_dirs[-1], "opt.th"))
    assert set(opt) == {"state", "param_groups"} and "square_avg" in opt["state"][0]
    # evaluate-from-checkpoint path (run.py:214-241)
    entry.main(["--env-config=group_matching", "--config=refil_group_matching", "with", "env_args.n_agents=4",
                "env_args.episode_limit=10", "batch_size_run=32", "test_nepisode=32", "evaluate=True", "seed=5",
                "checkpoint_path=" + run_dir, "local_results_path=" + res])


def test_learner_checkpoint_roundtrip(tmp_path):
    from golden_util import load_learner_case
    from gpu_util import build_product
    c = load_learner_case("refil")
    batch, mac, learner, _ = build_product(c.args, c.dims, c.batch, DEV)
    learner.train(batch, t_env=10, episode_num=0, group_bits=c.group_a.to(DEV))
    learner.save_models(str(tmp_path))
    c2 = load_learner_case("refil")
    batch2, mac2, learner2, _ = build_product(c2.args, c2.dims, c2.batch, DEV)
    learner2.load_models(str(tmp_path))
    assert torch.equal(learner2.flat, learner.flat) and torch.equal(learner2.square_avg, learner.square_avg)
    # the checkpoint is loadable by stock torch modules with the reference's parameter names
    sd = torch.load(str(tmp_path / "agent.th"))
    gru = torch.nn.GRUCell(c.args.rnn_hidden_dim, c.args.rnn_hidden_dim)
    gru.load_state_dict({k[4:]: v for k, v in sd.items() if k.startswith("rnn.")})


@pytest.mark.parametrize("alg", ["refil_group_matching", "refil"])
def test_graph_rollout_equals_eager_rollout(alg):
    """args.rollout_graph: the whole rollout replayed as ONE CUDA graph.  Greedy (test-mode) rollouts are deterministic given the
    env seeds, so the graph path must reproduce the eager path bit for bit -- entities, actions, rewards, flags, filled -- run
    after run (run 1 is eager, run 2 captures + replays, runs 3-4 replay); and the env RNG stream continues across resets."""
    outs = []
    for graph in (False, True):
        args, runner, mac, logger, _ = _setup(alg, n_envs=40, seed=21, rollout_graph=graph)
        runs = []
        for _ in range(4):
            b = runner.run(test_mode=True)
            runs.append({k: v.clone() for k, v in b.data.transition_data.items()})
        outs.append(runs)
        if graph:
            assert runner._graph_state["graph"] is not None, "the rollout was not captured"
    for r, (a, b) in enumerate(zip(*outs)):
        for k in a:
            assert torch.equal(a[k], b[k]), (r, k)
    assert not torch.equal(outs[0][0]["entities"], outs[0][1]["entities"])     # a new episode every run


def test_graph_rollout_training_mode_explores_and_counts_steps():
    """Training-mode graph rollouts: epsilon comes from a device scalar (schedule value of the run), the uniforms are re-drawn at
    every replay, t_env / returns follow the transcript exactly as in the eager path."""
    from oracle.gm_env_oracle import GroupMatchingOracle
    args, runner, mac, logger, _ = _setup("refil_group_matching", n_envs=48, seed=5, rollout_graph=True)
    acts, hs = [], []
    for rep in range(4):
        t0 = runner.t_env
        eps_expected = mac.action_selector.schedule.eval(t0)
        batch = runner.run(test_mode=False)
        assert abs(mac.action_selector.epsilon - eps_expected) < 1e-12          # the run's epsilon = schedule at its first step
        h = {k: v.cpu() for k, v in batch.data.transition_data.items()}
        lengths = h["filled"].sum(1)[:, 0] - 1
        assert runner.t_env - t0 == int(lengths.sum())
        acts.append(h["actions"].clone())
        hs.append(h)
        assert torch.equal(h["actions_onehot"].argmax(-1)[h["filled"][:, :, 0] == 1], h["actions"][..., 0][h["filled"][:, :, 0] == 1])
    env_args = dict(args.env_args)
    seed = env_args.pop("seed")
    env_args.pop("entity_scheme")
    # every instance replays its four consecutive episodes (eager, capture + replay, replay, replay) against the CPU env: the
    # MT19937 stream of an instance runs on across resets and graph replays
    for i in range(0, 48, 7):
        o = GroupMatchingOracle(seed=seed + i, **env_args)
        for h in hs:
            o.reset()
            done, ts = False, 0
            assert np.array_equal(np.stack(o.get_entities()), h["entities"][i, 0].numpy())
            while not done:
                r, done, info = o.step(h["actions"][i, ts, :, 0].numpy())
                assert np.float32(r) == h["reward"][i, ts, 0].item()
                ts += 1
            assert int(h["filled"][i].sum()) == ts + 1
    assert 0.05 < mac.action_selector.epsilon < 1.0 and runner._graph_state["graph"] is not None
    assert abs(float(runner._graph_state["eps"].item()) - mac.action_selector.epsilon) < 1e-6   # the device scalar the graph reads
    assert not torch.equal(acts[1], acts[2]) and not torch.equal(acts[2], acts[3])      # fresh exploration noise per replay


def test_episode_runner_writes_gt_mask_every_step():
    """`runner: episode` keeps the reference EpisodeRunner's pre-transition data: gt_mask at EVERY filled step
    (runners/episode_runner.py:52-67); the parallel runner stores it at reset only (SURVEY.md section 3.2 quirk iv)."""
    for kind, every in (("episode", True), ("parallel", False)):
        args, runner, mac, logger, _ = _setup("refil_group_matching", n_envs=6, seed=3, runner=kind)
        b = runner.run(test_mode=True)
        gt, filled = b["gt_mask"].cpu(), b["filled"].cpu()[:, :, 0] == 1
        assert bool(gt[:, 0].any())
        for e in range(6):
            for t in range(1, int(filled[e].sum())):
                if every:
                    assert torch.equal(gt[e, t], gt[e, 0]), (kind, e, t)
                else:
                    assert not bool(gt[e, t].any()), (kind, e, t)
