"""Recorded-trajectory ingestion: a synthetic file in the StarCraft II tensor layout
(starcraft2custom.py:1024-1135: 8 ally + 8 enemy slots, 39 features, 14 actions = 3-8csz) round-trips through
save_episodes / load_episodes into the ReplayBuffer; layout violations are rejected.  The CPU tests exercise the host logic
on a CPU-resident buffer; the GPU test loads into the device-resident buffer and trains one step from it."""
import numpy as np
import pytest
import torch

from refil_b200.components import replay_io
from refil_b200.components.episode_buffer import EpisodeBatch, ReplayBuffer
from refil_b200.utils.synthetic import entity_scheme, synthetic_replay

NA, NE, ED, A, T = 8, 16, 39, 14, 21


def _make_file(path, E=10, device="cpu"):
    scheme, groups, preprocess = entity_scheme(NA, NE, ED, A)
    syn = synthetic_replay(E, T, NA, NE, ED, A, seed=5, pad=True)
    batch = EpisodeBatch(scheme, groups, E, T, preprocess=preprocess, device=device)
    for k, v in syn.items():
        batch.data.transition_data[k][:] = v.to(device)
    assert replay_io.save_episodes(batch, path, episode_limit=T - 1) == E
    return syn, (scheme, groups, preprocess)


def test_round_trip_cpu_buffer(tmp_path):
    path = str(tmp_path / "sc2_3-8csz.npz")
    syn, (scheme, groups, preprocess) = _make_file(path)
    z, info = replay_io.read_header(path)
    assert info == {"n_episodes": 10, "max_seq_length": T, "n_agents": NA, "n_entities": NE, "n_actions": A,
                    "entity_shape": ED, "episode_limit": T - 1, "gt_mask_avail": False}
    buf = ReplayBuffer(scheme, groups, 16, T + 4, preprocess=preprocess, device="cpu")      # longer episodes allowed: padded
    assert replay_io.load_episodes(path, buf, chunk=4) == 10
    assert buf.episodes_in_buffer == 10 and buf.can_sample(10)
    got = buf.sample(10)
    for k, v in syn.items():
        assert torch.equal(got[k][:, :T].cpu(), v), k
        assert not bool(got[k][:, T:].any()), k                  # the padding steps are empty / unfilled
    # a second file wraps around the ring exactly like insert_episode_batch (episode_buffer.py:213-228)
    assert replay_io.load_episodes(path, buf, chunk=7) == 10
    assert buf.episodes_in_buffer == 16 and buf.buffer_index == 4


@pytest.mark.parametrize("what", ["visible_absent", "nonzero_absent", "bad_action", "filled_gap", "shape", "format"])
def test_layout_violations_are_rejected(tmp_path, what):
    path = str(tmp_path / "x.npz")
    _make_file(path)
    z = dict(np.load(path))
    em = z["entity_mask"].astype(bool)
    e, t, j = np.argwhere(em)[0]
    if what == "visible_absent":
        z["obs_mask"][e, t, 0, j] = 0
    elif what == "nonzero_absent":
        z["entities"][e, t, j, 3] = 0.5
    elif what == "bad_action":
        z["actions"][0, 0, 0, 0] = A
    elif what == "filled_gap":
        z["filled"][0, 1, 0] = 0
    elif what == "shape":
        z["reward"] = z["reward"][:, :-1]
    elif what == "format":
        z["meta_format"] = np.array("something else")
    np.savez_compressed(path, **z)
    scheme, groups, preprocess = entity_scheme(NA, NE, ED, A)
    buf = ReplayBuffer(scheme, groups, 16, T, preprocess=preprocess, device="cpu")
    with pytest.raises(replay_io.ReplayFormatError):
        replay_io.load_episodes(path, buf)
    assert buf.episodes_in_buffer == 0


def test_buffer_mismatch_is_rejected(tmp_path):
    path = str(tmp_path / "x.npz")
    _make_file(path)
    scheme, groups, preprocess = entity_scheme(NA, NE + 1, ED, A)
    with pytest.raises(replay_io.ReplayFormatError):
        replay_io.load_episodes(path, ReplayBuffer(scheme, groups, 16, T, preprocess=preprocess, device="cpu"))
    scheme, groups, preprocess = entity_scheme(NA, NE, ED, A)
    with pytest.raises(replay_io.ReplayFormatError):
        replay_io.load_episodes(path, ReplayBuffer(scheme, groups, 16, T - 1, preprocess=preprocess, device="cpu"))


@pytest.mark.gpu
def test_load_into_device_buffer_and_train(tmp_path):
    """File -> pinned host chunks -> device ReplayBuffer -> sample -> one REFIL train step equal to the oracle's on the same
    episodes (loss within 1e-4): recorded trajectories replace config 5's synthetic tensors with no change to the learner."""
    from gpu_util import LoggerStub
    from oracle import learner_oracle as lo
    from refil_b200.controllers import REGISTRY as mac_REGISTRY
    from refil_b200.learners import REGISTRY as le_REGISTRY
    dev = "cuda:0"
    path = str(tmp_path / "sc2.npz")
    E = 6
    syn, (scheme, groups, preprocess) = _make_file(path, E=E)
    buf = ReplayBuffer(scheme, groups, 8, T, preprocess=preprocess, device=dev)
    assert replay_io.load_episodes(path, buf, chunk=4) == E
    sample = buf.sample(E)                                       # == all episodes, buffer order
    for k, v in syn.items():
        assert torch.equal(sample[k].cpu(), v), k
    args = lo.default_args()
    args.n_agents, args.n_actions, args.n_entities, args.entity_shape = NA, A, NE, ED
    args.mac, args.learner, args.agent_output_type, args.action_selector = "entity_mac", "q_learner", "q", "epsilon_greedy"
    args.target_update_interval, args.learner_log_interval, args.gt_mask_avail, args.device = 200, 1, False, dev
    gen = torch.Generator().manual_seed(3)
    ap, mp = lo.init_agent_params(gen, args, ED + A), lo.init_mixer_params(gen, args, ED + A)
    group_a = (torch.rand(E, NE, generator=gen) < 0.5).to(torch.uint8)
    ref = lo.train_step(ap, mp, ap, mp, syn, args, group_a=group_a)
    mac = mac_REGISTRY[args.mac](buf.scheme, groups, args)
    logger = LoggerStub()
    learner = le_REGISTRY[args.learner](mac, buf.scheme, logger, args)
    mac.agent.load_state_dict(ap)
    learner.target_mac.agent.load_state_dict(ap)
    learner.mixer.load_state_dict(mp)
    learner.target_mixer.load_state_dict(mp)
    learner.train(sample, t_env=10, episode_num=0, group_bits=group_a.to(dev))
    torch.cuda.synchronize()
    loss = float(ref["loss"])
    assert abs(logger.stats["loss"][0] - loss) <= 1e-4 * max(1.0, abs(loss))
