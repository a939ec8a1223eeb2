"""GPU parity: the batched Group Matching kernel vs the CPU oracle (oracle/gm_env.c) and the reference transcripts
(tests/golden/gm_transcripts.npz), bit-exact, through the C ABI."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _batch(E, T, na, ne, ed, dev):
    return dict(
        entities=torch.zeros(E, T, ne, ed, device=dev), obs_mask=torch.zeros(E, T, ne, ne, dtype=torch.uint8, device=dev),
        entity_mask=torch.zeros(E, T, ne, dtype=torch.uint8, device=dev),
        gt_mask=torch.zeros(E, T, na, ne, dtype=torch.uint8, device=dev),
        avail_actions=torch.zeros(E, T, na, 3, dtype=torch.int32, device=dev),
        actions=torch.zeros(E, T, na, 1, dtype=torch.int64, device=dev), reward=torch.zeros(E, T, 1, device=dev),
        terminated=torch.zeros(E, T, 1, dtype=torch.uint8, device=dev),
        filled=torch.zeros(E, T, 1, dtype=torch.int64, device=dev))


CONFIGS = [dict(n_agents=4, n_states=6, n_groups=2, rand_trans=0.1, episode_limit=50),
           dict(n_agents=8, n_states=6, n_groups=2, rand_trans=0.1, episode_limit=50),
           dict(n_agents=8, n_states=6, n_groups=3, rand_trans=0.3, episode_limit=20),
           dict(n_agents=1, n_states=1, n_groups=1, rand_trans=1.0, episode_limit=3),
           dict(n_agents=5, n_states=3, n_groups=4, rand_trans=0.5, episode_limit=7),
           dict(n_agents=16, n_states=5, n_groups=2, rand_trans=0.0, episode_limit=12)]


@pytest.mark.parametrize("cfg", CONFIGS)
@pytest.mark.parametrize("E,seed", [(1, 7), (150, 1234)])
def test_batched_env_matches_oracle(cfg, E, seed):
    from oracle.gm_env_oracle import GroupMatchingOracle
    from refil_b200.envs.group_matching import GroupMatchingBatch
    dev = torch.device("cuda:0")
    na, lim = cfg["n_agents"], cfg["episode_limit"]
    T = lim + 1
    env = GroupMatchingBatch(E, seed=seed, device=dev, **cfg)
    ed = env.get_entity_size()
    oracles = [GroupMatchingOracle(seed=seed + i, **cfg) for i in range(E)]
    rng = np.random.RandomState(99)
    for episode in range(3):            # RNG streams must continue across resets
        b = _batch(E, T, na, na, ed, dev)
        env.reset(b)
        tape = rng.randint(0, 3, size=(E, T, na, 1))
        b["actions"].copy_(torch.from_numpy(tape))
        for ts in range(lim):
            env.step(b, ts)
        torch.cuda.synchronize()
        h = {k: v.cpu().numpy() for k, v in b.items()}
        est = env.est.cpu().numpy()
        ret = env.ep_ret.cpu().numpy()
        n_steps = 0
        for i, o in enumerate(oracles):
            o.reset()
            assert np.array_equal(np.stack(o.get_entities()), h["entities"][i, 0]), (i, "reset entities")
            assert np.array_equal(o.get_masks()[2], h["gt_mask"][i, 0]), (i, "gt_mask")
            assert h["filled"][i, 0, 0] == 1 and (h["avail_actions"][i, 0] == 1).all()
            done, ts, total = False, 0, 0.0
            while not done:
                r, done, info = o.step(tape[i, ts, :, 0])
                total += r
                assert np.float32(r) == h["reward"][i, ts, 0], (i, ts, r, h["reward"][i, ts, 0])
                lim_hit = bool(info.get("episode_limit", False))
                assert h["terminated"][i, ts, 0] == int(done and not lim_hit), (i, ts)
                assert h["filled"][i, ts + 1, 0] == 1
                assert np.array_equal(np.stack(o.get_entities()), h["entities"][i, ts + 1]), (i, ts, "entities")
                assert (h["avail_actions"][i, ts + 1] == 1).all()
                ts += 1
            n_steps += ts
            assert (h["filled"][i, ts + 1:] == 0).all() and (h["entities"][i, ts + 1:] == 0).all()
            assert est[3, i] == ts and est[1, i] == ts
            assert bool(est[2, i] & 4) == bool(info["solved"])
            assert bool(est[2, i] & 8) == bool(info.get("episode_limit", False))
            assert ret[i] == total
            assert np.array_equal(o.get_locs(), env.loc[:, i].cpu().numpy())
        assert int(env.step_counter.item()) >= n_steps
    assert (h["obs_mask"] == 0).all() and (h["entity_mask"] == 0).all()


def test_padded_entities_config4():
    """8 agents / 24 entity slots (BASELINE config 4): slots 8..23 are masked padding, the first 8 equal the ne=8 env."""
    from refil_b200.envs.group_matching import GroupMatchingBatch
    dev = torch.device("cuda:0")
    cfg = dict(n_agents=8, n_states=6, n_groups=2, rand_trans=0.1, episode_limit=10)
    E, T = 33, 11
    a = GroupMatchingBatch(E, seed=5, device=dev, **cfg)
    p = GroupMatchingBatch(E, seed=5, device=dev, n_entities=24, **cfg)
    ba, bp = _batch(E, T, 8, 8, a.get_entity_size(), dev), _batch(E, T, 8, 24, p.get_entity_size(), dev)
    a.reset(ba)
    p.reset(bp)
    tape = torch.randint(0, 3, (E, T, 8, 1), device=dev)
    ba["actions"].copy_(tape)
    bp["actions"].copy_(tape)
    for ts in range(10):
        a.step(ba, ts)
        p.step(bp, ts)
    assert torch.equal(ba["entities"], bp["entities"][:, :, :8])
    assert torch.equal(ba["reward"], bp["reward"]) and torch.equal(ba["terminated"], bp["terminated"])
    assert (bp["entities"][:, :, 8:] == 0).all()
    f = bp["filled"].bool().squeeze(-1)
    assert (bp["entity_mask"][f][:, 8:] == 1).all() and (bp["entity_mask"][f][:, :8] == 0).all()
    om = bp["obs_mask"][f]
    assert (om[:, 8:, :] == 1).all() and (om[:, :, 8:] == 1).all() and (om[:, :8, :8] == 0).all()
    assert torch.equal(ba["gt_mask"][:, 0], bp["gt_mask"][:, 0, :, :8]) and (bp["gt_mask"][:, 0, :, 8:] == 1).all()


def test_single_env_api_matches_reference_transcripts(golden_dir):
    """The reference-facing single-env class replays the transcripts recorded from the reference's GroupMatching."""
    from refil_b200.envs import REGISTRY
    z = np.load(os.path.join(golden_dir, "gm_transcripts.npz"))
    cases = [str(c) for c in z["cases"]]
    for key in cases[::3]:
        na, ns, ng, lim = [int(x) for x in z[key + "_cfg"]]
        env = REGISTRY["group_matching"](n_agents=na, n_states=ns, n_groups=ng, rand_trans=float(z[key + "_rt"]),
                                         episode_limit=lim, seed=int(z[key + "_seed"]))
        acts, rew, flags = z[key + "_actions"], z[key + "_reward"], z[key + "_flags"]
        locs, ents, gts, resets = z[key + "_locs"], z[key + "_entities"], z[key + "_gt"], set(z[key + "_resets"].tolist())
        row = 0
        for i in range(len(acts) + 1):
            if i in resets:
                env.reset()
                assert np.array_equal(env.get_locs(), locs[row]), (key, i)
                assert np.array_equal(np.stack(env.get_entities()), ents[row])
                assert np.array_equal(env.get_masks()[2], gts[row])
                row += 1
            if i == len(acts):
                break
            r, done, info = env.step(acts[i])
            f = int(done) | (int(info["solved"]) << 1) | (int(info.get("episode_limit", False)) << 2)
            assert r == rew[i] and f == flags[i], (key, i, r, rew[i], f, flags[i])
            assert np.array_equal(env.get_locs(), locs[row])
            assert np.array_equal(np.stack(env.get_entities()), ents[row])
            row += 1
        assert row == len(locs)
