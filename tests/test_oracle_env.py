"""Pins oracle/gm_env.c: (1) the numpy legacy RandomState draw model, (2) the reference GroupMatching transcripts
committed in tests/golden/gm_transcripts.npz (generated from /root/reference by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle.gm_env_oracle import GroupMatchingOracle, MT19937Oracle


@pytest.mark.parametrize("seed", [0, 1, 42, 12345, 2 ** 32 - 1])
def test_mt19937_matches_numpy_randomstate(seed):
    rs = np.random.RandomState(seed)
    o = MT19937Oracle(seed)
    ctl = np.random.RandomState((seed + 5) % 2 ** 32)
    for _ in range(3000):
        k = ctl.randint(0, 4)
        if k == 0:
            assert rs.uniform() == o.uniform()
        elif k == 1:
            m = int(ctl.randint(1, 70))
            assert rs.randint(0, m) == o.bounded(m - 1)
        elif k == 2:
            n = int(ctl.randint(1, 12))
            a = list(range(n))
            rs.shuffle(a)
            b = list(range(n))
            for i in range(n - 1, 0, -1):
                j = o.bounded(i)
                b[i], b[j] = b[j], b[i]
            assert a == b
        else:
            m = int(ctl.randint(1, 9))
            assert list(rs.randint(0, m, size=6)) == [o.bounded(m - 1) for _ in range(6)]


def _cases(golden_dir):
    z = np.load(os.path.join(golden_dir, "gm_transcripts.npz"))
    return z, [str(c) for c in z["cases"]]


def test_env_oracle_matches_reference_transcripts(golden_dir):
    z, cases = _cases(golden_dir)
    assert len(cases) >= 20
    for key in cases:
        na, ns, ng, lim = [int(x) for x in z[key + "_cfg"]]
        env = GroupMatchingOracle(n_agents=na, n_states=ns, n_groups=ng, rand_trans=float(z[key + "_rt"]),
                                  episode_limit=lim, seed=int(z[key + "_seed"]))
        acts, rew, flags = z[key + "_actions"], z[key + "_reward"], z[key + "_flags"]
        locs, ents, gts, resets = z[key + "_locs"], z[key + "_entities"], z[key + "_gt"], set(z[key + "_resets"].tolist())
        row = 0
        for i in range(len(acts) + 1):
            if i in resets or i == 0:
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
            assert np.array_equal(env.get_masks()[2], gts[row])
            row += 1
        assert row == len(locs)


def test_env_oracle_edge_cases():
    # single agent / single state / single group: always solved at reset, first step terminates
    env = GroupMatchingOracle(n_agents=1, n_states=1, n_groups=1, rand_trans=1.0, episode_limit=5, seed=3)
    env.reset()
    r, done, info = env.step([0])
    assert done and info["solved"] and r == -0.1
    # episode limit wins when nothing is solved
    env = GroupMatchingOracle(n_agents=6, n_states=9, n_groups=2, rand_trans=0.0, episode_limit=2, seed=5)
    env.reset()
    r, d1, i1 = env.step([1] * 6)
    r, d2, i2 = env.step([1] * 6)
    assert d2 and i2.get("episode_limit", False)
    assert env.get_env_info()["n_entities"] == 6 and env.get_entity_size() == 9 + 2 + 6
