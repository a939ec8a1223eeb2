"""Synthetic replay tensors in the EpisodeBatch layout (SURVEY.md section 8d, config-3 recipe), generated on the host
with a seeded torch generator.  Used by bench.py and the multi-GPU tests; the product path only ever sees the tensors.

entities ~ U[0,1); obs_mask ~ Bernoulli(0.2) with a clear diagonal; per-episode padded suffix of ally / other slots
(entity_mask = 1, never observable, zero features); avail ~ Bernoulli(0.7) with action 1 always available; actions
uniform over the available ones; reward ~ N(0,1); each episode ends at a random step in [T/2, T) -- 70 % by
termination, the rest by the time limit."""
import torch


def synthetic_replay(B, T, na, ne, ed, A, seed=0, gt_mask=False, pad=True):
    g = torch.Generator().manual_seed(int(seed))
    ents = torch.rand(B, T, ne, ed, generator=g)
    obs = (torch.rand(B, T, ne, ne, generator=g) < 0.2).to(torch.uint8)
    idx = torch.arange(ne)
    obs[:, :, idx, idx] = 0
    em = torch.zeros(B, T, ne, dtype=torch.uint8)
    if pad:
        ka = torch.randint(max(1, na - 2), na + 1, (B,), generator=g)
        slot = torch.arange(ne).unsqueeze(0)
        dead = (slot >= ka.unsqueeze(1)) & (slot < na)
        if ne > na:
            ke = torch.randint(max(1, ne - na - 3), ne - na + 1, (B,), generator=g)
            dead |= slot >= (na + ke).unsqueeze(1)
        em[:] = dead.to(torch.uint8).unsqueeze(1)
    obs = ((obs + em.unsqueeze(2) + em.unsqueeze(3)) > 0).to(torch.uint8)
    ents = ents * (1 - em).unsqueeze(-1).float()
    avail = (torch.rand(B, T, na, A, generator=g) < 0.7).to(torch.int32)
    avail[..., min(1, A - 1)] = 1
    actions = torch.multinomial(avail.float().reshape(-1, A) + 1e-9, 1, generator=g).reshape(B, T, na, 1)
    onehot = torch.zeros(B, T, na, A).scatter_(3, actions, 1.0)
    reward = torch.randn(B, T, 1, generator=g)
    L = torch.randint(max(1, T // 2), T, (B,), generator=g)            # env steps per episode, <= T-1
    tt = torch.arange(T).unsqueeze(0)
    filled = (tt <= L.unsqueeze(1)).to(torch.int64).unsqueeze(-1)
    ended = torch.rand(B, generator=g) < 0.7
    term = ((tt == (L - 1).unsqueeze(1)) & ended.unsqueeze(1)).to(torch.uint8).unsqueeze(-1)
    batch = dict(entities=ents, obs_mask=obs, entity_mask=em, actions=actions, actions_onehot=onehot,
                 avail_actions=avail, reward=reward, terminated=term, filled=filled)
    if gt_mask:
        batch["gt_mask"] = (torch.rand(B, T, na, ne, generator=g) < 0.5).to(torch.uint8)
    return batch


def entity_scheme(na, ne, ed, A, gt_mask=False):
    """The scheme / groups / preprocess triple of run.py:178-196 for entity environments."""
    from ..components.transforms import OneHot
    scheme = {
        "entities": {"vshape": ed, "group": "entities"},
        "obs_mask": {"vshape": ne, "group": "entities", "dtype": torch.uint8},
        "entity_mask": {"vshape": ne, "dtype": torch.uint8},
        "actions": {"vshape": (1,), "group": "agents", "dtype": torch.long},
        "avail_actions": {"vshape": (A,), "group": "agents", "dtype": torch.int},
        "reward": {"vshape": (1,)},
        "terminated": {"vshape": (1,), "dtype": torch.uint8},
    }
    if gt_mask:
        scheme["gt_mask"] = {"vshape": ne, "group": "agents", "dtype": torch.uint8}
    groups = {"agents": na, "entities": ne}
    preprocess = {"actions": ("actions_onehot", [OneHot(out_dim=A)])}
    return scheme, groups, preprocess
