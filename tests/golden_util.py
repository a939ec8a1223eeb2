"""Load tests/golden/learner_*.npz (produced by the reference, see tests/golden/make_golden.py)."""
import glob
import os
from types import SimpleNamespace

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def learner_cases():
    return sorted(os.path.basename(p)[len("learner_"):-4] for p in glob.glob(os.path.join(GOLDEN, "learner_*.npz")))


def _pyval(a):
    v = a.item() if a.shape == () else a
    if isinstance(v, str) and v == "None":
        return None
    return v


def load_learner_case(name):
    z = np.load(os.path.join(GOLDEN, "learner_%s.npz" % name), allow_pickle=False)
    args = {k[4:]: _pyval(z[k]) for k in z.files if k.startswith("arg_")}
    args = SimpleNamespace(**args)
    B, T, na, ne, ed, A = [int(x) for x in z["meta_dims"]]

    def grab(prefix):
        return {k[len(prefix):]: torch.from_numpy(z[k].copy()) for k in z.files if k.startswith(prefix)}

    agent, mixer = grab("agent_"), grab("mixer_")
    tagent, tmixer = grab("tagent_"), grab("tmixer_")
    if "meta_compact" in z.files and int(z["meta_compact"]):
        # compact fixture: the target networks are the online ones plus the generator's seeded perturbation
        # (make_golden.py: 0.05 * randn, parameter order of target_mac then target_mixer, generator seed + 99)
        gen = torch.Generator().manual_seed(int(z["meta_seed"]) + 99)
        tagent, tmixer = {}, {}
        for src, dst in ((agent, tagent), (mixer, tmixer)):
            for k, v in src.items():
                dst[k] = v.clone() if k.endswith("scale_factor") else v + 0.05 * torch.randn(v.shape, generator=gen)
    case = SimpleNamespace(
        name=name, args=args, dims=(B, T, na, ne, ed, A), compact="meta_compact" in z.files and bool(int(z["meta_compact"])),
        batch=grab("in_"), agent=agent, tagent=tagent, mixer=mixer, tmixer=tmixer,
        group_a=torch.from_numpy(z["group_a"].copy()),
        fwd=grab("fwd_"), greedy_actions=torch.from_numpy(z["greedy_actions"].copy()),
        greedy_q=torch.from_numpy(z["greedy_q"].copy()),
        stats={k[5:]: float(z[k]) for k in z.files if k.startswith("stat_")},
        grad_agent=grab("grad_agent_"), grad_mixer=grab("grad_mixer_"),
        new_agent=grab("new_agent_"), new_mixer=grab("new_mixer_"),
    )
    return case
