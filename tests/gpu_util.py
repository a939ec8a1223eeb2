"""Helpers shared by the GPU parity tests, __graft_entry__.smoke() and bench.py: build the product-side
EpisodeBatch / EntityMAC / QLearner from a golden case or from synthetic tensors."""
from collections import defaultdict

import torch


class ConsoleStub:
    def info(self, *a, **k):
        pass


class LoggerStub:
    def __init__(self):
        self.stats = defaultdict(list)
        self.console_logger = ConsoleStub()

    def log_stat(self, key, value, t, to_sacred=True):
        self.stats[key].append(float(value))


def make_scheme(na, ne, ed, A, gt):
    scheme = {
        "entities": {"vshape": ed, "group": "entities"},
        "obs_mask": {"vshape": ne, "group": "entities", "dtype": torch.uint8},
        "entity_mask": {"vshape": ne, "dtype": torch.uint8},
        "actions": {"vshape": (1,), "group": "agents", "dtype": torch.long},
        "avail_actions": {"vshape": (A,), "group": "agents", "dtype": torch.int},
        "reward": {"vshape": (1,)},
        "terminated": {"vshape": (1,), "dtype": torch.uint8},
    }
    if gt:
        scheme["gt_mask"] = {"vshape": ne, "group": "agents", "dtype": torch.uint8}
    return scheme, {"agents": na, "entities": ne}


def build_product(args, dims, tensors, device="cuda:0"):
    """-> (batch, mac, learner, logger) on `device` from a dict of CPU tensors in EpisodeBatch layout."""
    from refil_b200.components.episode_buffer import EpisodeBatch
    from refil_b200.components.transforms import OneHot
    from refil_b200.controllers import REGISTRY as mac_REGISTRY
    from refil_b200.learners import REGISTRY as le_REGISTRY
    B, T, na, ne, ed, A = dims
    args.device = device
    scheme, groups = make_scheme(na, ne, ed, A, "gt_mask" in tensors)
    preprocess = {"actions": ("actions_onehot", [OneHot(out_dim=A)])}
    batch = EpisodeBatch(scheme, groups, B, T, preprocess=preprocess, device=device)
    for k, v in tensors.items():
        batch.data.transition_data[k][:] = v.to(device)
    mac = mac_REGISTRY[args.mac](batch.scheme, groups, args)
    logger = LoggerStub()
    learner = le_REGISTRY[args.learner](mac, batch.scheme, logger, args)
    return batch, mac, learner, logger
