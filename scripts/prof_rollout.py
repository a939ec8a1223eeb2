"""One eager rollout of BASELINE config 2 (group_matching 4 agents, refil_group_matching, 4096 envs) after a warm-up: target of an
ncu launch list (`--metrics gpu__time_duration.sum`), to see which kernels a rollout timestep spends its time in."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from types import SimpleNamespace
from refil_b200.components.episode_buffer import EpisodeBatch
from refil_b200.config import build_config
from refil_b200.controllers import REGISTRY as mac_REGISTRY
from refil_b200.runners import REGISTRY as r_REGISTRY
from refil_b200.utils.synthetic import entity_scheme


class Log:
    class console_logger:
        @staticmethod
        def info(*a, **k):
            pass

    def log_stat(self, *a, **k):
        pass


alg, n_envs, na, ne = (sys.argv[1:] + ["refil_group_matching", "4096", "4", "4"])[:4]
over = ["batch_size_run=%s" % n_envs, "env_args.n_agents=%s" % na, "env_args.episode_limit=%s" % os.environ.get("LIMIT", "50")]
if ne != na:
    over.append("env_args.n_entities=%s" % ne)
cfg = build_config("group_matching", alg, over)
cfg["env_args"]["seed"] = 0
args = SimpleNamespace(**cfg)
args.device, args.rank = "cuda:0", 0
runner = r_REGISTRY[args.runner](args=args, logger=Log())
info = runner.get_env_info()
args.n_agents, args.n_actions, args.entity_shape, args.n_entities = info["n_agents"], info["n_actions"], info["entity_shape"], info["n_entities"]
args.gt_mask_avail, args.entity_scheme = True, True
scheme, groups, preprocess = entity_scheme(args.n_agents, args.n_entities, args.entity_shape, args.n_actions, gt_mask=True)
proto = EpisodeBatch(scheme, groups, 1, 2, preprocess=preprocess, device="cuda:0")
torch.manual_seed(0)
mac = mac_REGISTRY[args.mac](proto.scheme, groups, args)
runner.setup(scheme=scheme, groups=groups, preprocess=preprocess, mac=mac)
runner.run(test_mode=False)
torch.cuda.synchronize()
torch.cuda.profiler.start()
runner.run(test_mode=False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
