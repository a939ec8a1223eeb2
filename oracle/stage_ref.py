"""TEST / BENCH INFRASTRUCTURE -- not product code.  Recipe that stages the reference's own hot-path sources where the GPU box
can run them.

The reference is pure Python (nothing to compile), and /root/reference does not exist on the GPU box.  `stage()` copies the
files SURVEY.md section 8c lists -- learner, controllers, agents, mixers, attention layer, replay buffer, Group Matching env,
config YAMLs -- from /root/reference/src into `oracle/_ref/src/`, byte for byte.  `oracle/_ref/` is git-ignored (reference
sources never enter this repository's history) but not gpurun-ignored, so the staged copy travels with the snapshot like the
built .so files.  `install_shims()` is the import recipe of SURVEY.md section 8c / BASELINE.md section 3 (two run-time shims,
no reference file is modified):
  1. an empty `envs` package whose __path__ points at the staged `envs/` (skips `envs/__init__.py`, which imports pysc2);
  2. `torch.Tensor.masked_fill` casts uint8 masks to bool (torch >= 2 rejects the reference's uint8 masks,
     modules/layers/attention.py:57).
Users: `__graft_entry__.build()` (staging), `bench.py --impl reference` / `cpu_baseline` (timing the reference's own
QLearner.train and GroupMatching.step on the host cores).  Nothing under refil_b200/ imports this.
"""
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"
STAGE = os.path.join(HERE, "_ref")
STAGE_SRC = os.path.join(STAGE, "src")

# (directory under src/, file names) -- the hot path only: no StarCraft II wrappers, no sacred / tensorboard plumbing
FILES = [
    ("components", ["__init__.py", "episode_buffer.py", "transforms.py", "action_selectors.py", "epsilon_schedules.py"]),
    ("controllers", ["__init__.py", "basic_controller.py", "entity_controller.py"]),
    ("learners", ["__init__.py", "q_learner.py"]),
    ("modules", ["__init__.py"]),
    ("modules/agents", ["__init__.py", "rnn_agent.py", "ff_agent.py", "entity_rnn_agent.py", "entity_ff_agent.py"]),
    ("modules/layers", ["__init__.py", "attention.py"]),
    ("modules/mixers", ["__init__.py", "vdn.py", "qmix.py", "flex_qmix.py"]),
    ("envs", ["multiagentenv.py"]),
    ("envs/group_matching", ["__init__.py", "group_matching.py"]),
    ("config", ["default.yaml"]),
    ("config/algs", ["refil.yaml", "qmix_atten.yaml", "refil_group_matching.yaml", "qmix_atten_group_matching.yaml",
                     "refil_vdn.yaml", "vdn_atten.yaml"]),
    ("config/envs", ["group_matching.yaml"]),
]


def stage(force=False):
    """Copy the listed files from /root/reference/src to oracle/_ref/src.  Returns the staged path, or None when the reference
    is not mounted (GPU box: the copy made in the build container is used as it is)."""
    src_root = os.path.join(REF_ROOT, "src")
    if not os.path.isdir(src_root):
        return STAGE_SRC if os.path.isdir(STAGE_SRC) else None
    for sub, names in FILES:
        dst_dir = os.path.join(STAGE_SRC, sub)
        os.makedirs(dst_dir, exist_ok=True)
        for n in names:
            s, d = os.path.join(src_root, sub, n), os.path.join(dst_dir, n)
            if not os.path.exists(s):
                if n == "__init__.py":          # namespace marker only
                    open(d, "a").close()
                    continue
                raise FileNotFoundError(s)
            if force or not os.path.exists(d) or os.path.getmtime(s) > os.path.getmtime(d):
                shutil.copyfile(s, d)
    with open(os.path.join(STAGE, "README"), "w") as f:
        f.write("Staged copy of shariqiqbal2810/REFIL hot-path sources (oracle/stage_ref.py); git-ignored, bench/test use only.\n")
    return STAGE_SRC


def available():
    return os.path.exists(os.path.join(STAGE_SRC, "learners", "q_learner.py"))


def install_shims():
    """Make `controllers`, `learners`, `components`, `modules`, `envs.group_matching` importable from the staged copy."""
    import torch
    if not available():
        raise RuntimeError("oracle/_ref is not staged (run __graft_entry__.build() where /root/reference is mounted)")
    if STAGE_SRC not in sys.path:
        sys.path.insert(0, STAGE_SRC)
    if "envs" not in sys.modules:
        pkg = types.ModuleType("envs")
        pkg.__path__ = [os.path.join(STAGE_SRC, "envs")]
        sys.modules["envs"] = pkg
    if not getattr(torch.Tensor, "_refil_mf_patched", False):
        _orig = torch.Tensor.masked_fill

        def masked_fill(self, mask, value):
            if mask.dtype == torch.uint8:
                mask = mask.bool()
            return _orig(self, mask, value)

        torch.Tensor.masked_fill = masked_fill
        torch.Tensor._refil_mf_patched = True


class _Console:
    def info(self, *a, **k):
        pass


class _Logger:
    def __init__(self):
        self.stats = {}
        self.console_logger = _Console()

    def log_stat(self, key, value, t, to_sacred=True):
        self.stats.setdefault(key, []).append(float(value))


def ref_args(alg, **over):
    """default.yaml <- algs/<alg>.yaml <- overrides, as src/main.py:72-84 merges them."""
    import yaml
    from types import SimpleNamespace
    cfg_dir = os.path.join(STAGE_SRC, "config")
    with open(os.path.join(cfg_dir, "default.yaml")) as f:
        cfg = yaml.safe_load(f)
    with open(os.path.join(cfg_dir, "algs", alg + ".yaml")) as f:
        cfg.update(yaml.safe_load(f))
    cfg.update(over)
    return SimpleNamespace(**cfg)


def build_reference_learner(alg, tensors, dims, gt=False, seed=0, **over):
    """The reference's own EpisodeBatch + EntityMAC + QLearner on the CPU for a dict of EpisodeBatch-layout tensors.
    -> (learner, batch, logger).  Scheme as src/run.py:178-196."""
    import torch as th
    install_shims()
    from components.episode_buffer import EpisodeBatch
    from components.transforms import OneHot
    from controllers import REGISTRY as mac_REGISTRY
    from learners import REGISTRY as le_REGISTRY
    B, T, na, ne, ed, A = dims
    args = ref_args(alg, **over)
    args.n_agents, args.n_actions, args.n_entities, args.entity_shape = na, A, ne, ed
    args.entity_scheme, args.gt_mask_avail, args.device = True, gt, "cpu"
    th.manual_seed(seed)
    scheme = {
        "entities": {"vshape": ed, "group": "entities"},
        "obs_mask": {"vshape": ne, "group": "entities", "dtype": th.uint8},
        "entity_mask": {"vshape": ne, "dtype": th.uint8},
        "actions": {"vshape": (1,), "group": "agents", "dtype": th.long},
        "avail_actions": {"vshape": (A,), "group": "agents", "dtype": th.int},
        "reward": {"vshape": (1,)},
        "terminated": {"vshape": (1,), "dtype": th.uint8},
    }
    if gt:
        scheme["gt_mask"] = {"vshape": ne, "group": "agents", "dtype": th.uint8}
    groups = {"agents": na, "entities": ne}
    preprocess = {"actions": ("actions_onehot", [OneHot(out_dim=A)])}
    batch = EpisodeBatch(scheme, groups, B, T, preprocess=preprocess, device="cpu")
    for k, v in tensors.items():            # entities ... actions, actions_onehot, filled: EpisodeBatch layout already
        batch.data.transition_data[k][:] = v
    mac = mac_REGISTRY[args.mac](batch.scheme, groups, args)
    logger = _Logger()
    learner = le_REGISTRY[args.learner](mac, batch.scheme, logger, args)
    return learner, batch, logger


def group_matching_env(**kw):
    install_shims()
    from envs.group_matching.group_matching import GroupMatching
    return GroupMatching(**kw)


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
