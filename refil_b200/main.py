"""Command-line entry point with the reference's surface (src/main.py:66-102):

    python src/main.py --env-config=group_matching --config=refil_group_matching with env_args.n_agents=4 t_max=20000 seed=1

`with k=v` overrides use sacred's dotted-key syntax; `seed` seeds numpy, torch and the environments (main.py:29-31)."""
import os
import random
import sys

import numpy as np
import torch

from .config import build_config
from .run import run
from .utils.logging import get_logger


def _pop_option(params, name):
    for i, v in enumerate(params):
        if v.split("=")[0] == name:
            if "=" in v:
                value = v.split("=", 1)[1]
                del params[i]
            else:
                value = params[i + 1]
                del params[i:i + 2]
            return value
    return None


def main(argv=None):
    params = list(sys.argv[1:] if argv is None else argv)
    env_config = _pop_option(params, "--env-config")
    alg_config = _pop_option(params, "--config")
    if env_config is None or alg_config is None:
        raise SystemExit("usage: main.py --env-config=<env> --config=<alg> [with key=value ...]")
    overrides = [p for p in params if p != "with"]
    config = build_config(env_config, alg_config, overrides)
    seed = config.get("seed")
    if seed is None:
        seed = random.SystemRandom().randrange(1, 2 ** 31 - 1)     # sacred draws a seed when none is given
    config["seed"] = int(seed)
    np.random.seed(config["seed"])
    torch.manual_seed(config["seed"])
    config["env_args"]["seed"] = config["seed"]
    console = get_logger()
    jsonl = None
    if not config.get("evaluate", False):
        os.makedirs(os.path.join(config["local_results_path"], "stats"), exist_ok=True)
        jsonl = os.path.join(config["local_results_path"], "stats", "%s_%s_seed%d.jsonl" % (config["env"], alg_config, config["seed"]))
    run(config, console, jsonl)


if __name__ == "__main__":
    main()
