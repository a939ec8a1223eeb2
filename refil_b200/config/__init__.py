"""Config assembly: default -> env preset -> alg preset -> `with k=v` overrides (reference: src/main.py:40-87)."""
import copy
import os

import yaml

_PRESETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "presets.yaml")


def _merge(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = copy.deepcopy(v)
    return dst


def _parse_value(text):
    try:
        return yaml.safe_load(text)
    except yaml.YAMLError:
        return text


def build_config(env_config, alg_config, overrides=()):
    with open(_PRESETS) as f:
        presets = yaml.safe_load(f)
    if env_config not in presets["envs"]:
        raise KeyError("unknown --env-config %r (available: %s; StarCraft II environments are out of scope)"
                       % (env_config, sorted(presets["envs"])))
    algs = {k: v for k, v in presets["algs"].items() if not k.startswith("_")}
    if alg_config not in algs:
        raise KeyError("unknown --config %r (available: %s)" % (alg_config, sorted(algs)))
    cfg = copy.deepcopy(presets["default"])
    _merge(cfg, presets["envs"][env_config])
    _merge(cfg, algs[alg_config])
    for item in overrides:                       # sacred-style `with a.b=c`
        if "=" not in item:
            raise ValueError("override %r is not of the form key=value" % item)
        key, val = item.split("=", 1)
        node = cfg
        parts = key.split(".")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = _parse_value(val)
    return cfg


def alg_names():
    with open(_PRESETS) as f:
        return sorted(k for k in yaml.safe_load(f)["algs"] if not k.startswith("_"))
