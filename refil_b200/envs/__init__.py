"""Environment registry (reference: src/envs/__init__.py:15-21).  StarCraft II wrappers are out of scope (they drive an
external game binary); `group_matching` is the device-batched implementation."""
from functools import partial

from .group_matching import GroupMatching, GroupMatchingBatch


def env_fn(env, **kwargs):
    return env(**kwargs)


REGISTRY = {"group_matching": partial(env_fn, env=GroupMatching)}
BATCHED_REGISTRY = {"group_matching": GroupMatchingBatch}
