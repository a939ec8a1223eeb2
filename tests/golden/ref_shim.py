"""Import shim for the mounted reference (/root/reference) -- used ONLY by tests/golden/make_golden.py
in the build container.  Nothing in tests/, bench.py or smoke() imports this at run time: the GPU box has
no /root/reference.  No reference file is modified:
  shim 1: pre-register an empty ``envs`` package so group_matching imports without pysc2;
  shim 2: torch>=2 rejects uint8 masks in masked_fill -> cast uint8 masks to bool.
"""
import sys
import types

import torch

REF_SRC = "/root/reference/src"


def install():
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    if "envs" not in sys.modules:
        pkg = types.ModuleType("envs")
        pkg.__path__ = [REF_SRC + "/envs"]
        sys.modules["envs"] = pkg
    if not getattr(torch.Tensor, "_refil_mf_patched", False):
        _orig = torch.Tensor.masked_fill

        def masked_fill(self, mask, value):
            if mask.dtype == torch.uint8:
                mask = mask.bool()
            return _orig(self, mask, value)

        torch.Tensor.masked_fill = masked_fill
        torch.Tensor._refil_mf_patched = True


def group_matching_cls():
    install()
    from envs.group_matching.group_matching import GroupMatching
    return GroupMatching
