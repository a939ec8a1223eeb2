"""Flat fp32 parameter / gradient storage with named views (reference `state_dict` key names).

All parameters of the learner (agent + mixer) live in ONE contiguous buffer so that the optimiser is one kernel and
the multi-GPU gradient exchange is one all-reduce.  A store can be created stand-alone (the controller creates the
agent before the learner exists) and later re-bound onto a slice of a larger buffer."""
import math
from collections import OrderedDict

import torch


def _numel(shape):
    n = 1
    for s in shape:
        n *= int(s)
    return n


class ParamStore:
    def __init__(self, specs, device):
        """specs: OrderedDict name -> shape (declaration order == reference `parameters()` order)."""
        self.specs = OrderedDict((k, tuple(int(x) for x in v)) for k, v in specs.items())
        self.device = torch.device(device)
        self.size = sum(_numel(s) for s in self.specs.values())
        self.padded_size = (self.size + 3) // 4 * 4          # keep every store 16-byte aligned inside a flat buffer
        self.flat = torch.zeros(self.padded_size, dtype=torch.float32, device=self.device)
        self.grad = torch.zeros(self.padded_size, dtype=torch.float32, device=self.device)
        self.buffers = OrderedDict()      # non-trainable state_dict entries (e.g. attn.scale_factor), kept on the host
        self._make_views()

    def _make_views(self):
        self.p, self.g = OrderedDict(), OrderedDict()
        off = 0
        for k, shp in self.specs.items():
            n = _numel(shp)
            self.p[k] = self.flat[off:off + n].view(shp)
            self.g[k] = self.grad[off:off + n].view(shp)
            off += n

    def rebind(self, flat_slice, grad_slice):
        """Move the storage onto slices of a larger flat buffer (values are carried over)."""
        assert flat_slice.numel() == self.padded_size and grad_slice.numel() == self.padded_size
        flat_slice.copy_(self.flat)
        grad_slice.copy_(self.grad)
        self.flat, self.grad = flat_slice, grad_slice
        self._make_views()

    def to(self, device):
        device = torch.device(device)
        if device != self.flat.device:
            self.device = device
            self.flat = self.flat.to(device)
            self.grad = self.grad.to(device)
            self._make_views()
        return self

    # ---- reference-compatible views -------------------------------------------------------------------------
    def state_dict(self):
        sd = OrderedDict((k, v) for k, v in self.p.items())
        sd.update(self.buffers)
        return sd

    def load_state_dict(self, sd, strict=True):
        missing = [k for k in self.p if k not in sd]
        extra = [k for k in sd if k not in self.p and k not in self.buffers]
        if strict and (missing or extra):
            raise KeyError("state_dict mismatch: missing %s, unexpected %s" % (missing, extra))
        with torch.no_grad():
            for k, v in self.p.items():
                if k in sd:
                    src = sd[k]
                    if tuple(src.shape) != tuple(v.shape):
                        raise ValueError("shape mismatch for %s: %s vs %s" % (k, tuple(src.shape), tuple(v.shape)))
                    v.copy_(src.to(device=v.device, dtype=v.dtype))

    def parameters(self):
        return list(self.p.values())

    def named_parameters(self):
        return list(self.p.items())

    def named_grads(self):
        return list(self.g.items())


def init_linear_(gen, w, b=None):
    """nn.Linear default init: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias
    (kaiming_uniform_(a=sqrt(5)) reduces to that bound)."""
    bound = 1.0 / math.sqrt(w.shape[1])
    w.copy_(((torch.rand(w.shape, generator=gen) * 2 - 1) * bound).to(w.device))
    if b is not None:
        b.copy_(((torch.rand(b.shape, generator=gen) * 2 - 1) * bound).to(b.device))


def init_uniform_(gen, t, bound):
    t.copy_(((torch.rand(t.shape, generator=gen) * 2 - 1) * bound).to(t.device))


class Workspace:
    """Tag-keyed cache of device scratch tensors.  A tag owns ONE allocation sized for the largest request seen so far and
    hands out a view of its head, so (a) the same (tag, shape) always returns the same address -- a whole training step can be
    captured in a CUDA graph and re-launched -- and (b) a run whose batches change length (run.py truncates every sample to
    its longest filled episode) holds one activation set, not one per distinct T.  `generation` counts re-allocations: an
    address that moved invalidates every graph captured against it (QLearner drops its graph cache when it changes).
    Requests of different shapes under one tag alias each other: tags are only shared by uses that are ordered on one
    stream and do not outlive the step."""

    def __init__(self, device):
        self.device = torch.device(device)
        self._bufs = {}
        self.generation = 0

    def get(self, tag, shape, dtype=torch.float32, zero=False):
        shape = tuple(int(s) for s in shape)
        n = _numel(shape)
        key = (tag, dtype)
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < n:
            if buf is not None:
                self.generation += 1
            buf = torch.empty(max(n, 1), dtype=dtype, device=self.device)
            self._bufs[key] = buf
        t = buf[:n].view(shape)
        if zero:
            t.zero_()
        return t

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self._bufs.values())
