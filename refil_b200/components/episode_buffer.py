"""Device-resident episode storage with the reference's EpisodeBatch / ReplayBuffer interface.

Interface mirror of /root/reference/src/components/episode_buffer.py:6-246: a dict of dense tensors
`(batch, time, [group members], *vshape)` plus the reserved `filled` mask; `batch[key]`, `batch[(k1, k2)]`,
`batch[bs_index, time_slice]`, `update`, `max_t_filled`, `to`; the ReplayBuffer is a ring over episodes with uniform
sampling without replacement (`np.random.choice(n, B, replace=False)`, :239).

Differences by design (B200-first): storage lives on the CUDA device from the start (the env kernel writes rollouts
straight into it), `update` accepts device tensors without a host round trip, and sampling is an on-device row gather."""
from types import SimpleNamespace

import numpy as np
import torch


def _as_shape(v):
    return (v,) if isinstance(v, int) else tuple(v)


class EpisodeBatch:
    def __init__(self, scheme, groups, batch_size, max_seq_length, data=None, preprocess=None, device="cpu", zero_init=True):
        self.scheme = dict(scheme)
        self.groups = groups
        self.batch_size = batch_size
        self.max_seq_length = max_seq_length
        self.preprocess = {} if preprocess is None else preprocess
        self.device = device
        self._alloc = torch.zeros if zero_init else torch.empty       # zero_init=False: the caller overwrites every tensor at once
        if data is not None:
            self.data = data
        else:
            self.data = SimpleNamespace(transition_data={}, episode_data={})
            self._allocate(self.scheme, groups, batch_size, max_seq_length, self.preprocess)

    # ---- allocation -------------------------------------------------------------------------------------------
    def _allocate(self, scheme, groups, batch_size, max_seq_length, preprocess):
        for src, (dst, transforms) in (preprocess or {}).items():
            if src not in scheme:
                raise KeyError("preprocess source %r is not in the scheme" % src)
            vshape, dtype = scheme[src]["vshape"], scheme[src].get("dtype", torch.float32)
            for tr in transforms:
                vshape, dtype = tr.infer_output_info(vshape, dtype)
            entry = {"vshape": vshape, "dtype": dtype}
            for inherit in ("group", "episode_const"):
                if inherit in scheme[src]:
                    entry[inherit] = scheme[src][inherit]
            self.scheme[dst] = entry
        if "filled" in scheme:
            raise KeyError('"filled" is a reserved key for masking.')
        scheme = dict(self.scheme)
        scheme["filled"] = {"vshape": (1,), "dtype": torch.long}
        self.scheme = scheme
        for key, info in scheme.items():
            if "vshape" not in info:
                raise KeyError("Scheme must define vshape for %s" % key)
            shape = _as_shape(info["vshape"])
            group = info.get("group")
            if group:
                if group not in groups:
                    raise KeyError("Group %s must have its number of members defined in groups" % group)
                shape = (groups[group],) + shape
            dtype = info.get("dtype", torch.float32)
            if info.get("episode_const", False):
                self.data.episode_data[key] = self._alloc((batch_size,) + shape, dtype=dtype, device=self.device)
            else:
                self.data.transition_data[key] = self._alloc((batch_size, max_seq_length) + shape, dtype=dtype,
                                                             device=self.device)

    def to(self, device):
        for store in (self.data.transition_data, self.data.episode_data):
            for k in store:
                store[k] = store[k].to(device)
        self.device = device

    # ---- writes -----------------------------------------------------------------------------------------------
    def update(self, data, bs=slice(None), ts=slice(None), mark_filled=True):
        idx = self._parse_slices((bs, ts))
        for k, v in data.items():
            if k in self.data.transition_data:
                target, where = self.data.transition_data, tuple(idx)
                if mark_filled:
                    target["filled"][where] = 1
                    mark_filled = False
            elif k in self.data.episode_data:
                target, where = self.data.episode_data, idx[0]
            else:
                raise KeyError("%s not found in transition or episode data" % k)
            dtype = self.scheme[k].get("dtype", torch.float32)
            if torch.is_tensor(v):
                v = v.to(device=self.device, dtype=dtype)
            else:
                v = torch.as_tensor(np.asarray(v), dtype=dtype, device=self.device)
            dest = target[k][where]
            self._check_safe_view(v, dest)
            target[k][where] = v.reshape(dest.shape)
            if k in self.preprocess:
                new_k, transforms = self.preprocess[k]
                out = target[k][where]
                for tr in transforms:
                    out = tr.transform(out)
                target[new_k][where] = out.reshape(target[new_k][where].shape)

    @staticmethod
    def _check_safe_view(v, dest):
        i = v.dim() - 1
        for s in reversed(dest.shape):
            if i < 0 or v.shape[i] != s:
                if s != 1:
                    raise ValueError("Unsafe reshape of %s to %s" % (tuple(v.shape), tuple(dest.shape)))
            else:
                i -= 1

    # ---- reads ------------------------------------------------------------------------------------------------
    def __getitem__(self, item):
        if isinstance(item, str):
            if item in self.data.episode_data:
                return self.data.episode_data[item]
            if item in self.data.transition_data:
                return self.data.transition_data[item]
            raise ValueError(item)
        if isinstance(item, tuple) and all(isinstance(it, str) for it in item):
            nd = SimpleNamespace(transition_data={}, episode_data={})
            for key in item:
                if key in self.data.transition_data:
                    nd.transition_data[key] = self.data.transition_data[key]
                elif key in self.data.episode_data:
                    nd.episode_data[key] = self.data.episode_data[key]
                else:
                    raise KeyError("Unrecognised key %s" % key)
            scheme = {k: self.scheme[k] for k in item}
            groups = {self.scheme[k]["group"]: self.groups[self.scheme[k]["group"]] for k in item
                      if "group" in self.scheme[k]}
            return EpisodeBatch(scheme, groups, self.batch_size, self.max_seq_length, data=nd, device=self.device)
        idx = self._parse_slices(item)
        nd = SimpleNamespace(transition_data={}, episode_data={})
        b_idx = idx[0]
        if isinstance(b_idx, (list, np.ndarray)):
            b_idx = torch.as_tensor(np.asarray(b_idx), dtype=torch.long, device=self.device)
        for k, v in self.data.transition_data.items():
            nd.transition_data[k] = v[b_idx][:, idx[1]] if torch.is_tensor(b_idx) else v[b_idx, idx[1]]
        for k, v in self.data.episode_data.items():
            nd.episode_data[k] = v[b_idx]
        nb = self._count(idx[0], self.batch_size)
        nt = self._count(idx[1], self.max_seq_length)
        return EpisodeBatch(self.scheme, self.groups, nb, nt, data=nd, device=self.device)

    @staticmethod
    def _count(index, size):
        if isinstance(index, (list, np.ndarray)):
            return len(index)
        if torch.is_tensor(index):
            return int(index.numel())
        lo, hi, step = index.indices(size)
        return max(0, 1 + (hi - lo - 1) // step)

    @staticmethod
    def _parse_slices(items):
        if isinstance(items, (slice, int, list, np.ndarray)) or torch.is_tensor(items):
            items = (items, slice(None))
        if isinstance(items[1], list):
            raise IndexError("Indexing across Time must be contiguous")
        return [slice(it, it + 1) if isinstance(it, int) else it for it in items]

    def max_t_filled(self):
        return torch.sum(self.data.transition_data["filled"], 1).max(0)[0]

    def __repr__(self):
        return "EpisodeBatch. Batch Size:{} Max_seq_len:{} Keys:{} Groups:{}".format(
            self.batch_size, self.max_seq_length, self.scheme.keys(), self.groups.keys())


class ReplayBuffer(EpisodeBatch):
    def __init__(self, scheme, groups, buffer_size, max_seq_length, preprocess=None, device="cpu"):
        super().__init__(scheme, groups, buffer_size, max_seq_length, preprocess=preprocess, device=device)
        self.buffer_size = buffer_size
        self.buffer_index = 0
        self.episodes_in_buffer = 0

    def insert_episode_batch(self, ep_batch):
        n = ep_batch.batch_size
        if self.buffer_index + n > self.buffer_size:       # wrap: split at the end of the ring
            left = self.buffer_size - self.buffer_index
            self.insert_episode_batch(ep_batch[0:left, :])
            self.insert_episode_batch(ep_batch[left:, :])
            return
        rows = slice(self.buffer_index, self.buffer_index + n)
        for k, v in ep_batch.data.transition_data.items():
            self.data.transition_data[k][rows, :ep_batch.max_seq_length] = v
        for k, v in ep_batch.data.episode_data.items():
            self.data.episode_data[k][rows] = v
        self.buffer_index += n
        self.episodes_in_buffer = max(self.episodes_in_buffer, self.buffer_index)
        self.buffer_index %= self.buffer_size

    def can_sample(self, batch_size):
        return self.episodes_in_buffer >= batch_size

    def sample(self, batch_size):
        if not self.can_sample(batch_size):
            raise AssertionError("not enough episodes in the buffer")
        if self.episodes_in_buffer == batch_size:
            return self[:batch_size]
        ep_ids = np.random.choice(self.episodes_in_buffer, batch_size, replace=False)
        return self[ep_ids]

    def __repr__(self):
        return "ReplayBuffer. {}/{} episodes. Keys:{} Groups:{}".format(
            self.episodes_in_buffer, self.buffer_size, self.scheme.keys(), self.groups.keys())
