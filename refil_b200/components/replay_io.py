"""Recorded-trajectory ingestion: StarCraft II (or any entity-scheme) episodes recorded with the reference into the
device-resident ReplayBuffer.

The StarCraft II environment itself is out of scope (external game binary).  What the learner consumes from it is the
tensor layout of `StarCraft2CustomEnv.get_entities / get_masks / get_avail_actions`
(/root/reference/src/envs/starcraft2/starcraft2custom.py:1024-1135, 1137-1150), as the reference's runner stores it in an
EpisodeBatch (src/runners/parallel_runner.py:140-197, scheme src/run.py:178-196):

    entities       f32 [E, T, ne, ed]    ne = max_n_agents + max_n_enemies slots; rows of absent units are all-zero
    obs_mask       u8  [E, T, ne, ne]    1 = "row entity cannot see column entity"; rows / columns of absent slots are 1
    entity_mask    u8  [E, T, ne]        1 = slot not in the scenario (allies [n_agents, max_n_agents), enemies likewise)
    avail_actions  i32 [E, T, na, A]
    actions        i64 [E, T, na, 1]
    reward         f32 [E, T, 1]
    terminated     u8  [E, T, 1]         excludes time-limit endings (parallel_runner.py:180-183)
    filled         i64 [E, T, 1]
    [gt_mask       u8  [E, T, na, ne]]   only environments that provide it

File format: one `.npz` (numpy, no pickling) holding those arrays plus `meta_n_agents`, `meta_n_entities`, `meta_n_actions`,
`meta_entity_shape`, `meta_episode_limit` and a format tag.  On the reference side the recorder is four lines around the
replay buffer it already has (INTEGRATION.md shows them); `save_episodes` below is the same thing for batches of this package.

`load_episodes` streams a file into the ReplayBuffer in chunks through pinned host memory (the buffer can be far larger than a
chunk: 5000 episodes x 150 steps of 3-8csz is ~4 GB on the device), validating the layout invariants above first."""
import numpy as np
import torch

FORMAT = "refil_b200.replay.v1"
KEYS = ("entities", "obs_mask", "entity_mask", "avail_actions", "actions", "reward", "terminated", "filled")
OPTIONAL = ("gt_mask",)
_DTYPES = {"entities": np.float32, "obs_mask": np.uint8, "entity_mask": np.uint8, "avail_actions": np.int32,
           "actions": np.int64, "reward": np.float32, "terminated": np.uint8, "filled": np.int64, "gt_mask": np.uint8}


class ReplayFormatError(ValueError):
    pass


def save_episodes(batch, path, episode_limit=None):
    """Write the transition tensors of an EpisodeBatch / ReplayBuffer (only its filled episodes) to `path` (.npz)."""
    n = getattr(batch, "episodes_in_buffer", batch.batch_size)
    td = batch.data.transition_data
    out = {}
    for k in KEYS + OPTIONAL:
        if k in td:
            out[k] = td[k][:n].detach().cpu().numpy().astype(_DTYPES[k], copy=False)
        elif k in KEYS:
            raise ReplayFormatError("batch has no %r tensor" % k)
    E, T, ne, ed = out["entities"].shape
    na, A = out["avail_actions"].shape[2], out["avail_actions"].shape[3]
    out.update(meta_format=np.array(FORMAT), meta_n_agents=np.int64(na), meta_n_entities=np.int64(ne),
               meta_n_actions=np.int64(A), meta_entity_shape=np.int64(ed),
               meta_episode_limit=np.int64(episode_limit if episode_limit is not None else T - 1))
    np.savez_compressed(path, **out)
    return E


def read_header(path):
    z = np.load(path, allow_pickle=False)
    if "meta_format" not in z.files or str(z["meta_format"]) != FORMAT:
        raise ReplayFormatError("%s is not a %s file" % (path, FORMAT))
    miss = [k for k in KEYS if k not in z.files]
    if miss:
        raise ReplayFormatError("%s lacks %s" % (path, miss))
    E, T, ne, ed = z["entities"].shape
    info = {"n_episodes": E, "max_seq_length": T, "n_agents": int(z["meta_n_agents"]), "n_entities": int(z["meta_n_entities"]),
            "n_actions": int(z["meta_n_actions"]), "entity_shape": int(z["meta_entity_shape"]),
            "episode_limit": int(z["meta_episode_limit"]), "gt_mask_avail": "gt_mask" in z.files}
    if (ne, ed) != (info["n_entities"], info["entity_shape"]):
        raise ReplayFormatError("entities %s disagree with the header %s" % ((ne, ed), info))
    return z, info


def validate_layout(arrs, info, sc2=True):
    """Shape / dtype checks, then the invariants of get_masks / get_entities that the kernels rely on.  Raises ReplayFormatError."""
    E, T, ne, ed = arrs["entities"].shape
    na, A = info["n_agents"], info["n_actions"]
    want = {"entities": (E, T, ne, ed), "obs_mask": (E, T, ne, ne), "entity_mask": (E, T, ne), "avail_actions": (E, T, na, A),
            "actions": (E, T, na, 1), "reward": (E, T, 1), "terminated": (E, T, 1), "filled": (E, T, 1)}
    if "gt_mask" in arrs:
        want["gt_mask"] = (E, T, na, ne)
    for k, shp in want.items():
        if tuple(arrs[k].shape) != shp:
            raise ReplayFormatError("%s has shape %s, expected %s" % (k, tuple(arrs[k].shape), shp))
        if arrs[k].dtype != _DTYPES[k]:
            raise ReplayFormatError("%s has dtype %s, expected %s" % (k, arrs[k].dtype, np.dtype(_DTYPES[k])))
    if not (0 < na <= ne <= 32):
        raise ReplayFormatError("n_agents=%d / n_entities=%d outside the kernels' range (<= 32 entities)" % (na, ne))
    filled = arrs["filled"][..., 0].astype(bool)
    if not np.isfinite(arrs["entities"]).all() or not np.isfinite(arrs["reward"]).all():
        raise ReplayFormatError("non-finite entities / rewards")
    if (np.diff(filled.astype(np.int8), axis=1) > 0).any() or not filled[:, 0].all():
        raise ReplayFormatError("`filled` must be a prefix of every episode")
    acts = arrs["actions"][..., 0]
    if (acts < 0).any() or (acts >= A).any():
        raise ReplayFormatError("actions outside [0, %d)" % A)
    steps = filled & np.concatenate([filled[:, 1:], np.zeros((E, 1), bool)], axis=1)     # steps that have a successor
    em = arrs["entity_mask"].astype(bool)
    picked = np.take_along_axis(arrs["avail_actions"], arrs["actions"].astype(np.int64), axis=3)[..., 0]
    if ((picked == 0) & steps[..., None] & ~em[:, :, :na]).any():
        raise ReplayFormatError("a stored action of a present agent is not among its available actions")
    if sc2:
        # starcraft2custom.py:1033-1055: absent slots are unobservable both ways; :1117-1126: and all-zero feature rows
        f = filled[..., None]
        if ((arrs["obs_mask"] == 0) & (em[:, :, :, None] | em[:, :, None, :]) & f[..., None]).any():
            raise ReplayFormatError("obs_mask leaves an absent entity slot observable")
        if (np.abs(arrs["entities"]).sum(-1) * (em & f) != 0).any():
            raise ReplayFormatError("an absent entity slot has non-zero features")
        if (em != em[:, :1]).any():
            raise ReplayFormatError("entity_mask changes inside an episode (the scenario's slots are fixed at reset)")
    return True


def load_episodes(path, buffer, chunk=256, validate=True, sc2=True):
    """Stream the episodes of `path` into `buffer` (a ReplayBuffer of this package, normally device-resident).  Episodes shorter
    than the buffer's max_seq_length are zero-padded (unfilled).  Returns the number of episodes inserted."""
    z, info = read_header(path)
    arrs = {k: z[k] for k in KEYS + OPTIONAL if k in z.files}
    if validate:
        validate_layout(arrs, info, sc2=sc2)
    sch = buffer.scheme
    ne_b = buffer.groups.get("entities")
    na_b = buffer.groups.get("agents")
    if (na_b, ne_b) != (info["n_agents"], info["n_entities"]):
        raise ReplayFormatError("file has %d agents / %d entities, buffer %s / %s" % (info["n_agents"], info["n_entities"], na_b, ne_b))
    ed_b = sch["entities"]["vshape"]
    ed_b = ed_b if isinstance(ed_b, int) else ed_b[0]
    if ed_b != info["entity_shape"]:
        raise ReplayFormatError("file has entity_shape %d, buffer %d" % (info["entity_shape"], ed_b))
    E, T = info["n_episodes"], info["max_seq_length"]
    if T > buffer.max_seq_length:
        raise ReplayFormatError("episodes of %d steps do not fit max_seq_length=%d" % (T, buffer.max_seq_length))
    use_gt = "gt_mask" in arrs and "gt_mask" in buffer.data.transition_data
    cuda = torch.device(buffer.device).type == "cuda"
    for lo in range(0, E, chunk):
        hi = min(E, lo + chunk)
        data = {}
        for k in KEYS + (("gt_mask",) if use_gt else ()):
            t = torch.from_numpy(np.ascontiguousarray(arrs[k][lo:hi]))
            if cuda:
                t = t.pin_memory().to(buffer.device, non_blocking=True)
            data[k] = t
        part = _batch_like(buffer, hi - lo, T)
        filled = data.pop("filled")
        part.update(data, mark_filled=False)             # runs the preprocess (actions -> actions_onehot) on the device
        part.data.transition_data["filled"][:] = filled
        buffer.insert_episode_batch(part)
    return E


def _batch_like(buffer, n, T):
    """An empty EpisodeBatch with the buffer's (already pre-processed) scheme."""
    from .episode_buffer import EpisodeBatch
    scheme = {k: v for k, v in buffer.scheme.items() if k != "filled" and k not in {d for d, _ in buffer.preprocess.values()}}
    return EpisodeBatch(scheme, buffer.groups, n, T, preprocess=buffer.preprocess, device=buffer.device)
