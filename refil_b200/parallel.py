"""Multi-GPU plumbing: one process per GPU, `torch.distributed` (NCCL over NVLink / NVSwitch) for the single collective
the path needs.

The hot path shards naturally: environment instances by index, replay by episodes.  Each rank runs
forward / backward on its own episodes and contributes [flat grads of sum((mask*td)^2) terms | sum(mask), loss sums]; ONE
all-reduce later every rank divides by the global sum(mask), clips and applies RMSprop to its replica
(`QLearner.train`).  The reference has no distributed code at all (single `RMSprop` over one parameter list,
/root/reference/src/learners/q_learner.py:37)."""
import os

import torch
import torch.distributed as dist


def init_distributed(backend=None, device=None):
    """Initialise the default process group from the torchrun environment (RANK / WORLD_SIZE / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1 or (dist.is_available() and dist.is_initialized()):
        return rank_world()
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kw = {}
    if backend == "nccl" and device is not None:
        kw["device_id"] = torch.device(device)
    dist.init_process_group(backend, **kw)
    return rank_world()


def rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n, rank, world):
    """Contiguous [lo, hi) slice of n items for `rank`; the first n % world ranks get one extra item."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_sum_(flat):
    """In-place sum over ranks of the [grads | stats] buffer; no-op on a single rank."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


def broadcast_(flat, src=0):
    """Make every replica start from rank `src`'s parameters."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat, src=src)
    return flat


def all_reduce_sum_int(value, device=None):
    """Sum of a host integer over the ranks (rollout step counts: `t_env` must be the same on every rank, or the ranks leave
    the training loop at different iterations and dead-lock in the gradient all-reduce)."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return int(value)
    dev = device if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([int(value)], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def all_reduce_max_int(value, device=None):
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return int(value)
    dev = device if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([int(value)], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item())
