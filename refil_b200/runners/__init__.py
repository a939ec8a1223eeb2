"""Runner registry (reference: src/runners/__init__.py:4-7).  Both runners are the device-batched loop; "episode" keeps the
reference EpisodeRunner's per-step gt_mask write (normally with batch_size_run = 1)."""
from .parallel_runner import EpisodeRunner, ParallelRunner

REGISTRY = {"parallel": ParallelRunner, "episode": EpisodeRunner}
