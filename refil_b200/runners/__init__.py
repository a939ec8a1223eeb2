"""Runner registry (reference: src/runners/__init__.py:4-7).  Both keys map to the device-batched runner; "episode" is the
same loop with batch_size_run = 1."""
from .parallel_runner import ParallelRunner

REGISTRY = {"parallel": ParallelRunner, "episode": ParallelRunner}
