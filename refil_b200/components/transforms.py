"""Batch preprocessing transforms (reference: src/components/transforms.py:11-23)."""
import torch


class Transform:
    def transform(self, tensor):
        raise NotImplementedError

    def infer_output_info(self, vshape_in, dtype_in):
        raise NotImplementedError


class OneHot(Transform):
    def __init__(self, out_dim):
        self.out_dim = out_dim

    def transform(self, tensor):
        y = torch.zeros(*tensor.shape[:-1], self.out_dim, dtype=torch.float32, device=tensor.device)
        y.scatter_(-1, tensor.long(), 1.0)
        return y

    def infer_output_info(self, vshape_in, dtype_in):
        return (self.out_dim,), torch.float32
