"""The three in_trans dense kernels of the north-star step, once each after a warm-up (target of `ncu --set full`)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from refil_b200 import ops

DEV = "cuda:0"
M, N, K = 184320, 384, 128
A, W, b = torch.randn(M, K, device=DEV), torch.randn(N, K, device=DEV), torch.randn(N, device=DEV)
out, dA = torch.empty(M, N, device=DEV), torch.empty(M, K, device=DEV)
dW, db = torch.zeros(N, K, device=DEV), torch.zeros(N, device=DEV)
for _ in range(2):
    ops.linear_fwd(A, W, None, out)            # in_trans forward
    ops.linear_bwd_data(out, W, dA)            # in_trans backward-data (K = 3d, split-K)
    ops.linear_bwd_weight(out, A, dW, None)    # in_trans weight gradient
torch.cuda.synchronize()
