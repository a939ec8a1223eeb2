"""Micro-benchmark of the dense-layer kernels on the shapes of the north-star step (CUDA events, L2 flushed between reps)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from refil_b200 import ops

DEV = "cuda:0"
flush = torch.empty(256 * 1024 * 1024 // 4, device=DEV)


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


shapes = [("in_trans", 184320, 384, 128), ("out_trans C=3", 184320, 128, 128), ("out_trans C=1", 61440, 128, 128),
          ("fc2 agent", 184320, 64, 128), ("gi", 184320, 192, 64), ("fc2 hyper", 61440, 32, 128)]
only = sys.argv[1] if len(sys.argv) > 1 else None
for name, M, N, K in shapes:
    if only and only not in name:
        continue
    A, W, b = torch.randn(M, K, device=DEV), torch.randn(N, K, device=DEV), torch.randn(N, device=DEV)
    out = torch.empty(M, N, device=DEV)
    dA = torch.empty(M, K, device=DEV)
    res = {}
    for tc in (False, True):
        ops.USE_TENSOR_CORES = tc
        res[("fwd", tc)] = timeit(lambda: ops.linear_fwd(A, W, b, out))
        res[("bwd", tc)] = timeit(lambda: ops.linear_bwd_data(out, W, dA))
    fl = 2.0 * M * N * K
    byts = 4.0 * (M * K + M * N)
    print("%-14s M=%6d N=%3d K=%3d | fwd simt %.3f ms (%.1f TF)  tc %.3f ms (%.1f TF, %.0f GB/s) | bwd-data simt %.3f  tc %.3f ms (%.1f TF)" % (
        name, M, N, K, res[("fwd", False)], fl / res[("fwd", False)] / 1e9, res[("fwd", True)], fl / res[("fwd", True)] / 1e9,
        byts / res[("fwd", True)] / 1e6, res[("bwd", False)], res[("bwd", True)], fl / res[("bwd", True)] / 1e9))

print("--- weight gradient dW[P,Q] = dC[M,P]^T A[M,Q] (+db) ---")
for name, M, P, Q in [("in_trans", 184320, 384, 128), ("out_trans C=3", 184320, 128, 128), ("fc1 (packed)", 184320, 128, 64),
                      ("gi", 184320, 192, 64), ("fc2 hyper", 61440, 32, 128)]:
    if only and only not in name:
        continue
    dC, A, y = torch.randn(M, P, device=DEV), torch.randn(M, Q, device=DEV), torch.randn(M, P, device=DEV)
    dW, db = torch.zeros(P, Q, device=DEV), torch.zeros(P, device=DEV)
    res = {}
    for tc in (False, True):
        ops.USE_TENSOR_CORES = tc
        res[tc] = timeit(lambda: ops.linear_bwd_weight(dC, A, dW, db))
        res[(tc, "relu")] = timeit(lambda: ops.linear_bwd_weight(dC, A, dW, db, relu_y=y))
    byts = 4.0 * (M * P + M * Q)
    print("%-14s M=%6d P=%3d Q=%3d | simt %.3f ms  tc %.3f ms (%.0f GB/s) | with relu': simt %.3f  tc %.3f ms" % (
        name, M, P, Q, res[False], res[True], byts / res[True] / 1e6, res[(False, "relu")], res[(True, "relu")]))
