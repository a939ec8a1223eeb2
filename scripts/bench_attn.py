"""Micro-benchmark of the masked-attention kernels on the north-star shapes (CUDA events, L2 flushed between reps)."""
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


B, T, ne, na, d, H = 128, 60, 24, 8, 128, 4
N = B * T
qkv = torch.randn(N * ne, 3 * d, device=DEV)
obs = (torch.rand(N, ne, ne, device=DEV) < 0.2).to(torch.uint8)
em = torch.zeros(N, ne, dtype=torch.uint8, device=DEV)
gb = (torch.rand(B, ne, device=DEV) < 0.5).to(torch.uint8)
unit = 4 * d * (2 * ne + 2 * na) + na * ne
for name, copies in (("agent C=3", [(obs, ne * ne, 0), (obs, ne * ne, 1), (obs, ne * ne, 2)]),
                     ("hyper C=1", [(None, 0, 8)]), ("hyper_w1 C=3", [(None, 0, 8), (None, 0, 5), (None, 0, 6)])):
    C = len(copies)
    out = torch.empty(C, N, na, d, device=DEV)
    dout = torch.randn(C, N, na, d, device=DEV)
    dqkv = torch.empty(N * ne, 3 * d, device=DEV)
    tf = timeit(lambda: ops.masked_attn_fwd(qkv, out, copies, gb, em, N, T, ne, na, d, H))
    tb = timeit(lambda: ops.masked_attn_bwd(qkv, dout, dqkv, copies, gb, em, N, T, ne, na, d, H))
    act_f = N * (ne * 3 * d * 4 + C * na * d * 4)
    act_b = N * (2 * ne * 3 * d * 4 + C * na * d * 4)
    print("%-13s fwd %.3f ms: algorithmic %.0f GB/s (%.2f of 6554), actual traffic %.0f GB/s | bwd %.3f ms: actual %.0f GB/s" % (
        name, tf, C * N * unit / tf / 1e6, C * N * unit / tf / 1e6 / 6554.6, act_f / tf / 1e6, tb, act_b / tb / 1e6))
