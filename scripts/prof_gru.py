"""The GRU scans of the north-star step (online: 3 mask copies x 128 episodes x 8 agents = 3072 sequences; the 16-episode
strong-scaling shard: 384), once each after a warm-up: target of `ncu --set full`, and a CUDA-event timing table."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from refil_b200 import ops

DEV = "cuda:0"
T, na, r = 60, 8, 64
Whh, bhh = torch.randn(3 * r, r, device=DEV) * 0.1, torch.randn(3 * r, device=DEV) * 0.1
for n_seq in (3072, 384, 1024, 128):
    R = n_seq * T
    GI = torch.randn(R, 3 * r, device=DEV)
    HS, gates = torch.empty(R, r, device=DEV), torch.empty(R, 4 * r, device=DEV)
    dHS, dGI, dGH = torch.randn(R, r, device=DEV), torch.empty(R, 3 * r, device=DEV), torch.empty(R, 3 * r, device=DEV)
    for it in range(3):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        ops.gru_scan_fwd(GI, Whh, bhh, None, HS, gates, n_seq, T, na)
        e[1].record()
        ops.gru_scan_bwd(dHS, gates, HS, None, Whh, dGI, dGH, n_seq, T, na)
        e[2].record()
        torch.cuda.synchronize()
    print("n_seq %5d T %d: fwd %.1f us (%.2f us/step)  bwd %.1f us (%.2f us/step)" % (
        n_seq, T, 1e3 * e[0].elapsed_time(e[1]), 1e3 * e[0].elapsed_time(e[1]) / T, 1e3 * e[1].elapsed_time(e[2]),
        1e3 * e[1].elapsed_time(e[2]) / T))
