"""The Group Matching step kernel alone (32768 and 4096 instances x 8 agents, random action tape) after a warm-up: target of
`ncu --set full -k regex:gm_step`, plus a CUDA-event timing of graph-replayed rollouts."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from refil_b200.envs.group_matching import GroupMatchingBatch

dev = "cuda:0"
for E in (32768, 4096):
    na, lim = 8, 50
    T = lim + 1
    env = GroupMatchingBatch(E, n_agents=na, n_states=6, n_groups=2, rand_trans=0.1, episode_limit=lim, seed=0, device=dev)
    ed = env.get_entity_size()
    b = dict(entities=torch.zeros(E, T, na, ed, device=dev), gt_mask=torch.zeros(E, T, na, na, dtype=torch.uint8, device=dev),
             obs_mask=torch.zeros(E, T, na, na, dtype=torch.uint8, device=dev), entity_mask=torch.zeros(E, T, na, dtype=torch.uint8, device=dev),
             avail_actions=torch.zeros(E, T, na, 3, dtype=torch.int32, device=dev), actions=torch.randint(0, 3, (E, T, na, 1), device=dev),
             reward=torch.zeros(E, T, 1, device=dev), terminated=torch.zeros(E, T, 1, dtype=torch.uint8, device=dev),
             filled=torch.zeros(E, T, 1, dtype=torch.int64, device=dev))

    def rollout():
        env.reset(b)
        for ts in range(lim):
            env.step(b, ts)

    rollout()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        rollout()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print("E=%d: %.1f us per step launch (graph replay)" % (E, 1e3 * e0.elapsed_time(e1) / 10 / (lim + 1)))
