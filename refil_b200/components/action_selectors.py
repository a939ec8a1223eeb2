"""Action selection on device (reference: src/components/action_selectors.py:36-63).

Greedy index parity: unavailable actions are -inf and the FIRST maximum wins (torch `max(dim)[1]` on CPU), computed
by the `refil_select_actions` kernel.  Exploration draws two uniforms per (env, agent) with torch's device generator
(u_pick < eps -> the floor(u_act * n_avail)-th available action), which is Categorical(avail) in distribution."""
import torch

from .. import ops
from .epsilon_schedules import DecayThenFlatSchedule

REGISTRY = {}


class EpsilonGreedyActionSelector:
    def __init__(self, args):
        self.args = args
        self.schedule = DecayThenFlatSchedule(args.epsilon_start, args.epsilon_finish, args.epsilon_anneal_time,
                                              decay="linear")
        self.epsilon = self.schedule.eval(0)

    def select_action(self, agent_inputs, avail_actions, t_env, test_mode=False, est_flags=None, out=None, eps_dev=None,
                      uniforms=None):
        """agent_inputs [bs, na, A] f32, avail_actions [bs, na, A] int32 -> actions [bs, na] int64.
        avail_actions / out may be time slices of the EpisodeBatch tensors (no copies).  eps_dev: optional device scalar holding
        epsilon (the graph-captured rollout sets it per run: schedule value, or 0 in test mode) -- then t_env is not consulted."""
        q = agent_inputs.contiguous()
        bs, na, A = q.shape
        avail = avail_actions if avail_actions.dtype == torch.int32 else avail_actions.to(torch.int32)
        if avail.stride(-1) != 1 or avail.stride(-2) != A:
            avail = avail.contiguous()
        if out is None:
            out = torch.zeros(bs, na, dtype=torch.int64, device=q.device)
        u_pick = u_act = None
        if eps_dev is None:
            self.epsilon = 0.0 if test_mode else self.schedule.eval(t_env)
        if eps_dev is not None or self.epsilon > 0.0:
            u = uniforms if uniforms is not None else torch.rand(2, bs, na, device=q.device)
            u_pick, u_act = u[0], u[1]
        ops.select_actions(q, avail, u_pick, u_act, est_flags, 0.0 if eps_dev is not None else self.epsilon, out, bs, na, A,
                           eps_dev=eps_dev)
        return out


REGISTRY["epsilon_greedy"] = EpsilonGreedyActionSelector
