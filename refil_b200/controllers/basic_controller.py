"""Shared-parameter multi-agent controller over the sm_100a agent kernels.

API mirror of /root/reference/src/controllers/basic_controller.py:7-121 (BasicMAC): ctor (scheme, groups, args);
select_actions / forward / init_hidden / parameters / load_state / cuda / eval / train / save_models / load_models.
Only the entity scheme is on the hot path (every shipped alg YAML uses an `entity_attend` agent); the flat
`obs` agents (`rnn`, `ff`) are out of scope and rejected loudly."""
import os

import torch

from .. import ops
from ..components.action_selectors import REGISTRY as action_REGISTRY
from ..modules.agents import REGISTRY as agent_REGISTRY
from ..modules.nets import MaskSpec


class BasicMAC:
    def __init__(self, scheme, groups, args):
        self.n_agents = args.n_agents
        self.args = args
        self.scheme, self.groups = scheme, groups
        self.device = torch.device(getattr(args, "device", "cuda"))
        input_shape = self._get_input_shape(scheme)
        self._build_agents(input_shape)
        self.agent_output_type = args.agent_output_type
        if self.agent_output_type != "q":
            raise NotImplementedError("agent_output_type=%r: only the Q-learning path is implemented (no shipped "
                                      "config uses pi_logits)" % (self.agent_output_type,))
        self.action_selector = action_REGISTRY[args.action_selector](args)
        self.hidden_states = None

    # ---- acting ----------------------------------------------------------------------------------------------
    def select_actions(self, ep_batch, t_ep, t_env, bs=slice(None), test_mode=False, ret_agent_outs=False):
        avail_actions = ep_batch["avail_actions"][:, t_ep]
        agent_outputs = self.forward(ep_batch, t_ep, test_mode=test_mode)
        chosen = self.action_selector.select_action(agent_outputs[bs], avail_actions[bs], t_env, test_mode=test_mode)
        if ret_agent_outs:
            return chosen, agent_outputs[bs]
        return chosen

    def forward(self, ep_batch, t, test_mode=False, **kwargs):
        raise NotImplementedError

    def init_hidden(self, batch_size):
        # basic_controller.py:69-70: zeros (bs, n_agents, r)
        self.hidden_states = torch.zeros(batch_size, self.n_agents, self.agent.r, device=self.agent.store.flat.device)

    def parameters(self):
        return self.agent.parameters()

    def load_state(self, other_mac):
        self.agent.load_state_dict(other_mac.agent.state_dict())

    def cuda(self):
        self.agent.cuda()
        self.device = self.agent.store.flat.device

    def eval(self):
        self.agent.eval()

    def train(self):
        self.agent.train()

    def save_models(self, path):
        torch.save({k: v.detach().cpu() for k, v in self.agent.state_dict().items()}, os.path.join(path, "agent.th"))

    def load_models(self, path):
        self.agent.load_state_dict(torch.load(os.path.join(path, "agent.th"), map_location="cpu"))

    def _build_agents(self, input_shape):
        self.agent = agent_REGISTRY[self.args.agent](input_shape, self.args, self.device)

    def _get_input_shape(self, scheme):
        raise NotImplementedError("flat-observation agents are out of scope; use mac: entity_mac")
