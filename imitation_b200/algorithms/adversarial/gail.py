"""GAIL (mirror of imitation.algorithms.adversarial.gail:14-168)."""
from typing import Optional

import torch as th
from torch.nn import functional as F

from ...rewards import reward_nets
from . import common


class RewardNetFromDiscriminatorLogit(reward_nets.RewardNet):
    """reward = -logsigmoid(-logit) = softplus(logit)  (gail.py:14-83)."""

    def __init__(self, base: reward_nets.RewardNet):
        super().__init__(observation_space=base.observation_space, action_space=base.action_space,
                         normalize_images=base.normalize_images)
        self.base = base

    def forward(self, state, action, next_state, done) -> th.Tensor:
        logits = self.base.forward(state, action, next_state, done)
        return -F.logsigmoid(-logits)


class GAIL(common.AdversarialTrainer):
    def __init__(self, *, demonstrations, demo_batch_size: int, venv, gen_algo, reward_net, **kwargs):
        reward_net = reward_net.to(gen_algo.device)
        self._processed_reward = RewardNetFromDiscriminatorLogit(reward_net)
        super().__init__(demonstrations=demonstrations, demo_batch_size=demo_batch_size, venv=venv,
                         gen_algo=gen_algo, reward_net=reward_net, **kwargs)

    def logits_expert_is_high(self, state, action, next_state, done, log_policy_act_prob: Optional[th.Tensor] = None
                              ) -> th.Tensor:
        del log_policy_act_prob
        logits = self._reward_net(state, action, next_state, done)
        assert logits.shape == state.shape[:1]
        return logits

    @property
    def reward_train(self) -> reward_nets.RewardNet:
        return self._processed_reward

    @property
    def reward_test(self) -> reward_nets.RewardNet:
        return self._processed_reward
