"""Reward-network arithmetic restated on torch-CPU.  TEST INFRASTRUCTURE.

Follows (reference file:line, /root/reference/src/imitation):
  util/networks.py:47-95   BaseNorm  (train: update then normalise with UPDATED stats)
  util/networks.py:98-134  RunningNorm.update_stats (Chan merge, biased batch var, int32 count)
  util/networks.py:204-283 build_mlp (norm? -> [Linear,ReLU]* -> Linear -> squeeze)
  rewards/reward_nets.py:383-457  BasicRewardNet (concat of selected inputs)
  rewards/reward_nets.py:674-736  ShapedRewardNet.forward (base, Phi(s'), Phi(s) order)
  rewards/reward_nets.py:739-839  BasicShapedRewardNet / BasicPotentialMLP
  rewards/reward_nets.py:613-671  NormalizedRewardNet.predict_processed
  algorithms/adversarial/gail.py:75-83  RewardNetFromDiscriminatorLogit
Written in this repo's own words; numerics use the same torch ops so CPU results are
bit-comparable with the reference on the same machine.
"""
from typing import Optional, Sequence

import numpy as np
import torch as th
from torch import nn


class RunningNormPort(nn.Module):
    def __init__(self, n: int, eps: float = 1e-5):
        super().__init__()
        self.eps = eps
        self.register_buffer("running_mean", th.zeros(n))
        self.register_buffer("running_var", th.ones(n))
        self.register_buffer("count", th.zeros((), dtype=th.int))

    @th.no_grad()
    def update_stats(self, batch: th.Tensor) -> None:
        b_mean = th.mean(batch, dim=0)
        b_var = th.var(batch, dim=0, unbiased=False)
        b_n = batch.shape[0]
        delta = b_mean - self.running_mean
        tot = self.count + b_n
        self.running_mean += delta * b_n / tot
        self.running_var *= self.count
        self.running_var += b_var * b_n
        self.running_var += th.square(delta) * self.count * b_n / tot
        self.running_var /= tot
        self.count += b_n

    def forward(self, x):
        if self.training:
            self.update_stats(x)
        return (x - self.running_mean) / th.sqrt(self.running_var + self.eps)


def mlp_port(in_size: int, hid_sizes: Sequence[int], normalize: bool, flatten_input: bool = False) -> nn.Sequential:
    layers = []
    if flatten_input:
        layers.append(("flatten", nn.Flatten()))
    if normalize:
        layers.append(("normalize_input", RunningNormPort(in_size)))
    prev = in_size
    for i, h in enumerate(hid_sizes):
        layers.append((f"dense{i}", nn.Linear(prev, h)))
        layers.append((f"act{i}", nn.ReLU()))
        prev = h
    layers.append(("dense_final", nn.Linear(prev, 1)))
    import collections
    return nn.Sequential(collections.OrderedDict(layers))


class BasicRewardNetPort(nn.Module):
    def __init__(self, d_obs, d_act, hid_sizes=(32, 32), use_state=True, use_action=True,
                 use_next_state=False, use_done=False, normalize_input=False):
        super().__init__()
        self.use = (use_state, use_action, use_next_state, use_done)
        din = d_obs * use_state + d_act * use_action + d_obs * use_next_state + 1 * use_done
        self.mlp = mlp_port(din, hid_sizes, normalize_input)

    def forward(self, state, action, next_state, done):
        parts = []
        if self.use[0]:
            parts.append(th.flatten(state, 1))
        if self.use[1]:
            parts.append(th.flatten(action, 1))
        if self.use[2]:
            parts.append(th.flatten(next_state, 1))
        if self.use[3]:
            parts.append(done.reshape(-1, 1))
        return self.mlp(th.cat(parts, dim=1)).squeeze(1)


class ShapedRewardNetPort(nn.Module):
    """r(s,a,s',d) + gamma (1-d) Phi(s') - Phi(s); Phi(s') is evaluated BEFORE Phi(s), so a
    RunningNorm inside Phi is updated twice per training forward."""

    def __init__(self, d_obs, d_act, reward_hid_sizes=(32,), potential_hid_sizes=(32, 32),
                 use_state=True, use_action=True, use_next_state=False, use_done=False,
                 discount_factor=0.99, normalize_input=False):
        super().__init__()
        self.base = BasicRewardNetPort(d_obs, d_act, reward_hid_sizes, use_state, use_action,
                                       use_next_state, use_done, normalize_input)
        self.potential = mlp_port(d_obs, potential_hid_sizes, normalize_input, flatten_input=True)
        self.discount_factor = discount_factor

    def forward(self, state, action, next_state, done):
        base = self.base(state, action, next_state, done)
        new_phi = self.potential(next_state).squeeze(1).flatten()
        old_phi = self.potential(state).squeeze(1).flatten()
        return base + self.discount_factor * (1 - done.float()) * new_phi - old_phi


def preprocess_port(obs, acts, next_obs, dones, n_actions: Optional[int] = None):
    """rewards/reward_nets.py:52-118: float cast for Box, one-hot for Discrete, done->f32."""
    s = th.as_tensor(np.array(obs)).float()
    ns = th.as_tensor(np.array(next_obs)).float()
    a = th.as_tensor(np.array(acts))
    a = nn.functional.one_hot(a.long(), n_actions).float() if n_actions is not None else a.float()
    d = th.as_tensor(np.array(dones)).to(th.float32)
    return s, a, ns, d


@th.no_grad()
def predict_port(net: nn.Module, obs, acts, next_obs, dones, n_actions=None, gail_transform=False) -> np.ndarray:
    """RewardNet.predict (reward_nets.py:120-176): eval mode, no grad, numpy out.
    gail_transform applies -logsigmoid(-logit) (gail.py:83)."""
    was = net.training
    net.eval()
    try:
        out = net(*preprocess_port(obs, acts, next_obs, dones, n_actions))
        if gail_transform:
            out = -nn.functional.logsigmoid(-out)
    finally:
        net.train(was)
    return out.detach().cpu().numpy().flatten()


class OutputNormPort:
    """NormalizedRewardNet.predict_processed (reward_nets.py:637-671): normalise with the
    CURRENT stats in eval mode, then update the stats with the raw rewards."""

    def __init__(self):
        self.norm = RunningNormPort(1)
        self.norm.eval()

    def __call__(self, raw: np.ndarray, update_stats: bool = True) -> np.ndarray:
        rew_th = th.tensor(raw)
        out = self.norm(rew_th).detach().numpy().flatten()
        if update_stats:
            self.norm.update_stats(rew_th)
        return out
