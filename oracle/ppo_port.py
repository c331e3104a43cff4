"""PPO generator restated from stable-baselines3 ~=2.2.1 semantics on torch-CPU.

TEST INFRASTRUCTURE.  **PARITY UNPINNED**: SB3 is a third-party dependency of the
reference (setup.py:206) that is neither vendored under /root/reference nor installed
here, and no reference test pins PPO numerics.  This
file restates SB3's published algorithm (ActorCriticPolicy with separate 32x32 tanh
pi/vf nets = imitation's FeedForward32Policy, policies/base.py:92-104; optional
NormalizeFeaturesExtractor = Flatten->RunningNorm, policies/base.py:123-149;
OnPolicyAlgorithm.collect_rollouts; RolloutBuffer GAE; PPO.train) and is this repo's
declared oracle for the generator.  Reference call sites it serves:
algorithms/adversarial/common.py:414 (gen_algo.learn), :490-496 (evaluate_actions).
"""
import math
from typing import Optional

import numpy as np
import torch as th
from torch import nn

from .nets_port import RunningNormPort


def _ortho(layer: nn.Linear, gain: float):
    nn.init.orthogonal_(layer.weight, gain=gain)
    nn.init.zeros_(layer.bias)


class ActorCriticPort(nn.Module):
    def __init__(self, d_obs: int, d_act: int, discrete: bool = False, hidden=(32, 32), normalize_features=False):
        super().__init__()
        self.discrete = discrete
        self.d_obs, self.d_act = d_obs, d_act
        self.feat_norm = RunningNormPort(d_obs) if normalize_features else None

        def tower():
            mods, prev = [], d_obs
            for h in hidden:
                mods += [nn.Linear(prev, h), nn.Tanh()]
                prev = h
            return nn.Sequential(*mods)

        self.pi = tower()
        self.vf = tower()
        self.action_net = nn.Linear(hidden[-1], d_act)
        self.value_net = nn.Linear(hidden[-1], 1)
        self.log_std = None if discrete else nn.Parameter(th.zeros(d_act))
        for m in list(self.pi) + list(self.vf):
            if isinstance(m, nn.Linear):
                _ortho(m, math.sqrt(2))
        _ortho(self.action_net, 0.01)
        _ortho(self.value_net, 1.0)

    def features(self, obs):
        x = th.flatten(obs.float(), 1)
        return self.feat_norm(x) if self.feat_norm is not None else x

    def _dist(self, f):
        out = self.action_net(self.pi(f))
        if self.discrete:
            return th.distributions.Categorical(logits=out)
        return th.distributions.Normal(out, th.ones_like(out) * self.log_std.exp())

    def _logp(self, dist, actions):
        return dist.log_prob(actions) if self.discrete else dist.log_prob(actions).sum(dim=1)

    def forward(self, obs, noise: Optional[th.Tensor] = None, deterministic=False):
        """Returns actions, values, log_prob.  `noise` (standard normal / uniform) pins sampling."""
        f = self.features(obs)
        dist = self._dist(f)
        if self.discrete:
            if deterministic:
                actions = dist.probs.argmax(dim=1)
            elif noise is not None:  # inverse-CDF sampling from supplied uniforms
                cdf = dist.probs.cumsum(dim=1)
                actions = (noise.unsqueeze(1) >= cdf).sum(dim=1).clamp(max=self.d_act - 1)
            else:
                actions = dist.sample()
        else:
            if deterministic:
                actions = dist.mean
            elif noise is not None:
                actions = dist.mean + dist.stddev * noise
            else:
                actions = dist.sample()
        values = self.value_net(self.vf(f))
        return actions, values, self._logp(dist, actions)

    def evaluate_actions(self, obs, actions):
        f = self.features(obs)
        dist = self._dist(f)
        if self.discrete:
            actions = actions.long().flatten()
            ent = dist.entropy()
        else:
            ent = dist.entropy().sum(dim=1)
        return self.value_net(self.vf(f)), self._logp(dist, actions), ent

    def predict_values(self, obs):
        return self.value_net(self.vf(self.features(obs)))


class PPOPort:
    def __init__(self, policy: ActorCriticPort, venv, n_steps: int, batch_size=64, n_epochs=10, learning_rate=3e-4,
                 gamma=0.99, gae_lambda=0.95, clip_range=0.2, ent_coef=0.0, vf_coef=0.5, max_grad_norm=0.5,
                 normalize_advantage=True, noise_fn=None, perm_fn=None):
        self.policy, self.env = policy, venv
        self.n_steps, self.batch_size, self.n_epochs = n_steps, batch_size, n_epochs
        self.gamma, self.gae_lambda, self.clip_range = gamma, gae_lambda, clip_range
        self.ent_coef, self.vf_coef, self.max_grad_norm = ent_coef, vf_coef, max_grad_norm
        self.normalize_advantage = normalize_advantage
        self.opt = th.optim.Adam(policy.parameters(), lr=learning_rate, eps=1e-5)
        self.device = th.device("cpu")
        self.noise_fn = noise_fn      # (global_step) -> np.ndarray [E, d_act] or [E]; None -> torch RNG
        self.perm_fn = perm_fn        # (epoch_counter, n) -> permutation; None -> np.random.permutation
        self.num_timesteps = 0
        self._last_obs = None
        self._last_starts = None
        self._epoch_counter = 0
        self._step_counter = 0
        self.loss_log = []            # per-minibatch (pg, vf, ent, total) for parity tests

    def set_env(self, env):
        self.env = env

    def get_env(self):
        return self.env

    # -- OnPolicyAlgorithm.collect_rollouts + RolloutBuffer ------------------------------
    def collect_rollouts(self):
        E, T, p = self.env.num_envs, self.n_steps, self.policy
        p.eval()
        buf = dict(
            obs=np.zeros((T, E, p.d_obs), np.float32),
            actions=np.zeros((T, E) if p.discrete else (T, E, p.d_act), np.float32),
            rewards=np.zeros((T, E), np.float32), starts=np.zeros((T, E), np.float32),
            values=np.zeros((T, E), np.float32), log_probs=np.zeros((T, E), np.float32))
        for t in range(T):
            with th.no_grad():
                noise = None
                if self.noise_fn is not None:
                    noise = th.as_tensor(self.noise_fn(self._step_counter))
                actions, values, logp = p(th.as_tensor(self._last_obs), noise)
            actions = actions.numpy()
            env_actions = actions
            if not p.discrete:
                env_actions = np.clip(actions, self.env.action_space.low, self.env.action_space.high)
            new_obs, rewards, dones, infos = self.env.step(env_actions)
            rewards = np.array(rewards, dtype=np.float32)
            self.num_timesteps += E
            self._step_counter += 1
            for i, d in enumerate(dones):
                if d and infos[i].get("terminal_observation") is not None and infos[i].get("TimeLimit.truncated", False):
                    with th.no_grad():
                        tv = p.predict_values(th.as_tensor(infos[i]["terminal_observation"][None]))[0]
                    rewards[i] += self.gamma * float(tv)
            buf["obs"][t] = self._last_obs
            buf["actions"][t] = actions
            buf["rewards"][t] = rewards
            buf["starts"][t] = self._last_starts
            buf["values"][t] = values.numpy().flatten()
            buf["log_probs"][t] = logp.numpy()
            self._last_obs = new_obs
            self._last_starts = dones
        with th.no_grad():
            last_values = p.predict_values(th.as_tensor(new_obs)).numpy().flatten()
        adv = np.zeros((T, E), np.float32)
        last = 0
        for t in reversed(range(T)):
            if t == T - 1:
                nonterm, nextv = 1.0 - dones.astype(np.float32), last_values
            else:
                nonterm, nextv = 1.0 - buf["starts"][t + 1], buf["values"][t + 1]
            delta = buf["rewards"][t] + self.gamma * nextv * nonterm - buf["values"][t]
            last = delta + self.gamma * self.gae_lambda * nonterm * last
            adv[t] = last
        buf["advantages"] = adv
        buf["returns"] = adv + buf["values"]
        return buf

    # -- PPO.train ----------------------------------------------------------------------
    def train(self, buf):
        p = self.policy
        p.train()
        T, E = buf["rewards"].shape

        def flat(x):  # swap_and_flatten: env-major, index = env * T + step
            return th.as_tensor(np.ascontiguousarray(x.swapaxes(0, 1)).reshape(T * E, *x.shape[2:]))

        obs, actions = flat(buf["obs"]), flat(buf["actions"])
        values_old, logp_old = flat(buf["values"]), flat(buf["log_probs"])
        adv_all, ret_all = flat(buf["advantages"]), flat(buf["returns"])
        n = T * E
        for _ in range(self.n_epochs):
            perm = self.perm_fn(self._epoch_counter, n) if self.perm_fn is not None else np.random.permutation(n)
            self._epoch_counter += 1
            for start in range(0, n, self.batch_size):
                idx = th.as_tensor(perm[start:start + self.batch_size]).long()
                values, logp, ent = p.evaluate_actions(obs[idx], actions[idx])
                values = values.flatten()
                adv = adv_all[idx]
                if self.normalize_advantage and len(adv) > 1:
                    adv = (adv - adv.mean()) / (adv.std() + 1e-8)
                ratio = th.exp(logp - logp_old[idx])
                pl1 = adv * ratio
                pl2 = adv * th.clamp(ratio, 1 - self.clip_range, 1 + self.clip_range)
                pg_loss = -th.min(pl1, pl2).mean()
                v_loss = nn.functional.mse_loss(ret_all[idx], values)
                ent_loss = -th.mean(ent)
                loss = pg_loss + self.ent_coef * ent_loss + self.vf_coef * v_loss
                self.opt.zero_grad()
                loss.backward()
                nn.utils.clip_grad_norm_(p.parameters(), self.max_grad_norm)
                self.opt.step()
                self.loss_log.append(tuple(float(x.detach()) for x in (pg_loss, v_loss, ent_loss, loss)))

    def learn(self, total_timesteps, reset_num_timesteps=False, callback=None):
        if self._last_obs is None:
            self._last_obs = self.env.reset()
            self._last_starts = np.ones((self.env.num_envs,), dtype=bool)
        done_steps = 0
        while done_steps < total_timesteps:
            buf = self.collect_rollouts()
            done_steps += self.env.num_envs * self.n_steps
            self.train(buf)
            self.last_rollout = buf
        return self
