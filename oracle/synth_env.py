"""Synthetic MuJoCo-shaped batched environment (NumPy twin of csrc/imb_rollout.cu).

TEST INFRASTRUCTURE.  seals/MuJoCo are not available offline, so the benchmark env is
defined by this repo:  obs' = tanh(A obs + Bm u + c),
reward = w.obs' - 0.1|u|^2, u = clip(act,-1,1) (Box) or one-hot(act) (Discrete, no
penalty); fixed horizon H then auto-reset with the SB3 VecEnv contract the reference
relies on (data/rollout.py:161-186, rewards/reward_wrapper.py:98-109): on done, the
returned obs is the reset obs, infos[i]["terminal_observation"] is the true last obs and
infos[i]["TimeLimit.truncated"] is True (seals envs end by time limit).
Reset obs: 0.1 * N(0,1) from Philox(seed, STREAM_ENV_RESET, counter=(env_id, episode)).
"""
import dataclasses

import numpy as np

from . import philox


@dataclasses.dataclass
class SynthEnvSpec:
    d_obs: int
    d_act: int  # Box: action dim; Discrete: number of actions
    discrete: bool = False
    horizon: int = 1000
    seed: int = 0

    def __post_init__(self):
        rng = np.random.default_rng(self.seed)
        A = rng.standard_normal((self.d_obs, self.d_obs))
        A *= 0.9 / np.max(np.abs(np.linalg.eigvals(A)))
        self.A = A.astype(np.float32)
        self.Bm = (0.5 * rng.standard_normal((self.d_obs, self.d_act))).astype(np.float32)
        self.c = (0.1 * rng.standard_normal(self.d_obs)).astype(np.float32)
        self.w = (rng.standard_normal(self.d_obs) / np.sqrt(self.d_obs)).astype(np.float32)

    def reset_obs(self, env_ids, episodes):
        z = philox.normals(self.seed, philox.STREAM_ENV_RESET, np.asarray(env_ids, np.uint32),
                           np.asarray(episodes, np.uint32), self.d_obs)
        return (np.float32(0.1) * z).astype(np.float32)

    def control(self, acts):
        if self.discrete:
            u = np.zeros((len(acts), self.d_act), np.float32)
            u[np.arange(len(acts)), np.asarray(acts).astype(np.int64)] = 1.0
            return u
        return np.clip(np.asarray(acts, np.float32), -1.0, 1.0)

    def dynamics(self, obs, acts):
        u = self.control(acts)
        pre = obs @ self.A.T + u @ self.Bm.T + self.c
        nobs = np.tanh(pre).astype(np.float32)
        rew = nobs @ self.w
        if not self.discrete:
            rew = rew - np.float32(0.1) * np.sum(u * u, axis=1)
        return nobs, rew.astype(np.float32)


class _Box:
    def __init__(self, low, high, shape):
        self.shape = tuple(shape)
        self.dtype = np.dtype(np.float32)
        self.low = np.full(shape, low, np.float32)
        self.high = np.full(shape, high, np.float32)


class _Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)


class SynthVecEnv:
    """DummyVecEnv-style host env: per-env info dicts built every step, like SB3's."""

    def __init__(self, spec: SynthEnvSpec, num_envs: int, env_id_offset: int = 0, spaces_mod=None):
        self.spec = spec
        self.num_envs = num_envs
        self.env_ids = np.arange(num_envs, dtype=np.uint32) + np.uint32(env_id_offset)
        if spaces_mod is not None:  # real (shim) gymnasium spaces for the reference's isinstance checks
            self.observation_space = spaces_mod.Box(-np.inf, np.inf, (spec.d_obs,), np.float32)
            self.action_space = (spaces_mod.Discrete(spec.d_act) if spec.discrete
                                 else spaces_mod.Box(-1.0, 1.0, (spec.d_act,), np.float32))
        else:
            self.observation_space = _Box(-np.inf, np.inf, (spec.d_obs,))
            self.action_space = _Discrete(spec.d_act) if spec.discrete else _Box(-1.0, 1.0, (spec.d_act,))
        self.episode = np.zeros(num_envs, np.uint32)
        self.t = np.zeros(num_envs, np.int64)
        self.obs = None
        self._acts = None

    def reset(self):
        self.t[:] = 0
        self.obs = self.spec.reset_obs(self.env_ids, self.episode)
        return self.obs.copy()

    def step_async(self, actions):
        self._acts = actions

    def step_wait(self):
        nobs, rew = self.spec.dynamics(self.obs, self._acts)
        self.t += 1
        dones = self.t >= self.spec.horizon
        infos = [{} for _ in range(self.num_envs)]
        if dones.any():
            idx = np.nonzero(dones)[0]
            self.episode[idx] += 1
            fresh = self.spec.reset_obs(self.env_ids[idx], self.episode[idx])
            for j, i in enumerate(idx):
                infos[i]["terminal_observation"] = nobs[i].copy()
                infos[i]["TimeLimit.truncated"] = True
                nobs[i] = fresh[j]
            self.t[idx] = 0
        self.obs = nobs
        return nobs.copy(), rew, dones.copy(), infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        pass
