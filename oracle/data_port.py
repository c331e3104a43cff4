"""Host data plane of the GAIL/AIRL round restated in NumPy/Python.  TEST INFRASTRUCTURE.

Follows (reference file:line, /root/reference/src/imitation):
  data/buffer.py:30-237    Buffer (FIFO ring, store with <=2 slices, randint sample)
  data/buffer.py:240-416   ReplayBuffer (Transitions facade; store truncates to last `capacity`)
  data/rollout.py:57-187   TrajectoryAccumulator (per-env step lists, terminal_observation splice)
  data/rollout.py:563-621  flatten_trajectories[_with_rew]
  data/wrappers.py:13-169  BufferingWrapper (record, pop_trajectories order: finished, then partial)
  rewards/reward_wrapper.py:40-133  RewardVecEnvWrapper (terminal-obs splice, reward relabel)
  algorithms/base.py:77-110  fixed-horizon check
The per-env / per-step Python object handling is kept on purpose: it is the reference's
cost structure, and this port is what `bench.py --impl reference` times.
"""
import collections
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


class RingBufferPort:
    """FIFO ring of named arrays (buffer.py:30-237)."""

    def __init__(self, capacity: int, shapes: Dict[str, tuple], dtypes: Dict[str, np.dtype]):
        if shapes.keys() != dtypes.keys():
            raise KeyError("sample_shape and dtypes keys don't match")
        self.capacity = capacity
        self.shapes = {k: tuple(s) for k, s in shapes.items()}
        self._arrays = {k: np.zeros((capacity,) + s, dtype=dtypes[k]) for k, s in self.shapes.items()}
        self._n_data = 0
        self._idx = 0

    def size(self) -> int:
        return self._n_data

    def store(self, data: Dict[str, np.ndarray], truncate_ok: bool = False) -> None:
        missing = set(self.shapes) - set(data)
        extra = set(data) - set(self.shapes)
        if missing:
            raise ValueError(f"Missing keys {missing}")
        if extra:
            raise ValueError(f"Unexpected keys {extra}")
        lens = {a.shape[0] for a in data.values()}
        if len(lens) > 1:
            raise ValueError("Keys map to different length values.")
        n = lens.pop()
        if n == 0:
            raise ValueError("Trying to store empty data.")
        if n > self.capacity:
            if not truncate_ok:
                raise ValueError("Not enough capacity to store data.")
            data = {k: a[-self.capacity:] for k, a in data.items()}
            n = self.capacity
        for k, a in data.items():
            if a.shape[1:] != self.shapes[k]:
                raise ValueError(f"Wrong data shape for {k}")
        first = min(n, self.capacity - self._idx)
        for lo, hi in ((0, first), (first, n)):
            if hi > lo:
                for k, a in data.items():
                    self._arrays[k][self._idx:self._idx + (hi - lo)] = a[lo:hi]
                self._idx = (self._idx + (hi - lo)) % self.capacity
                self._n_data = min(self._n_data + (hi - lo), self.capacity)

    def sample_indices(self, n: int) -> np.ndarray:
        if self.size() == 0:
            raise ValueError("Buffer is empty")
        return np.random.randint(self.size(), size=n)  # global legacy RNG, with replacement

    def sample(self, n: int) -> Dict[str, np.ndarray]:
        ind = self.sample_indices(n)
        return {k: arr[ind] for k, arr in self._arrays.items()}


class ReplayBufferPort:
    def __init__(self, capacity, obs_shape, act_shape, obs_dtype=np.float32, act_dtype=np.float32):
        self.capacity = capacity
        self._buffer = RingBufferPort(
            capacity,
            {"obs": obs_shape, "acts": act_shape, "next_obs": obs_shape, "dones": (), "infos": ()},
            {"obs": obs_dtype, "acts": act_dtype, "next_obs": obs_dtype, "dones": np.dtype(bool),
             "infos": np.dtype(object)},
        )

    def store(self, trans: Dict[str, np.ndarray], truncate_ok: bool = True) -> None:
        self._buffer.store({k: trans[k] for k in self._buffer.shapes}, truncate_ok=truncate_ok)

    def sample(self, n: int) -> Dict[str, np.ndarray]:
        return self._buffer.sample(n)

    def size(self) -> int:
        return self._buffer.size()


class TrajAccumulatorPort:
    """rollout.py:57-187: per-env list of step dicts -> stacked trajectory dict."""

    def __init__(self):
        self.partial = collections.defaultdict(list)

    def add_step(self, step: dict, key) -> None:
        self.partial[key].append(step)

    def finish(self, key, terminal: bool) -> dict:
        parts = self.partial.pop(key)
        cols = collections.defaultdict(list)
        for p in parts:
            for k, v in p.items():
                cols[k].append(v)
        traj = {k: np.stack(v, axis=0) for k, v in cols.items()}
        traj["terminal"] = terminal
        assert traj["rews"].shape[0] == traj["acts"].shape[0] == len(traj["obs"]) - 1
        return traj

    def add_steps_and_auto_finish(self, acts, obs, rews, dones, infos) -> List[dict]:
        done_trajs = []
        for i, (a, o, r, d, info) in enumerate(zip(acts, obs, rews, dones, infos)):
            real_ob = info["terminal_observation"] if d else o
            self.add_step(dict(acts=a, rews=r, obs=real_ob, infos=info), i)
            if d:
                done_trajs.append(self.finish(i, terminal=True))
                self.add_step(dict(obs=o), i)
        return done_trajs


def flatten_port(trajs: Sequence[dict]) -> Dict[str, np.ndarray]:
    """rollout.py:563-621: obs=obs[:-1], next_obs=obs[1:], dones only at a terminal traj's last row."""
    cols = {k: [] for k in ("obs", "next_obs", "acts", "dones", "infos", "rews")}
    for t in trajs:
        cols["acts"].append(t["acts"])
        cols["obs"].append(t["obs"][:-1])
        cols["next_obs"].append(t["obs"][1:])
        d = np.zeros(len(t["acts"]), dtype=bool)
        d[-1] = t["terminal"]
        cols["dones"].append(d)
        cols["infos"].append(t["infos"] if "infos" in t else np.array([{}] * len(t["acts"])))
        cols["rews"].append(t["rews"])
    return {k: np.concatenate(v) for k, v in cols.items()}


class BufferingPort:
    """wrappers.py:13-169.  Wraps a host VecEnv; records every step."""

    def __init__(self, venv, error_on_premature_reset: bool = True):
        self.venv = venv
        self.num_envs = venv.num_envs
        self.observation_space = venv.observation_space
        self.action_space = venv.action_space
        self.error_on_premature_reset = error_on_premature_reset
        self._trajs: List[dict] = []
        self._ep_lens: List[int] = []
        self._init_reset = False
        self._acc: Optional[TrajAccumulatorPort] = None
        self._saved_acts = None
        self._timesteps = None
        self.n_transitions = None

    def reset(self):
        if self._init_reset and self.error_on_premature_reset and self.n_transitions > 0:
            raise RuntimeError("BufferingWrapper reset() before samples were accessed")
        self._init_reset = True
        self.n_transitions = 0
        obs = self.venv.reset()
        self._acc = TrajAccumulatorPort()
        for i, ob in enumerate(obs):
            self._acc.add_step({"obs": ob}, i)
        self._timesteps = np.zeros((len(obs),), dtype=int)
        return obs

    def step_async(self, actions):
        assert self._init_reset and self._saved_acts is None
        self.venv.step_async(actions)
        self._saved_acts = actions

    def step_wait(self):
        acts, self._saved_acts = self._saved_acts, None
        obs, rews, dones, infos = self.venv.step_wait()
        self.n_transitions += self.num_envs
        self._timesteps += 1
        ep = self._timesteps[dones]
        if len(ep) > 0:
            self._ep_lens += list(ep)
        self._timesteps[dones] = 0
        self._trajs.extend(self._acc.add_steps_and_auto_finish(acts, obs, rews, dones, infos))
        return obs, rews, dones, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def pop_trajectories(self) -> Tuple[List[dict], List[int]]:
        if self.n_transitions == 0:
            return [], []
        for i in range(self.num_envs):
            if len(self._acc.partial[i]) - 1 >= 1:
                t = self._acc.finish(i, terminal=False)
                self._trajs.append(t)
                self._acc.add_step({"obs": t["obs"][-1]}, i)
        trajs, lens = self._trajs, self._ep_lens
        self._trajs, self._ep_lens = [], []
        self.n_transitions = 0
        return trajs, lens

    def pop_transitions(self) -> Dict[str, np.ndarray]:
        if self.n_transitions == 0:
            raise RuntimeError("Called pop_transitions on an empty BufferingWrapper")
        n = self.n_transitions
        trajs, _ = self.pop_trajectories()
        out = flatten_port(trajs)
        assert len(out["obs"]) == n
        return out


class RewardRelabelPort:
    """reward_wrapper.py:40-133: replace env reward by reward_fn(old_obs, acts, obs_fixed, dones)."""

    def __init__(self, venv, reward_fn, ep_history: int = 100):
        self.venv = venv
        self.num_envs = venv.num_envs
        self.observation_space = venv.observation_space
        self.action_space = venv.action_space
        self.episode_rewards = collections.deque(maxlen=ep_history)
        self._cum = np.zeros((venv.num_envs,))
        self.reward_fn = reward_fn
        self._actions = None
        self._old_obs = None
        self.reset()

    def reset(self):
        self._old_obs = self.venv.reset()
        return self._old_obs

    def step_async(self, actions):
        self._actions = actions
        self.venv.step_async(actions)

    def step_wait(self):
        obs, old_rews, dones, infos = self.venv.step_wait()
        fixed = []
        for ob, d, info in zip(obs, dones, infos):
            fixed.append(info["terminal_observation"] if d else ob)
        fixed = np.stack(fixed)
        rews = self.reward_fn(self._old_obs, self._actions, fixed, np.array(dones))
        assert len(rews) == len(obs)
        self._cum += rews
        for d, r in zip(dones, self._cum):
            if d:
                self.episode_rewards.append(r)
        self._cum[np.asarray(dones, dtype=bool)] = 0
        self._old_obs = obs
        for info, r in zip(infos, old_rews):
            info["original_env_rew"] = r
        return obs, rews, dones, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()


class FixedHorizonCheckPort:
    """algorithms/base.py:77-110."""

    def __init__(self, allow_variable_horizon=False):
        self.allow = allow_variable_horizon
        self._horizon = None

    def check(self, horizons):
        if self.allow:
            return
        hs = set(int(h) for h in horizons)
        if self._horizon is not None:
            hs.add(self._horizon)
        if len(hs) > 1:
            raise ValueError(f"Episodes of different length detected: {hs}.")
        if len(hs) == 1:
            self._horizon = hs.pop()
