"""Device-resident replay buffer with the reference's `ReplayBuffer` API and FIFO semantics.

Mirrors imitation.data.buffer (Buffer :30-237, ReplayBuffer :240-416): `store` keeps only the
last `capacity` rows of an oversized chunk and wraps in at most two pieces, `sample` draws with
replacement among the stored rows.  Storage is one AoS table [capacity][2*d_obs+d_act+1] in HBM
(imb_table_store / imb_gather_rows, csrc/imb_ring.cu); `_idx` / `_n_data` live in the device
counter block so captured CUDA graphs keep them consistent.
"""
from typing import Optional

import numpy as np
import torch as th

from .. import _desc, _lib, spaces
from . import types


class ReplayBuffer:
    def __init__(self, capacity: int, venv=None, *, obs_shape=None, act_shape=None, obs_dtype=None, act_dtype=None,
                 n_actions: Optional[int] = None, device="cuda", state: Optional[th.Tensor] = None):
        if venv is not None:
            if obs_shape is not None or act_shape is not None or obs_dtype is not None or act_dtype is not None:
                raise ValueError("Cannot specify both shape/dtype and also environment")
            obs_shape = tuple(venv.observation_space.shape)
            act_shape = tuple(venv.action_space.shape)
            obs_dtype, act_dtype = venv.observation_space.dtype, venv.action_space.dtype
            if spaces.is_discrete(venv.action_space):
                n_actions = int(venv.action_space.n)
            state = getattr(venv, "state", state)
            device = getattr(venv, "device", device)
        elif any(x is None for x in (obs_shape, act_shape, obs_dtype, act_dtype)):
            raise ValueError("Shape or dtype missing and no environment specified.")
        self.capacity = int(capacity)
        self.obs_shape, self.act_shape = tuple(obs_shape), tuple(act_shape)
        self.obs_dtype, self.act_dtype = np.dtype(obs_dtype), np.dtype(act_dtype)
        self.discrete = n_actions is not None
        if np.issubdtype(self.act_dtype, np.integer) and n_actions is None:
            raise ValueError("integer actions need n_actions (Discrete space) for the one-hot table layout")
        self.d_obs = int(np.prod(self.obs_shape))
        self.d_act = int(n_actions) if self.discrete else int(np.prod(self.act_shape))
        self.tw = _desc.table_width(self.d_obs, self.d_act)
        self.device = th.device(device)
        self.table = th.zeros(self.capacity, self.tw, device=self.device)
        self.state = state if state is not None else th.zeros(_lib.ST_WORDS, dtype=th.int64, device=self.device)
        self._idx_host, self._n_host = 0, 0  # host mirrors of the device ring header

    @classmethod
    def from_data(cls, transitions: types.Transitions, capacity: Optional[int] = None, truncate_ok: bool = False,
                  n_actions: Optional[int] = None, device="cuda") -> "ReplayBuffer":
        obs = transitions.obs
        if capacity is None:
            capacity = obs.shape[0]
        inst = cls(capacity, obs_shape=obs.shape[1:], act_shape=transitions.acts.shape[1:], obs_dtype=obs.dtype,
                   act_dtype=transitions.acts.dtype, n_actions=n_actions, device=device)
        inst.store(transitions, truncate_ok=truncate_ok)
        return inst

    # -- reference API ---------------------------------------------------------------------------------------
    def size(self) -> int:
        return self._n_host

    @property
    def _idx(self) -> int:
        return self._idx_host

    def note_stored(self, n: int) -> None:
        """Host mirror of imb_ring_advance / imb_rollout_advance (Buffer._store_easy bookkeeping)."""
        kept = min(n, self.capacity)
        self._idx_host = (self._idx_host + kept) % self.capacity
        self._n_host = min(self._n_host + kept, self.capacity)

    def store(self, transitions, truncate_ok: bool = True) -> None:
        tr = types.as_transition_arrays(transitions)
        n = len(tr["obs"])
        if n == 0:
            raise ValueError("Trying to store empty data.")
        if n > self.capacity and not truncate_ok:
            raise ValueError("Not enough capacity to store data.")
        if tuple(tr["obs"].shape[1:]) != self.obs_shape or tuple(tr["next_obs"].shape[1:]) != self.obs_shape:
            raise ValueError("Wrong data shape for obs")
        if tuple(tr["acts"].shape[1:]) != self.act_shape:
            raise ValueError("Wrong data shape for acts")
        dev = self.device
        obs = th.as_tensor(np.ascontiguousarray(tr["obs"], dtype=np.float32)).to(dev).reshape(n, self.d_obs)
        nobs = th.as_tensor(np.ascontiguousarray(tr["next_obs"], dtype=np.float32)).to(dev).reshape(n, self.d_obs)
        dones = th.as_tensor(np.ascontiguousarray(tr["dones"]).astype(np.uint8)).to(dev)
        if self.discrete:
            acts_i = th.as_tensor(np.ascontiguousarray(tr["acts"]).astype(np.int64)).to(dev)
            _lib.table_store(self.table, self.capacity, self.d_obs, self.d_act, obs, None, acts_i, nobs, dones, n,
                             True, self.state)
        else:
            acts = th.as_tensor(np.ascontiguousarray(tr["acts"], dtype=np.float32)).to(dev).reshape(n, self.d_act)
            _lib.table_store(self.table, self.capacity, self.d_obs, self.d_act, obs, acts, None, nobs, dones, n, True,
                             self.state)
        _lib.ring_advance(self.state, self.capacity, n)
        self.note_stored(n)

    def sample_indices(self, n_samples: int, mode: str = "numpy", seed: int = 0) -> th.Tensor:
        """mode "numpy": np.random.randint on the global legacy RNG exactly like Buffer.sample
        (data/buffer.py:231) -- bit-exact parity; mode "device": Philox on the GPU."""
        if mode == "numpy":
            size = self.size()
            if size == 0:
                raise ValueError("Buffer is empty")
            return th.as_tensor(np.random.randint(size, size=n_samples)).to(self.device)
        idx = th.empty(n_samples, dtype=th.int64, device=self.device)
        _lib.sample_indices(0, idx, n_samples, 0, seed, self.state)
        return idx

    def rows(self, idx: th.Tensor) -> th.Tensor:
        return self.table[idx]

    def sample(self, n_samples: int) -> types.Transitions:
        if self.size() == 0:
            raise ValueError("Buffer is empty")
        rows = self.table[self.sample_indices(n_samples)].cpu().numpy()
        return rows_to_transitions(rows, self.d_obs, self.d_act, self.obs_shape, self.act_shape, self.obs_dtype,
                                   self.act_dtype, self.discrete)


def rows_to_transitions(rows: np.ndarray, d_obs, d_act, obs_shape, act_shape, obs_dtype, act_dtype, discrete,
                        rews: Optional[np.ndarray] = None):
    n = len(rows)
    obs = rows[:, :d_obs].reshape((n,) + tuple(obs_shape)).astype(obs_dtype)
    nobs = rows[:, d_obs + d_act:2 * d_obs + d_act].reshape((n,) + tuple(obs_shape)).astype(obs_dtype)
    a = rows[:, d_obs:d_obs + d_act]
    acts = a.argmax(1).astype(act_dtype) if discrete else a.reshape((n,) + tuple(act_shape)).astype(act_dtype)
    dones = rows[:, -1] > 0.5
    infos = np.array([{}] * n)
    if rews is not None:
        return types.TransitionsWithRew(obs=obs, acts=acts, infos=infos, next_obs=nobs, dones=dones, rews=rews)
    return types.Transitions(obs=obs, acts=acts, infos=infos, next_obs=nobs, dones=dones)
