"""`BufferingWrapper` over the GPU-resident VecEnv (mirror of imitation.data.wrappers:13-169).

The reference records every env step in per-env Python lists and rebuilds trajectories at pop
time.  Here the rollout kernel writes each transition straight to its FINAL position in the
flattened order `pop_trajectories()` + `flatten_trajectories()` would produce (finished
trajectories in completion order, then partial ones in env order -- closed form for lock-step
fixed-horizon envs, csrc/imb_rollout.cu:flat_index), and, when a ReplayBuffer is attached,
also into the generator ring.  The host API below materialises NumPy views on demand.
"""
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch as th

from .. import _desc, _lib
from . import buffer as buffer_mod
from . import types


class BufferingWrapper:
    def __init__(self, venv, error_on_premature_reset: bool = True):
        self.venv = venv
        self.num_envs = venv.num_envs
        self.observation_space, self.action_space = venv.observation_space, venv.action_space
        self.error_on_premature_reset = error_on_premature_reset
        self.n_transitions: Optional[int] = None
        self._init_reset = False
        self._ring: Optional[buffer_mod.ReplayBuffer] = None
        self._flat: Optional[th.Tensor] = None      # [E*T][tw] of the last rollout, reference order
        self._env_rews: Optional[th.Tensor] = None  # [E][T] ground-truth env rewards of the last rollout
        self._last: Optional[Tuple[int, int, int]] = None  # (T, t0, H) of the last rollout
        self._ep_lens: List[int] = []

    # -- plumbing used by DevicePPO / AdversarialTrainer ------------------------------------------------------
    @property
    def state(self):
        return self.venv.state

    @property
    def device(self):
        return self.venv.device

    def attach_ring(self, ring: buffer_mod.ReplayBuffer) -> None:
        """Generator samples go straight into `ring` (AdversarialTrainer.train_gen's
        flatten_trajectories_with_rew + ReplayBuffer.store, common.py:422-425, fused)."""
        self._ring = ring

    def rollout_targets(self, n_steps: int):
        E = self.num_envs
        tw = _desc.table_width(self.venv.d_obs, self.venv.d_act)
        if self._flat is None or self._flat.shape[0] != E * n_steps:
            self._flat = th.zeros(E * n_steps, tw, device=self.venv.device)
        if self.n_transitions:
            raise NotImplementedError("collecting a second rollout before pop_trajectories()/pop_transitions() is "
                                      "not supported by the device BufferingWrapper")
        return self._flat, self._ring

    def after_rollout(self, n_steps: int, t0: int, env_rews: th.Tensor) -> None:
        H = self.venv.horizon
        self._last = (n_steps, t0, H)
        self._env_rews = env_rews
        self.n_transitions = (self.n_transitions or 0) + self.num_envs * n_steps
        self._ep_lens += [H] * (self.num_envs * ((t0 + n_steps) // H))
        if self._ring is not None:
            self._ring.note_stored(self.num_envs * n_steps)

    # -- reference API -----------------------------------------------------------------------------------------------
    def reset(self, **kwargs):
        if self._init_reset and self.error_on_premature_reset and self.n_transitions:
            raise RuntimeError("BufferingWrapper reset() before samples were accessed")
        self._init_reset = True
        self.n_transitions = 0
        return self.venv.reset(**kwargs)

    def _segments(self):
        T, t0, H = self._last
        bounds = [0] + [b for b in range(H - t0, T, H)] + [T]
        return [(bounds[i], bounds[i + 1]) for i in range(len(bounds) - 1) if bounds[i + 1] > bounds[i]]

    def _flat_rews(self) -> np.ndarray:
        T, t0, H = self._last
        E = self.num_envs
        r = self._env_rews.cpu().numpy().reshape(E, T)
        return np.concatenate([r[:, a:b].reshape(-1) for a, b in self._segments()])

    def pop_transitions(self) -> types.TransitionsWithRew:
        if not self.n_transitions:
            raise RuntimeError("Called pop_transitions on an empty BufferingWrapper")
        v = self.venv
        out = buffer_mod.rows_to_transitions(self._flat.cpu().numpy(), v.d_obs, v.d_act, v.observation_space.shape,
                                             v.action_space.shape, v.observation_space.dtype, v.action_space.dtype,
                                             v.discrete, rews=self._flat_rews().astype(np.float32))
        assert len(out.obs) == self.n_transitions
        self.n_transitions = 0
        self._ep_lens = []
        return out

    def pop_trajectories(self) -> Tuple[Sequence[types.TrajectoryWithRew], Sequence[int]]:
        if not self.n_transitions:
            return [], []
        ep_lens = list(self._ep_lens)
        tr = self.pop_transitions()
        T, t0, H = self._last
        E = self.num_envs
        trajs, off = [], 0
        segs = self._segments()
        for si, (a, b) in enumerate(segs):
            L = b - a
            terminal = (t0 + b) % H == 0
            for e in range(E):
                sl = slice(off + e * L, off + (e + 1) * L)
                obs = np.concatenate([tr.obs[sl], tr.next_obs[sl][-1:]])
                trajs.append(types.TrajectoryWithRew(obs=obs, acts=tr.acts[sl], infos=None, terminal=bool(terminal),
                                                     rews=tr.rews[sl]))
            off += E * L
        return trajs, ep_lens

    def pop_finished_trajectories(self):
        trajs, lens = self.pop_trajectories()
        return [t for t in trajs if t.terminal], lens
