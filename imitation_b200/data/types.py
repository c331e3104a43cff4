"""Host-side contract types of the data plane (mirror of imitation.data.types:335-638 for
Box/Discrete spaces): frozen dataclasses with read-only arrays and the reference's validation
errors.  These cross the API boundary only; the hot path keeps transitions in HBM tables."""
import dataclasses
from typing import Any, Dict, Mapping, Optional, Sequence

import numpy as np


def dataclass_quick_asdict(obj) -> Dict[str, Any]:
    return {f.name: getattr(obj, f.name) for f in dataclasses.fields(obj)}


@dataclasses.dataclass(frozen=True)
class Trajectory:
    obs: np.ndarray
    acts: np.ndarray
    infos: Optional[np.ndarray]
    terminal: bool

    def __len__(self) -> int:
        return len(self.acts)

    def __post_init__(self):
        if len(self.obs) != len(self.acts) + 1:
            raise ValueError(f"expected one more observations than actions: {len(self.obs)} != {len(self.acts)} + 1")
        if self.infos is not None and len(self.infos) != len(self.acts):
            raise ValueError(f"infos when present must be present for each action: {len(self.infos)} != {len(self.acts)}")
        if len(self.acts) == 0:
            raise ValueError("Degenerate trajectory: must have at least one action.")


def _rews_validation(rews: np.ndarray, acts: np.ndarray):
    if rews.shape != (len(acts),):
        raise ValueError(f"rewards must be 1D array, one entry for each action: {rews.shape} != ({len(acts)},)")
    if not np.issubdtype(rews.dtype, np.floating):
        raise ValueError(f"rewards dtype {rews.dtype} not a float")


@dataclasses.dataclass(frozen=True)
class TrajectoryWithRew(Trajectory):
    rews: np.ndarray

    def __post_init__(self):
        super().__post_init__()
        _rews_validation(self.rews, self.acts)


@dataclasses.dataclass(frozen=True)
class TransitionsMinimal:
    obs: np.ndarray
    acts: np.ndarray
    infos: np.ndarray

    def __len__(self) -> int:
        return len(self.obs)

    def __post_init__(self):
        for val in vars(self).values():
            if isinstance(val, np.ndarray):
                val.setflags(write=False)
        if len(self.obs) != len(self.acts):
            raise ValueError(f"obs and acts must have same number of timesteps: {len(self.obs)} != {len(self.acts)}")
        if len(self.infos) != len(self.obs):
            raise ValueError(f"obs and infos must have same number of timesteps: {len(self.obs)} != {len(self.infos)}")

    def __getitem__(self, key):
        d = {k: v[key] for k, v in dataclass_quick_asdict(self).items()}
        if isinstance(key, slice):
            return dataclasses.replace(self, **d)
        return d


@dataclasses.dataclass(frozen=True)
class Transitions(TransitionsMinimal):
    next_obs: np.ndarray
    dones: np.ndarray

    def __post_init__(self):
        super().__post_init__()
        if self.obs.shape != self.next_obs.shape:
            raise ValueError(f"obs and next_obs must have same shape: {self.obs.shape} != {self.next_obs.shape}")
        if self.obs.dtype != self.next_obs.dtype:
            raise ValueError(f"obs and next_obs must have the same dtype: {self.obs.dtype} != {self.next_obs.dtype}")
        if self.dones.shape != (len(self.acts),):
            raise ValueError(f"dones must be 1D array, one entry for each timestep: {self.dones.shape} != ({len(self.acts)},)")
        if self.dones.dtype != bool:
            raise ValueError(f"dones must be boolean, not {self.dones.dtype}")


@dataclasses.dataclass(frozen=True)
class TransitionsWithRew(Transitions):
    rews: np.ndarray

    def __post_init__(self):
        super().__post_init__()
        _rews_validation(self.rews, self.acts)


def flatten_trajectories(trajectories: Sequence[Trajectory]) -> Transitions:
    """data/rollout.py:563-610."""
    parts = {k: [] for k in ("obs", "next_obs", "acts", "dones", "infos")}
    for t in trajectories:
        parts["acts"].append(t.acts)
        parts["obs"].append(t.obs[:-1])
        parts["next_obs"].append(t.obs[1:])
        d = np.zeros(len(t.acts), dtype=bool)
        d[-1] = t.terminal
        parts["dones"].append(d)
        parts["infos"].append(np.array([{}] * len(t)) if t.infos is None else t.infos)
    return Transitions(**{k: np.concatenate(v) for k, v in parts.items()})


def flatten_trajectories_with_rew(trajectories: Sequence[TrajectoryWithRew]) -> TransitionsWithRew:
    tr = flatten_trajectories(trajectories)
    return TransitionsWithRew(**dataclass_quick_asdict(tr), rews=np.concatenate([t.rews for t in trajectories]))


def as_transition_arrays(demos) -> Mapping[str, np.ndarray]:
    """AnyTransitions (algorithms/base.py:125-129) -> dict of obs/acts/next_obs/dones arrays."""
    if isinstance(demos, TransitionsMinimal):
        return {k: v for k, v in dataclass_quick_asdict(demos).items() if k != "infos"}
    if isinstance(demos, Mapping):
        return {k: np.asarray(v) for k, v in demos.items() if k != "infos"}
    demos = list(demos)
    if demos and isinstance(demos[0], Trajectory):
        return as_transition_arrays(flatten_trajectories(demos))
    if demos and isinstance(demos[0], Mapping):  # iterable of transition batches
        keys = [k for k in demos[0] if k != "infos"]
        return {k: np.concatenate([np.asarray(b[k]) for b in demos]) for k in keys}
    raise TypeError(f"`demonstrations` unexpected type {type(demos)}")
