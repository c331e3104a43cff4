"""Demonstration ingest: trajectories on disk -> `types.Trajectory[WithRew]` -> (via `flatten_trajectories`) the
device-resident expert table the discriminator samples from.

Mirror of imitation.data.serialize (serialize.py:27-93): `load` understands the legacy compressed `.npz` layout
(concatenated `obs`/`acts`/`infos`/`rews` split at `indices`, one extra observation per trajectory), the older
pickle of a trajectory sequence, and -- when the optional `datasets` package is importable -- a HuggingFace
datasets directory.  `save` writes the `.npz` layout (readable by the reference's `load`); the reference itself
writes a datasets directory, which needs the `datasets` package that this offline image does not ship.
"""
import os
import warnings
from typing import Mapping, Sequence

import numpy as np

from .types import Trajectory, TrajectoryWithRew


def save(path, trajectories: Sequence[Trajectory]) -> None:
    """Save trajectories in the `.npz` layout `load` (and the reference's `load`) reads."""
    trajectories = list(trajectories)
    if not trajectories:
        raise ValueError("no trajectories to save")
    path = os.fspath(path)
    os.makedirs(os.path.dirname(os.path.abspath(path)) or ".", exist_ok=True)
    lens = np.asarray([len(t) for t in trajectories])
    out = dict(obs=np.concatenate([t.obs for t in trajectories]), acts=np.concatenate([t.acts for t in trajectories]),
               infos=np.concatenate([t.infos if t.infos is not None else np.array([{}] * len(t)) for t in trajectories]),
               terminal=np.asarray([t.terminal for t in trajectories]), indices=np.cumsum(lens[:-1]))
    if all(isinstance(t, TrajectoryWithRew) for t in trajectories):
        out["rews"] = np.concatenate([t.rews for t in trajectories])
    with open(path, "wb") as f:  # (np.savez would append ".npz" to a bare path)
        np.savez_compressed(f, **out)


def load(path) -> Sequence[Trajectory]:
    """Loads a sequence of trajectories saved by `save()` (or by the reference) from `path`."""
    path = os.fspath(path)
    if os.path.isdir(path):  # huggingface datasets format (serialize.py:37-45)
        try:
            import datasets  # noqa: F401
        except ImportError as e:
            raise ImportError("loading a HuggingFace-datasets demonstration directory needs the `datasets` package; "
                              "convert it to the .npz layout with the reference's tooling first") from e
        return _load_hf(path)
    data = np.load(path, allow_pickle=True)  # works for both .npz and .pkl
    if isinstance(data, Sequence):  # pickle format
        warnings.warn("Loading old pickle version of Trajectories", DeprecationWarning)
        return data
    if isinstance(data, Mapping):  # .npz format
        num_trajs = len(data["indices"]) + 1 if len(data["terminal"]) == len(data["indices"]) + 1 else len(data["indices"])
        idx = np.asarray(data["indices"])
        obs = np.split(data["obs"], idx + np.arange(len(idx)) + 1)  # account for the extra obs in each trajectory
        acts = np.split(data["acts"], idx)
        infos = np.split(data["infos"], idx)
        terminal = data["terminal"]
        if "rews" in data:
            rews = np.split(data["rews"], idx)
            out = [TrajectoryWithRew(obs=o, acts=a, infos=i, terminal=bool(t), rews=r)
                   for o, a, i, t, r in zip(obs, acts, infos, terminal, rews)]
        else:
            out = [Trajectory(obs=o, acts=a, infos=i, terminal=bool(t)) for o, a, i, t in zip(obs, acts, infos, terminal)]
        assert len(out) in (num_trajs, len(idx) + 1)
        return out
    raise ValueError("Expected either an .npz file or a pickled sequence of trajectories; "
                     f"got a pickled object of type {type(data).__name__}")


def _load_hf(path) -> Sequence[Trajectory]:
    import datasets

    ds = datasets.load_from_disk(str(path))
    out = []
    for row in ds:
        kw = dict(obs=np.asarray(row["obs"]), acts=np.asarray(row["acts"]), infos=None, terminal=bool(row["terminal"]))
        if "rews" in row:
            out.append(TrajectoryWithRew(rews=np.asarray(row["rews"], dtype=np.float32), **kw))
        else:
            out.append(Trajectory(**kw))
    return out


def load_with_rewards(path) -> Sequence[TrajectoryWithRew]:
    """Loads a sequence of trajectories with rewards from a file (serialize.py:77-93)."""
    data = load(path)
    mismatched = [type(t) for t in data if not isinstance(t, TrajectoryWithRew)]
    if mismatched:
        raise ValueError(f"Expected all trajectories to be of type `TrajectoryWithRew`, but found {mismatched[0].__name__}")
    return data
