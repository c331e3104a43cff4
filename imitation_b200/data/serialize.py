"""Demonstration ingest: trajectories on disk -> `types.Trajectory[WithRew]` -> (via `flatten_trajectories`) the
device-resident expert table the discriminator samples from.

Mirror of imitation.data.serialize (serialize.py:15-93): `save` writes a HuggingFace `datasets` directory like the
reference's (one row per trajectory, `huggingface_utils.trajectories_to_dataset`) -- or, for a path ending in `.npz`,
the legacy compressed layout the reference still reads; `load` understands the datasets directory, the legacy `.npz`
layout (concatenated `obs`/`acts`/`infos`/`rews` split at `indices`, one extra observation per trajectory) and the older
pickle of a trajectory sequence.
"""
import logging
import os
import warnings
from typing import Mapping, Sequence

import numpy as np

from . import huggingface_utils
from .types import Trajectory, TrajectoryWithRew


def save(path, trajectories: Sequence[Trajectory]) -> None:
    """Save a sequence of trajectories: a HuggingFace datasets directory (serialize.py:15-24), or the legacy `.npz`
    layout when `path` ends in `.npz`."""
    path = os.fspath(path)
    if not path.endswith(".npz"):
        huggingface_utils.trajectories_to_dataset(trajectories).save_to_disk(path)
        logging.info(f"Dumped demonstrations to {path}.")
        return
    trajectories = list(trajectories)
    if not trajectories:
        raise ValueError("no trajectories to save")
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
            import datasets
        except ImportError as e:
            raise ImportError("loading a HuggingFace-datasets demonstration directory needs the `datasets` package; "
                              "convert it to the .npz layout with the reference's tooling first") from e
        dataset = datasets.load_from_disk(path)
        if not isinstance(dataset, datasets.Dataset):
            raise ValueError(f"Expected to load a `datasets.Dataset` but got {type(dataset)}")
        return huggingface_utils.TrajectoryDatasetSequence(dataset)
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


def load_with_rewards(path) -> Sequence[TrajectoryWithRew]:
    """Loads a sequence of trajectories with rewards from a file (serialize.py:77-93)."""
    data = load(path)
    mismatched = [type(t) for t in data if not isinstance(t, TrajectoryWithRew)]
    if mismatched:
        raise ValueError(f"Expected all trajectories to be of type `TrajectoryWithRew`, but found {mismatched[0].__name__}")
    return data
