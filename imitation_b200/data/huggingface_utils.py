"""Trajectories <-> HuggingFace `datasets` (the on-disk format the reference's `data.serialize.save` writes).

Interface of imitation.data.huggingface_utils (huggingface_utils.py:11-157): one dataset row per trajectory with the
columns `obs` [T + 1, ...], `acts` [T, ...], `infos` (T strings, each the JSON pickle of the step's info dict), `terminal`
and, for trajectories with rewards, `rews` [T].  `TrajectoryDatasetSequence` presents such a dataset as a read-only sequence
of `Trajectory` / `TrajectoryWithRew`; `trajectories_to_dict` / `trajectories_to_dataset` go the other way.

Differences from the reference, both invisible to the device path (the expert table is float32 either way): rows are
materialised through the dataset's numpy formatter restricted to the array columns, so arrays keep their stored dtype
(float32 observations stay float32; the reference's python-list round trip widens them to float64), and info strings are
decoded with `jsonpickle` when it is installed, else with `json` (identical for the plain-JSON infos gym environments
emit; this image ships `datasets` but not `jsonpickle`).
"""
import json
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from . import types

try:  # the reference's codec for info dicts; plain json covers the same strings for JSON-representable infos
    import jsonpickle as _codec

    _encode, _decode = _codec.encode, _codec.decode
except ImportError:  # pragma: no cover (depends on the image)
    _encode, _decode = (lambda o: json.dumps(o)), json.loads

_ARRAY_COLUMNS = ("obs", "acts", "rews")


class _LazyDecodedList(Sequence):
    """The info dicts of one trajectory, decoded from their strings on first access."""

    def __init__(self, encoded: Sequence[str]):
        self._encoded = encoded
        self._cache: Dict[int, Any] = {}

    def __len__(self) -> int:
        return len(self._encoded)

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            return [self[i] for i in range(*idx.indices(len(self)))]
        i = idx if idx >= 0 else idx + len(self)
        if i not in self._cache:
            self._cache[i] = _decode(self._encoded[i])
        return self._cache[i]


class TrajectoryDatasetSequence(Sequence):
    """A `datasets.Dataset` of trajectory rows as a sequence of trajectories (converted on access)."""

    def __init__(self, dataset):
        self._dataset = dataset
        cols = [c for c in _ARRAY_COLUMNS if c in dataset.features]
        self._arrays = dataset.with_format("numpy", columns=cols)
        self._trajectory_class = types.TrajectoryWithRew if "rews" in dataset.features else types.Trajectory

    def __len__(self) -> int:
        return len(self._dataset)

    def __getitem__(self, idx):
        if isinstance(idx, slice):  # (rows may have different lengths: convert one by one)
            return [self[i] for i in range(*idx.indices(len(self)))]
        arrays = self._arrays[idx]
        plain = self._dataset[idx]
        kwargs = dict(obs=np.asarray(arrays["obs"]), acts=np.asarray(arrays["acts"]),
                      infos=_LazyDecodedList(plain["infos"]), terminal=bool(plain["terminal"]))
        if self._trajectory_class is types.TrajectoryWithRew:
            kwargs["rews"] = np.asarray(arrays["rews"])
        return self._trajectory_class(**kwargs)

    @property
    def dataset(self):
        """The underlying dataset (unformatted: it can be saved to disk again)."""
        return self._dataset


def trajectories_to_dict(trajectories: Sequence[types.Trajectory]) -> Dict[str, List[Any]]:
    """One list per dataset column, one entry per trajectory (huggingface_utils.py:91-144)."""
    with_rew = [isinstance(t, types.TrajectoryWithRew) for t in trajectories]
    if any(with_rew) and not all(with_rew):
        raise ValueError("Some trajectories have rewards but not all")
    out: Dict[str, List[Any]] = dict(
        obs=[t.obs for t in trajectories],
        acts=[t.acts for t in trajectories],
        infos=[[_encode(info) for info in (t.infos if t.infos is not None else [{}] * len(t))] for t in trajectories],
        terminal=[bool(t.terminal) for t in trajectories],
    )
    if trajectories and all(with_rew):
        out["rews"] = [t.rews for t in trajectories]
    return out


def trajectories_to_dataset(trajectories: Sequence[types.Trajectory], info: Optional[Any] = None):
    """A `datasets.Dataset` with one row per trajectory (huggingface_utils.py:147-157)."""
    import datasets

    if isinstance(trajectories, TrajectoryDatasetSequence):
        return trajectories.dataset
    return datasets.Dataset.from_dict(trajectories_to_dict(trajectories), info=info)
