"""Minimal observation/action spaces (gymnasium is not a dependency of this package).

Any object with the same attributes works (duck typing): Box -> .shape/.dtype/.low/.high,
Discrete -> .n.  gymnasium.spaces.Box/Discrete instances are accepted everywhere.
"""
import numpy as np


class Space:
    shape = None
    dtype = None


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.asarray(low).shape
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()

    def __eq__(self, other):
        return (is_box(other) and tuple(other.shape) == self.shape and np.array_equal(other.low, self.low)
                and np.array_equal(other.high, self.high))

    def __repr__(self):
        return f"Box({self.shape}, {self.dtype})"


class Discrete(Space):
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def __eq__(self, other):
        return is_discrete(other) and int(other.n) == self.n

    def __repr__(self):
        return f"Discrete({self.n})"


def is_discrete(space) -> bool:
    return hasattr(space, "n") and not hasattr(space, "nvec")


def is_box(space) -> bool:
    return hasattr(space, "low") and hasattr(space, "high") and hasattr(space, "shape")


def flat_dim(space) -> int:
    """SB3 preprocessing.get_flattened_obs_dim for Box / Discrete (one-hot width)."""
    if is_discrete(space):
        return int(space.n)
    if is_box(space):
        return int(np.prod(space.shape))
    raise NotImplementedError(f"unsupported space {space!r} (Box and Discrete only; SURVEY.md section 2 row 5)")
