"""Minimal `gymnasium.spaces` names (oracle-side shim; see package docstring)."""
import numpy as np


class Space:
    shape = None
    dtype = None

    def sample(self):
        raise NotImplementedError


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.asarray(low).shape
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
        self._rng = np.random.default_rng(0)

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return self._rng.uniform(lo, hi).astype(self.dtype)

    def __eq__(self, other):
        return (
            isinstance(other, Box)
            and self.shape == other.shape
            and np.array_equal(self.low, other.low)
            and np.array_equal(self.high, other.high)
        )


class Discrete(Space):
    def __init__(self, n, start=0):
        self.n = int(n)
        self.start = int(start)
        self.shape = ()
        self.dtype = np.dtype(np.int64)
        self._rng = np.random.default_rng(0)

    def sample(self):
        return np.int64(self.start + self._rng.integers(self.n))

    def __eq__(self, other):
        return isinstance(other, Discrete) and self.n == other.n


class Dict(Space, dict):
    pass


class MultiDiscrete(Space):
    pass


class MultiBinary(Space):
    pass
