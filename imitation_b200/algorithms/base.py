"""Base classes (mirror of imitation.algorithms.base:24-223): logger plumbing, the fixed-horizon
safety check with the reference's error text, demonstration handling."""
import abc
from typing import Iterable, Optional

from ..util import logger as imit_logger


class BaseImitationAlgorithm(abc.ABC):
    def __init__(self, *, custom_logger: Optional[imit_logger.HierarchicalLogger] = None,
                 allow_variable_horizon: bool = False):
        self._logger = custom_logger or imit_logger.configure()
        self.allow_variable_horizon = allow_variable_horizon
        if allow_variable_horizon:
            self.logger.warn("Running with `allow_variable_horizon` set to True. Some algorithms are biased towards "
                             "shorter or longer episodes, which may significantly confound results.")
        self._horizon = None

    @property
    def logger(self) -> imit_logger.HierarchicalLogger:
        return self._logger

    @logger.setter
    def logger(self, value) -> None:
        self._logger = value

    def _check_fixed_horizon(self, horizons: Iterable[int]) -> None:
        if self.allow_variable_horizon:
            return
        horizons = set(int(h) for h in horizons)
        if self._horizon is not None:
            horizons.add(self._horizon)
        if len(horizons) > 1:
            raise ValueError(
                f"Episodes of different length detected: {horizons}. Variable horizon environments are discouraged "
                "-- termination conditions leak information about reward. See "
                "https://imitation.readthedocs.io/en/latest/getting-started/variable-horizon.html for more "
                "information. If you are SURE you want to run imitation on a variable horizon task, then please "
                "pass in the flag: `allow_variable_horizon=True`.")
        elif len(horizons) == 1:
            self._horizon = horizons.pop()

    def __getstate__(self):
        state = self.__dict__.copy()
        del state["_logger"]
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._logger = state.get("_logger") or imit_logger.configure()


class DemonstrationAlgorithm(BaseImitationAlgorithm):
    def __init__(self, *, demonstrations, custom_logger=None, allow_variable_horizon: bool = False):
        super().__init__(custom_logger=custom_logger, allow_variable_horizon=allow_variable_horizon)
        if demonstrations is not None:
            self.set_demonstrations(demonstrations)

    @abc.abstractmethod
    def set_demonstrations(self, demonstrations) -> None:
        """Sets the demonstration data."""

    @property
    @abc.abstractmethod
    def policy(self):
        """Returns a policy imitating the demonstration data."""
