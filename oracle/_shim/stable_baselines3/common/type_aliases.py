"""Names-only stub (oracle/_shim): type aliases the reference's signatures mention."""
from typing import Any, Callable, Dict, Union

import numpy as np

Schedule = Callable[[float], float]
GymObs = Union[np.ndarray, Dict[str, np.ndarray]]
GymEnv = Any
MaybeCallback = Any
