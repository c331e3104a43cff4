"""Load serialized reward functions of different types (mirror of imitation.rewards.serialize:1-282).

`scripts/train_adversarial.py:25-35` checkpoints with `th.save(reward_net, path)` (whole-module pickle);
`load_reward(reward_type, path, venv)` turns such a file back into a `RewardFn(state, action, next_state, done)
-> np.ndarray[batch]`.  The fused reward networks pickle like any `nn.Module`; their engines re-alias the loaded
parameters on first use, so a loaded network predicts through the same CUDA kernels.
"""
from typing import Any, Callable, Dict, Iterable, Optional, Sequence, Type

import numpy as np
import torch as th

from . import reward_nets

RewardFn = Callable[[np.ndarray, np.ndarray, np.ndarray, np.ndarray], np.ndarray]


class ValidateRewardFn:
    """Checks that the reward vector has one entry per input row (serialize.py:18-47)."""

    def __init__(self, reward_fn: RewardFn) -> None:
        self.reward_fn = reward_fn

    def __call__(self, state, action, next_state, done) -> np.ndarray:
        rew = self.reward_fn(state, action, next_state, done)
        assert rew.shape == (len(state),)
        return rew


def _strip_wrappers(reward_net, wrapper_types: Iterable[Type[reward_nets.RewardNetWrapper]]):
    """Remove the listed wrapper types, outermost first, for as long as they match (serialize.py:50-78)."""
    for wrapper_type in wrapper_types:
        assert issubclass(wrapper_type, reward_nets.RewardNetWrapper), f"trying to remove non-wrapper type {wrapper_type}"
        if isinstance(reward_net, wrapper_type):
            reward_net = reward_net.base
        else:
            break
    return reward_net


def _make_functional(net, attr: str = "predict", default_kwargs: Optional[Dict[str, Any]] = None, **kwargs) -> RewardFn:
    kw = dict(default_kwargs or {})
    kw.update(kwargs)
    return lambda *args: getattr(net, attr)(*args, **kw)


def _wrapper_chain(reward_net) -> Sequence[type]:
    chain, w = [], reward_net
    while hasattr(w, "base"):
        chain.append(w.__class__)
        w = w.base
    chain.append(w.__class__)
    return chain


def _validate_wrapper_structure(reward_net, prefixes: Iterable[Sequence[type]]):
    """The outermost wrappers must match one of `prefixes` (serialize.py:107-161)."""
    chain = _wrapper_chain(reward_net)

    def matches(prefix):
        return len(prefix) <= len(chain) and all(issubclass(c, p) for c, p in zip(chain, prefix))

    prefixes = list(prefixes)
    if any(matches(p) for p in prefixes):
        return reward_net
    want = " or ".join("[" + ",".join(t.__name__ for t in p) + "]" for p in prefixes)
    raise TypeError(f"Wrapper structure should match {want} but found [" + ",".join(t.__name__ for t in chain) + "]")


def _load(path):
    return th.load(str(path), weights_only=False)


def load_zero(path, venv) -> RewardFn:
    del path, venv
    return lambda state, action, next_state, done: np.zeros(state.shape[0])


reward_registry: Dict[str, Callable[..., RewardFn]] = {
    "zero": load_zero,
    "RewardNet_shaped": lambda path, _, **kw: ValidateRewardFn(_make_functional(
        _validate_wrapper_structure(_load(path), [(reward_nets.ShapedRewardNet,)]))),
    "RewardNet_unshaped": lambda path, _, **kw: ValidateRewardFn(_make_functional(
        _strip_wrappers(_load(path), (reward_nets.ShapedRewardNet,)))),
    "RewardNet_normalized": lambda path, _, **kw: ValidateRewardFn(_make_functional(
        _validate_wrapper_structure(_load(path), [(reward_nets.NormalizedRewardNet,)]),
        attr="predict_processed", default_kwargs={"update_stats": False}, **kw)),
    "RewardNet_unnormalized": lambda path, _, **kw: ValidateRewardFn(_make_functional(
        _strip_wrappers(_load(path), (reward_nets.NormalizedRewardNet,)))),
    "RewardNet_std_added": lambda path, _, **kw: ValidateRewardFn(_make_functional(
        _strip_wrappers(_validate_wrapper_structure(
            _load(path), [(reward_nets.AddSTDRewardWrapper,),
                          (reward_nets.NormalizedRewardNet, reward_nets.AddSTDRewardWrapper)]),
            (reward_nets.NormalizedRewardNet,)),
        attr="predict_processed", default_kwargs={}, **kw)),
}


def load_reward(reward_type: str, reward_path: str, venv, **kwargs: Any) -> RewardFn:
    """Load a serialized reward of `reward_type` (serialize.py:263-282)."""
    if reward_type not in reward_registry:
        raise KeyError(f"Key '{reward_type}' is not registered")
    return reward_registry[reward_type](reward_path, venv, **kwargs)
