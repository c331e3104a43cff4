"""Names-only stand-in for `gymnasium` (oracle process only; never imported by the product).

TEST INFRASTRUCTURE.  `gymnasium` is not installed in this image; the reference's
modules import it for type names and for `spaces.Box/Discrete` attribute access only
(SURVEY.md Appendix B).  Nothing here carries arithmetic under test.
"""
from . import spaces  # noqa: F401
from .spaces import Space  # noqa: F401


class Env:
    observation_space = None
    action_space = None

    def reset(self, **kwargs):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self.observation_space = getattr(env, "observation_space", None)
        self.action_space = getattr(env, "action_space", None)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def step(self, action):
        return self.env.step(action)


class ObservationWrapper(Wrapper):
    pass


def make(*args, **kwargs):
    raise RuntimeError("gymnasium shim: no environments registered")
