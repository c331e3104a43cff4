from torch import nn


class BasePolicy(nn.Module):
    squash_output = False

    def predict(self, observation, state=None, episode_start=None, deterministic=False):
        raise NotImplementedError


class ActorCriticPolicy(BasePolicy):
    def evaluate_actions(self, obs, actions):
        raise NotImplementedError
