"""SB3 2.2.x behaviour for NON-image Box / Discrete spaces only."""
import numpy as np
import torch as th
from gymnasium import spaces


def is_image_space(observation_space, check_channels=False, normalized_image=False):
    return False


def get_flattened_obs_dim(space):
    if isinstance(space, spaces.Discrete):
        return int(space.n)
    return int(np.prod(space.shape))


def preprocess_obs(obs, observation_space, normalize_images=True):
    if isinstance(observation_space, spaces.Box):
        return obs.float()
    if isinstance(observation_space, spaces.Discrete):
        return th.nn.functional.one_hot(obs.long(), num_classes=int(observation_space.n)).float()
    raise NotImplementedError(type(observation_space))
