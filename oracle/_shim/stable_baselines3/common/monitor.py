import gymnasium as gym


class Monitor(gym.Wrapper):
    def __init__(self, env, filename=None, **kwargs):
        super().__init__(env)
