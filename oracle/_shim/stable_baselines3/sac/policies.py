from stable_baselines3.common.policies import BasePolicy


class SACPolicy(BasePolicy):
    pass
