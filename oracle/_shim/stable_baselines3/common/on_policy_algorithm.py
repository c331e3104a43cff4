from stable_baselines3.common.base_class import BaseAlgorithm


class OnPolicyAlgorithm(BaseAlgorithm):
    n_steps = 0
