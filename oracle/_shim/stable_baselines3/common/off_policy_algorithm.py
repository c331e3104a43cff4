from stable_baselines3.common.base_class import BaseAlgorithm


class OffPolicyAlgorithm(BaseAlgorithm):
    pass
