from stable_baselines3.common.on_policy_algorithm import OnPolicyAlgorithm


class PPO(OnPolicyAlgorithm):
    pass
