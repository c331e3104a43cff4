"""`RewardVecEnvWrapper` (mirror of imitation.rewards.reward_wrapper:40-133).

In the reference this wrapper calls `reward_fn(old_obs, acts, terminal-fixed obs, dones)` on the
host after every env step (H2D, MLP, D2H).  Here it only DESCRIBES the relabel: the rollout
kernel evaluates the reward network in place (csrc/imb_rollout.cu), so `reward_fn` must be the
`predict_processed` of a fusable reward network.
"""
import collections

from . import reward_nets


class WrappedRewardCallback:
    def __init__(self, episode_rewards):
        self.episode_rewards = episode_rewards
        self.logger = None

    def init_callback(self, model):
        self.logger = model.logger

    def on_rollout_start(self):
        if len(self.episode_rewards) == 0 or self.logger is None:
            return
        self.logger.record("rollout/ep_rew_wrapped_mean", sum(self.episode_rewards) / len(self.episode_rewards))


class RewardVecEnvWrapper:
    def __init__(self, venv, reward_fn, ep_history: int = 100):
        assert not isinstance(venv, RewardVecEnvWrapper)
        self.venv = venv
        self.num_envs = venv.num_envs
        self.observation_space, self.action_space = venv.observation_space, venv.action_space
        self.episode_rewards = collections.deque(maxlen=ep_history)
        self.reward_fn = reward_fn
        self.reset()

    def make_log_callback(self) -> WrappedRewardCallback:
        return WrappedRewardCallback(self.episode_rewards)

    def reset(self):
        return self.venv.reset()

    def resolve(self):
        """-> (fused net with engine, reward_mode, NormalizedRewardNet or None) for the rollout kernel."""
        net = getattr(self.reward_fn, "__self__", None)
        if not isinstance(net, reward_nets.RewardNet) or getattr(self.reward_fn, "__name__", "") != "predict_processed":
            raise NotImplementedError("RewardVecEnvWrapper on the GPU path needs reward_fn = <RewardNet>.predict_processed")
        mode, out_norm = 2, None
        from ..algorithms.adversarial import gail

        if isinstance(net, gail.RewardNetFromDiscriminatorLogit):
            mode, net = 1, net.base
        if isinstance(net, reward_nets.NormalizedRewardNet):
            if mode == 1:
                net = net.base  # GAIL bypasses the output normaliser (gail.py:82-83)
            else:
                out_norm, net = net, net.base
        while isinstance(net, reward_nets.RewardNetWrapper) and not hasattr(net, "_engine"):
            net = net.base
        if not hasattr(net, "_engine"):
            raise NotImplementedError(f"reward net {type(net).__name__} has no fused sm_100a implementation")
        return net, mode, out_norm
