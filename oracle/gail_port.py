"""GAIL/AIRL round driver restated on the CPU.  TEST INFRASTRUCTURE.

Follows algorithms/adversarial/common.py:112-266 (constructor wiring), :306-315 (expert
iterator), :391-425 (train_gen), :427-461 (train); algorithms/base.py:226-288
(make_data_loader: DataLoader(shuffle=True, drop_last=True) + transitions_collate_fn,
data/types.py:447-474) and util/util.py:215-241 (endless_iter, which builds and drops
iterators before the first batch and therefore advances the global torch RNG).

This is the loop `bench.py --impl reference` times: host VecEnv -> BufferingPort ->
RewardRelabelPort -> PPOPort (SB3 restatement) -> flatten -> ReplayBufferPort ->
DiscTrainerPort.
"""
import itertools
from typing import Dict, Optional

import numpy as np
import torch as th
from torch.utils import data as th_data

from . import data_port, disc_port, nets_port, ppo_port


class _TransitionsDataset(th_data.Dataset):
    def __init__(self, trans: Dict[str, np.ndarray]):
        self.trans = trans

    def __len__(self):
        return len(self.trans["obs"])

    def __getitem__(self, i):
        return {k: v[i] for k, v in self.trans.items()}


def _collate(batch):
    small = [{k: np.array(v) for k, v in s.items() if k in ("acts", "dones")} for s in batch]
    out = th_data.dataloader.default_collate(small)
    out["infos"] = [s["infos"] for s in batch]
    out["obs"] = np.stack([s["obs"] for s in batch])
    out["next_obs"] = np.stack([s["next_obs"] for s in batch])
    return out


def expert_iterator_port(trans: Dict[str, np.ndarray], batch_size: int):
    if batch_size <= 0:
        raise ValueError(f"batch_size={batch_size} must be positive.")
    if len(trans["obs"]) < batch_size:
        raise ValueError(f"Number of transitions in `demonstrations` {len(trans['obs'])} "
                         f"is smaller than batch size {batch_size}.")
    if "infos" not in trans:
        trans = dict(trans, infos=np.array([{}] * len(trans["obs"])))
    loader = th_data.DataLoader(_TransitionsDataset(trans), batch_size=batch_size, shuffle=True,
                                drop_last=True, collate_fn=_collate)
    # endless_iter: `iter(it) == it` probe, then get_first_iter_element -> one more iter()+next()
    _probe = iter(loader)
    del _probe
    next(iter(loader))
    return itertools.chain.from_iterable(itertools.repeat(loader))


class AdversarialPort:
    def __init__(self, *, venv, expert: Dict[str, np.ndarray], demo_batch_size: int, gen: ppo_port.PPOPort,
                 reward_net: th.nn.Module, airl: bool = False, demo_minibatch_size: Optional[int] = None,
                 n_disc_updates_per_round: int = 2, gen_replay_buffer_capacity: Optional[int] = None,
                 normalize_output: bool = False, disc_opt_kwargs=None):
        self.venv, self.gen, self.net, self.airl = venv, gen, reward_net, airl
        self.n_disc = n_disc_updates_per_round
        n_actions = venv.action_space.n if hasattr(venv.action_space, "n") else None
        self.n_actions = n_actions
        logp_fn = (lambda o, a: gen.policy.evaluate_actions(o, a)[1])
        self.disc = disc_port.DiscTrainerPort(reward_net, demo_batch_size, demo_minibatch_size, airl=airl,
                                              n_actions=n_actions, logp_fn=logp_fn, opt_kwargs=disc_opt_kwargs)
        self.expert_iter = expert_iterator_port(expert, demo_batch_size)
        self.buffering = data_port.BufferingPort(venv)
        # GAIL trains the generator on -logsigmoid(-logit) without output norm (gail.py:82-83;
        # SURVEY Appendix A.6); AIRL on the (optionally output-normalised) shaped net.
        self.out_norm = nets_port.OutputNormPort() if (airl and normalize_output) else None

        def reward_fn(obs, acts, next_obs, dones):
            r = nets_port.predict_port(self.net, obs, acts, next_obs, dones, n_actions, gail_transform=not airl)
            return self.out_norm(r) if self.out_norm is not None else r

        self.venv_train = data_port.RewardRelabelPort(self.buffering, reward_fn)
        gen.set_env(self.venv_train)
        self.gen_train_timesteps = venv.num_envs * gen.n_steps
        cap = gen_replay_buffer_capacity or self.gen_train_timesteps
        act_shape = venv.action_space.shape
        self.replay = data_port.ReplayBufferPort(cap, venv.observation_space.shape, act_shape,
                                                 venv.observation_space.dtype, venv.action_space.dtype)
        self.horizon = data_port.FixedHorizonCheckPort()
        self._global_step = 0
        self.stats = []

    def train_gen(self):
        self.gen.learn(self.gen_train_timesteps, reset_num_timesteps=False)
        self._global_step += 1
        trajs, ep_lens = self.buffering.pop_trajectories()
        self.horizon.check(ep_lens)
        self.last_gen_samples = data_port.flatten_port(trajs)
        self.replay.store(self.last_gen_samples)

    def train_disc(self, expert=None, gen=None):
        if expert is None:
            expert = next(self.expert_iter)
        if gen is None:
            if self.replay.size() == 0:
                raise RuntimeError("No generator samples for training. Call `train_gen()` first.")
            gen = self.replay.sample(self.disc.B)
        expert, gen = dict(expert), dict(gen)
        for d in (expert, gen):
            for k in ("obs", "acts", "next_obs", "dones"):
                if isinstance(d[k], th.Tensor):
                    d[k] = d[k].detach().numpy()
        return self.disc.train_disc(expert, gen)

    def train(self, total_timesteps: int, callback=None):
        n_rounds = total_timesteps // self.gen_train_timesteps
        assert n_rounds >= 1
        for r in range(n_rounds):
            self.train_gen()
            for _ in range(self.n_disc):
                self.net.train()
                try:
                    self.stats.append(self.train_disc())
                finally:
                    self.net.eval()
            if callback:
                callback(r)
