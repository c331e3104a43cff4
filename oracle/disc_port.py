"""Discriminator update restated on torch-CPU.  TEST INFRASTRUCTURE.

Follows algorithms/adversarial/common.py:27-92 (compute_train_stats), :317-389
(train_disc), :521-632 (_make_disc_train_batches); gail.py:135-160; airl.py:67-119.
"""
from typing import Callable, Dict, Mapping, Optional

import numpy as np
import torch as th
from torch.nn import functional as F

from .nets_port import preprocess_port


def train_stats_port(logits: th.Tensor, labels: th.Tensor, loss: th.Tensor) -> Dict[str, float]:
    """common.py:27-92.  labels: 1 = expert, 0 = generator; logit > 0 predicts expert."""
    with th.no_grad():
        pred_gen = logits < 0
        true_gen = labels == 0
        true_exp = ~true_gen
        n_gen = float(true_gen.long().sum())
        n_lab = float(len(labels))
        n_exp = n_lab - n_gen
        pct_exp = n_exp / n_lab if n_lab > 0 else float("nan")
        n_exp_pred = int(n_lab - pred_gen.long().sum())
        pct_exp_pred = n_exp_pred / n_lab if n_lab > 0 else float("nan")
        correct = pred_gen == true_gen
        acc = correct.float().mean()
        n_ok_exp = (true_exp & correct).sum()
        exp_acc = float("nan") if n_exp < 1 else n_ok_exp.item() / n_exp
        n_ok_gen = (true_gen & correct).sum()
        gen_acc = n_ok_gen / float(max(1, n_gen))
        ent = th.distributions.Bernoulli(logits=logits).entropy().mean()
    return {
        "disc_loss": float(loss.mean()), "disc_acc": float(acc), "disc_acc_expert": float(exp_acc),
        "disc_acc_gen": float(gen_acc), "disc_entropy": float(ent),
        "disc_proportion_expert_true": float(pct_exp), "disc_proportion_expert_pred": float(pct_exp_pred),
        "n_expert": float(n_exp), "n_generated": float(n_gen),
    }


class DiscTrainerPort:
    """One discriminator + optimiser; `train_disc` = one optimiser step over B expert +
    B generator rows split into minibatches (gradient accumulation)."""

    def __init__(self, net: th.nn.Module, demo_batch_size: int, demo_minibatch_size: Optional[int] = None,
                 airl: bool = False, n_actions: Optional[int] = None,
                 logp_fn: Optional[Callable[[th.Tensor, th.Tensor], th.Tensor]] = None,
                 opt_kwargs: Optional[Mapping] = None):
        self.net = net
        self.B = demo_batch_size
        self.mb = demo_minibatch_size or demo_batch_size
        if self.B % self.mb != 0:
            raise ValueError("Batch size must be a multiple of minibatch size.")
        self.airl = airl
        self.n_actions = n_actions
        self.logp_fn = logp_fn
        self.opt = th.optim.Adam(net.parameters(), **(opt_kwargs or {}))
        self.last_logits = None

    def logits(self, s, a, ns, d, logp):
        out = self.net(s, a, ns, d)
        return out - logp if self.airl else out

    def train_disc(self, expert: Mapping[str, np.ndarray], gen: Mapping[str, np.ndarray]) -> Dict[str, float]:
        if not (len(gen["obs"]) == len(expert["obs"]) == self.B):
            raise ValueError("Need to have exactly `demo_batch_size` number of expert and generator samples, each. "
                             f"(n_gen={len(gen['obs'])} n_expert={len(expert['obs'])} demo_batch_size={self.B})")
        self.opt.zero_grad()
        for start in range(0, self.B, self.mb):
            sl = slice(start, start + self.mb)
            obs = np.concatenate([expert["obs"][sl], gen["obs"][sl]])
            acts = np.concatenate([expert["acts"][sl], gen["acts"][sl]])
            nobs = np.concatenate([expert["next_obs"][sl], gen["next_obs"][sl]])
            dones = np.concatenate([expert["dones"][sl], gen["dones"][sl]])
            labels = np.concatenate([np.ones(self.mb, dtype=int), np.zeros(self.mb, dtype=int)])
            logp = None
            if self.logp_fn is not None:
                with th.no_grad():
                    logp = self.logp_fn(th.as_tensor(obs), th.as_tensor(acts)).reshape(2 * self.mb)
            s, a, ns, d = preprocess_port(obs, acts, nobs, dones, self.n_actions)
            labels_th = th.as_tensor(labels)
            logits = self.logits(s, a, ns, d, logp)
            loss = F.binary_cross_entropy_with_logits(logits, labels_th.float())
            loss = loss * (self.mb / self.B)
            loss.backward()
        self.opt.step()
        self.last_logits = logits.detach()
        return train_stats_port(logits.detach(), labels_th, loss.detach())
