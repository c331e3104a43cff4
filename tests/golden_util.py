"""Helpers shared by the parity tests: load tests/golden/*.npz (produced by
oracle/make_golden.py from the reference's own modules) and rebuild nets from them."""
import os

import numpy as np
import torch as th

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

DISC_CASES = {
    # name: (algo, shaped, net kwargs in *port* vocabulary)
    "disc_gail_hc": ("gail", False, dict(hid_sizes=(32, 32), normalize_input=True)),
    "disc_gail_hc_minibatch": ("gail", False, dict(hid_sizes=(32, 32), normalize_input=True)),
    "disc_gail_nonorm": ("gail", False, dict(hid_sizes=(32,), normalize_input=False)),
    "disc_gail_cartpole": ("gail", False, dict(hid_sizes=(64, 64), normalize_input=False)),
    "disc_gail_allinputs": ("gail", False, dict(hid_sizes=(32, 32), normalize_input=True, use_next_state=True,
                                                use_done=True)),
    "disc_airl_hc": ("airl", True, dict(reward_hid_sizes=(32,), potential_hid_sizes=(32, 32), normalize_input=True)),
    "disc_airl_nonorm": ("airl", True, dict(reward_hid_sizes=(32, 32), potential_hid_sizes=(32,),
                                            normalize_input=False)),
}


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=True)


def sub(z, prefix):
    """{'a/b/c': v} -> {'c': v} for keys under `prefix/`."""
    p = prefix.rstrip("/") + "/"
    return {k[len(p):]: z[k] for k in z.files if k.startswith(p)}


def port_key(k: str) -> str:
    """reference state_dict key -> oracle-port state_dict key."""
    return (k.replace("_base._base.", "base.").replace("_base.", "base.")
            .replace("potential._potential_net.", "potential."))


def state_to_torch(state: dict, rename=port_key) -> dict:
    return {rename(k): th.as_tensor(np.array(v)) for k, v in state.items()}


def fixed_logp(M: np.ndarray, obs, acts) -> th.Tensor:
    """Same closed form as make_golden._FixedLogpPolicy.evaluate_actions."""
    obs = th.as_tensor(np.asarray(obs)).float()
    acts = th.as_tensor(np.asarray(acts)).float()
    if acts.ndim == 1:
        acts = acts[:, None]
    mean = obs @ th.as_tensor(M).T[:, : acts.shape[1]]
    return -0.5 * ((acts - mean) ** 2).sum(1) - 1.0
