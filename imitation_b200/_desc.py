"""Builders for the C-ABI descriptors (include/imb.h) + flat parameter layouts.

Parameter layout = torch nn.Linear order (weight [out][in] row-major, then bias) so that
nn.Parameters can alias slices of the flat vectors the kernels read.
"""
from typing import List, Sequence, Tuple

import numpy as np

from . import _lib


def mlp_param_shapes(din: int, hid_sizes: Sequence[int], n_out: int = 1) -> List[Tuple[str, Tuple[int, ...]]]:
    """[(name, shape)] in flat order for build_mlp's Linear layers (util/networks.py:264-279)."""
    shapes, prev = [], din
    for i, h in enumerate(hid_sizes):
        shapes += [(f"dense{i}.weight", (h, prev)), (f"dense{i}.bias", (h,))]
        prev = h
    shapes += [("dense_final.weight", (n_out, prev)), ("dense_final.bias", (n_out,))]
    return shapes


def mlp_desc(din: int, hid_sizes: Sequence[int], has_norm: bool, param_off: int, norm_off: int, count_idx: int,
             n_out: int = 1, eps: float = 1e-5) -> Tuple[_lib.Mlp, int]:
    hid = list(hid_sizes)
    if len(hid) > 2:
        raise NotImplementedError("fused sm_100a kernels support at most 2 hidden layers (reference configs use 1-2)")
    if any(h < 1 or h > _lib.IMB_MAX_HIDDEN for h in hid):
        raise NotImplementedError(f"hidden widths must be in [1, {_lib.IMB_MAX_HIDDEN}]")
    if not 1 <= din <= _lib.IMB_MAX_DIN:
        raise NotImplementedError(f"MLP input width must be in [1, {_lib.IMB_MAX_DIN}], got {din}")
    m = _lib.Mlp(din=din, n_hidden=len(hid), h1=hid[0] if hid else 0, h2=hid[1] if len(hid) > 1 else 0,
                 n_out=n_out, has_norm=int(has_norm), param_off=param_off, norm_off=norm_off, count_idx=count_idx,
                 norm_eps=eps)
    n = sum(int(np.prod(s)) for _, s in mlp_param_shapes(din, hid, n_out))
    return m, n


def disc_desc(d_obs: int, d_act: int, *, hid_sizes=(32, 32), use_state=True, use_action=True, use_next_state=False,
              use_done=False, normalize_input=False, shaped=False, potential_hid_sizes=(32, 32), gamma=0.99,
              subtract_logp=False) -> _lib.DiscDesc:
    """d_act is the FLATTENED action width (Discrete(n) -> n, one-hot)."""
    din = d_obs * use_state + d_act * use_action + d_obs * use_next_state + int(use_done)
    if din < 1:
        raise ValueError("reward net needs at least one input")
    base, nb = mlp_desc(din, hid_sizes, normalize_input, 0, 0, 0)
    d = _lib.DiscDesc(d_obs=d_obs, d_act=d_act, use_state=int(use_state), use_action=int(use_action),
                      use_next_state=int(use_next_state), use_done=int(use_done), base=base, shaped=int(shaped),
                      gamma=gamma, subtract_logp=int(subtract_logp), n_params=nb)
    if shaped:
        pot, npot = mlp_desc(d_obs, potential_hid_sizes, normalize_input, nb, 2 * din, 1)
        d.potential = pot
        d.n_params = nb + npot
    return d


def disc_norm_floats(d: _lib.DiscDesc) -> int:
    return 2 * d.base.din + (2 * d.potential.din if d.shaped else 0)


def batch_ld(n: int) -> int:
    return max(_lib.IMB_TILE_ROWS, (n + _lib.IMB_TILE_ROWS - 1) // _lib.IMB_TILE_ROWS * _lib.IMB_TILE_ROWS)


def batch_rows(d_obs: int, d_act: int) -> int:
    """feature rows of a disc batch: obs | act | next_obs | done | logp"""
    return 2 * d_obs + d_act + 2


def table_width(d_obs: int, d_act: int) -> int:
    return 2 * d_obs + d_act + 1


def policy_param_shapes(d_obs: int, d_act: int, discrete: bool, hidden: int):
    """SB3 ActorCriticPolicy(net_arch=[h, h]) parameter names/shapes (separate pi / vf towers)."""
    s = [
        ("mlp_extractor.policy_net.0.weight", (hidden, d_obs)), ("mlp_extractor.policy_net.0.bias", (hidden,)),
        ("mlp_extractor.policy_net.2.weight", (hidden, hidden)), ("mlp_extractor.policy_net.2.bias", (hidden,)),
        ("mlp_extractor.value_net.0.weight", (hidden, d_obs)), ("mlp_extractor.value_net.0.bias", (hidden,)),
        ("mlp_extractor.value_net.2.weight", (hidden, hidden)), ("mlp_extractor.value_net.2.bias", (hidden,)),
        ("action_net.weight", (d_act, hidden)), ("action_net.bias", (d_act,)),
        ("value_net.weight", (1, hidden)), ("value_net.bias", (1,)),
    ]
    if not discrete:
        s.append(("log_std", (d_act,)))
    return s


def policy_desc(d_obs: int, d_act: int, discrete: bool, hidden: int = 32, has_norm: bool = False,
                eps: float = 1e-5) -> _lib.PolicyDesc:
    if hidden < 1 or hidden > 64:
        raise NotImplementedError("policy tower width must be in [1, 64]")
    offs, o = {}, 0
    for name, shape in policy_param_shapes(d_obs, d_act, discrete, hidden):
        offs[name] = o
        o += int(np.prod(shape))
    return _lib.PolicyDesc(
        d_obs=d_obs, d_act=d_act, discrete=int(discrete), hidden=hidden, has_norm=int(has_norm), norm_eps=eps,
        off_pi_w1=offs["mlp_extractor.policy_net.0.weight"], off_pi_b1=offs["mlp_extractor.policy_net.0.bias"],
        off_pi_w2=offs["mlp_extractor.policy_net.2.weight"], off_pi_b2=offs["mlp_extractor.policy_net.2.bias"],
        off_vf_w1=offs["mlp_extractor.value_net.0.weight"], off_vf_b1=offs["mlp_extractor.value_net.0.bias"],
        off_vf_w2=offs["mlp_extractor.value_net.2.weight"], off_vf_b2=offs["mlp_extractor.value_net.2.bias"],
        off_act_w=offs["action_net.weight"], off_act_b=offs["action_net.bias"],
        off_val_w=offs["value_net.weight"], off_val_b=offs["value_net.bias"],
        off_log_std=offs.get("log_std", 0), n_params=o)


def synth_env_params(d_obs: int, d_act: int, seed: int) -> np.ndarray:
    """Flat [A(Do x Do) | Bm(Do x Da) | c(Do) | w(Do)] of the synthetic MuJoCo-shaped env
    (SURVEY.md section 8d); same draw order as the CPU twin used by the parity tests."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((d_obs, d_obs))
    A *= 0.9 / np.max(np.abs(np.linalg.eigvals(A)))
    Bm = 0.5 * rng.standard_normal((d_obs, d_act))
    c = 0.1 * rng.standard_normal(d_obs)
    w = rng.standard_normal(d_obs) / np.sqrt(d_obs)
    return np.concatenate([A.ravel(), Bm.ravel(), c, w]).astype(np.float32)
