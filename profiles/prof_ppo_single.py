"""Two launches of the persistent PPO update (4096 rows, mb 64, 5 epochs, FeedForward32Policy + feature norm)
for `ncu --set full -k regex:k_ppo_update -s 1 -c 1`."""
import os
import sys

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imitation_b200 import _desc, _lib  # noqa: E402

pd = _desc.policy_desc(17, 6, False, 32, True)
N = 4096
rw = _lib.rollout_row_width(pd)
tbl = th.randn(N, rw, device="cuda")
tbl[:, 17 + 6] = -8.0 + 0.1 * th.randn(N, device="cuda")
P = (th.rand(pd.n_params, device="cuda") - 0.5) * 0.3
PN = th.cat([th.zeros(17), th.ones(17)]).cuda()
PC = th.zeros(1, dtype=th.int32, device="cuda")
M, V = th.zeros_like(P), th.zeros_like(P)
st = th.zeros(_lib.ST_WORDS, dtype=th.int64, device="cuda")
hp = _lib.PpoHparams(gamma=0.95, gae_lambda=0.95, clip_range=0.1, ent_coef=4e-6, vf_coef=0.11, max_grad_norm=0.8,
                     lr=2.6e-4, adam_eps=1e-5, n_epochs=5, batch_size=64, normalize_advantage=1)
for _ in range(2):
    _lib.ppo_update(pd, P, PN, PC, M, V, tbl, N, hp, None, 0, None, st)
th.cuda.synchronize()
print("ok", float(P.abs().mean()))
