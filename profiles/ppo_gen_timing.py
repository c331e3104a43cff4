"""us per optimiser step of the two PPO update kernels (CUDA events over whole launches, 4096 rollout rows, 5 epochs):
k_ppo_update (tower 32, minibatch 64) against k_ppo_update_gen on the same shape and on the shapes only it covers."""
import os
import sys

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imitation_b200 import _desc, _lib  # noqa: E402


def run(Do, Da, hidden, mb, force_general=False, N=4096, epochs=5, reps=5):
    os.environ["IMB_PPO_FORCE_GENERAL"] = "1" if force_general else "0"
    pd = _desc.policy_desc(Do, Da, False, hidden, True)
    rw = _lib.rollout_row_width(pd)
    tbl = th.randn(N, rw, device="cuda")
    tbl[:, Do + Da] = -8.0 + 0.1 * th.randn(N, device="cuda")
    P = (th.rand(pd.n_params, device="cuda") - 0.5) * 0.3
    PN = th.cat([th.zeros(Do), th.ones(Do)]).cuda()
    PC = th.zeros(1, dtype=th.int32, device="cuda")
    M, V = th.zeros_like(P), th.zeros_like(P)
    st = th.zeros(_lib.ST_WORDS, dtype=th.int64, device="cuda")
    hp = _lib.PpoHparams(gamma=0.95, gae_lambda=0.95, clip_range=0.1, ent_coef=4e-6, vf_coef=0.11, max_grad_norm=0.8,
                         lr=2.6e-4, adam_eps=1e-5, n_epochs=epochs, batch_size=mb, normalize_advantage=1)
    for _ in range(2):
        _lib.ppo_update(pd, P, PN, PC, M, V, tbl, N, hp, None, 0, None, st)
    th.cuda.synchronize()
    a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        _lib.ppo_update(pd, P, PN, PC, M, V, tbl, N, hp, None, 0, None, st)
    b.record()
    b.synchronize()
    ms = a.elapsed_time(b) / reps
    steps = epochs * ((N + mb - 1) // mb)
    assert bool(th.isfinite(P).all())
    kern = "k_ppo_update_gen" if (force_general or hidden > 32 or mb > 64) else "k_ppo_update"
    print(f"obs {Do:2d} act {Da} tower {hidden:2d} minibatch {mb:4d}  {kern:<17s} {ms:8.3f} ms / launch  {steps:4d} steps  "
          f"{ms * 1e3 / steps:7.2f} us / step  {N * epochs / ms / 1e3:7.2f} M rows/s")


if __name__ == "__main__":
    run(17, 6, 32, 64)
    run(17, 6, 32, 64, force_general=True)
    run(17, 6, 64, 64)
    run(17, 6, 32, 128)
    run(17, 6, 32, 512)
    run(17, 6, 64, 512)
    run(11, 3, 32, 512)
    run(27, 8, 64, 128)
    run(27, 8, 32, 16)
    run(27, 8, 32, 16, force_general=True)
