"""Warm per-kernel times of one discriminator update at the tuned sizes (CUDA events over 200 back-to-back
launches each; GAIL 17/6, 32x32 + RunningNorm, demo_batch 8192 -> 16 384 rows)."""
import os
import sys

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imitation_b200 import _desc, _lib  # noqa: E402

Do, Da, B = 17, 6, 8192
d = _desc.disc_desc(Do, Da, hid_sizes=(32, 32), normalize_input=True)
tw, n = _desc.table_width(Do, Da), 2 * B
ld = _desc.batch_ld(n)
bw = tw + 1
P = th.randn(d.n_params, device="cuda") * 0.1
M, V = th.zeros_like(P), th.zeros_like(P)
NS = th.cat([th.zeros(d.base.din), th.ones(d.base.din)]).cuda()
NC = th.zeros(4, dtype=th.int32, device="cuda")
ws = th.zeros(_lib.disc_workspace_floats(d), device="cuda")
e_table = th.randn(60000, tw, device="cuda")
ring = th.randn(512, tw, device="cuda")
batch = th.zeros(bw, ld, device="cuda")
logits = th.zeros(n, device="cuda")
st_e = th.zeros(_lib.ST_WORDS, dtype=th.int64, device="cuda")
st_g = th.zeros(_lib.ST_WORDS, dtype=th.int64, device="cuda")
st_g[_lib.ST_RING_N] = 512
opt = _lib.Adam(lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8)
out = th.zeros(16, device="cuda")


def t(fn, reps=200):
    for _ in range(5):
        fn()
    th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    th.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


steps = {
    "sample_gather": lambda: _lib.disc_sample_gather(e_table, 60000, ring, 512, tw, B, 0, 1, st_e, st_g, batch, ld),
    "norm_stats": lambda: _lib.disc_norm_update(d, batch, ld, n, NS, NC, ws),
    "fwd_bwd": lambda: _lib.disc_fwd_bwd(d, P, NS, batch, ld, n, B, 1.0 / n, None, logits,
                                         _lib.IMB_F_ZERO_GRAD | _lib.IMB_F_TRAIN_NORM, ws),
    "reduce_adam": lambda: _lib.disc_reduce_adam(d, opt, P, M, V, 1.0, ws, st_g, out),
    "reduce": lambda: _lib.disc_reduce(d, ws, None),
    "adam": lambda: _lib.disc_adam(d, opt, P, M, V, None, 1.0, ws, st_g, out),
    "advance2": lambda: _lib.sample_advance2(B, 60000, st_e, st_g),
}
steps["fwd_bwd"]()
for k, f in steps.items():
    print(f"{k:<14s} {t(f):7.2f} us per launch (back-to-back, includes launch overhead)")
