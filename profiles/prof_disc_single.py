"""A few launches of the fused discriminator kernel at 2^20 rows (GAIL 17/6, 32x32) for
`ncu --set full -k regex:k_disc_fwdbwd -s 2 -c 1`."""
import os
import sys

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imitation_b200 import _desc, _lib  # noqa: E402

d = _desc.disc_desc(17, 6)
n = 1 << 20
ld = _desc.batch_ld(n)
P = (th.rand(d.n_params, device="cuda") - 0.5) * 0.5
NS = th.ones(128, device="cuda")
ws = th.zeros(_lib.disc_workspace_floats(d), device="cuda")
batch = th.randn(_desc.batch_rows(17, 6), ld, device="cuda")
logits = th.empty(n, device="cuda")
for _ in range(4):
    _lib.disc_fwd_bwd(d, P, NS, batch, ld, n, n // 2, 1.0 / n, None, logits, _lib.IMB_F_ZERO_GRAD, ws)
th.cuda.synchronize()
print("ok")
