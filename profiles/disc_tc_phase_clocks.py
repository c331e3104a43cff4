"""Per-phase cycle counts of the tcgen05 discriminator kernel (CTA 0, epilogue thread of row 0 / column half 0).
Build the timing variant first:
  IMB_VARIANT=_timing IMB_EXTRA_NVCC_FLAGS=-DIMB_TC_TIMING python imitation_b200/_build.py
then run with IMB_VARIANT=_timing."""
import os
import sys

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imitation_b200 import _desc, _lib as L  # noqa: E402

NAMES = ["wait: tile staged (bulk copy)", "E0 load + normalise + split + TMEM store + signal", "wait: M1 (z1)",
         "E1 ld + relu + split + stores + signal", "wait: M2 (z2)", "E2 ld + head + BCE + dz2 + split + TMEM store + signal",
         "wait: previous tile's weight-gradient MMAs drained", "E2 tail: dz2 atoms + x recompute + x atoms", "wait: M3 (dz1')",
         "E3 ld + mask + split + stores + proxy fence + signal"]
Do, Da = 17, 6
d = _desc.disc_desc(Do, Da)
P = (th.rand(d.n_params, device="cuda") - 0.5) * 0.6
NS = th.zeros(2, device="cuda")
ws = th.zeros(L.disc_workspace_floats(d), device="cuda")
al = (d.n_params + 31) // 32 * 32
part_off = al + 32 + 32 + 128 + 32 + 16384 * 132
for lg2 in (14, 20):
    n = 1 << lg2
    ld = _desc.batch_ld(n)
    batch = th.randn(Do + Da, ld, device="cuda")
    logits = th.empty(n, device="cuda")
    for _ in range(3):
        L.disc_fwd_bwd(d, P, NS, batch, ld, n, n // 2, 1.0 / n, None, logits, L.IMB_F_ZERO_GRAD, ws)
    th.cuda.synchronize()
    t = ws[part_off + d.n_params + 5: part_off + d.n_params + 16].cpu().numpy()
    tiles = t[10]
    print(f"rows {n}: CTA 0 processed {int(tiles)} tiles; cycles per tile by phase (total {t[:10].sum() / tiles:.0f}):")
    for i, nm in enumerate(NAMES):
        print(f"  {t[i] / tiles:8.0f}  {nm}")
