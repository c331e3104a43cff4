"""Rows-per-launch sweep of the fused discriminator kernel (SURVEY.md section 8d caveat: at the tuned
sizes a disc update is below launch latency, so the roofline fraction is quoted on a size sweep) and
per-launch timing of the other hot kernels.  Run on the GPU box:  python profiles/kernel_sweep.py"""
import json
import os
import sys

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imitation_b200 import _desc, _lib  # noqa: E402


def ev_ms(fn, reps):
    a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps):
        fn(i)
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / reps


def main():
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    out = []
    for name, d, bytes_row, flop_row in (
            ("gail_23_32x32", _desc.disc_desc(17, 6), 96, 9280),
            ("gail_23_32x32_norm", _desc.disc_desc(17, 6, normalize_input=True), 96, 9280),
            ("airl_shaped", _desc.disc_desc(17, 6, hid_sizes=(32,), shaped=True, subtract_logp=True), 169, 20160),
            ("cartpole_6_64x64", _desc.disc_desc(4, 2, hid_sizes=(64, 64)), 28, 26496)):
        P = (th.rand(d.n_params, device="cuda") - 0.5) * 0.5
        NS = th.ones(4 * 64, device="cuda")
        ws = th.zeros(_lib.disc_workspace_floats(d), device="cuda")
        bw = _desc.batch_rows(d.d_obs, d.d_act)
        for lg in (14, 16, 18, 20, 22):
            n = 1 << lg
            ld = _desc.batch_ld(n)
            nbuf = max(1, min(4, (1 << 29) // (bw * ld * 4)))
            bufs = [th.randn(bw, ld, device="cuda") for _ in range(nbuf)]
            logits = th.empty(n, device="cuda")
            f = lambda i: _lib.disc_fwd_bwd(d, P, NS, bufs[i % nbuf], ld, n, n // 2, 1.0 / n, None, logits,
                                            _lib.IMB_F_ZERO_GRAD, ws)
            ev_ms(f, 3)
            ms = ev_ms(f, 10 if lg < 22 else 5)
            gbs = n * bytes_row / (ms / 1e3) / 1e9
            out.append(dict(kernel="k_disc_fwdbwd", net=name, rows=n, ms=ms, rows_per_s=n / (ms / 1e3), GBps=gbs,
                            hbm_frac=gbs / peak, tflops=n * flop_row / (ms / 1e3) / 1e12, rotating_buffers=nbuf))
            print(json.dumps(out[-1]), flush=True)
            del bufs
    json.dump(out, open(os.path.join(os.path.dirname(__file__), "..", "gpurun_out", "kernel_sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    os.makedirs(os.path.join(os.path.dirname(__file__), "..", "gpurun_out"), exist_ok=True)
    main()
