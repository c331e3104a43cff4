"""rows/s of the fused discriminator fwd/BCE/bwd kernel: tcgen05 path vs the fp32-FFMA path of round 1, rows-per-launch
sweep (2^14 .. 2^22), three rotating input buffers so the inputs never sit in L2 (2^20 rows x 23 features x 4 B
= 96 MB per buffer).  Usage: python profiles/disc_tc_sweep.py [d_obs d_act]"""
import json
import os
import sys

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imitation_b200 import _desc, _lib as L  # noqa: E402

PEAK = 6574.5
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def main():
    Do, Da = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (17, 6)
    d = _desc.disc_desc(Do, Da)
    din = Do + Da
    bpr = 4 * din + 4
    flop = 2 * (3 * (din * 32 + 32 * 32 + 32)) - 2 * din * 32
    g = th.Generator(device="cuda").manual_seed(0)
    P = (th.rand(d.n_params, device="cuda", generator=g) - 0.5) * 0.6
    NS = th.zeros(2, device="cuda")
    ws = th.zeros(L.disc_workspace_floats(d), device="cuda")
    res = []
    for lg2 in (14, 16, 18, 20, 21, 22):
        n = 1 << lg2
        ld = _desc.batch_ld(n)
        nb = 3 if n * din * 4 * 3 < 6e9 else 2
        bufs = [th.randn(din, ld, device="cuda") for _ in range(nb)]  # only the rows the kernel reads
        # the descriptor addresses feature rows 0..din-1 of the batch (obs | act): a [din][ld] buffer is enough
        logits = th.empty(n, device="cuda")
        row = {"rows": n}
        for name, fl in (("tc", 0), ("ffma", L.IMB_F_NO_TENSOR)):
            def run(i):
                L.disc_fwd_bwd(d, P, NS, bufs[i % nb], ld, n, n // 2, 1.0 / n, None, logits, L.IMB_F_ZERO_GRAD | fl, ws)
            for i in range(3):
                run(i)
            th.cuda.synchronize()
            reps = 12 if lg2 <= 20 else 6
            a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
            a.record()
            for i in range(reps):
                run(i)
            b.record()
            b.synchronize()
            ms = a.elapsed_time(b) / reps
            gbs = n * bpr / (ms / 1e3) / 1e9
            row[name] = {"ms": round(ms, 4), "Grows_s": round(n / (ms / 1e3) / 1e9, 3), "GB_s": round(gbs, 1),
                         "frac_hbm": round(gbs / PEAK, 4), "fp32_equiv_tflops": round(n * flop / (ms / 1e3) / 1e12, 2)}
        row["speedup"] = round(row["ffma"]["ms"] / row["tc"]["ms"], 2)
        res.append(row)
        print(json.dumps(row), flush=True)
        del bufs
    out = {"d_obs": Do, "d_act": Da, "bytes_per_row": bpr, "flop_per_row": flop, "peak_gbs": PEAK, "sweep": res}
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/disc_tc_sweep_{Do}_{Da}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
