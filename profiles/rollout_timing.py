"""ms per collect_rollouts() (k_rollout + k_gae + k_rollout_advance) at the bench configurations, CUDA events over 20
back-to-back calls.  IMB_VARIANT selects the library build.  (rollout_calls_probe_r02.txt: a build whose seven tiled layers
were ONE out-of-line routine per activation -- 11.5 k -> 9.8 k SASS instructions -- ran the launch in 72.3 instead of
74.4 us: instruction fetch is not what binds it; the experiment was reverted.)"""
import os
import sys

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

for name in sys.argv[1:] or ["hc", "cartpole", "airl_hc", "ant"]:
    cfg = bench.CONFIGS[name]
    tr, _ = bench.build_trainer(cfg, 0, 1, th.device("cuda", 0))
    gen = tr.gen_algo
    for _ in range(3):
        gen.collect_rollouts()
        tr.venv_buffering.discard()
    th.cuda.synchronize()
    reps = 20
    a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        gen.collect_rollouts()
    b.record()
    b.synchronize()
    tr.venv_buffering.discard()
    ms = a.elapsed_time(b) / reps
    E, T = cfg["envs_per_gpu"], cfg["ppo_batch"] // cfg["envs_per_gpu"]
    print(f"{os.environ.get('IMB_VARIANT', '') or 'default':<8s} {name:<9s} {E:5d} envs x {T:3d} steps  {ms * 1e3:8.1f} us / rollout  "
          f"{ms * 1e3 / T:7.2f} us / step")
