"""A few rollouts at the headline configuration (1024 envs x 4 steps, GAIL relabel) for
`ncu --set full --import-source on -k regex:k_rollout -s 2 -c 1`."""
import os
import sys

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "hc"]
tr, _ = bench.build_trainer(cfg, 0, 1, th.device("cuda", 0))
for _ in range(4):
    tr.gen_algo.collect_rollouts()
    tr.venv_buffering.discard()
th.cuda.synchronize()
print("ok")
