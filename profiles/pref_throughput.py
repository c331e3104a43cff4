"""Reward-model training throughput of the preference-comparison row (BASELINE config 5 shape: Hopper-sized
obs 11 / act 3, fragment length 100, 2048 fragment pairs, 5-member ensemble): one pass (forward + BCE + backward,
no optimiser step) over all pairs per member, fused path vs the fragment-by-fragment CPU restatement (bounded sample).
"""
import os
import sys
import time

import numpy as np
import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imitation_b200 import spaces  # noqa: E402
from imitation_b200.algorithms import preference_comparisons as pc  # noqa: E402
from imitation_b200.data import types  # noqa: E402
from imitation_b200.rewards import reward_nets  # noqa: E402

Do, Da, L, P, M, MB = 11, 3, 100, 2048, 5, 256
rng = np.random.default_rng(0)
frags = []
for _ in range(P):
    pair = []
    for _ in range(2):
        pair.append(types.TrajectoryWithRew(obs=rng.standard_normal((L + 1, Do)).astype(np.float32),
                                            acts=rng.uniform(-1, 1, (L, Da)).astype(np.float32), infos=None,
                                            terminal=False, rews=rng.standard_normal(L).astype(np.float32)))
    frags.append(tuple(pair))
prefs = (rng.random(P) < 0.5).astype(np.float32)
obs_space, act_space = spaces.Box(-np.inf, np.inf, (Do,)), spaces.Box(-1.0, 1.0, (Da,))
members = [reward_nets.BasicRewardNet(obs_space, act_space, hid_sizes=(32, 32)).cuda() for _ in range(M)]
loss_fn = pc.CrossEntropyRewardLoss()
pms = [pc.PreferenceModel(m) for m in members]


def one_pass():
    for pm in pms:
        for s in range(0, P, MB):
            out = loss_fn(frags[s:s + MB], prefs[s:s + MB], pm)
            out.loss.backward()


one_pass()
th.cuda.synchronize()
t0 = time.perf_counter()
reps = 3
for _ in range(reps):
    one_pass()
th.cuda.synchronize()
dt = (time.perf_counter() - t0) / reps
rows = M * 2 * P * L
print(f"fused path: {dt * 1e3:.1f} ms per pass over {P} pairs x {M} members ({rows / dt / 1e6:.1f} M transition rows/s, "
      f"{M * P / dt:.0f} pair-evaluations/s; includes the host-side stacking and H2D copy of every minibatch)")

# CPU restatement, fragment by fragment like the reference (bounded sample: 64 pairs, one member) -- a bench-style
# baseline leg, run only on request: the oracle is test infrastructure, nothing in the product path imports it
if "--cpu-baseline" not in sys.argv:
    sys.exit(0)
from oracle import nets_port, pref_port  # noqa: E402

th.set_num_threads(8)
net = nets_port.BasicRewardNetPort(Do, Da, hid_sizes=(32, 32))
sample = [tuple(dict(obs=f.obs, acts=f.acts, rews=f.rews, terminal=f.terminal) for f in pr) for pr in frags[:64]]
t0 = time.perf_counter()
probs, gt = pref_port.preference_probs_port(net, sample)
loss, _, _ = pref_port.cross_entropy_loss_port(probs, gt, prefs[:64])
loss.backward()
dtc = time.perf_counter() - t0
print(f"CPU restatement (8 threads): {64 / dtc:.0f} pair-evaluations/s on a 64-pair sample "
      f"({2 * 64 * L / dtc / 1e6:.3f} M rows/s) -> ratio {M * P / dt / (64 / dtc):.0f}x")
