"""TEST INFRASTRUCTURE.  Cut small demonstration fixtures in the reference's legacy `.npz` layout (data/serialize.py:50-65:
concatenated obs / acts / infos / rews split at `indices`, one extra observation per trajectory, `terminal` per trajectory)
out of the reference's own on-disk expert rollouts, so that the ingest tests can run where /root/reference does not exist:

    tests/testdata/expert_models/cartpole_0/rollouts/final.npz  -> tests/golden/demo_cartpole_legacy.npz  (first 4 trajectories)
    tests/testdata/expert_models/pendulum_0/rollouts/final.npz  -> tests/golden/demo_pendulum_legacy.npz  (first 3 trajectories)

Run in the build container:  python oracle/make_demo_fixture.py
"""
import os

import numpy as np

REF = "/root/reference/tests/testdata/expert_models"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def cut(src, dst, k):
    z = np.load(src, allow_pickle=True)
    idx = np.asarray(z["indices"])
    n_act = int(idx[k - 1]) if k <= len(idx) else len(z["acts"])
    n_obs = n_act + k
    out = dict(obs=z["obs"][:n_obs], acts=z["acts"][:n_act], infos=z["infos"][:n_act], terminal=z["terminal"][:k],
               indices=idx[:k - 1])
    if "rews" in z.files:
        out["rews"] = z["rews"][:n_act]
    with open(dst, "wb") as f:
        np.savez_compressed(f, **out)
    print("wrote", dst, {a: out[a].shape for a in out})


if __name__ == "__main__":
    cut(os.path.join(REF, "cartpole_0/rollouts/final.npz"), os.path.join(OUT, "demo_cartpole_legacy.npz"), 4)
    cut(os.path.join(REF, "pendulum_0/rollouts/final.npz"), os.path.join(OUT, "demo_pendulum_legacy.npz"), 3)
