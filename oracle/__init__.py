"""oracle/ -- TEST INFRASTRUCTURE ONLY (never imported by `imitation_b200`).

CPU restatement (torch-CPU eager + NumPy, plain Python loops where the reference has
them) of the one hot path this repo accelerates: the GAIL/AIRL round of
HumanCompatibleAI/imitation -- generator rollouts -> replay-buffer sampling ->
discriminator update.  Every function cites the reference
file:line it follows.

Who may import this package: `tests/`, `__graft_entry__.smoke()`, and `bench.py`'s
`cpu_baseline` / `--impl reference` legs.  The product path (`imitation_b200/`) must
never route through it; it fails loudly if the CUDA extension is missing.

Pinning status
--------------
* discriminator side, RunningNorm, reward nets, buffers, wrappers, rollout flattening:
  PINNED -- `oracle/make_golden.py` runs the reference's own modules (imported from
  /root/reference/src through the names-only shim in `oracle/_shim/`) on seeded inputs
  and stores input/output vectors in `tests/golden/*.npz`; `tests/test_oracle_golden.py`
  checks this restatement against them, and `tests/test_oracle_vs_reference.py`
  re-checks live against the reference whenever /root/reference is present.
* generator (PPO): the arithmetic lives in stable-baselines3 ~=2.2.1, which is NOT
  under /root/reference and not installed -> `oracle/ppo_port.py` restates its published
  algorithm; PARITY UNPINNED for that part (no reference test pins PPO numerics either,
  SURVEY.md section 8c).
* synthetic MuJoCo-shaped env + Philox streams: defined by this repo (seals/MuJoCo are
  not available), implemented twice (NumPy here, CUDA in the product).
"""
