"""Worker of tests/test_multi_gpu.py: one process per GPU under torch.distributed.run (NCCL).

Preference comparisons, ensemble members over GPUs (`EnsembleTrainer.set_distributed`): every rank trains member k iff
k % W == rank, then the owners broadcast; the result on EVERY rank must be bit-identical to the single-process training of
all members, which every rank also runs locally as its own reference (same seeds, same dataset)."""
import os
import sys

import numpy as np
import torch as th
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(distributed: bool):
    from imitation_b200 import spaces
    from imitation_b200.algorithms import preference_comparisons as pc
    from imitation_b200.data import types
    from imitation_b200.rewards import reward_nets
    from imitation_b200.util import networks

    Do, Da, L, P, M = 11, 3, 20, 96, 3
    rng = np.random.default_rng(0)
    obs_space, act_space = spaces.Box(-np.inf, np.inf, (Do,)), spaces.Box(-1.0, 1.0, (Da,))

    def frag():
        return types.TrajectoryWithRew(obs=rng.standard_normal((L + 1, Do)).astype(np.float32),
                                       acts=rng.uniform(-1, 1, (L, Da)).astype(np.float32), infos=None, terminal=False,
                                       rews=rng.standard_normal(L).astype(np.float32))

    ds = pc.PreferenceDataset()
    ds.push([(frag(), frag()) for _ in range(P)], (rng.random(P) < 0.5).astype(np.float32))
    th.manual_seed(3)
    members = [reward_nets.BasicRewardNet(obs_space, act_space, hid_sizes=(32, 32),
                                          normalize_input_layer=networks.RunningNorm).cuda() for _ in range(M)]
    ens = reward_nets.RewardEnsemble(obs_space, act_space, members)
    et = pc.EnsembleTrainer(pc.PreferenceModel(ens), pc.CrossEntropyRewardLoss(), rng=np.random.default_rng(1),
                            batch_size=32, minibatch_size=16, epochs=2, lr=1e-3)
    if distributed:
        et.set_distributed()
    th.manual_seed(11)
    et.train(ds)
    et.train(ds, epoch_multiplier=1.5)
    th.cuda.synchronize()
    state = []
    for m, t in zip(members, et.member_trainers):
        e = m.engine()
        state.append(th.cat([e.params, e.norm_state, e.norm_count.float(), t._fused_opt["m"], t._fused_opt["v"],
                             th.tensor([float(t.optim.state[e._param_list()[0]]["step"])], device="cuda")]).clone())
    return state, dict(et.last_epoch_stats), float(th.rand(1))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    th.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=th.device("cuda", local))
    want, want_stats, want_probe = run(distributed=False)   # every rank: all members, single-process semantics
    got, got_stats, got_probe = run(distributed=True)       # member k on rank k % W, then broadcasts
    assert len({float(w.sum()) for w in want}) == len(want), "the members did not train differently"
    for k, (a, b) in enumerate(zip(got, want)):
        assert th.equal(a, b), f"rank {rank}: member {k} differs from the single-process run (max |d| = {(a - b).abs().max()})"
    assert got_probe == want_probe, "torch's global RNG ended in a different state"
    for key in want_stats:
        assert abs(got_stats[key] - want_stats[key]) < 1e-6, (key, got_stats, want_stats)
    # and the replicas agree with each other
    x = th.cat(got)
    all_x = [th.empty_like(x) for _ in range(world)]
    dist.all_gather(all_x, x)
    assert all(th.equal(all_x[0], y) for y in all_x[1:])
    dist.barrier()
    if rank == 0:
        print("DIST_PREF_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
