"""Worker of tests/test_multi_gpu.py: one process per GPU under torch.distributed.run (NCCL).

Checks: (1) a distributed discriminator step on per-rank batches equals the single-GPU step on the concatenated
(global) batch, statistics included; (2) after every discriminator step the replicas are BIT-identical (parameters, Adam
moments, RunningNorm statistics, step counter); (3) whole rounds with device-side sampling -- eager and as one captured CUDA
graph containing the NCCL collectives -- keep discriminator and (after the per-round generator sync) policy replicas
bit-identical, and every rank's ring holds its own env slice."""
import os
import sys

import numpy as np
import torch as th
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build(rank, world, device, B, E=16, T=8, seed=0, distributed=True):
    from imitation_b200.algorithms import ppo
    from imitation_b200.algorithms.adversarial import gail
    from imitation_b200.envs import synth
    from imitation_b200.rewards import reward_nets
    from imitation_b200.util import networks

    Do, Da = 17, 6
    th.manual_seed(seed)  # identical initial networks on every rank
    venv = synth.DeviceVecEnv(Do, Da, E, horizon=1000, seed=seed, env_id_offset=rank * E, device=device)
    gen = ppo.DevicePPO("FeedForward32Policy", venv, n_steps=T, batch_size=32, n_epochs=2, seed=seed + rank,
                        policy_kwargs=dict(normalize_features=True), device=device)
    net = reward_nets.BasicRewardNet(venv.observation_space, venv.action_space, normalize_input_layer=networks.RunningNorm)
    rng = np.random.default_rng(seed)
    n = 8 * B
    demos = dict(obs=rng.standard_normal((n, Do)).astype(np.float32), acts=rng.uniform(-1, 1, (n, Da)).astype(np.float32),
                 next_obs=rng.standard_normal((n, Do)).astype(np.float32), dones=np.zeros(n, bool))
    tr = gail.GAIL(demonstrations=demos, demo_batch_size=B, venv=venv, gen_algo=gen, reward_net=net,
                   n_disc_updates_per_round=3, gen_replay_buffer_capacity=96, sampling="device", seed=seed + 17 * rank)
    if distributed:
        tr.set_distributed()
    return tr


def replica_state(tr):
    eng, opt = tr._fused_net.engine(), tr._disc_opt
    return th.cat([eng.params, opt.exp_avg, opt.exp_avg_sq, eng.norm_state, eng.norm_count.float(),
                   tr.venv.state[9:10].float()])


def assert_replicas_identical(x, what, world):
    got = [th.empty_like(x) for _ in range(world)]
    dist.all_gather(got, x.contiguous())
    for r in range(1, world):
        assert th.equal(got[0], got[r]), f"{what}: rank {r} differs from rank 0 (max |d| = {(got[0] - got[r]).abs().max()})"


def samples(step, rank, B):
    rng = np.random.default_rng(1000 * step + rank)
    return (dict(obs=rng.standard_normal((B, 17)).astype(np.float32), acts=rng.uniform(-1, 1, (B, 6)).astype(np.float32),
                 next_obs=rng.standard_normal((B, 17)).astype(np.float32), dones=np.zeros(B, bool)),
            dict(obs=(rng.standard_normal((B, 17)) + 0.5).astype(np.float32), acts=rng.uniform(-1, 1, (B, 6)).astype(np.float32),
                 next_obs=rng.standard_normal((B, 17)).astype(np.float32), dones=np.zeros(B, bool)))


def main():
    from imitation_b200 import distributed
    from imitation_b200.util import networks

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    th.cuda.set_device(local)
    device = th.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    B = 64
    # ---- (1) + (2): explicit batches ------------------------------------------------------------------------------
    tr = build(rank, world, device, B)
    init = {k: v.clone() for k, v in tr._reward_net.state_dict().items()}
    stats = []
    for step in range(3):
        ex, ge = samples(step, rank, B)
        with networks.training(tr.reward_train):
            stats.append(tr.train_disc(expert_samples=ex, gen_samples=ge))
        tr.join()
        assert_replicas_identical(replica_state(tr), f"disc step {step}", world)
    if rank == 0:
        ref = build(0, 1, device, world * B, distributed=False)
        ref.reproduce_evaluate_actions_side_effect = False
        ref._reward_net.load_state_dict(init)
        for step in range(3):
            parts = [samples(step, r, B) for r in range(world)]
            ex = {k: np.concatenate([p[0][k] for p in parts]) for k in parts[0][0]}
            ge = {k: np.concatenate([p[1][k] for p in parts]) for k in parts[0][1]}
            with networks.training(ref.reward_train):
                want = ref.train_disc(expert_samples=ex, gen_samples=ge)
            for k in want:
                np.testing.assert_allclose(stats[step][k], want[k], rtol=2e-5, atol=1e-6, err_msg=f"step {step} {k}")
        a, b = tr._fused_net.engine(), ref._fused_net.engine()
        np.testing.assert_allclose(a.params.cpu().numpy(), b.params.cpu().numpy(), rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(a.norm_state.cpu().numpy(), b.norm_state.cpu().numpy(), rtol=1e-5, atol=1e-6)
        assert int(a.norm_count[0]) == int(b.norm_count[0]) == 3 * 2 * world * B
        del ref
    dist.barrier()
    # ---- (3): whole rounds, device sampling, eager then one captured graph per round -------------------------------------
    tr = build(rank, world, device, B, seed=1)
    sync = distributed.trainer_round_sync(tr)
    sync.broadcast_initial(0)

    def check(tag):
        th.cuda.synchronize()
        assert_replicas_identical(replica_state(tr), f"{tag}: discriminator", world)
        pp, pn, pc = tr.policy.flat_vectors()
        assert_replicas_identical(th.cat([pp, tr.gen_algo.exp_avg, tr.gen_algo.exp_avg_sq, pn, pc.float()]),
                                  f"{tag}: generator", world)

    for r in range(3):
        sync.begin_round()
        tr.train_gen()
        tr.disc_train_mode = True
        for _ in range(tr.n_disc_updates_per_round):
            tr.train_disc_async(check_ring=False)
        tr.disc_train_mode = False
        tr.join()
        sync.end_round()
        check(f"eager round {r}")
    graph_ok = True
    try:
        tr.capture_round()
    except Exception as e:  # (NCCL inside a captured graph needs a recent torch / NCCL pair)
        graph_ok = False
        sys.stderr.write(f"rank {rank}: capture_round with collectives failed: {type(e).__name__}: {e}\n")
    flags = th.tensor([int(graph_ok)], device=device)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if int(flags):
        for r in range(3):
            sync.begin_round()
            tr.replay_round()
            sync.end_round()
            check(f"graph round {r}")
    # every rank's ring holds transitions of its OWN env slice: the first observation column differs across ranks
    ring = tr._gen_replay_buffer.table[:, :17].contiguous()
    got = [th.empty_like(ring) for _ in range(world)]
    dist.all_gather(got, ring)
    assert not th.equal(got[0], got[1]), "ranks hold the same generator samples (env slices not sharded)"
    dist.barrier()
    th.cuda.synchronize()
    if rank == 0:
        print(f"DIST_OK world={world} graph_with_collectives={bool(int(flags))}", flush=True)
    # (a process group whose collectives were captured in CUDA graphs can block in destroy_process_group while the
    #  graph objects are still alive: leave without the teardown)
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
