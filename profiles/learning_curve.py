"""Does the round learn, and does the multi-GPU scheme (global-batch discriminator + generator averaged once per round)
learn like one GPU on the same number of environments?

GAIL on the synthetic HalfCheetah-shaped env with the tuned hyper-parameters (bench.py config `hc`); the measure is the
ground-truth per-step environment reward of the generator's own rollouts (the env's `w . obs' - 0.1 |a|^2`, which the
learner never sees), averaged over all envs and steps of a round, next to the expert's and a random policy's.

    python profiles/learning_curve.py [--rounds 600]                                  # 1 GPU, 2048 envs
    python -m torch.distributed.run --nproc-per-node 2 ... profiles/learning_curve.py # 2 GPUs x 1024 envs

Writes gpurun_out/learning_curve_n{world}.json."""
import argparse
import json
import os
import sys

import torch as th

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=600)
    ap.add_argument("--total-envs", type=int, default=2048)
    ap.add_argument("--every", type=int, default=20)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    th.cuda.set_device(local)
    device = th.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=device)
    from imitation_b200 import _desc, distributed

    cfg = dict(bench.CONFIGS["hc"])
    cfg["envs_per_gpu"] = args.total_envs // world
    cfg["ppo_batch"] = 4 * cfg["envs_per_gpu"]          # n_steps = 4 like the tuned 4096 / 1024
    tr, expert = bench.build_trainer(cfg, rank, world, device)
    sync = None
    if world > 1:
        tr.set_distributed()
        sync = distributed.trainer_round_sync(tr)
        sync.broadcast_initial(0)
    # reference points: the expert's per-step reward, from its own demonstrations
    ep = _desc.synth_env_params(cfg["d_obs"], cfg["d_act"], cfg["seed"])
    Do, Da = cfg["d_obs"], cfg["d_act"]
    w = ep[Do * Do + Do * Da + Do:Do * Do + Do * Da + 2 * Do]
    expert_rew = float((expert["next_obs"] @ w - 0.1 * (expert["acts"] ** 2).sum(1)).mean())
    curve = []
    acc, n_acc = 0.0, 0
    for r in range(args.rounds):
        if sync:
            sync.begin_round()
        tr.train_gen()
        rew = tr.gen_algo._aux[2 * tr.venv.num_envs + tr.venv.num_envs * tr.gen_algo.n_steps:]  # ground-truth env rewards
        m = rew.mean()
        tr.disc_train_mode = True
        for _ in range(tr.n_disc_updates_per_round):
            tr.train_disc_async(check_ring=False)
        tr.disc_train_mode = False
        tr.join()
        if sync:
            sync.end_round()
            import torch.distributed as dist

            dist.all_reduce(m)
            m = m / world
        acc += float(m)
        n_acc += 1
        if (r + 1) % args.every == 0:
            curve.append((r + 1, acc / n_acc))
            acc, n_acc = 0.0, 0
    if rank == 0:
        out = {"world": world, "envs_total": args.total_envs, "rounds": args.rounds,
               "env_steps": args.rounds * args.total_envs * 4, "expert_reward_per_step": expert_rew,
               "first": curve[0][1], "last": curve[-1][1], "curve": curve}
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"learning_curve_n{world}.json"), "w"))
        print(json.dumps({k: out[k] for k in ("world", "expert_reward_per_step", "first", "last")}),
              "curve", [round(c[1], 4) for c in curve[::max(1, len(curve) // 12)]], flush=True)
    if world > 1:
        sys.stdout.flush()
        th.cuda.synchronize()
        os._exit(0)


if __name__ == "__main__":
    main()
