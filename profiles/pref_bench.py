"""bench.py --config pref: reward-model training of preference comparisons (BASELINE config 5: Hopper-shaped obs 11 /
act 3, fragment length 100, 2048 fragment pairs, 5-member RewardEnsemble; reference: algorithms/preference_comparisons.py
:441-454 per-fragment loop, :1417-1424 per-member loop, scripts/config/train_preference_comparisons.py:33-44).

A "step" = one epoch of `EnsembleTrainer` over the 2048 pairs for all 5 members (forward through the fused reward kernels,
Boltzmann preference probabilities, cross-entropy, backward, AdamW).  Unit: fragment-pair evaluations / s (pairs x members
per epoch / time).  The reference arm runs the CPU restatement fragment by fragment on a bounded sample of pairs."""
import json
import os
import sys
import time

import numpy as np
import torch as th

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

Do, Da, L, P, M, MB = 11, 3, 100, 2048, 5, 256


def _fragments(rng, n_pairs):
    from imitation_b200.data import types

    frags = []
    for _ in range(n_pairs):
        pair = []
        for _ in range(2):
            pair.append(types.TrajectoryWithRew(obs=rng.standard_normal((L + 1, Do)).astype(np.float32),
                                                acts=rng.uniform(-1, 1, (L, Da)).astype(np.float32), infos=None,
                                                terminal=False, rews=rng.standard_normal(L).astype(np.float32)))
        frags.append(tuple(pair))
    return frags


def main(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": "preference comparisons reward training, Hopper-shaped (obs11/act3), 2048 fragment pairs x length "
                          "100, 5-member ensemble of BasicRewardNet 32x32, minibatch 256 pairs", "name": "pref",
              "pairs": P, "fragment_length": L, "members": M,
              "parallelism": ("one GPU: the members are trained one after the other" if world == 1 else
                              f"members over {world} GPUs (member k on rank k % {world}: {-(-M // world)} members on the busiest "
                              "rank; owners broadcast parameters + AdamW state after every training call; result "
                              "bit-identical to the single-GPU run)")}
    rng = np.random.default_rng(0)
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import nets_port, pref_port

        th.set_num_threads(min(8, os.cpu_count() or 1))
        n = 64
        frags = _fragments(rng, n)
        prefs = (rng.random(n) < 0.5).astype(np.float32)
        net = nets_port.BasicRewardNetPort(Do, Da, hid_sizes=(32, 32))
        sample = [tuple(dict(obs=f.obs, acts=f.acts, rews=f.rews, terminal=f.terminal) for f in pr) for pr in frags]
        steps = max(1, min(args.steps, 20))
        t0 = time.perf_counter()
        for _ in range(steps):
            probs, gt = pref_port.preference_probs_port(net, sample)
            loss, _, _ = pref_port.cross_entropy_loss_port(probs, gt, prefs)
            net.zero_grad()
            loss.backward()
        dt = (time.perf_counter() - t0) / steps
        v = n / dt
        print(json.dumps({"impl": "reference", "metric": "preference reward-model training, fragment-pair evaluations/sec",
                          "value": v, "unit": "pair-evaluations/s", "n_gpus": args.gpus, "steps": steps, "warmup": 0,
                          "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f32", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": "pair-evaluations/s", "cores": th.get_num_threads(), "kind": "port",
                                           "sample": f"{steps} x (forward + loss + backward) over {n} pairs, one member, "
                                                     "oracle/pref_port.py (the reference's fragment-by-fragment loop)"},
                          "e2e": {"value": v, "unit": "pair-evaluations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    local = int(os.environ.get("LOCAL_RANK", "0"))
    th.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=th.device("cuda", local))
    from imitation_b200 import _lib, spaces
    from imitation_b200.algorithms import preference_comparisons as pc
    from imitation_b200.rewards import reward_nets

    frags = _fragments(rng, P)
    prefs = (rng.random(P) < 0.5).astype(np.float32)
    obs_space, act_space = spaces.Box(-np.inf, np.inf, (Do,)), spaces.Box(-1.0, 1.0, (Da,))
    members = [reward_nets.BasicRewardNet(obs_space, act_space, hid_sizes=(32, 32)).cuda() for _ in range(M)]
    ens = reward_nets.RewardEnsemble(obs_space, act_space, members)
    dataset = pc.PreferenceDataset()
    dataset.push(frags, prefs)
    trainer = pc.EnsembleTrainer(pc.PreferenceModel(ens), pc.CrossEntropyRewardLoss(), rng=np.random.default_rng(1),
                                 batch_size=MB, epochs=1, lr=1e-3)
    if world > 1:
        trainer.set_distributed()
    if os.environ.get("IMB_PREF_AUTOGRAD") == "1":  # the per-minibatch autograd + torch AdamW path, for comparison
        for t in trainer.member_trainers:
            t.use_fused_step = False

    def epoch():
        trainer.train(dataset)  # one epoch of every member on its own bagging subset (reads back the epoch's statistics)
        return trainer.last_epoch_stats["loss"]

    W, K = max(3, args.warmup), max(1, min(args.steps, 10))
    for _ in range(W):
        epoch()
    th.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = _lib.LAUNCHES["count"]
    a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(K):
        epoch()
    b.record()
    b.synchronize()
    ms = a.elapsed_time(b)
    if world > 1:  # max over ranks (device time)
        t = th.tensor([ms], device="cuda", dtype=th.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    launches = _lib.LAUNCHES["count"] - l0
    v = K * P * M / (ms / 1e3)  # whole job: the ONE ensemble all ranks train together
    rows = K * M * 2 * P * L
    config["fused_step"] = all("_fused_opt" in t.__dict__ for t in trainer.member_trainers)
    ach = rows * (2 * 4 * (Do + Da) + 8) / (ms / 1e3) / 1e9
    peak, peak_src = getattr(args, "hbm_peak", None), getattr(args, "hbm_peak_source", None)
    h2d = M * P * 8  # per epoch and member: the minibatches' item indices (int64); the pool uploads happened in warm-up
    if rank == 0:
        print(json.dumps({"metric": "preference reward-model training, fragment-pair evaluations/sec", "value": v,
                          "unit": "pair-evaluations/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
                          "higher_is_better": True, "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": config,
                          "e2e": {"value": v, "unit": "pair-evaluations/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": M * 32,
                                  "path": "EnsembleTrainer.train(PreferenceDataset of host TrajectoryWithRew pairs): bagging "
                                          "subsets + DataLoader index stream on the host, every minibatch as device work "
                                          "(gather from the fragment pool -> imb_reward_forward -> imb_pref_loss -> "
                                          "imb_disc_fwd_bwd -> reduce + AdamW), one read-back of the epoch's statistics per "
                                          "member; every call starts from the host dataset and ends with host floats, so "
                                          "value == e2e for this row (the fragments' transitions are uploaded on their first "
                                          "use; the warm-up epochs contain those uploads)"},
                          "gpu_launches": launches,
                          "roofline": {"kernel": "k_disc_fwdbwd / k_reward_fwd over 2 * 256 * 100 = 51 200 transition rows per "
                                                 "minibatch", "bound": "hbm",
                                       "achieved": ach, "peak": peak, "unit": "GB/s",
                                       "frac": (ach / peak) if peak else None, "traffic": None, "peak_source": peak_src,
                                       "algorithmic_bytes": "per transition row: 4 (Do + Da) read by the forward + the same "
                                                            "again by the fused forward/backward + 4 reward + 4 gradient "
                                                            "= 2 x 56 + 8 = 120 B at 11/3",
                                       "note": "fragments are device-resident after their first use (FragmentPool) and a minibatch "
                                               "is ~10 launches without a host round trip; what remains is the host's launch "
                                               "rate (the members are trained one after the other like the reference's); "
                                               "transition rows/s = "
                                               f"{rows / (ms / 1e3) / 1e6:.1f} M"},
                          "cpu_baseline": None}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    main(ap.parse_args())
