"""Preference comparisons: host-side fragmenting / synthetic preferences bit-exact against the
reference's own classes (tests/golden/preference.npz, oracle/make_golden.py), the CPU restatement of the preference
model against the same goldens, and -- on the GPU -- the batched fused-kernel implementation against both."""
import numpy as np
import pytest
import torch as th

from tests import golden_util as G


def _golden():
    z = G.load("preference")
    Do, Da, L, P = [int(v) for v in z["cfg"]]
    trajs = []
    i = 0
    while f"traj{i}/obs" in z.files:
        trajs.append(dict(obs=z[f"traj{i}/obs"], acts=z[f"traj{i}/acts"], rews=z[f"traj{i}/rews"],
                          terminal=bool(z[f"traj{i}/terminal"])))
        i += 1
    pairs = [tuple(dict(obs=z[f"pair{k}/{s}/obs"], acts=z[f"pair{k}/{s}/acts"], rews=z[f"pair{k}/{s}/rews"],
                        terminal=bool(z[f"pair{k}/{s}/terminal"])) for s in ("a", "b")) for k in range(P)]
    return z, (Do, Da, L, P), trajs, pairs


def _as_traj(d):
    from imitation_b200.data import types

    return types.TrajectoryWithRew(obs=d["obs"], acts=d["acts"], infos=None, terminal=d["terminal"], rews=d["rews"])


def test_fragmenter_and_gatherer_bit_exact_with_reference():
    from imitation_b200.algorithms import preference_comparisons as pc

    z, (Do, Da, L, P), trajs, pairs = _golden()
    frags = pc.RandomFragmenter(rng=np.random.default_rng(3), warning_threshold=0)([_as_traj(t) for t in trajs], L, P)
    assert len(frags) == P
    for (a, b), (wa, wb) in zip(frags, pairs):
        for got, want in ((a, wa), (b, wb)):
            np.testing.assert_array_equal(got.obs, want["obs"])
            np.testing.assert_array_equal(got.acts, want["acts"])
            np.testing.assert_array_equal(got.rews, want["rews"])
            assert got.terminal == want["terminal"]
    np.testing.assert_allclose(pc.SyntheticGatherer(sample=False, temperature=0.5, discount_factor=0.9)(frags),
                               z["prefs_prob"], rtol=1e-6, atol=1e-7)
    np.testing.assert_array_equal(pc.SyntheticGatherer(sample=True, rng=np.random.default_rng(4))(frags),
                                  z["prefs_sampled"])
    np.testing.assert_array_equal(pc.SyntheticGatherer(sample=False, temperature=0)(frags), z["prefs_t0"])
    ds = pc.PreferenceDataset(max_size=5)
    ds.push(frags, z["prefs_sampled"].astype(np.float32))
    assert len(ds) == 5 and ds[0][0][0] is frags[P - 5][0]
    with pytest.raises(ValueError, match="float32"):
        ds.push(frags, z["prefs_sampled"].astype(np.float64))


def test_preference_model_port_matches_reference_golden():
    from oracle import nets_port, pref_port

    z, (Do, Da, L, P), _, pairs = _golden()
    net = nets_port.BasicRewardNetPort(Do, Da, hid_sizes=(32, 32))
    net.load_state_dict(G.state_to_torch(G.sub(z, "net")))
    probs, gt = pref_port.preference_probs_port(net, pairs, noise_prob=0.1, discount_factor=0.95)
    np.testing.assert_allclose(probs.detach().numpy(), z["probs"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gt.numpy(), z["gt_probs"], rtol=1e-6, atol=1e-7)
    loss, acc, gt_loss = pref_port.cross_entropy_loss_port(probs, gt, z["prefs_sampled"])
    np.testing.assert_allclose(loss.item(), z["loss"], rtol=1e-5)
    assert acc.item() == z["accuracy"]
    np.testing.assert_allclose(gt_loss.item(), z["gt_reward_loss"], rtol=1e-5)
    loss.backward()
    for k, p in net.named_parameters():
        np.testing.assert_allclose(p.grad.numpy(), z[f"grad/{k}"], rtol=1e-4, atol=1e-6, err_msg=k)


@pytest.mark.gpu
def test_preference_model_fused_matches_reference_golden():
    """One batch of 2 * P * L rows through the fused kernels == the reference's fragment-by-fragment loop:
    probabilities, loss, metrics and the gradients of every reward-network parameter."""
    from imitation_b200 import spaces
    from imitation_b200.algorithms import preference_comparisons as pc
    from imitation_b200.rewards import reward_nets

    z, (Do, Da, L, P), _, pairs = _golden()
    obs_space, act_space = spaces.Box(-np.inf, np.inf, (Do,)), spaces.Box(-1.0, 1.0, (Da,))
    net = reward_nets.BasicRewardNet(obs_space, act_space, hid_sizes=(32, 32)).cuda()
    net.load_state_dict({k: th.as_tensor(v) for k, v in G.sub(z, "net").items()})
    frags = [(_as_traj(a), _as_traj(b)) for a, b in pairs]
    pm = pc.PreferenceModel(net, noise_prob=0.1, discount_factor=0.95)
    prefs = z["prefs_sampled"].astype(np.float32)
    res = pc.CrossEntropyRewardLoss()(frags, prefs, pm)
    probs, gt = pm(frags)
    np.testing.assert_allclose(probs.detach().cpu().numpy(), z["probs"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gt.numpy(), z["gt_probs"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(res.loss.item(), z["loss"], rtol=1e-5)
    assert res.metrics["accuracy"].item() == z["accuracy"]
    np.testing.assert_allclose(res.metrics["gt_reward_loss"].item(), z["gt_reward_loss"], rtol=1e-5)
    net.zero_grad()
    res.loss.backward()
    for k, p in net.named_parameters():
        np.testing.assert_allclose(p.grad.cpu().numpy(), z[f"grad/{k}"], rtol=1e-4, atol=1e-6, err_msg=k)
    # ensemble: per-member probabilities [P, members] (reference: rewards() + probability() per pair)
    members = []
    for i in range(3):
        m = reward_nets.BasicRewardNet(obs_space, act_space, hid_sizes=(32, 32)).cuda()
        m.load_state_dict({k: th.as_tensor(v) for k, v in G.sub(z, f"member{i}").items()})
        members.append(m)
    ens = reward_nets.RewardEnsemble(obs_space, act_space, members)
    pe, _ = pc.PreferenceModel(ens, discount_factor=1.0)(frags)
    np.testing.assert_allclose(pe.cpu().numpy(), z["ensemble_probs"], rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_reward_trainer_matches_reference_after_two_epochs():
    """BasicRewardTrainer (AdamW, batch 8, minibatch 4, 2 epochs, DataLoader shuffling from torch's global RNG) ends at
    the reference's weights; EnsembleTrainer + PreferenceComparisons run and improve the training accuracy."""
    from imitation_b200 import spaces
    from imitation_b200.algorithms import preference_comparisons as pc
    from imitation_b200.rewards import reward_nets

    z, (Do, Da, L, P), trajs, pairs = _golden()
    obs_space, act_space = spaces.Box(-np.inf, np.inf, (Do,)), spaces.Box(-1.0, 1.0, (Da,))
    net = reward_nets.BasicRewardNet(obs_space, act_space, hid_sizes=(32, 32)).cuda()
    net.load_state_dict({k: th.as_tensor(v) for k, v in G.sub(z, "net").items()})
    frags = [(_as_traj(a), _as_traj(b)) for a, b in pairs]
    pm = pc.PreferenceModel(net, noise_prob=0.1, discount_factor=0.95)
    ds = pc.PreferenceDataset()
    ds.push(frags, z["prefs_sampled"].astype(np.float32))
    th.manual_seed(10)
    trainer = pc.BasicRewardTrainer(pm, pc.CrossEntropyRewardLoss(), rng=np.random.default_rng(5), batch_size=8,
                                    minibatch_size=4, epochs=2, lr=1e-3)
    trainer.train(ds)
    for k, v in net.state_dict().items():
        if k == "mlp.dense_final.bias":
            # the output bias cancels in r(fragment 2) - r(fragment 1): its gradient is pure rounding noise (~1e-9),
            # which Adam normalises into +-lr steps -- not reproducible between any two summation orders
            # (the CPU restatement differs from the reference on this one parameter in the same way)
            continue
        np.testing.assert_allclose(v.cpu().numpy(), z[f"net_trained/{k}"], rtol=2e-4, atol=2e-6, err_msg=k)
    # the outer loop on an ensemble, synthetic preferences from the ground-truth rewards
    members = [reward_nets.BasicRewardNet(obs_space, act_space, hid_sizes=(32, 32)).cuda() for _ in range(3)]
    ens = reward_nets.RewardEnsemble(obs_space, act_space, members)
    rng = np.random.default_rng(1)
    gen = pc.TrajectoryDataset([_as_traj(t) for t in trajs], rng)
    algo = pc.PreferenceComparisons(gen, ens, num_iterations=2, fragment_length=5, rng=rng,
                                    initial_epoch_multiplier=4.0,
                                    reward_trainer=pc.EnsembleTrainer(pc.PreferenceModel(ens), pc.CrossEntropyRewardLoss(),
                                                                     rng=rng, batch_size=16, epochs=3, lr=3e-3), allow_variable_horizon=True)
    out = algo.train(total_timesteps=0, total_comparisons=20)
    assert out["reward_accuracy"] is not None and 0.0 <= out["reward_accuracy"] <= 1.0 and np.isfinite(out["reward_loss"])
    assert len(algo.dataset) == 20


@pytest.mark.gpu
def test_agent_trainer_on_device_generator():
    """AgentTrainer (:127-316) over DevicePPO + DeviceVecEnv: `train` runs the generator on the learned reward (fused into
    the rollout kernel), `sample` returns finished trajectories carrying the ENVIRONMENT's rewards (checked against the env's
    closed form), unfinished episodes stay buffered across pops, and PreferenceComparisons runs an iteration on top."""
    from imitation_b200 import _desc
    from imitation_b200.algorithms import ppo
    from imitation_b200.algorithms import preference_comparisons as pc
    from imitation_b200.envs import synth
    from imitation_b200.rewards import reward_nets

    Do, Da, E, T, H = 5, 2, 4, 4, 6
    th.manual_seed(0)
    venv = synth.DeviceVecEnv(Do, Da, E, horizon=H, seed=3)
    net = reward_nets.BasicRewardNet(venv.observation_space, venv.action_space).cuda()
    algo = ppo.DevicePPO("FeedForward32Policy", venv, n_steps=T, batch_size=8, n_epochs=1, seed=0)
    rng = np.random.default_rng(0)
    agent = pc.AgentTrainer(algo, net, venv, rng)
    agent.train(steps=3 * E * T)  # 12 env steps per env = two finished episodes of 6
    trajs = agent.sample(2 * E * H)
    assert len(trajs) == 2 * E and all(t.terminal and len(t) == H for t in trajs)
    with pytest.raises(RuntimeError, match="transitions left in the buffer"):
        algo.collect_rollouts()
        agent.buffering_wrapper.after_rollout  # noqa: B018  (the rollout above left transitions behind)
        agent.train(steps=E * T)
    agent.buffering_wrapper.discard()
    # environment rewards: w . obs' - 0.1 |clip(a)|^2 of the synthetic env
    ep = _desc.synth_env_params(Do, Da, 3)
    w = ep[Do * Do + Do * Da + Do:Do * Do + Do * Da + 2 * Do]
    for t in trajs:
        want = t.obs[1:] @ w - 0.1 * (np.clip(t.acts, -1, 1) ** 2).sum(1)
        np.testing.assert_allclose(t.rews, want, rtol=1e-4, atol=1e-5)
    # asking for more than is buffered rolls the policy further (no training) until enough episodes finished
    more = agent.sample(3 * E * H)
    assert sum(len(t) for t in more) >= 3 * E * H and all(t.terminal for t in more)
    # the whole pipeline on top
    agent.buffering_wrapper.discard()
    frag = pc.RandomFragmenter(warning_threshold=0, rng=rng)
    pcs = pc.PreferenceComparisons(agent, net, num_iterations=1, fragmenter=frag, fragment_length=3,
                                   transition_oversampling=1, initial_comparison_frac=0.5, initial_epoch_multiplier=1.0, rng=rng)
    res = pcs.train(total_timesteps=2 * E * H, total_comparisons=8)
    assert np.isfinite(res["reward_loss"]) and 0.0 <= res["reward_accuracy"] <= 1.0


def test_index_dataloader_consumes_the_rng_like_the_fragment_dataloader():
    """The fused reward-trainer step draws its minibatches from DataLoader(range(n), shuffle=True): the same batches, in
    the same order, with the same consumption of torch's global RNG as the reference's DataLoader over the fragment pairs
    (algorithms/preference_comparisons.py:1207-1216)."""
    from torch.utils import data as data_th

    from imitation_b200.algorithms import preference_comparisons as pc

    z, (Do, Da, L, P), trajs, pairs = _golden()
    ds = pc.PreferenceDataset()
    ds.push([(_as_traj(a), _as_traj(b)) for a, b in pairs], np.arange(P, dtype=np.float32))
    th.manual_seed(21)
    want = [prefs.tolist() for _ in range(2)
            for _, prefs in data_th.DataLoader(ds, batch_size=3, shuffle=True, collate_fn=pc.preference_collate_fn)]
    after_want = th.rand(2)
    th.manual_seed(21)
    got = [ids.tolist() for _ in range(2) for ids in data_th.DataLoader(range(len(ds)), batch_size=3, shuffle=True)]
    after_got = th.rand(2)
    assert got == [[int(v) for v in b] for b in want]
    assert th.equal(after_want, after_got)
    # and the whole-epoch shortcut the trainer uses (one upload per epoch instead of a host collate per minibatch)
    th.manual_seed(21)
    perms = [pc._epoch_permutation(len(ds)).tolist() for _ in range(2)]
    after_perm = th.rand(2)
    assert pc._PERM_FAST is True, "torch's DataLoader draws its seeds differently now: update _epoch_permutation"
    assert perms == [[i for b in want[e * 4:(e + 1) * 4] for i in (int(v) for v in b)] for e in range(2)]
    assert th.equal(after_want, after_perm)


@pytest.mark.parametrize("noise,discount,threshold", [(0.0, 1.0, 50.0), (0.1, 0.95, 50.0), (0.2, 0.9, 1.5), (0.0, 1.0, 5.0)])
def test_pref_loss_closed_form_matches_torch_autograd(noise, discount, threshold):
    """The closed forms the `imb_pref_loss` kernel evaluates (oracle/pref_port.pref_loss_closed_form, a NumPy float32 twin)
    against torch autograd through probability_port + binary_cross_entropy: probabilities, loss, accuracy and the gradient
    with respect to every transition reward, including clipped pairs and soft preferences."""
    from oracle import pref_port

    th.manual_seed(1)
    P, L, scale = 41, 11, 0.5
    rews = (th.randn(2, P, L) * 2.5).requires_grad_()
    y = (th.rand(P) < 0.5).float()
    y[::5] = th.rand(len(y[::5]))
    probs = th.stack([pref_port.probability_port(rews[0, k], rews[1, k], noise, discount, threshold) for k in range(P)])
    loss = th.nn.functional.binary_cross_entropy(probs, y)
    (loss * scale).backward()
    p, l, acc, grad = pref_port.pref_loss_closed_form(rews.detach().numpy(), y.numpy(), noise, discount, threshold, scale)
    np.testing.assert_allclose(p, probs.detach().numpy(), rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(l, float(loss.detach()), rtol=1e-5)
    assert abs(acc - float(((probs > 0.5) == (y > 0.5)).float().mean())) < 1e-6
    g = rews.grad.numpy()
    np.testing.assert_allclose(grad, g, rtol=3e-5, atol=1e-9 + 1e-6 * float(np.abs(g).max()))
    assert (grad[:, np.abs(((rews[1] - rews[0]).detach().numpy() * discount ** np.arange(L)).sum(1)) > threshold] == 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("noise,discount,threshold", [(0.0, 1.0, 50.0), (0.1, 0.95, 50.0), (0.2, 0.9, 1.5)])
def test_pref_loss_kernel_matches_torch_autograd(noise, discount, threshold):
    """imb_pref_loss = PreferenceModel.probability + F.binary_cross_entropy + autograd's d loss / d rewards
    (:487-530, :1043-1090), including clipped pairs (zero gradient), soft preferences and the statistics accumulator."""
    from imitation_b200 import _lib

    th.manual_seed(0)
    P, L, scale = 77, 13, 0.25
    rews = (th.randn(2, P, L) * 2.0).cuda().requires_grad_()
    y = (th.rand(P) < 0.5).float()
    y[::5] = th.rand(len(y[::5]))
    y = y.cuda()
    w = discount ** th.arange(L, device="cuda")
    d = th.clip(((rews[1] - rews[0]) * w).sum(1), -threshold, threshold)
    p = noise * 0.5 + (1 - noise) / (1 + d.exp())
    loss = th.nn.functional.binary_cross_entropy(p, y)
    (loss * scale).backward()
    grad, probs, stats = th.zeros(2 * P * L, device="cuda"), th.zeros(P, device="cuda"), th.zeros(8, device="cuda")
    for _ in range(2):  # two minibatches accumulate
        _lib.pref_loss(rews.detach().reshape(-1).contiguous(), P, L, y, noise, discount, threshold, scale, grad, probs, stats, 1)
    th.cuda.synchronize()
    np.testing.assert_allclose(probs.cpu().numpy(), p.detach().cpu().numpy(), rtol=2e-6, atol=1e-7)
    g = rews.grad.reshape(-1).cpu().numpy()
    np.testing.assert_allclose(grad.cpu().numpy(), g, rtol=2e-5, atol=1e-9 + 1e-6 * float(np.abs(g).max()))
    acc = float(((p > 0.5) == (y > 0.5)).float().mean())
    np.testing.assert_allclose(stats.cpu().numpy(), [0, 0, 0, 0, 2 * float(loss), 2 * acc, 2, 0], rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_fused_reward_trainer_step_equals_the_autograd_path():
    """BasicRewardTrainer's device-only step (imb_gather_rows -> imb_reward_forward -> imb_pref_loss -> imb_disc_fwd_bwd
    -> reduce + AdamW) against the per-minibatch autograd + torch.optim.AdamW path on the same data and seeds: same
    minibatches, same statistics, same weights and optimiser state; then the ensemble trainer on bagging subsets."""
    from imitation_b200 import spaces
    from imitation_b200.algorithms import preference_comparisons as pc
    from imitation_b200.rewards import reward_nets
    from imitation_b200.util import networks

    z, (Do, Da, L, P), trajs, pairs = _golden()
    obs_space, act_space = spaces.Box(-np.inf, np.inf, (Do,)), spaces.Box(-1.0, 1.0, (Da,))
    frags = [(_as_traj(a), _as_traj(b)) for a, b in pairs]
    ds = pc.PreferenceDataset()
    ds.push(frags, z["prefs_sampled"].astype(np.float32))
    nets, trainers = [], []
    for fused in (True, False):
        th.manual_seed(4)
        net = reward_nets.BasicRewardNet(obs_space, act_space, hid_sizes=(32, 32),
                                         normalize_input_layer=networks.RunningNorm).cuda()
        pm = pc.PreferenceModel(net, noise_prob=0.05, discount_factor=0.97)
        tr = pc.BasicRewardTrainer(pm, pc.CrossEntropyRewardLoss(), rng=np.random.default_rng(5), batch_size=6,
                                   minibatch_size=3, epochs=3, lr=2e-3)
        tr.use_fused_step = fused
        th.manual_seed(10)
        before = __import__("imitation_b200")._lib.LAUNCHES["count"]
        tr.train(ds)
        tr.launches = __import__("imitation_b200")._lib.LAUNCHES["count"] - before
        nets.append(net)
        trainers.append(tr)
    a, b = trainers
    assert a._fused_target(ds) is not None and "_fused_opt" in a.__dict__ and "_fused_opt" not in b.__dict__
    for k in ("loss", "accuracy"):
        np.testing.assert_allclose(a.last_epoch_stats[k], b.last_epoch_stats[k], rtol=1e-4, atol=1e-6)
    for key in ("mean/reward/epoch-0/train/loss", "mean/reward/epoch-2/train/gt_reward_loss",
                "mean/reward/epoch-1/train/accuracy"):
        np.testing.assert_allclose(a.logger.name_to_value[key], b.logger.name_to_value[key], rtol=1e-4, atol=1e-6)
    sa, sb = nets[0].state_dict(), nets[1].state_dict()
    for k in sa:
        if k == "mlp.dense_final.bias":  # gradient = rounding noise (cancels in r2 - r1); Adam turns it into +-lr steps
            continue
        np.testing.assert_allclose(sa[k].cpu().numpy(), sb[k].cpu().numpy(), rtol=2e-4, atol=2e-6, err_msg=k)
    pa, pb = list(nets[0].parameters()), list(nets[1].parameters())
    for x, y in zip(pa[:-1], pb[:-1]):
        assert float(a.optim.state[x]["step"]) == float(b.optim.state[y]["step"]) > 0
        np.testing.assert_allclose(a.optim.state[x]["exp_avg"].cpu().numpy(), b.optim.state[y]["exp_avg"].cpu().numpy(),
                                   rtol=1e-3, atol=1e-7)
    # an ensemble: every member on its own bagging subset, all on the fused step, the fragment pool shared
    members = [reward_nets.BasicRewardNet(obs_space, act_space, hid_sizes=(32, 32)).cuda() for _ in range(3)]
    ens = reward_nets.RewardEnsemble(obs_space, act_space, members)
    et = pc.EnsembleTrainer(pc.PreferenceModel(ens), pc.CrossEntropyRewardLoss(), rng=np.random.default_rng(2),
                            batch_size=4, epochs=2, lr=1e-3)
    w0 = [m.mlp.dense0.weight.detach().clone() for m in members]
    et.train(ds)
    assert all("_fused_opt" in t.__dict__ for t in et.member_trainers)
    assert all(not th.equal(m.mlp.dense0.weight, w) for m, w in zip(members, w0))
    assert np.isfinite(et.last_epoch_stats["loss"]) and 0.0 <= et.last_epoch_stats["accuracy"] <= 1.0
    assert et._preference_model._pool is not None and et._preference_model._pool.table is not None
