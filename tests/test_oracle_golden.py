"""CPU: the oracle restatement (oracle/*_port.py) reproduces the golden vectors that
oracle/make_golden.py generated from the reference's own modules.  No GPU, no /root/reference."""
import numpy as np
import pytest
import torch as th

from oracle import data_port, disc_port, gail_port, nets_port, synth_env
from tests import golden_util as G


def _build_port(name, z):
    algo, shaped, kw = G.DISC_CASES[name]
    d_obs, d_act, discrete, B, mb, steps, seed = [int(v) for v in z["meta"]]
    if shaped:
        net = nets_port.ShapedRewardNetPort(d_obs, d_act, **kw)
    else:
        net = nets_port.BasicRewardNetPort(d_obs, d_act, **kw)
    net.load_state_dict(G.state_to_torch(G.sub(z, "init")))
    M = z["policy_M"]
    tr = disc_port.DiscTrainerPort(net, B, mb, airl=(algo == "airl"), n_actions=d_act if discrete else None,
                                   logp_fn=lambda o, a: G.fixed_logp(M, o, a))
    return net, tr, steps, (d_act if discrete else None), M


@pytest.mark.parametrize("name", sorted(G.DISC_CASES))
def test_disc_port_matches_reference_golden(name):
    z = G.load(name)
    net, tr, steps, n_actions, M = _build_port(name, z)
    keys = [str(k) for k in z["stats_keys"]]
    for s in range(steps):
        net.train()
        stats = tr.train_disc(G.sub(z, f"step{s}/expert"), G.sub(z, f"step{s}/gen"))
        net.eval()
        want = dict(zip(keys, z[f"step{s}/stats"]))
        for k in keys:
            np.testing.assert_allclose(stats[k], want[k], rtol=1e-6, atol=1e-7, equal_nan=True, err_msg=f"{k} step{s}")
        got_state = {k: v.numpy() for k, v in net.state_dict().items()}
        for k, v in G.state_to_torch(G.sub(z, f"step{s}/state")).items():
            np.testing.assert_allclose(got_state[k], v.numpy(), rtol=1e-6, atol=1e-7, err_msg=f"{k} step{s}")
        q = G.sub(z, f"step{s}/query")
        st, a, ns, d = nets_port.preprocess_port(q["obs"], q["acts"], q["next_obs"], q["dones"], n_actions)
        with th.no_grad():
            logits = tr.logits(st, a, ns, d, G.fixed_logp(M, q["obs"], q["acts"]))
        np.testing.assert_allclose(logits.numpy(), z[f"step{s}/query_logits"], rtol=1e-6, atol=1e-6)
        rt = nets_port.predict_port(net, q["obs"], q["acts"], q["next_obs"], q["dones"], n_actions,
                                    gail_transform=not tr.airl)
        np.testing.assert_allclose(rt, z[f"step{s}/reward_train"], rtol=1e-6, atol=1e-6)


def test_running_norm_port():
    z = G.load("running_norm")
    rn = nets_port.RunningNormPort(5)
    rn.train()
    for i in range(4):
        y = rn(th.as_tensor(z[f"x{i}"]))
        np.testing.assert_allclose(y.numpy(), z[f"y{i}"], rtol=1e-6, atol=1e-7)
        for k in ("running_mean", "running_var", "count"):
            np.testing.assert_allclose(getattr(rn, k).numpy(), z[f"s{i}/{k}"], rtol=1e-6)
    assert rn.count.dtype == th.int32 and int(rn.count) == 75
    rn.eval()
    np.testing.assert_allclose(rn(th.as_tensor(z["x_eval"])).numpy(), z["y_eval"], rtol=1e-6, atol=1e-7)
    assert int(rn.count) == 75


def test_replay_buffer_port_bit_exact():
    z = G.load("replay_buffer")
    np.random.seed(11)
    buf = data_port.ReplayBufferPort(10, (3,), (2,))
    for i in range(5):
        t = G.sub(z, f"store{i}")
        n = len(t["obs"])
        buf.store(dict(obs=t["obs"], acts=t["acts"], next_obs=t["next_obs"], dones=t["dones"],
                       infos=np.array([{}] * n)))
        assert [buf._buffer._idx, buf._buffer._n_data] == list(z[f"store{i}/idx"])
        np.testing.assert_array_equal(buf._buffer._arrays["obs"], z[f"store{i}/obs_arr"])
        np.testing.assert_array_equal(buf._buffer._arrays["dones"], z[f"store{i}/dones_arr"])
        s = buf.sample(6)
        np.testing.assert_array_equal(s["obs"], z[f"sample{i}/obs"])
        np.testing.assert_array_equal(s["acts"], z[f"sample{i}/acts"])
        np.testing.assert_array_equal(s["dones"], z[f"sample{i}/dones"])


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_rollout_order_port_bit_exact(name):
    z = G.load("rollout_order")
    E, T, H, cap, discrete = [int(v) for v in z[f"{name}/cfg"]]
    spec = synth_env.SynthEnvSpec(4, 2, discrete=bool(discrete), horizon=H, seed=7)
    venv = synth_env.SynthVecEnv(spec, E)
    bw = data_port.BufferingPort(venv)
    bw.reset()
    acts = z[f"{name}/acts_fed"]
    k = 0
    for rnd in range(2):
        for _ in range(T):
            bw.step(acts[k])
            k += 1
        trajs, ep_lens = bw.pop_trajectories()
        tr = data_port.flatten_port(trajs)
        for key in ("obs", "acts", "next_obs", "dones", "rews"):
            np.testing.assert_array_equal(tr[key], z[f"{name}/round{rnd}/{key}"], err_msg=key)
        np.testing.assert_array_equal(np.array(ep_lens), z[f"{name}/round{rnd}/ep_lens"])
        rb = data_port.ReplayBufferPort(cap, (4,), () if discrete else (2,), np.float32,
                                        np.int64 if discrete else np.float32)
        rb.store(tr)
        np.testing.assert_array_equal(rb._buffer._arrays["obs"], z[f"{name}/round{rnd}/ring_obs"])
        assert [rb._buffer._idx, rb._buffer._n_data] == list(z[f"{name}/round{rnd}/ring_idx"])


def test_reward_relabel_port():
    z = G.load("reward_relabel")
    E, T, H = [int(v) for v in z["cfg"]]
    spec = synth_env.SynthEnvSpec(5, 3, horizon=H, seed=9)
    venv = synth_env.SynthVecEnv(spec, E)
    net = nets_port.ShapedRewardNetPort(5, 3, normalize_input=True)
    st = G.sub(z, "net")
    out_norm = nets_port.OutputNormPort()
    out_norm.norm.load_state_dict({k.split(".")[-1]: th.as_tensor(np.array(v)) for k, v in st.items()
                                   if k.startswith("normalize_output_layer")})
    net.load_state_dict(G.state_to_torch({k[len("_base."):]: v for k, v in st.items()
                                          if not k.startswith("normalize_output")}))
    net.eval()
    wrapped = data_port.RewardRelabelPort(
        data_port.BufferingPort(venv),
        lambda o, a, no, d: out_norm(nets_port.predict_port(net, o, a, no, d)))
    for t in range(T):
        o, r, d, infos = wrapped.step(z["acts"][t])
        np.testing.assert_allclose(r, z["rews"][t], rtol=2e-6, atol=2e-6)
        np.testing.assert_array_equal(d, z["dones"][t])
        np.testing.assert_array_equal(o, z["obs"][t])
    for k in ("running_mean", "running_var", "count"):
        np.testing.assert_allclose(getattr(out_norm.norm, k).numpy(), z[f"net_after/normalize_output_layer.{k}"],
                                   rtol=1e-6)


def test_expert_loader_port_bit_exact():
    z = G.load("expert_loader")
    for c in range(3):
        n, B, seed = [int(v) for v in z[f"case{c}/cfg"]]
        th.manual_seed(seed)
        trans = dict(obs=np.arange(n, dtype=np.float32)[:, None], acts=np.arange(n, dtype=np.float32)[:, None],
                     next_obs=np.zeros((n, 1), np.float32), dones=np.zeros(n, bool))
        it = gail_port.expert_iterator_port(trans, B)
        want = z[f"case{c}/idx"]
        for i in range(len(want)):
            got = next(it)["acts"].numpy()[:, 0].astype(np.int64)
            np.testing.assert_array_equal(got, want[i])


def test_train_stats_port():
    z = G.load("train_stats")
    keys = [str(k) for k in z["keys"]]
    for i in range(4):
        st = disc_port.train_stats_port(th.as_tensor(z[f"c{i}/logits"]), th.as_tensor(z[f"c{i}/labels"]),
                                        th.as_tensor(z[f"c{i}/loss"]))
        assert all(isinstance(v, float) for v in st.values())
        np.testing.assert_allclose([st[k] for k in keys], z[f"c{i}/stats"], rtol=1e-6, equal_nan=True)
