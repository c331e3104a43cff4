"""GPU parity tests at the C-ABI level: every kernel of libimb.so against the CPU oracle
(oracle/*_port.py) and the golden vectors generated from the reference's own modules.

Tolerances: integer / index / done-mask work is bit-exact; fp32 results 1e-5 relative
(north_star), looser only where a documented chain of optimiser steps amplifies rounding.
"""
import numpy as np
import pytest
import torch as th

from tests import golden_util as G

pytestmark = pytest.mark.gpu

RTOL = 1e-5
LOGIT_RTOL, LOGIT_ATOL = 1e-5, 1e-6   # north_star: logits within 1e-5 relative (fp32)


@pytest.fixture(scope="module")
def L():
    from imitation_b200 import _lib

    _lib.lib()
    return _lib


def dev(x, dtype=None):
    t = th.as_tensor(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


def new_state(L):
    st = th.zeros(L.ST_WORDS, dtype=th.int64, device="cuda")
    return st


# ---------------------------------------------------------------------------------------------
# Philox streams / index sampling: bit-exact against oracle/philox.py
# ---------------------------------------------------------------------------------------------
def test_env_reset_matches_philox_twin(L):
    from oracle import philox

    E, Do, seed = 1000, 17, 12345678901
    env = L.EnvDesc(d_obs=Do, d_act=6, discrete=0, horizon=10, seed=seed, env_id_offset=7)
    st = new_state(L)
    st[L.ST_EPISODE] = 3
    obs = th.empty(Do, E, device="cuda")
    L.env_reset(obs, E, env, st)
    want = 0.1 * philox.normals(seed, philox.STREAM_ENV_RESET, np.arange(E, dtype=np.uint32) + 7, np.uint32(3), Do)
    np.testing.assert_allclose(obs.cpu().numpy().T, want, rtol=2e-6, atol=2e-7)


def test_sample_indices_bit_exact(L):
    from oracle import philox

    st = new_state(L)
    st[L.ST_RING_N] = 512
    seed = 99
    for draw in range(3):
        idx = th.empty(8192, dtype=th.int64, device="cuda")
        L.sample_indices(0, idx, 8192, 0, seed, st)
        want = philox.randint(seed, philox.STREAM_REPLAY, draw, 8192, 512)
        np.testing.assert_array_equal(idx.cpu().numpy(), want)
    assert int(st[L.ST_REPLAY_DRAW]) == 3
    # expert stream: endless Feistel permutations, drop_last
    n, B = 1000, 96
    got, want = [], []
    st = new_state(L)
    per_epoch = n // B
    for k in range(3 * per_epoch + 2):
        idx = th.empty(B, dtype=th.int64, device="cuda")
        L.sample_indices(1, idx, B, n, seed, st)
        got.append(idx.cpu().numpy())
        ep, pos = divmod(k, per_epoch)
        want.append(philox.feistel_perm(seed, philox.STREAM_EXPERT, ep, n)[pos * B:(pos + 1) * B])
    np.testing.assert_array_equal(np.stack(got), np.stack(want))
    first_epoch = np.concatenate(got[:per_epoch])
    assert len(np.unique(first_epoch)) == len(first_epoch)  # without replacement inside an epoch


# ---------------------------------------------------------------------------------------------
# tables / ring / gather: bit-exact against the reference's Buffer semantics (golden)
# ---------------------------------------------------------------------------------------------
def test_ring_store_matches_reference_buffer(L):
    z = G.load("replay_buffer")
    cap, Do, Da = 10, 3, 2
    tw = 2 * Do + Da + 1
    table = th.zeros(cap, tw, device="cuda")
    st = new_state(L)
    for i in range(5):
        t = G.sub(z, f"store{i}")
        n = len(t["obs"])
        L.table_store(table, cap, Do, Da, dev(t["obs"]), dev(t["acts"]), None, dev(t["next_obs"]),
                      dev(t["dones"].astype(np.uint8)), n, True, st)
        L.ring_advance(st, cap, n)
        idx_n = [int(st[L.ST_RING_IDX]), int(st[L.ST_RING_N])]
        assert idx_n == list(z[f"store{i}/idx"])
        tab = table.cpu().numpy()
        nd = idx_n[1]
        # rows the reference has written so far (its array is zero elsewhere, ours too)
        np.testing.assert_array_equal(tab[:, :Do], z[f"store{i}/obs_arr"])
        np.testing.assert_array_equal(tab[:, -1] > 0.5, z[f"store{i}/dones_arr"])
        assert nd <= cap


def test_gather_rows_and_onehot(L):
    rng = np.random.default_rng(0)
    cap, Do, n_act = 777, 5, 3
    tw = 2 * Do + n_act + 1
    obs = rng.standard_normal((cap, Do)).astype(np.float32)
    nobs = rng.standard_normal((cap, Do)).astype(np.float32)
    acts = rng.integers(0, n_act, cap)
    dones = rng.random(cap) < 0.3
    table = th.zeros(cap, tw, device="cuda")
    st = new_state(L)
    L.table_store(table, cap, Do, n_act, dev(obs), None, dev(acts, th.int64), dev(nobs), dev(dones.astype(np.uint8)),
                  cap, False, st)
    want_rows = np.concatenate([obs, np.eye(n_act, dtype=np.float32)[acts], nobs, dones[:, None].astype(np.float32)], 1)
    np.testing.assert_array_equal(table.cpu().numpy(), want_rows)
    for n in (1, 31, 32, 33, 1000):
        idx = rng.integers(0, cap, n)
        ld = max(128, (n + 200 + 127) // 128 * 128)
        batch = th.zeros(tw + 1, ld, device="cuda")
        L.gather_rows(table, cap, tw, dev(idx, th.int64), n, batch, ld, 64)
        got = batch.cpu().numpy()
        np.testing.assert_array_equal(got[:tw, 64:64 + n], want_rows[idx].T)
        assert not got[:, :64].any() and not got[:, 64 + n:].any() and not got[tw].any()
    batch = th.zeros(tw + 1, 896, device="cuda")
    L.gather_rows(table, cap, tw, None, cap, batch, 896, 0)  # identity gather
    np.testing.assert_array_equal(batch.cpu().numpy()[:tw, :cap], want_rows.T)


# ---------------------------------------------------------------------------------------------
# discriminator: golden vectors from the reference's GAIL/AIRL train_disc
# ---------------------------------------------------------------------------------------------
def _flat_from_state(state, shaped):
    """reference state_dict (numpy) -> (flat params, norm floats, norm counts)."""
    def mlp(prefix):
        ps, i = [], 0
        while f"{prefix}dense{i}.weight" in state:
            ps += [state[f"{prefix}dense{i}.weight"].ravel(), state[f"{prefix}dense{i}.bias"].ravel()]
            i += 1
        ps += [state[f"{prefix}dense_final.weight"].ravel(), state[f"{prefix}dense_final.bias"].ravel()]
        nm = []
        cnt = 0
        if f"{prefix}normalize_input.running_mean" in state:
            nm = [state[f"{prefix}normalize_input.running_mean"], state[f"{prefix}normalize_input.running_var"]]
            cnt = int(state[f"{prefix}normalize_input.count"])
        return ps, nm, cnt
    if shaped:
        p0, n0, c0 = mlp("_base.mlp.")
        p1, n1, c1 = mlp("potential._potential_net.")
        return np.concatenate(p0 + p1), (np.concatenate(n0 + n1) if n0 else np.zeros(0, np.float32)), [c0, c1]
    p0, n0, c0 = mlp("mlp.")
    return np.concatenate(p0), (np.concatenate(n0) if n0 else np.zeros(0, np.float32)), [c0, 0]


def _make_desc(name, z):
    from imitation_b200 import _desc

    algo, shaped, kw = G.DISC_CASES[name]
    d_obs, d_act, discrete, B, mb, steps, seed = [int(v) for v in z["meta"]]
    if shaped:
        d = _desc.disc_desc(d_obs, d_act, hid_sizes=kw["reward_hid_sizes"], potential_hid_sizes=kw["potential_hid_sizes"],
                            normalize_input=kw["normalize_input"], shaped=True, subtract_logp=(algo == "airl"))
    else:
        d = _desc.disc_desc(d_obs, d_act, hid_sizes=kw["hid_sizes"], normalize_input=kw["normalize_input"],
                            use_next_state=kw.get("use_next_state", False), use_done=kw.get("use_done", False),
                            subtract_logp=(algo == "airl"))
    return d, (d_obs, d_act, bool(discrete), B, mb, steps)


def _upload_rows(L, tr, Do, Da, discrete):
    """host transitions -> device AoS table via imb_table_store (covers the one-hot path)."""
    n = len(tr["obs"])
    table = th.zeros(n, 2 * Do + Da + 1, device="cuda")
    st = th.zeros(L.ST_WORDS, dtype=th.int64, device="cuda")
    L.table_store(table, n, Do, Da, dev(tr["obs"]), None if discrete else dev(tr["acts"], th.float32),
                  dev(tr["acts"], th.int64) if discrete else None, dev(tr["next_obs"]),
                  dev(np.asarray(tr["dones"]).astype(np.uint8)), n, False, st)
    return table


@pytest.mark.parametrize("name", sorted(G.DISC_CASES))
def test_disc_update_matches_reference_golden(L, name):
    from imitation_b200 import _desc

    z = G.load(name)
    d, (Do, Da, discrete, B, mb, steps) = _make_desc(name, z)
    algo, shaped, _ = G.DISC_CASES[name]
    params, norm, counts = _flat_from_state(G.sub(z, "init"), shaped)
    assert len(params) == d.n_params
    P = dev(params)
    NS = dev(norm) if len(norm) else th.zeros(2, device="cuda")
    NC = dev(np.array(counts, np.int32))
    M, V = th.zeros_like(P), th.zeros_like(P)
    ws = th.zeros(L.disc_workspace_floats(d), device="cuda")
    st = new_state(L)
    stats_out = th.zeros(16, device="cuda")
    opt = L.Adam(lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8)
    tw, bw = _desc.table_width(Do, Da), _desc.batch_rows(Do, Da)
    Mpol = z["policy_M"]
    keys = [str(k) for k in z["stats_keys"]]
    order = ["disc_loss", "disc_acc", "disc_acc_expert", "disc_acc_gen", "disc_entropy",
             "disc_proportion_expert_true", "disc_proportion_expert_pred", "n_expert", "n_generated"]
    has_norm = bool(d.base.has_norm)
    for s in range(steps):
        ex, ge = G.sub(z, f"step{s}/expert"), G.sub(z, f"step{s}/gen")
        t_ex, t_ge = _upload_rows(L, ex, Do, Da, discrete), _upload_rows(L, ge, Do, Da, discrete)
        n = 2 * mb
        ld = _desc.batch_ld(n)
        logits = th.zeros(n, device="cuda")
        for i, start in enumerate(range(0, B, mb)):
            batch = th.zeros(bw, ld, device="cuda")
            L.gather_rows(t_ex[start:start + mb], mb, tw, None, mb, batch, ld, 0)
            L.gather_rows(t_ge[start:start + mb], mb, tw, None, mb, batch, ld, mb)
            if algo == "airl":
                obs = np.concatenate([ex["obs"][start:start + mb], ge["obs"][start:start + mb]])
                acts = np.concatenate([ex["acts"][start:start + mb], ge["acts"][start:start + mb]])
                batch[bw - 1, :n] = G.fixed_logp(Mpol, obs, acts).cuda()
            if has_norm:
                L.disc_norm_update(d, batch, ld, n, NS, NC, ws)
            flags = (L.IMB_F_ZERO_GRAD if i == 0 else 0) | (L.IMB_F_TRAIN_NORM if has_norm else 0)
            L.disc_fwd_bwd(d, P, NS, batch, ld, n, mb, 1.0 / (2 * B), None, logits, flags, ws)
            # odd steps take the fused reduce + Adam launch of the last minibatch, even steps the two separate ones
            if s % 2 == 1 and start + mb >= B:
                L.disc_reduce_adam(d, opt, P, M, V, 1.0, ws, st, stats_out)
            else:
                L.disc_reduce(d, ws, None)
        if s % 2 == 0:
            L.disc_adam(d, opt, P, M, V, None, 1.0, ws, st, stats_out)
        th.cuda.synchronize()
        got = dict(zip(order, stats_out.cpu().numpy()[:9]))
        want = dict(zip(keys, z[f"step{s}/stats"]))
        for k in order:
            np.testing.assert_allclose(got[k], want[k], rtol=2e-5, atol=1e-6, err_msg=f"{name} step{s} {k}")
        wp, wn, wc = _flat_from_state(G.sub(z, f"step{s}/state"), shaped)
        # Adam's first steps move every weight by ~lr regardless of gradient scale, so a 1e-5
        # relative gradient difference can flip ~1e-8 absolute; compare with atol 2e-6.
        np.testing.assert_allclose(P.cpu().numpy(), wp, rtol=RTOL, atol=2e-6, err_msg=f"{name} params step{s}")
        if has_norm:
            np.testing.assert_allclose(NS.cpu().numpy(), wn, rtol=RTOL, atol=1e-6, err_msg=f"{name} norm step{s}")
            assert list(NC.cpu().numpy()[: (2 if shaped else 1)]) == wc[: (2 if shaped else 1)]
        # eval-mode logits / rewards on the query rows
        q = G.sub(z, f"step{s}/query")
        tq = _upload_rows(L, q, Do, Da, discrete)
        nq = len(q["obs"])
        ldq = _desc.batch_ld(nq)
        bq = th.zeros(bw, ldq, device="cuda")
        L.gather_rows(tq, nq, tw, None, nq, bq, ldq, 0)
        bq[bw - 1, :nq] = dev(z[f"step{s}/query_logp"])
        out = th.zeros(nq, device="cuda")
        # north_star: discriminator logits within 1e-5 relative.  The forward is evaluated with the REFERENCE's parameters
        # and statistics of this step (our own have drifted by up to the 1e-5 / 2e-6 asserted above, which would be
        # measured instead of the kernel); LOGIT_ATOL covers logits that cancel to ~0 (|logit| <~ 0.1 in these cases).
        Pq, NSq = dev(wp), (dev(wn) if len(wn) else NS)
        L.reward_forward(d, Pq, NSq, bq, ldq, nq, 1, out)
        np.testing.assert_allclose(out.cpu().numpy(), z[f"step{s}/query_logits"], rtol=LOGIT_RTOL, atol=LOGIT_ATOL)
        L.reward_forward(d, Pq, NSq, bq, ldq, nq, 2 if algo == "gail" else 0, out)
        np.testing.assert_allclose(out.cpu().numpy(), z[f"step{s}/reward_train"], rtol=LOGIT_RTOL, atol=LOGIT_ATOL)
        # and with our own parameters: the drift stays inside the parameter tolerance propagated through the net
        L.reward_forward(d, P, NS, bq, ldq, nq, 1, out)
        np.testing.assert_allclose(out.cpu().numpy(), z[f"step{s}/query_logits"], rtol=1e-4, atol=2e-5)
    assert int(st[L.ST_DISC_STEP]) == steps


def _torch_disc_reference(d_obs, d_act, hid, x_obs, x_act, params, n_exp):
    """plain torch fp32 reference of BasicRewardNet + BCE on the GPU (cuBLAS) for big sizes."""
    W1, b1, W2, b2, wf, bf = params
    x = th.cat([x_obs, x_act], 1)
    h1 = th.relu(x @ W1.T + b1)
    h2 = th.relu(h1 @ W2.T + b2)
    logit = (h2 @ wf.T + bf).squeeze(1)
    y = th.zeros_like(logit)
    y[:n_exp] = 1
    loss = th.nn.functional.binary_cross_entropy_with_logits(logit, y)
    return logit, loss


def test_disc_fwd_bwd_large_against_torch_fp32(L):
    """north_star size sweep point: 2^18 rows, Din=23, 32x32, checked against torch autograd
    in fp32 (TF32 off), plus the row-partition additivity property of the accumulator."""
    from imitation_b200 import _desc

    th.backends.cuda.matmul.allow_tf32 = False
    Do, Da, n = 17, 6, 1 << 18
    d = _desc.disc_desc(Do, Da)
    g = th.Generator(device="cuda").manual_seed(0)
    P = (th.rand(d.n_params, device="cuda", generator=g) - 0.5) * 0.6
    bw, ld = _desc.batch_rows(Do, Da), _desc.batch_ld(n)
    batch = th.zeros(bw, ld, device="cuda")
    batch[:, :n] = th.randn(bw, n, device="cuda", generator=g)
    ws = th.zeros(L.disc_workspace_floats(d), device="cuda")
    logits = th.zeros(n, device="cuda")
    NS = th.zeros(2, device="cuda")
    L.disc_fwd_bwd(d, P, NS, batch, ld, n, n // 2, 1.0 / n, None, logits, L.IMB_F_ZERO_GRAD, ws)
    grad = th.zeros(d.n_params, device="cuda")
    L.disc_reduce(d, ws, grad)
    shapes = [(32, 23), (32,), (32, 32), (32,), (1, 32), (1,)]
    ps, o = [], 0
    for s in shapes:
        k = int(np.prod(s))
        ps.append(P[o:o + k].view(s).clone().requires_grad_(True))
        o += k
    lg, loss = _torch_disc_reference(Do, Da, 32, batch[:Do, :n].T, batch[Do:Do + Da, :n].T, ps, n // 2)
    loss.backward()
    want = th.cat([p.grad.ravel() for p in ps])
    np.testing.assert_allclose(logits.cpu().numpy(), lg.detach().cpu().numpy(), rtol=LOGIT_RTOL, atol=2e-6)
    scale = float(want.abs().max())
    np.testing.assert_allclose(grad.cpu().numpy(), want.cpu().numpy(), rtol=1e-3, atol=1e-5 * max(scale, 1.0))


@pytest.mark.parametrize("n", [100, 128, 1000, 16384, 1 << 18])
@pytest.mark.parametrize("dims", [(17, 6), (4, 2), (20, 11)])
def test_disc_tensor_core_path_matches_ffma_and_fp64(L, n, dims):
    """The tcgen05 kernel (3xTF32 split, TMEM accumulators) against the fp32-FFMA kernel of round 1 and against
    a float64 evaluation of the same network: logits within north_star's 1e-5 relative, gradients alike, every
    statistic equal.  Sizes cover a partial tile, one tile, several tiles per CTA (persistent loop, weight-gradient
    accumulator kept in TMEM across tiles) and K1 = 8 / 24 / 32 input columns."""
    from imitation_b200 import _desc

    Do, Da = dims
    d = _desc.disc_desc(Do, Da)
    din = Do + Da
    g = th.Generator(device="cuda").manual_seed(n + Do)
    P = (th.rand(d.n_params, device="cuda", generator=g) - 0.5) * 0.6
    bw, ld = _desc.batch_rows(Do, Da), _desc.batch_ld(n)
    batch = th.zeros(bw, ld, device="cuda")
    batch[:, :n] = th.randn(bw, n, device="cuda", generator=g)
    NS = th.zeros(2, device="cuda")
    n_exp = n // 2
    out = {}
    for name, extra in (("tc", 0), ("ffma", L.IMB_F_NO_TENSOR)):
        ws = th.zeros(L.disc_workspace_floats(d), device="cuda")
        logits = th.zeros(n, device="cuda")
        grad = th.zeros(d.n_params, device="cuda")
        L.disc_fwd_bwd(d, P, NS, batch, ld, n, n_exp, 1.0 / n, None, logits, L.IMB_F_ZERO_GRAD | extra, ws)
        L.disc_reduce(d, ws, grad)
        th.cuda.synchronize()
        stats = ws[_desc_stats_offset(L, d):_desc_stats_offset(L, d) + 5].cpu().numpy()
        out[name] = (logits.cpu().numpy(), grad.cpu().numpy(), stats)
    # float64 reference
    Pd = P.double()
    o = 0
    W1 = Pd[o:o + 32 * din].view(32, din); o += 32 * din
    b1 = Pd[o:o + 32]; o += 32
    W2 = Pd[o:o + 1024].view(32, 32); o += 1024
    b2 = Pd[o:o + 32]; o += 32
    wf = Pd[o:o + 32]; o += 32
    bf = Pd[o]
    ps = [t.clone().requires_grad_(True) for t in (W1, b1, W2, b2, wf, bf)]
    x = batch[:din, :n].T.double()
    lg = th.relu(th.relu(x @ ps[0].T + ps[1]) @ ps[2].T + ps[3]) @ ps[4] + ps[5]
    y = th.zeros(n, device="cuda", dtype=th.float64)
    y[:n_exp] = 1
    loss = th.nn.functional.binary_cross_entropy_with_logits(lg, y)
    loss.backward()
    want_g = th.cat([p.grad.ravel() for p in ps]).cpu().numpy()
    want_l = lg.detach().cpu().numpy()
    for name in ("tc", "ffma"):
        lgt, grd, _ = out[name]
        # north_star: logits within 1e-5 relative (fp32); the floor of 1e-6 absolute covers logits that cancel to ~0
        np.testing.assert_allclose(lgt, want_l, rtol=1e-5, atol=1e-6, err_msg=f"{name} logits")
        scale = float(np.abs(want_g).max())
        np.testing.assert_allclose(grd, want_g, rtol=2e-4, atol=2e-6 * max(scale, 1.0), err_msg=f"{name} grads")
    # the two kernels agree far inside the tolerance against fp64, and on every statistic
    np.testing.assert_allclose(out["tc"][0], out["ffma"][0], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(out["tc"][2][2:], out["ffma"][2][2:], rtol=0, atol=0)       # counts: exact
    np.testing.assert_allclose(out["tc"][2][:2], out["ffma"][2][:2], rtol=2e-6)             # loss / entropy sums


def _desc_stats_offset(L, d):
    """float offset of the reduced statistics inside the workspace (mirror of ws_layout in csrc/imb_disc.cu)."""
    return (d.n_params + 31) // 32 * 32


def test_norm_batch_stats_and_fold_match_running_norm(L):
    """imb_norm_batch_stats / imb_norm_fold (the policy feature-norm side effect of SURVEY App. A.14 and the multi-GPU
    RunningNorm merge) against RunningNorm.update_stats (util/networks.py:111-134) restated in torch: immediate mode, deferred
    slots folded in order, and an explicit slot count (the all-gathered list of the multi-GPU step)."""
    from imitation_b200 import _desc
    from oracle import nets_port

    Do, Da = 7, 3
    d = _desc.disc_desc(Do, Da, normalize_input=True)
    ws = th.zeros(L.disc_workspace_floats(d), device="cuda")
    g = th.Generator(device="cuda").manual_seed(3)
    port = nets_port.RunningNormPort(Do)
    port.train()
    ns = th.cat([th.zeros(Do), th.ones(Do)]).cuda()
    nc = th.zeros(1, dtype=th.int32, device="cuda")
    cap = 8
    defer = th.zeros(4 + cap * (2 * Do + 1), device="cuda")
    batches = []
    for i, n in enumerate((100, 257, 64, 1000)):
        ld = _desc.batch_ld(n)
        b = th.zeros(_desc.batch_rows(Do, Da), ld, device="cuda")
        b[:, :n] = th.randn(b.shape[0], n, device="cuda", generator=g) * (1 + i) + i
        batches.append((b, ld, n))
    # immediate
    b, ld, n = batches[0]
    L.norm_batch_stats(d, b, ld, n, 0, Do, ns, nc, None, 0, ws)
    port.update_stats(b[:Do, :n].t().cpu())
    np.testing.assert_allclose(ns.cpu().numpy(), th.cat([port.running_mean, port.running_var]).numpy(), rtol=1e-5, atol=1e-6)
    assert int(nc) == int(port.count) == 100
    # deferred: two batches into slots, nothing changes until the fold, then both are applied in order
    before = ns.clone()
    for b, ld, n in batches[1:3]:
        L.norm_batch_stats(d, b, ld, n, 0, Do, ns, nc, defer, cap, ws)
        port.update_stats(b[:Do, :n].t().cpu())
    th.cuda.synchronize()
    assert th.equal(ns, before) and int(nc) == 100 and float(defer[0]) == 2.0
    L.norm_fold(Do, defer, ns, nc)
    np.testing.assert_allclose(ns.cpu().numpy(), th.cat([port.running_mean, port.running_var]).numpy(), rtol=1e-5, atol=1e-6)
    assert int(nc) == int(port.count) == 421 and float(defer[0]) == 0.0
    # explicit slot count: the list keeps its counter
    b, ld, n = batches[3]
    L.norm_batch_stats(d, b, ld, n, 0, Do, ns, nc, defer, cap, ws)
    L.norm_fold(Do, defer, ns, nc, 1)
    port.update_stats(b[:Do, :n].t().cpu())
    np.testing.assert_allclose(ns.cpu().numpy(), th.cat([port.running_mean, port.running_var]).numpy(), rtol=1e-5, atol=1e-6)
    assert int(nc) == int(port.count) == 1421 and float(defer[0]) == 1.0


# ---------------------------------------------------------------------------------------------
# NormalizedRewardNet scan
# ---------------------------------------------------------------------------------------------
def test_reward_norm_scan_matches_port(L):
    from oracle import nets_port

    rng = np.random.default_rng(3)
    E, T = 50, 7
    raw = (rng.standard_normal((T, E)) * 3 + 1).astype(np.float32)
    port = nets_port.OutputNormPort()
    want = np.stack([port(raw[t]) for t in range(T)])
    r = dev(raw)
    ns = th.tensor([0.0, 1.0], device="cuda")
    nc = th.zeros(1, dtype=th.int32, device="cuda")
    L.reward_norm_scan(r, E, T, E, 1, ns, nc, 1e-5, True)
    np.testing.assert_allclose(r.cpu().numpy(), want, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(ns.cpu().numpy(), [float(port.norm.running_mean), float(port.norm.running_var)], rtol=1e-5)
    assert int(nc) == E * T


# ---------------------------------------------------------------------------------------------
# generator: rollout kernel + GAE + PPO update against the SB3 restatement (oracle/ppo_port.py)
# ---------------------------------------------------------------------------------------------
def _policy_flat(pol):
    sd = pol.state_dict()
    parts = [sd["pi.0.weight"], sd["pi.0.bias"], sd["pi.2.weight"], sd["pi.2.bias"],
             sd["vf.0.weight"], sd["vf.0.bias"], sd["vf.2.weight"], sd["vf.2.bias"],
             sd["action_net.weight"], sd["action_net.bias"], sd["value_net.weight"], sd["value_net.bias"]]
    if not pol.discrete:
        parts.append(sd["log_std"])
    return th.cat([p.reshape(-1) for p in parts]).float()


def _oracle_rollout(Do, Da, discrete, E, T, H, hidden, normalize_features, reward_mode, n_rounds, seed):
    """CPU oracle: SynthVecEnv -> BufferingPort -> RewardRelabelPort -> PPOPort.collect_rollouts."""
    from oracle import data_port, nets_port, ppo_port, synth_env

    th.manual_seed(seed)
    spec = synth_env.SynthEnvSpec(Do, Da, discrete=discrete, horizon=H, seed=seed)
    venv = synth_env.SynthVecEnv(spec, E, env_id_offset=5)
    pol = ppo_port.ActorCriticPort(Do, Da, discrete=discrete, hidden=(hidden, hidden),
                                   normalize_features=normalize_features)
    with th.no_grad():
        for p in pol.parameters():  # orthogonal init with gain 0.01 gives degenerate actions; spread them
            p.add_(0.3 * th.randn_like(p))
        if normalize_features:
            pol.feat_norm.running_mean.normal_(0, 0.1)
            pol.feat_norm.running_var.uniform_(0.5, 1.5)
            pol.feat_norm.count.fill_(10)
    net = nets_port.BasicRewardNetPort(Do, Da, hid_sizes=(32, 32), normalize_input=True)
    with th.no_grad():
        net.mlp.normalize_input.running_mean.normal_(0, 0.2)
        net.mlp.normalize_input.running_var.uniform_(0.5, 2.0)
    net.eval()
    rng = np.random.default_rng(seed + 1)
    noise = (rng.random((n_rounds * T, E)).astype(np.float32) if discrete
             else rng.standard_normal((n_rounds * T, E, Da)).astype(np.float32))
    buffering = data_port.BufferingPort(venv)
    n_actions = Da if discrete else None
    if reward_mode == 0:
        train_env = buffering
        buffering_reset = buffering.reset
    else:
        train_env = data_port.RewardRelabelPort(
            buffering, lambda o, a, no, d: nets_port.predict_port(net, o, a, no, d, n_actions, gail_transform=True))
    gen = ppo_port.PPOPort(pol, train_env, n_steps=T, gamma=0.97, gae_lambda=0.9, noise_fn=lambda step: noise[step])
    if reward_mode == 0:
        gen._last_obs = buffering.reset()
    else:
        gen._last_obs = train_env._old_obs
    gen._last_starts = np.ones(E, dtype=bool)
    return spec, venv, pol, net, gen, buffering, noise


@pytest.mark.parametrize("cfg", [
    dict(Do=17, Da=6, discrete=False, E=37, T=9, H=5, hidden=32, norm=True, reward_mode=1),
    dict(Do=4, Da=2, discrete=True, E=64, T=6, H=4, hidden=32, norm=False, reward_mode=1),
    dict(Do=11, Da=3, discrete=False, E=33, T=4, H=1000, hidden=64, norm=False, reward_mode=0),
])
def test_rollout_gae_matches_oracle(L, cfg):
    from imitation_b200 import _desc
    from oracle import data_port

    Do, Da, discrete, E, T, H = cfg["Do"], cfg["Da"], cfg["discrete"], cfg["E"], cfg["T"], cfg["H"]
    n_rounds, seed = 2, 11
    spec, venv, pol, net, gen, buffering, noise = _oracle_rollout(Do, Da, discrete, E, T, H, cfg["hidden"],
                                                                  cfg["norm"], cfg["reward_mode"], n_rounds, seed)
    pd = _desc.policy_desc(Do, Da, discrete, cfg["hidden"], cfg["norm"])
    PP = _policy_flat(pol).cuda()
    PN = (th.cat([pol.feat_norm.running_mean, pol.feat_norm.running_var]).cuda() if cfg["norm"]
          else th.zeros(2, device="cuda"))
    dd = _desc.disc_desc(Do, Da, normalize_input=True)
    DP = th.cat([p.detach().reshape(-1) for p in net.mlp.parameters()]).cuda()
    DN = th.cat([net.mlp.normalize_input.running_mean, net.mlp.normalize_input.running_var]).cuda()
    env = L.EnvDesc(d_obs=Do, d_act=Da, discrete=int(discrete), horizon=H, seed=seed, env_id_offset=5)
    EP = dev(_desc.synth_env_params(Do, Da, seed))
    np.testing.assert_array_equal(EP.cpu().numpy()[:Do * Do], spec.A.ravel())
    hp = L.PpoHparams(gamma=0.97, gae_lambda=0.9, clip_range=0.2, ent_coef=0.0, vf_coef=0.5, max_grad_norm=0.5,
                      lr=3e-4, adam_eps=1e-5, n_epochs=1, batch_size=32, normalize_advantage=1)
    st = new_state(L)
    obs = th.empty(Do, E, device="cuda")
    L.env_reset(obs, E, env, st)
    rw = L.rollout_row_width(pd)
    tw = _desc.table_width(Do, Da)
    cap = E * T - 5
    ring = th.zeros(cap, tw, device="cuda")
    ring_ref = data_port.ReplayBufferPort(cap, (Do,), () if discrete else (Da,), np.float32,
                                          np.int64 if discrete else np.float32)
    da_store = 1 if discrete else Da
    for rnd in range(n_rounds):
        tbl = th.zeros(E * T, rw, device="cuda")
        flat = th.zeros(E * T, tw, device="cuda")
        aux = th.zeros(2 * E + 2 * E * T, device="cuda")
        nz = dev(noise[rnd * T:(rnd + 1) * T])
        L.rollout(env, EP, obs, pd, PP, PN, dd, DP, DN, cfg["reward_mode"], hp, E, T, tbl, ring, cap, flat, aux, nz, st)
        L.gae(tbl, rw, Do + da_store + 1, E, T, aux, 0.97, 0.9, st, H)
        L.rollout_advance(st, E, T, H, cap)
        th.cuda.synchronize()
        buf = gen.collect_rollouts()
        trajs, ep_lens = buffering.pop_trajectories()
        want_flat = data_port.flatten_port(trajs)
        ring_ref.store(want_flat)
        got = tbl.cpu().numpy().reshape(E, T, rw)

        def col(a):  # oracle [T, E, ...] -> [E, T, ...]
            return np.swapaxes(a, 0, 1)
        # closed-loop trajectories: fp32 rounding differences (FMA order, tanhf) compound through
        # policy -> action -> dynamics over the T steps and both rounds, hence atol 1e-4 on states;
        # the first step of the first round is checked tightly below.
        if rnd == 0:
            np.testing.assert_allclose(got[:, 0, :Do], buf["obs"][0], rtol=1e-5, atol=2e-6, err_msg="obs t=0")
            np.testing.assert_allclose(got[:, 0, Do + da_store + 1], buf["values"][0], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(got[:, :, :Do], col(buf["obs"]), rtol=1e-3, atol=1e-4, err_msg="obs")
        if discrete:
            np.testing.assert_array_equal(got[:, :, Do], col(buf["actions"]))
        else:
            np.testing.assert_allclose(got[:, :, Do:Do + Da], col(buf["actions"]), rtol=1e-3, atol=1e-4)
        c = Do + da_store
        np.testing.assert_allclose(got[:, :, c], col(buf["log_probs"]), rtol=1e-3, atol=2e-4, err_msg="logp")
        np.testing.assert_allclose(got[:, :, c + 1], col(buf["values"]), rtol=1e-3, atol=1e-4, err_msg="value")
        np.testing.assert_allclose(got[:, :, c + 2], col(buf["rewards"]), rtol=1e-3, atol=2e-4, err_msg="reward")
        np.testing.assert_allclose(got[:, :, c + 3], col(buf["advantages"]), rtol=1e-3, atol=5e-4, err_msg="adv")
        np.testing.assert_allclose(got[:, :, c + 4], col(buf["returns"]), rtol=1e-3, atol=5e-4, err_msg="ret")
        # flattened transitions: order and done mask bit-exact, floats to tolerance
        gf = flat.cpu().numpy()
        np.testing.assert_array_equal(gf[:, -1] > 0.5, want_flat["dones"])
        np.testing.assert_allclose(gf[:, :Do], want_flat["obs"], rtol=1e-3, atol=1e-4)
        np.testing.assert_allclose(gf[:, Do + Da:2 * Do + Da], want_flat["next_obs"], rtol=1e-3, atol=1e-4)
        if discrete:
            np.testing.assert_array_equal(gf[:, Do:Do + Da].argmax(1), want_flat["acts"])
        else:
            np.testing.assert_allclose(gf[:, Do:Do + Da], want_flat["acts"], rtol=1e-3, atol=1e-4)
        # ring: same rows at the same positions as the reference's Buffer.store
        gr = ring.cpu().numpy()
        assert [int(st[L.ST_RING_IDX]), int(st[L.ST_RING_N])] == [ring_ref._buffer._idx, ring_ref._buffer._n_data]
        np.testing.assert_allclose(gr[:, :Do], ring_ref._buffer._arrays["obs"], rtol=1e-3, atol=1e-4)
        np.testing.assert_array_equal(gr[:, -1] > 0.5, ring_ref._buffer._arrays["dones"])
        assert int(st[L.ST_EP_STEP]) == ((rnd + 1) * T) % H
    np.testing.assert_allclose(obs.cpu().numpy().T, gen._last_obs, rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("cfg", [
    dict(Do=17, Da=6, discrete=False, hidden=32, norm=True, N=512, mb=64, epochs=3),
    dict(Do=4, Da=2, discrete=True, hidden=32, norm=False, N=200, mb=64, epochs=2),
    dict(Do=9, Da=3, discrete=False, hidden=20, norm=False, N=256, mb=32, epochs=2),  # width < 32: zero padding
    dict(Do=30, Da=8, discrete=False, hidden=32, norm=True, N=128, mb=48, epochs=2),  # Ant-shaped, ragged last batch
    # k_ppo_update_gen: tower width 64 (SB3 MlpPolicy default) and / or minibatches of more than 64 rows (tuned
    # airl_seals_walker: 128, airl_seals_hopper: 512)
    dict(Do=17, Da=6, discrete=False, hidden=64, norm=True, N=512, mb=64, epochs=2),
    dict(Do=11, Da=3, discrete=False, hidden=32, norm=True, N=1024, mb=512, epochs=2),   # Hopper-shaped, 4 passes per step
    dict(Do=4, Da=2, discrete=True, hidden=64, norm=False, N=600, mb=200, epochs=2),     # ragged passes, ragged last batch
    dict(Do=27, Da=8, discrete=False, hidden=64, norm=True, N=256, mb=128, epochs=2),    # Ant-shaped 64x64
    dict(Do=9, Da=3, discrete=False, hidden=40, norm=True, N=256, mb=96, epochs=2),      # width between 32 and 64: padding
    dict(Do=17, Da=6, discrete=False, hidden=32, norm=True, N=512, mb=64, epochs=3, force_general=True),
    dict(Do=4, Da=2, discrete=True, hidden=32, norm=False, N=200, mb=64, epochs=2, force_general=True),
])
def test_ppo_update_matches_oracle(L, cfg, monkeypatch):
    from imitation_b200 import _desc
    from oracle import ppo_port

    if cfg.get("force_general"):  # the general kernel on a shape the specialised one covers
        monkeypatch.setenv("IMB_PPO_FORCE_GENERAL", "1")

    Do, Da, discrete, N, mb, epochs = cfg["Do"], cfg["Da"], cfg["discrete"], cfg["N"], cfg["mb"], cfg["epochs"]
    th.manual_seed(5)
    rng = np.random.default_rng(5)
    pol = ppo_port.ActorCriticPort(Do, Da, discrete=discrete, hidden=(cfg["hidden"],) * 2,
                                   normalize_features=cfg["norm"])
    with th.no_grad():
        for p in pol.parameters():
            p.add_(0.2 * th.randn_like(p))
    E, T = N // 8, 8
    buf = dict(obs=rng.standard_normal((T, E, Do)).astype(np.float32) * 1.3 + 0.2,
               actions=(rng.integers(0, Da, (T, E)).astype(np.float32) if discrete
                        else rng.standard_normal((T, E, Da)).astype(np.float32)),
               values=rng.standard_normal((T, E)).astype(np.float32),
               log_probs=(rng.standard_normal((T, E)).astype(np.float32) * 0.3 - (0.7 if discrete else 8.0)),
               advantages=rng.standard_normal((T, E)).astype(np.float32) * 2,
               returns=rng.standard_normal((T, E)).astype(np.float32), rewards=np.zeros((T, E), np.float32))
    perms = np.stack([rng.permutation(N) for _ in range(epochs)])
    hp = dict(gamma=0.99, gae_lambda=0.95, clip_range=0.2, ent_coef=0.01, vf_coef=0.5, max_grad_norm=0.5,
              learning_rate=3e-4)
    gen = ppo_port.PPOPort(pol, None, n_steps=T, batch_size=mb, n_epochs=epochs, perm_fn=lambda e, n: perms[e], **hp)
    pd = _desc.policy_desc(Do, Da, discrete, cfg["hidden"], cfg["norm"])
    PP = _policy_flat(pol).cuda()
    PN = (th.cat([pol.feat_norm.running_mean, pol.feat_norm.running_var]).cuda() if cfg["norm"]
          else th.zeros(2, device="cuda"))
    PC = th.zeros(1, dtype=th.int32, device="cuda")
    rw = L.rollout_row_width(pd)
    da_store = 1 if discrete else Da
    tbl = np.zeros((E, T, rw), np.float32)
    sw = lambda a: np.swapaxes(a, 0, 1)
    tbl[:, :, :Do] = sw(buf["obs"])
    tbl[:, :, Do:Do + da_store] = sw(buf["actions"]).reshape(E, T, da_store)
    c = Do + da_store
    tbl[:, :, c], tbl[:, :, c + 1] = sw(buf["log_probs"]), sw(buf["values"])
    tbl[:, :, c + 3], tbl[:, :, c + 4] = sw(buf["advantages"]), sw(buf["returns"])
    M, V = th.zeros_like(PP), th.zeros_like(PP)
    n_steps = epochs * ((N + mb - 1) // mb)
    loss_log = th.zeros(n_steps, 4, device="cuda")
    st = new_state(L)
    hpc = L.PpoHparams(gamma=0.99, gae_lambda=0.95, clip_range=0.2, ent_coef=0.01, vf_coef=0.5, max_grad_norm=0.5,
                       lr=3e-4, adam_eps=1e-5, n_epochs=epochs, batch_size=mb, normalize_advantage=1)
    L.ppo_update(pd, PP, PN, PC, M, V, dev(tbl.reshape(N, rw)), N, hpc, dev(perms, th.int64), 0, loss_log, st)
    th.cuda.synchronize()
    gen.train(buf)
    want_log = np.array(gen.loss_log, np.float32)
    got_log = loss_log.cpu().numpy()
    # first step: identical parameters -> tight; later steps drift with Adam's sign-like updates
    np.testing.assert_allclose(got_log[0], want_log[0], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(got_log, want_log, rtol=2e-3, atol=2e-4)
    np.testing.assert_allclose(PP.cpu().numpy(), _policy_flat(pol).numpy(), rtol=1e-3, atol=2e-4)
    assert int(st[L.ST_PPO_STEP]) == n_steps and int(st[L.ST_PPO_EPOCH]) == epochs
    if cfg["norm"]:
        np.testing.assert_allclose(PN.cpu().numpy(), th.cat([pol.feat_norm.running_mean,
                                                              pol.feat_norm.running_var]).numpy(), rtol=1e-4, atol=1e-5)
        assert int(PC) == int(pol.feat_norm.count)


def test_ppo_device_permutation_is_feistel_twin(L):
    """perf-mode minibatch order == oracle/philox.feistel_perm (bit-exact indices): run one epoch
    with lr=0 and recover the visited rows from the per-minibatch value loss."""
    from imitation_b200 import _desc
    from oracle import philox, ppo_port

    Do, Da, N, mb = 3, 2, 256, 32
    th.manual_seed(0)
    pol = ppo_port.ActorCriticPort(Do, Da)
    pd = _desc.policy_desc(Do, Da, False, 32, False)
    PP = _policy_flat(pol).cuda()
    rw = L.rollout_row_width(pd)
    tbl = np.zeros((N, rw), np.float32)
    tbl[:, Do + Da + 4] = np.arange(N) * 0.01  # returns encode the row id
    with th.no_grad():
        v = float(pol.predict_values(th.zeros(1, Do))[0])
    hpc = L.PpoHparams(gamma=0.99, gae_lambda=0.95, clip_range=0.2, ent_coef=0.0, vf_coef=0.5, max_grad_norm=0.5,
                       lr=0.0, adam_eps=1e-5, n_epochs=2, batch_size=mb, normalize_advantage=0)
    st = new_state(L)
    st[L.ST_PPO_EPOCH] = 4
    loss_log = th.zeros(2 * N // mb, 4, device="cuda")
    L.ppo_update(pd, PP, th.zeros(2, device="cuda"), th.zeros(1, dtype=th.int32, device="cuda"), th.zeros_like(PP),
                 th.zeros_like(PP), dev(tbl), N, hpc, None, 77, loss_log, st)
    got = loss_log.cpu().numpy()[:, 1]
    for ep in range(2):
        perm = philox.feistel_perm(77, philox.STREAM_PPO_PERM, 4 + ep, N)
        for k in range(N // mb):
            ids = perm[k * mb:(k + 1) * mb]
            want = np.mean((ids * 0.01 - v) ** 2)
            np.testing.assert_allclose(got[ep * (N // mb) + k], want, rtol=1e-4)


def test_policy_logp_matches_oracle(L):
    from imitation_b200 import _desc
    from oracle import ppo_port

    for discrete, Do, Da in ((False, 17, 6), (True, 4, 3)):
        th.manual_seed(1)
        pol = ppo_port.ActorCriticPort(Do, Da, discrete=discrete, normalize_features=True)
        with th.no_grad():
            for p in pol.parameters():
                p.add_(0.3 * th.randn_like(p))
            pol.feat_norm.running_mean.normal_()
            pol.feat_norm.running_var.uniform_(0.5, 2)
        pol.eval()
        n = 300
        obs = th.randn(n, Do)
        acts = th.randint(0, Da, (n,)) if discrete else th.randn(n, Da)
        with th.no_grad():
            want = pol.evaluate_actions(obs, acts)[1].numpy()
        pd = _desc.policy_desc(Do, Da, discrete, 32, True)
        bw, ld = _desc.batch_rows(Do, Da), _desc.batch_ld(n)
        batch = th.zeros(bw, ld, device="cuda")
        batch[:Do, :n] = obs.T.cuda()
        a = th.nn.functional.one_hot(acts, Da).float() if discrete else acts
        batch[Do:Do + Da, :n] = a.T.cuda()
        PN = th.cat([pol.feat_norm.running_mean, pol.feat_norm.running_var]).cuda()
        L.policy_logp(pd, _policy_flat(pol).cuda(), PN, batch, ld, n, bw - 1)
        np.testing.assert_allclose(batch[bw - 1, :n].cpu().numpy(), want, rtol=1e-4, atol=5e-5)


def test_sync_pack_unpack_matches_torch_formulation(L):
    """Fused replica-state pack/unpack (csrc/imb_sync.cu) == the torch formulation the gloo host-logic test
    covers (distributed.NormStat / RoundSync), with two ranks emulated in one process."""
    from imitation_b200.distributed import NormStat

    rng = np.random.default_rng(11)
    world, k = 2, 17

    def rank_state(seed):
        r = np.random.default_rng(seed)
        avg = [th.tensor(r.standard_normal(n), dtype=th.float32, device="cuda") for n in (3501, 7, 1)]
        return avg

    # common round-start norm state, then each rank applies its own batches (Chan updates)
    mean0, var0, n0 = rng.standard_normal(k), rng.random(k) + 0.5, 640
    ranks = []
    for rk in range(world):
        r = np.random.default_rng(100 + rk)
        x = r.standard_normal((256 * (rk + 1), k)) * 2 + 1
        n1 = n0 + len(x)
        d = x.mean(0) - mean0
        mean1 = mean0 + d * len(x) / n1
        var1 = (var0 * n0 + x.var(0) * len(x) + d * d * n0 * len(x) / n1) / n1
        ranks.append(dict(avg=rank_state(rk), mean=th.tensor(mean1, dtype=th.float32, device="cuda"),
                          var=th.tensor(var1, dtype=th.float32, device="cuda"),
                          count=th.tensor([n1], dtype=th.int32, device="cuda")))
    start = dict(mean=th.tensor(mean0, dtype=th.float32, device="cuda"), var=th.tensor(var0, dtype=th.float32, device="cuda"),
                 count=th.tensor([n0], dtype=th.int32, device="cuda"))
    # torch formulation on CPU copies
    want_avg = [sum(r["avg"][i].double().cpu() for r in ranks) / world for i in range(3)]
    ns = NormStat(ranks[0]["mean"].cpu().clone(), ranks[0]["var"].cpu().clone(), ranks[0]["count"].cpu().clone())
    ns.start = (start["mean"].cpu(), start["var"].cpu(), start["count"].cpu())
    summed = sum(NormStat(r["mean"].cpu(), r["var"].cpu(), r["count"].cpu()).pack() for r in ranks)
    ns.unpack(summed, world)
    # fused kernels: snapshot from the start state, pack per rank, sum, unpack into rank 0's tensors
    d_start = L.sync_desc([], [(start["mean"], start["var"], start["count"])])
    snap = th.zeros(1 + 2 * k, dtype=th.float64, device="cuda")
    L.sync_snapshot(d_start, snap)
    bufs = []
    for r in ranks:
        d = L.sync_desc(r["avg"], [(r["mean"], r["var"], r["count"])])
        b = th.zeros(L.sync_buffer_doubles(d), dtype=th.float64, device="cuda")
        L.sync_pack(d, b)
        bufs.append(b)
    d0 = L.sync_desc(ranks[0]["avg"], [(ranks[0]["mean"], ranks[0]["var"], ranks[0]["count"])])
    L.sync_unpack(d0, bufs[0] + bufs[1], snap, world)
    for got, want in zip(ranks[0]["avg"], want_avg):
        np.testing.assert_allclose(got.cpu().numpy(), want.float().numpy(), rtol=0, atol=0)
    np.testing.assert_allclose(ranks[0]["mean"].cpu().numpy(), ns.mean.numpy(), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(ranks[0]["var"].cpu().numpy(), ns.var.numpy(), rtol=1e-6, atol=1e-7)
    assert int(ranks[0]["count"]) == int(ns.count) == n0 + 256 + 512


def test_fused_sample_gather_bit_identical_to_unfused(L):
    """imb_disc_sample_gather (+ imb_sample_advance2) == 2 x imb_sample_indices + 2 x imb_gather_rows, over several
    updates (epoch roll-over of the expert stream included) and two minibatches per update."""
    rng = np.random.default_rng(5)
    n_e, cap, tw, B, mb, seed = 1000, 512, 41, 192, 96, 1234
    e_table = dev(rng.standard_normal((n_e, tw)).astype(np.float32))
    ring = dev(rng.standard_normal((cap, tw)).astype(np.float32))
    ld = 256
    st_e1, st_g1, st_e2, st_g2 = (new_state(L) for _ in range(4))
    for st in (st_g1, st_g2):
        st[L.ST_RING_N] = 300  # ring not full yet: randint over the stored prefix
    for upd in range(8):
        a = th.zeros(tw + 1, ld, device="cuda")
        b = th.zeros(tw + 1, ld, device="cuda")
        idx_e = th.empty(B, dtype=th.int64, device="cuda")
        idx_g = th.empty(B, dtype=th.int64, device="cuda")
        L.sample_indices(1, idx_e, B, n_e, seed, st_e1)
        L.sample_indices(0, idx_g, B, 0, seed, st_g1)
        for start in range(0, B, mb):
            L.gather_rows(e_table, n_e, tw, idx_e[start:start + mb], mb, a, ld, 0)
            L.gather_rows(ring, cap, tw, idx_g[start:start + mb], mb, a, ld, mb)
            L.disc_sample_gather(e_table, n_e, ring, cap, tw, mb, start, seed, st_e2, st_g2, b, ld)
            np.testing.assert_array_equal(a.cpu().numpy(), b.cpu().numpy(), err_msg=f"update {upd} start {start}")
        L.sample_advance2(B, n_e, st_e2, st_g2)
        np.testing.assert_array_equal(st_e1.cpu().numpy(), st_e2.cpu().numpy())
        np.testing.assert_array_equal(st_g1.cpu().numpy(), st_g2.cpu().numpy())
