"""PPO generator on the GPU (`gen_algo` of the adversarial trainers).

The reference passes a stable-baselines3 `PPO` (scripts/ingredients/rl.py:122-194) and calls
`gen_algo.learn(total_timesteps, reset_num_timesteps=False, callback=...)`
(algorithms/adversarial/common.py:414).  SB3 is a host-side library; this class keeps the
constructor arguments and the attributes the trainers read (`policy`, `n_steps`, `device`,
`set_env`, `get_env`, `set_logger`, `logger`, `num_timesteps`) and runs both halves of `learn`
as kernels: one rollout launch (csrc/imb_rollout.cu) and one persistent PPO-update launch
(csrc/imb_ppo.cu).  Arithmetic follows SB3 2.2.x (oracle/ppo_port.py).
"""
from typing import Optional

import torch as th

from .. import _lib
from ..data import wrappers
from ..policies import base as policies
from ..rewards import reward_wrapper
from ..util import logger as imit_logger


class DevicePPO:
    def __init__(self, policy, env, learning_rate: float = 3e-4, n_steps: int = 2048, batch_size: int = 64,
                 n_epochs: int = 10, gamma: float = 0.99, gae_lambda: float = 0.95, clip_range: float = 0.2,
                 normalize_advantage: bool = True, ent_coef: float = 0.0, vf_coef: float = 0.5,
                 max_grad_norm: float = 0.5, policy_kwargs: Optional[dict] = None, seed: Optional[int] = None,
                 device="cuda", sampling: str = "device", **unused):
        self.device = th.device(device)
        self.n_steps, self.batch_size, self.n_epochs = int(n_steps), int(batch_size), int(n_epochs)
        self._user_seed = None if seed is None else int(seed)  # SB3: the global RNGs are seeded only when a seed is given
        self.seed = 0 if seed is None else int(seed)        # the Philox streams of the kernels always need a key
        self.sampling = sampling
        self.hp = _lib.PpoHparams(gamma=gamma, gae_lambda=gae_lambda, clip_range=clip_range, ent_coef=ent_coef,
                                  vf_coef=vf_coef, max_grad_norm=max_grad_norm, lr=learning_rate, adam_eps=1e-5,
                                  n_epochs=n_epochs, batch_size=batch_size,
                                  normalize_advantage=int(normalize_advantage))
        self.env = None
        self._base_env = None
        self._logger = imit_logger.configure()
        self.num_timesteps = 0
        self.set_env(env)
        if isinstance(policy, policies.ActorCriticPolicy):
            self.policy = policy.to(self.device)
        else:
            cls = {"MlpPolicy": policies.ActorCriticPolicy, "FeedForward32Policy": policies.FeedForward32Policy}.get(
                policy, policy)
            kw = dict(policy_kwargs or {})
            if cls is policies.ActorCriticPolicy:
                kw.setdefault("net_arch", (64, 64))
            if self._user_seed is not None:
                th.manual_seed(self._user_seed)
            self.policy = cls(self._base_env.observation_space, self._base_env.action_space, **kw).to(self.device)
        n = self.policy.desc.n_params
        self.exp_avg = th.zeros(n, device=self.device)
        self.exp_avg_sq = th.zeros(n, device=self.device)
        self._tbl = None
        self._aux = None
        self.loss_log = None      # optional [n_minibatch_steps][4] device tensor (parity tests)
        self.noise = None         # optional pinned sampling noise for the next rollout (parity tests)
        self.perm = None          # optional host permutations [n_epochs][N] (parity tests)
        self._capturing = False   # True while a CUDA graph of the round is being captured
        self.use_cuda_graph = True  # learn(): replay [rollout -> GAE -> advance -> PPO update] as one CUDA graph
        self._graph = None
        self._graph_key = None
        self._graph_launches = 0
        self._eager_iters = 0
        self.ev_rollout = th.cuda.Event()  # recorded after every rollout: generator samples are in the ring

    # -- SB3 surface ---------------------------------------------------------------------------------------------
    def set_env(self, env, force_reset: bool = True) -> None:
        self.env = env
        e, self._rw_wrapper, self._buffering = env, None, None
        while not hasattr(e, "desc"):
            if isinstance(e, reward_wrapper.RewardVecEnvWrapper):
                self._rw_wrapper = e
            if isinstance(e, wrappers.BufferingWrapper):
                self._buffering = e
            if not hasattr(e, "venv"):
                raise TypeError("DevicePPO needs a DeviceVecEnv (optionally wrapped); host VecEnvs have no GPU path")
            e = e.venv
        self._base_env = e

    def get_env(self):
        return self.env

    def set_logger(self, logger) -> None:
        self._logger = logger

    @property
    def logger(self):
        return self._logger

    # -- learn = collect_rollouts + train ---------------------------------------------------------------------------
    def collect_rollouts(self) -> None:
        env = self._base_env
        env.ensure_reset()
        E, T = env.num_envs, self.n_steps
        pol = self.policy
        pp, pn, pc = pol.flat_vectors()
        rw = _lib.rollout_row_width(pol.desc)
        if self._tbl is None or self._tbl.shape[0] != E * T:
            self._tbl = th.zeros(E * T, rw, device=self.device)
            self._aux = th.zeros(2 * E + 2 * E * T, device=self.device)
        disc = dparams = dnorm = None
        mode, out_norm = 0, None
        if self._rw_wrapper is not None:
            net, mode, out_norm = self._rw_wrapper.resolve()
            eng = net.engine()
            disc, dparams, dnorm = eng.desc, eng.params, eng.norm_state
        flat, ring = (None, None)
        if self._buffering is not None:
            flat, ring = self._buffering.rollout_targets(T)
        t0 = env.host_ep_step
        _lib.rollout(env.desc, env.params, env.obs, pol.desc, pp, pn, disc, dparams, dnorm, mode, self.hp, E, T,
                     self._tbl, ring.table if ring is not None else None, ring.capacity if ring is not None else 0,
                     flat, self._aux, self.noise, env.state)
        da = 1 if pol.discrete else pol.d_act
        col_val = pol.d_obs + da + 1
        if out_norm is not None:
            ns, nc = out_norm.output_norm_vectors()
            _lib.reward_norm_scan(self._tbl.view(-1)[col_val + 1:], E, T, rw, T * rw, ns, nc,
                                  out_norm.normalize_output_layer.eps, True)
        _lib.gae(self._tbl, rw, col_val, E, T, self._aux, self.hp.gamma, self.hp.gae_lambda, env.state, env.horizon)
        _lib.rollout_advance(env.state, E, T, env.horizon, ring.capacity if ring is not None else 0)
        if not self._capturing:
            self.after_rollout_host(t0)

    def after_rollout_host(self, t0: int) -> None:
        """Host-side mirrors of what the rollout kernels just did (also called per graph replay)."""
        env = self._base_env
        E, T = env.num_envs, self.n_steps
        env.host_ep_step = (t0 + T) % env.horizon
        if self._buffering is not None:
            self._buffering.after_rollout(T, t0, self._aux[2 * E + E * T:2 * E + 2 * E * T])
        self.num_timesteps += E * T

    def train(self) -> None:
        pol = self.policy
        pp, pn, pc = pol.flat_vectors()
        N = self._tbl.shape[0]
        _lib.ppo_update(pol.desc, pp, pn, pc, self.exp_avg, self.exp_avg_sq, self._tbl, N, self.hp, self.perm,
                        self.seed, self.loss_log, self._base_env.state)

    def _pointer_key(self):
        """Everything a captured graph bakes in: device pointers of the vectors the kernels touch."""
        pp, pn, pc = self.policy.flat_vectors()
        key = [pp.data_ptr(), pn.data_ptr(), pc.data_ptr(), self._tbl.data_ptr() if self._tbl is not None else 0,
               self.n_steps, id(self._rw_wrapper), id(self._buffering)]
        if self._rw_wrapper is not None:
            net, mode, out_norm = self._rw_wrapper.resolve()
            eng = net.engine()
            key += [eng.params.data_ptr(), eng.norm_state.data_ptr(), mode, id(out_norm)]
        if self._buffering is not None and self._buffering._ring is not None:
            key.append(self._buffering._ring.table.data_ptr())
        return tuple(key)

    def _iteration(self) -> None:
        """One collect_rollouts + train, replayed from a CUDA graph once it is warm (the kernels read every
        per-call scalar from the device counter block, so the captured launch sequence is exact)."""
        plain = self.noise is None and self.perm is None and self.loss_log is None
        if self._buffering is not None:
            self._buffering.before_rollout()
        if not (self.use_cuda_graph and plain) or self._eager_iters < 1 or self._tbl is None:
            self.collect_rollouts()
            if not self._capturing:
                self.ev_rollout.record()
            self.train()
            self._eager_iters += 1
            return
        key = self._pointer_key()
        if self._graph is None or key != self._graph_key:
            before = _lib.LAUNCHES["count"]
            self._capturing = True
            try:
                # two graphs with the "rollout done" event between them: the adversarial trainer's discriminator
                # stream starts from that event while the PPO update is still running
                g_roll, g_train = th.cuda.CUDAGraph(), th.cuda.CUDAGraph()
                with th.cuda.graph(g_roll):
                    self.collect_rollouts()
                with th.cuda.graph(g_train, pool=g_roll.pool()):
                    self.train()
            finally:
                self._capturing = False
            self._graph, self._graph_key = (g_roll, g_train), key
            self._graph_launches = _lib.LAUNCHES["count"] - before
            _lib.LAUNCHES["count"] = before
        t0 = self._base_env.host_ep_step
        self._graph[0].replay()
        self.ev_rollout.record()
        self._graph[1].replay()
        _lib.LAUNCHES["count"] += self._graph_launches
        self.after_rollout_host(t0)

    def learn(self, total_timesteps: int, callback=None, reset_num_timesteps: bool = True, **kwargs):
        per = self._base_env.num_envs * self.n_steps
        done = 0
        if callback is not None and hasattr(callback, "init_callback"):
            callback.init_callback(self)
        while done < total_timesteps:
            if callback is not None and hasattr(callback, "on_rollout_start"):
                callback.on_rollout_start()
            self._iteration()
            done += per
        return self

    def predict(self, observation, state=None, episode_start=None, deterministic=False):
        return self.policy.predict(observation, state, episode_start, deterministic)


PPO = DevicePPO
