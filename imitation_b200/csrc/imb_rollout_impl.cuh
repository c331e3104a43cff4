// imb_rollout.cu -- stage 1 of the GAIL/AIRL round: generator rollouts, GPU resident.
//
// One launch runs T environment steps for E environments (thread per env; environments never
// interact inside a rollout because the reward net and the policy run in eval mode):
//   policy forward + sampling      SB3 OnPolicyAlgorithm.collect_rollouts (restated; see
//                                  oracle/ppo_port.py -- parity unpinned by the reference)
//   env step + auto-reset          synthetic MuJoCo-shaped env; VecEnv
//                                  contract of data/rollout.py:161-186
//   reward relabel                 rewards/reward_wrapper.py:92-133 -> RewardNet.predict_processed
//                                  (reward_nets.py:120-204), GAIL transform gail.py:83
//   trajectory bookkeeping         data/wrappers.py:69-148 + data/rollout.py:563-621: the flattened
//                                  transition order (finished trajectories in completion order,
//                                  then partial ones in env order) has a closed form because the
//                                  envs are fixed-horizon and run in lock-step; rows go straight
//                                  into the generator ring with Buffer.store truncation/wrap
//                                  (data/buffer.py:174-192).
// Env state is SoA [d_obs][E] in HBM; all weights (<50 KB) live in shared memory.
#pragma once
#include "imb_common.cuh"
#include "imb_mlp.cuh"

namespace {

constexpr int RT = 32;  // threads (= envs) per CTA: small CTAs spread the envs over many SMs

// shared-memory image of the actor-critic policy, tower width HP (two tanh layers)
template <int HP>
struct PolSm {
  int w1t_pi, b1_pi, w2t_pi, b2_pi, w1t_vf, b1_vf, w2t_vf, b2_vf, wa, ba, wv, bv, lstd, mean, istd, total;
  __host__ __device__ PolSm(int d_obs, int d_act) {
    int o = 0;
    w1t_pi = o; o += d_obs * HP;
    b1_pi = o; o += HP;
    w2t_pi = o; o += HP * HP;
    b2_pi = o; o += HP;
    w1t_vf = o; o += d_obs * HP;
    b1_vf = o; o += HP;
    w2t_vf = o; o += HP * HP;
    b2_vf = o; o += HP;
    wa = o; o += d_act * HP;
    ba = o; o += (d_act + 3) / 4 * 4;
    wv = o; o += HP;
    bv = o; o += 4;
    lstd = o; o += (d_act + 3) / 4 * 4;
    mean = o; o += (d_obs + 3) / 4 * 4;
    istd = o; o += (d_obs + 3) / 4 * 4;
    total = o;
  }
};

template <int HP>
__device__ void load_policy(float* sm, const PolSm<HP>& S, const imb_policy_desc& pd, const float* __restrict__ q,
                            const float* __restrict__ norm) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int Do = pd.d_obs, Da = pd.d_act, h = pd.hidden;
  for (int i = tid; i < S.total; i += nt) sm[i] = 0.f;
  __syncthreads();
  for (int i = tid; i < h * Do; i += nt) {
    const int j = i / Do, k = i - j * Do;
    sm[S.w1t_pi + k * HP + j] = q[pd.off_pi_w1 + i];
    sm[S.w1t_vf + k * HP + j] = q[pd.off_vf_w1 + i];
  }
  for (int i = tid; i < h * h; i += nt) {
    const int j = i / h, ii = i - j * h;
    sm[S.w2t_pi + ii * HP + j] = q[pd.off_pi_w2 + i];
    sm[S.w2t_vf + ii * HP + j] = q[pd.off_vf_w2 + i];
  }
  for (int i = tid; i < h; i += nt) {
    sm[S.b1_pi + i] = q[pd.off_pi_b1 + i];
    sm[S.b2_pi + i] = q[pd.off_pi_b2 + i];
    sm[S.b1_vf + i] = q[pd.off_vf_b1 + i];
    sm[S.b2_vf + i] = q[pd.off_vf_b2 + i];
    sm[S.wv + i] = q[pd.off_val_w + i];
  }
  for (int i = tid; i < Da * h; i += nt) {
    const int a = i / h, ii = i - a * h;
    sm[S.wa + a * HP + ii] = q[pd.off_act_w + i];
  }
  for (int i = tid; i < Da; i += nt) {
    sm[S.ba + i] = q[pd.off_act_b + i];
    if (!pd.discrete) sm[S.lstd + i] = q[pd.off_log_std + i];
  }
  if (tid == 0) sm[S.bv] = q[pd.off_val_b];
  for (int i = tid; i < Do; i += nt) {
    if (pd.has_norm) {
      sm[S.mean + i] = norm[i];
      sm[S.istd + i] = 1.0f / sqrtf(norm[Do + i] + pd.norm_eps);
    } else {
      sm[S.mean + i] = 0.f;
      sm[S.istd + i] = 1.f;
    }
  }
}

// two-layer tanh tower: lat = tanh(W2 tanh(W1 x + b1) + b2); x in shared memory (stride 1)
template <int HP>
__device__ __forceinline__ void tower_fwd(const float* __restrict__ W1t, const float* __restrict__ b1,
                                          const float* __restrict__ W2t, const float* __restrict__ b2,
                                          const float* __restrict__ x, int din, float (&lat)[HP]) {
  float h1[HP];
#pragma unroll
  for (int j = 0; j < HP; ++j) h1[j] = b1[j];
  for (int k = 0; k < din; ++k) {
    const float xv = x[k];
    const float4* w = reinterpret_cast<const float4*>(W1t + k * HP);
#pragma unroll
    for (int j4 = 0; j4 < HP / 4; ++j4) {
      const float4 ww = w[j4];
      h1[4 * j4 + 0] = fmaf(ww.x, xv, h1[4 * j4 + 0]);
      h1[4 * j4 + 1] = fmaf(ww.y, xv, h1[4 * j4 + 1]);
      h1[4 * j4 + 2] = fmaf(ww.z, xv, h1[4 * j4 + 2]);
      h1[4 * j4 + 3] = fmaf(ww.w, xv, h1[4 * j4 + 3]);
    }
  }
#pragma unroll
  for (int j = 0; j < HP; ++j) {
    h1[j] = tanhf(h1[j]);
    lat[j] = b2[j];
  }
#pragma unroll
  for (int i = 0; i < HP; ++i) {
    const float hv = h1[i];
    const float4* w = reinterpret_cast<const float4*>(W2t + i * HP);
#pragma unroll
    for (int j4 = 0; j4 < HP / 4; ++j4) {
      const float4 ww = w[j4];
      lat[4 * j4 + 0] = fmaf(ww.x, hv, lat[4 * j4 + 0]);
      lat[4 * j4 + 1] = fmaf(ww.y, hv, lat[4 * j4 + 1]);
      lat[4 * j4 + 2] = fmaf(ww.z, hv, lat[4 * j4 + 2]);
      lat[4 * j4 + 3] = fmaf(ww.w, hv, lat[4 * j4 + 3]);
    }
  }
#pragma unroll
  for (int j = 0; j < HP; ++j) lat[j] = tanhf(lat[j]);
}

template <int HP>
__device__ __forceinline__ float value_of(const float* psm, const PolSm<HP>& S, const float* feat, int d_obs) {
  float lat[HP];
  tower_fwd<HP>(psm + S.w1t_vf, psm + S.b1_vf, psm + S.w2t_vf, psm + S.b2_vf, feat, d_obs, lat);
  float v = psm[S.bv];
#pragma unroll
  for (int j = 0; j < HP; ++j) v = fmaf(psm[S.wv + j], lat[j], v);
  return v;
}

struct RolloutArgs {
  imb_env_desc env;
  imb_policy_desc pol;
  imb_ppo_hparams hp;
  int reward_mode;
  int deterministic;  // 1: act = mean (Box) / argmax (Discrete), like policy.predict(deterministic=True)
  int64_t E, T;
  int rw;             // rollout row width
  int64_t ring_capacity;
};

// flattened (reference-order) index of local step t of env e; see file header
__device__ __forceinline__ int64_t flat_index(int64_t e, int64_t t, int64_t E, int64_t T, int64_t t0, int64_t H) {
  const int64_t seg = (t0 + t) / H;
  const int64_t start = seg == 0 ? 0 : seg * H - t0;
  int64_t end = (seg + 1) * H - t0;
  if (end > T) end = T;
  return E * start + e * (end - start) + (t - start);
}

template <int HP, int HD>
__global__ void __launch_bounds__(RT) k_rollout(const RolloutArgs A, const DiscLaunch L,
                                                const float* __restrict__ env_params, float* __restrict__ env_obs,
                                                const float* __restrict__ pol_params,
                                                const float* __restrict__ pol_norm,
                                                const float* __restrict__ disc_params, float* __restrict__ rollout,
                                                float* __restrict__ ring, float* __restrict__ flat_out,
                                                float* __restrict__ aux, const float* __restrict__ noise,
                                                const int64_t* __restrict__ state, int img1_off, int env_off,
                                                int pol_off, int scr_off, int scr_ld) {
  extern __shared__ __align__(128) float smem[];
  const int tid = threadIdx.x;
  const int Do = A.env.d_obs, Da = A.env.d_act;
  const PolSm<HP> S(Do, Da);
  float* img[MAX_PASS] = {smem, smem + img1_off, smem + img1_off};
  float* esm = smem + env_off;  // A[Do][Do] | Bm[Do][Da] | c[Do] | w[Do]
  float* psm = smem + pol_off;
  float* scr = smem + scr_off + tid * scr_ld;  // per-thread scratch: obs | u | nobs | xn
  float* s_obs = scr;
  float* s_u = scr + Do;
  float* s_no = s_u + Da;
  float* s_xn = s_no + Do;

  // ---- one-time loads ------------------------------------------------------------------------------
  float* mean2 = nullptr;
  float* istd2 = nullptr;
  if (A.reward_mode != 0) {
    load_mlp<HD>(img[0], L.pass[0], disc_params);
    if (L.npass == 3) {
      load_mlp<HD>(img[1], L.pass[1], disc_params);
      const int din = L.pass[2].din;
      mean2 = img[1] + MlpSm<HD>::size(din);
      istd2 = mean2 + IMB_MAX_DIN;
      for (int i = tid; i < din; i += RT) {
        if (L.pass[2].has_norm) {
          mean2[i] = L.pass[2].norm[i];
          istd2[i] = 1.0f / sqrtf(L.pass[2].norm[din + i] + L.pass[2].eps);
        } else {
          mean2[i] = 0.f;
          istd2[i] = 1.f;
        }
      }
    }
  }
  load_policy<HP>(psm, S, A.pol, pol_params, pol_norm);
  const int n_env_p = Do * Do + Do * Da + 2 * Do;
  for (int i = tid; i < n_env_p; i += RT) esm[i] = env_params[i];
  __syncthreads();
  const float* eA = esm;
  const float* eB = esm + Do * Do;
  const float* eC = eB + Do * Da;
  const float* eW = eC + Do;

  const int64_t e = (int64_t)blockIdx.x * RT + tid;
  if (e >= A.E) return;  // no block-level sync below this point
  const int64_t E = A.E, T = A.T, H = A.env.horizon;
  const int64_t t0 = state[IMB_ST_EP_STEP];
  int64_t episode = state[IMB_ST_EPISODE];
  const int64_t gstep0 = state[IMB_ST_GLOBAL_STEP];
  const uint32_t egid = (uint32_t)(A.env.env_id_offset + e);
  const int rw = A.rw;
  const int tw = 2 * Do + Da + 1;
  const int da_store = A.pol.discrete ? 1 : Da;
  const int col_logp = Do + da_store, col_val = col_logp + 1, col_rew = col_logp + 2;
  const int64_t n_total = E * T;
  const int64_t skip = (A.ring_capacity > 0 && n_total > A.ring_capacity) ? n_total - A.ring_capacity : 0;
  const int64_t ring_idx0 = state[IMB_ST_RING_IDX];

  for (int k = 0; k < Do; ++k) s_obs[k] = env_obs[(int64_t)k * E + e];
  bool done = false;
  float lat[HP];
  for (int64_t t = 0; t < T; ++t) {
    // ---- policy: features, pi tower, value ------------------------------------------------------------
    for (int k = 0; k < Do; ++k) s_xn[k] = (s_obs[k] - psm[S.mean + k]) * psm[S.istd + k];
    const float value = value_of<HP>(psm, S, s_xn, Do);
    tower_fwd<HP>(psm + S.w1t_pi, psm + S.b1_pi, psm + S.w2t_pi, psm + S.b2_pi, s_xn, Do, lat);
    float* row = rollout + (e * T + t) * rw;
    float logp = 0.f;
    if (!A.pol.discrete) {
      for (int a = 0; a < Da; ++a) {
        float m = psm[S.ba + a];
#pragma unroll
        for (int j = 0; j < HP; ++j) m = fmaf(psm[S.wa + a * HP + j], lat[j], m);
        const float z = A.deterministic ? 0.f
                        : noise   ? noise[(t * E + e) * Da + a]
                                  : philox_normal(A.env.seed, IMB_STREAM_ACT_NOISE, egid, (uint32_t)(gstep0 + t), a);
        const float ls = psm[S.lstd + a];
        const float sd = expf(ls);
        const float act = fmaf(sd, z, m);
        const float diff = act - m;
        logp += -(diff * diff) / (2.0f * sd * sd) - ls - 0.9189385332046727f;
        row[Do + a] = act;                               // SB3 stores the UNCLIPPED action
        s_u[a] = fminf(fmaxf(act, -1.0f), 1.0f);         // the env (and the wrappers) see the clipped one
      }
    } else {
      // categorical over Da logits; inverse-CDF sampling from one uniform
      float mx = -INFINITY;
      for (int a = 0; a < Da; ++a) {
        float m = psm[S.ba + a];
#pragma unroll
        for (int j = 0; j < HP; ++j) m = fmaf(psm[S.wa + a * HP + j], lat[j], m);
        s_u[a] = m;
        mx = fmaxf(mx, m);
      }
      float se = 0.f;
      for (int a = 0; a < Da; ++a) se += expf(s_u[a] - mx);
      const float lse = mx + logf(se);
      float u;
      if (noise) {
        u = noise[t * E + e];
      } else {
        uint32_t k0, k1;
        philox_key(A.env.seed, IMB_STREAM_ACT_NOISE, k0, k1);
        u = u01(philox4x32(egid, (uint32_t)(gstep0 + t), 0u, 0u, k0, k1).x);
      }
      int chosen = Da - 1;
      float cdf = 0.f;
      bool found = false;
      for (int a = 0; a < Da; ++a) {
        cdf += expf(s_u[a] - lse);
        if (!found && !(u >= cdf)) {
          chosen = a;
          found = true;
        }
      }
      if (A.deterministic) {
        chosen = 0;
        for (int a = 1; a < Da; ++a)
          if (s_u[a] > s_u[chosen]) chosen = a;
      }
      logp = s_u[chosen] - lse;
      for (int a = 0; a < Da; ++a) s_u[a] = (a == chosen) ? 1.f : 0.f;
      row[Do] = (float)chosen;
    }
    for (int k = 0; k < Do; ++k) row[k] = s_obs[k];
    row[col_logp] = logp;
    row[col_val] = value;

    // ---- environment step -------------------------------------------------------------------------------
    float rew_env = 0.f;
    for (int i = 0; i < Do; ++i) {
      float pre = eC[i];
      for (int j = 0; j < Do; ++j) pre = fmaf(eA[i * Do + j], s_obs[j], pre);
      for (int a = 0; a < Da; ++a) pre = fmaf(eB[i * Da + a], s_u[a], pre);
      const float v = tanhf(pre);
      s_no[i] = v;
      rew_env = fmaf(eW[i], v, rew_env);
    }
    if (!A.env.discrete) {
      float pen = 0.f;
      for (int a = 0; a < Da; ++a) pen = fmaf(s_u[a], s_u[a], pen);
      rew_env -= 0.1f * pen;
    }
    done = ((t0 + t + 1) % H) == 0;
    const float donef = done ? 1.f : 0.f;

    // ---- learned reward on (obs, clipped act, terminal-fixed next obs, done) ---------------------------
    float reward = rew_env;
    if (A.reward_mode != 0) {
      float h1[HD], h2[HD];
      float out = 0.f;
      for (int p = 0; p < L.npass; ++p) {
        const PassDesc& P = L.pass[p];
        const float* mean = (p == 2) ? mean2 : img[p] + MlpSm<HD>::mean_off(P.din);
        const float* istd = (p == 2) ? istd2 : img[p] + MlpSm<HD>::istd_off(P.din);
        for (int k = 0; k < P.din; ++k) {
          const int r = L.stage_row[P.in_slot[k]];  // batch feature row -> source field
          float v;
          if (r < Do) v = s_obs[r];
          else if (r < Do + Da) v = s_u[r - Do];
          else if (r < 2 * Do + Da) v = s_no[r - Do - Da];
          else v = donef;
          s_xn[k] = (v - mean[k]) * istd[k];
        }
        const float o = mlp_forward_row<HD, false>(img[p], P, s_xn, h1, h2);
        out = fmaf(pass_coef(P.coef_kind, L.gamma, donef), o, out);
      }
      reward = (A.reward_mode == 1) ? softplus_f(out) : out;
    }
    row[col_rew] = reward;

    // ---- time-limit bootstrap term gamma * V(terminal obs) (added after reward normalisation) ----------
    float boot = 0.f;
    if (done) {
      for (int k = 0; k < Do; ++k) s_xn[k] = (s_no[k] - psm[S.mean + k]) * psm[S.istd + k];
      boot = A.hp.gamma * value_of<HP>(psm, S, s_xn, Do);
    }
    aux[2 * E + e * T + t] = boot;
    aux[2 * E + E * T + e * T + t] = rew_env;  // ground-truth env reward (BufferingWrapper records it)

    // ---- flattened transition row (reference order) -> ring / flat_out -----------------------------------
    const int64_t f = flat_index(e, t, E, T, t0, H);
    float* dst0 = flat_out ? flat_out + f * tw : nullptr;
    float* dst1 = nullptr;
    if (ring && f >= skip) dst1 = ring + ((ring_idx0 + (f - skip)) % A.ring_capacity) * tw;
#pragma unroll 1
    for (int q = 0; q < 2; ++q) {
      float* dst = q == 0 ? dst0 : dst1;
      if (!dst) continue;
      for (int k = 0; k < Do; ++k) dst[k] = s_obs[k];
      for (int a = 0; a < Da; ++a) dst[Do + a] = s_u[a];
      for (int k = 0; k < Do; ++k) dst[Do + Da + k] = s_no[k];
      dst[2 * Do + Da] = donef;
    }

    // ---- advance: on done the next observation is the reset observation ------------------------------------
    if (done) {
      ++episode;
      for (int k = 0; k < Do; ++k)
        s_obs[k] = 0.1f * philox_normal(A.env.seed, IMB_STREAM_ENV_RESET, egid, (uint32_t)episode, k);
    } else {
      for (int k = 0; k < Do; ++k) s_obs[k] = s_no[k];
    }
  }
  // ---- tail: state back to HBM, V(last obs) for GAE ------------------------------------------------------
  for (int k = 0; k < Do; ++k) env_obs[(int64_t)k * E + e] = s_obs[k];
  for (int k = 0; k < Do; ++k) s_xn[k] = (s_obs[k] - psm[S.mean + k]) * psm[S.istd + k];
  aux[e] = value_of<HP>(psm, S, s_xn, Do);
  aux[E + e] = done ? 1.f : 0.f;
}

__global__ void k_rollout_advance(int64_t* state, int64_t n_envs, int64_t n_steps, int horizon, int64_t ring_cap) {
  const int64_t t0 = state[IMB_ST_EP_STEP];
  state[IMB_ST_EPISODE] += (t0 + n_steps) / horizon;
  state[IMB_ST_EP_STEP] = (t0 + n_steps) % horizon;
  state[IMB_ST_GLOBAL_STEP] += n_steps;
  if (ring_cap > 0) {
    const int64_t n = n_envs * n_steps;
    const int64_t kept = n < ring_cap ? n : ring_cap;
    state[IMB_ST_RING_IDX] = (state[IMB_ST_RING_IDX] + kept) % ring_cap;
    const int64_t nd = state[IMB_ST_RING_N] + kept;
    state[IMB_ST_RING_N] = nd < ring_cap ? nd : ring_cap;
  }
}

// GAE(lambda) per env over its T rows (SB3 RolloutBuffer.compute_returns_and_advantage).
// reward += bootstrap term; episode_start[t+1] == done[t] for this lock-step env, and done[t] is
// recoverable from the bootstrap bookkeeping: done at local step t <=> (t0 + t + 1) % H == 0.
__global__ void __launch_bounds__(128) k_gae(float* __restrict__ rollout, int rw, int col_val, int64_t E,
                                             int64_t T, const float* __restrict__ aux, float gamma, float lam,
                                             const int64_t* __restrict__ state_before, int horizon) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int col_rew = col_val + 1, col_adv = col_val + 2, col_ret = col_val + 3;
  const int64_t t0 = state_before[IMB_ST_EP_STEP];
  float last = 0.f;
  float next_v = aux[e];
  float next_nonterm = 1.0f - aux[E + e];
  for (int64_t t = T - 1; t >= 0; --t) {
    float* row = rollout + (e * T + t) * rw;
    const float r = row[col_rew] + aux[2 * E + e * T + t];
    const float v = row[col_val];
    const float delta = r + gamma * next_v * next_nonterm - v;
    last = delta + gamma * lam * next_nonterm * last;
    row[col_rew] = r;
    row[col_adv] = last;
    row[col_ret] = last + v;
    next_v = v;
    // episode_start of step t  ==  done of step t-1
    next_nonterm = (t > 0 && ((t0 + t) % horizon) == 0) ? 0.f : 1.f;
  }
}

__global__ void k_env_reset(float* __restrict__ env_obs, int64_t E, int d_obs, uint64_t seed, int64_t id_off,
                            const int64_t* __restrict__ state) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const uint32_t ep = (uint32_t)state[IMB_ST_EPISODE];
  for (int k = 0; k < d_obs; ++k)
    env_obs[(int64_t)k * E + e] = 0.1f * philox_normal(seed, IMB_STREAM_ENV_RESET, (uint32_t)(id_off + e), ep, k);
}

}  // namespace

template <int HP, int HD>
static int launch_rollout(const RolloutArgs& A, const DiscLaunch& L, const float* env_params, float* env_obs,
                          const float* pol_params, const float* pol_norm, const float* disc_params, float* rollout,
                          float* ring, float* flat_out, float* aux, const float* noise, const int64_t* state,
                          cudaStream_t st) {
  auto al = [](int x) { return (x + 31) / 32 * 32; };
  int o = 0;
  int img1_off = 0;
  if (A.reward_mode != 0) {
    o += al(MlpSm<HD>::size(L.pass[0].din));
    img1_off = o;
    if (L.npass == 3) o += al(MlpSm<HD>::size(L.pass[1].din) + 2 * IMB_MAX_DIN);
  }
  const int Do = A.env.d_obs, Da = A.env.d_act;
  const int env_off = o;
  o += al(Do * Do + Do * Da + 2 * Do);
  const int pol_off = o;
  o += al(PolSm<HP>(Do, Da).total);
  int maxdin = Do;
  if (A.reward_mode != 0)
    for (int p = 0; p < L.npass; ++p) maxdin = L.pass[p].din > maxdin ? L.pass[p].din : maxdin;
  const int scr_ld = (2 * Do + Da + maxdin) | 1;
  const int scr_off = o;
  o += al(RT * scr_ld);
  const size_t bytes = (size_t)o * 4;
  IMB_REQUIRE(bytes <= IMB_SMEM_MAX, "rollout kernel needs %zu B of shared memory", bytes);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_rollout<HP, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, IMB_SMEM_MAX);
    if (e != cudaSuccess) IMB_FAIL(-2, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int blocks = (int)((A.E + RT - 1) / RT);
  k_rollout<HP, HD><<<blocks, RT, bytes, st>>>(A, L, env_params, env_obs, pol_params, pol_norm, disc_params, rollout,
                                               ring, flat_out, aux, noise, state, img1_off, env_off, pol_off, scr_off,
                                               scr_ld);
  IMB_CHECK_LAUNCH("k_rollout");
  return 0;
}

