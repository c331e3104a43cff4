// imb_rollout_impl.cuh -- stage 1 of the GAIL/AIRL round: generator rollouts, GPU resident.
//
// One launch runs T environment steps for E environments; a CTA owns 128 environments (= tile rows)
// and every per-step network evaluation is a shared-memory tiled GEMM over that tile
// (imb_tile.cuh), with the environment state, the policy, the reward network and the synthetic
// dynamics all resident in shared memory for the whole rollout:
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
// Env state is SoA [d_obs][E] in HBM (coalesced tile load/store); environments never interact inside
// a rollout because the reward net and the policy run in eval mode.
// (First version: one thread per env with register-resident MLPs -- 218 us for 1024 envs x 4 steps,
//  250 KB of unrolled SASS per variant; see profiles/r01_summary.md.)
#pragma once
#include "imb_common.cuh"
#include "imb_mlp.cuh"
#include "imb_tile.cuh"

namespace {

constexpr int RT = 128;              // threads per CTA
// envs (tile rows) per CTA: RR = 32 * RPL with RPL = 1, 2 or 4 rows per lane in the tiled layers; the launch picks
// the smallest tile that still fills the GPU (1024 envs -> 32 CTAs of 32 envs: the kernel's time is one CTA's
// latency, so smaller tiles are faster until the SMs run out); tile row stride RRS = RR + 4.

struct RolloutArgs {
  imb_env_desc env;
  imb_policy_desc pol;
  imb_ppo_hparams hp;
  int reward_mode;
  int deterministic;  // 1: act = mean (Box) / argmax (Discrete), like policy.predict(deterministic=True)
  int64_t E, T;
  int rw;             // rollout row width
  int64_t ring_capacity;
  // shared-memory plan (floats)
  int HP, JP, IP, KU;          // policy width, reward-net width, obs width (padded to 32), d_obs + d_act
  int pol_off, env_off, img_off, img_sz, obsu_off, xn_off, h1_off, h2_off, nobs_off, vec_off, total;
};

// policy image (runtime tower width HP, zero padded):
//   piW1t[Do][HP] pib1[HP] piW2t[HP][HP] pib2[HP] vfW1t vfb1 vfW2t vfb2 Wa[Da][HP] ba[64] wv[HP] bv[4] lstd[64] mean[64] istd[64]
struct PolImg {
  int w1p, b1p, w2p, b2p, w1v, b1v, w2v, b2v, wa, ba, wv, bv, lstd, mean, istd, total;
  __host__ __device__ PolImg(int Do, int Da, int HP) {
    int o = 0;
    w1p = o; o += Do * HP;
    b1p = o; o += HP;
    w2p = o; o += HP * HP;
    b2p = o; o += HP;
    w1v = o; o += Do * HP;
    b1v = o; o += HP;
    w2v = o; o += HP * HP;
    b2v = o; o += HP;
    wa = o; o += Da * HP;
    ba = o; o += 64;
    wv = o; o += HP;
    bv = o; o += 4;
    lstd = o; o += 64;
    mean = o; o += 64;
    istd = o; o += 64;
    total = o;
  }
};

__device__ void load_policy_img(float* sm, const PolImg& S, const imb_policy_desc& pd, int HP,
                                const float* __restrict__ q, const float* __restrict__ norm) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int Do = pd.d_obs, Da = pd.d_act, h = pd.hidden;
  for (int i = tid; i < S.total; i += nt) sm[i] = 0.f;
  __syncthreads();
#pragma unroll 4
  for (int i = tid; i < h * Do; i += nt) {
    const int j = i / Do, k = i - j * Do;
    sm[S.w1p + k * HP + j] = q[pd.off_pi_w1 + i];
    sm[S.w1v + k * HP + j] = q[pd.off_vf_w1 + i];
  }
#pragma unroll 4
  for (int i = tid; i < h * h; i += nt) {
    const int j = i / h, ii = i - j * h;
    sm[S.w2p + ii * HP + j] = q[pd.off_pi_w2 + i];
    sm[S.w2v + ii * HP + j] = q[pd.off_vf_w2 + i];
  }
  for (int i = tid; i < h; i += nt) {
    sm[S.b1p + i] = q[pd.off_pi_b1 + i];
    sm[S.b2p + i] = q[pd.off_pi_b2 + i];
    sm[S.b1v + i] = q[pd.off_vf_b1 + i];
    sm[S.b2v + i] = q[pd.off_vf_b2 + i];
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
    sm[S.mean + i] = pd.has_norm ? norm[i] : 0.f;
    sm[S.istd + i] = pd.has_norm ? 1.0f / sqrtf(norm[Do + i] + pd.norm_eps) : 1.f;
  }
}

// flattened (reference-order) index of local step t of env e; see file header
__device__ __forceinline__ int64_t flat_index(int64_t e, int64_t t, int64_t E, int64_t T, int64_t t0, int64_t H) {
  const int64_t seg = (t0 + t) / H;
  const int64_t start = seg == 0 ? 0 : seg * H - t0;
  int64_t end = (seg + 1) * H - t0;
  if (end > T) end = T;
  return E * start + e * (end - start) + (t - start);
}

enum { ACT_TANH = 0, ACT_RELU = 1 };

// OUT[j][r] = act(bias[j] + sum_k A[k][r] * Wk[k][j]) for j < JPx (multiple of 32); 128 threads:
// warp -> 8-column group, lane -> RPL consecutive rows.
// rows per tile for a rows-per-lane parameter (RPL = 0: the 8-row tile, lane = (row, column pair))
__host__ __device__ constexpr int rows_of(int rpl) { return rpl == 0 ? 8 : 32 * rpl; }

// 8-row tile: warp -> 8-column group, lane -> (row = lane % 8, columns 2 * (lane / 8), +1): two FMAs per input and
// lane, so a layer's latency is ~K x 10 cycles and the grid covers all SMs at 1024 envs (128 CTAs)
template <int ACT>
__device__ __forceinline__ void tile_layer8(const float* __restrict__ A, int K, const float* __restrict__ Wk, int wld,
                                            const float* __restrict__ bias, float* __restrict__ OUT, int JPx) {
  constexpr int RRS = 8 + TILE_PAD;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = lane & 7, c2 = 2 * (lane >> 3);
  for (int jh = 0; jh < JPx / 32; ++jh) {
    const int j0 = jh * 32 + warp * 8 + c2;
    float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;  // two accumulator pairs (even / odd k)
    int k = 0;
#pragma unroll 4
    for (; k + 2 <= K; k += 2) {
      const float x0 = A[k * RRS + r], x1 = A[(k + 1) * RRS + r];
      const float2 w0 = *reinterpret_cast<const float2*>(Wk + k * wld + j0);
      const float2 w1 = *reinterpret_cast<const float2*>(Wk + (k + 1) * wld + j0);
      a0 = fmaf(x0, w0.x, a0);
      a1 = fmaf(x0, w0.y, a1);
      b0 = fmaf(x1, w1.x, b0);
      b1 = fmaf(x1, w1.y, b1);
    }
    if (k < K) {
      const float x0 = A[k * RRS + r];
      const float2 w0 = *reinterpret_cast<const float2*>(Wk + k * wld + j0);
      a0 = fmaf(x0, w0.x, a0);
      a1 = fmaf(x0, w0.y, a1);
    }
    const float z0 = (a0 + b0) + bias[j0], z1 = (a1 + b1) + bias[j0 + 1];
    OUT[j0 * RRS + r] = ACT == ACT_TANH ? tanh_fast(z0) : fmaxf(z0, 0.f);
    OUT[(j0 + 1) * RRS + r] = ACT == ACT_TANH ? tanh_fast(z1) : fmaxf(z1, 0.f);
  }
}

template <int ACT, int RPL>
__device__ __forceinline__ void tile_layer(const float* __restrict__ A, int K, const float* __restrict__ Wk, int wld,
                                           const float* __restrict__ bias, float* __restrict__ OUT, int JPx) {
  if (RPL == 0) {
    tile_layer8<ACT>(A, K, Wk, wld, bias, OUT, JPx);
    return;
  }
  constexpr int RRS = 32 * RPL + TILE_PAD;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r0 = lane * RPL;
  for (int jh = 0; jh < JPx / 32; ++jh) {
    const int j0 = jh * 32 + warp * 8;
    float acc[RPL > 0 ? RPL : 1][8];
#pragma unroll
    for (int a = 0; a < RPL; ++a)
#pragma unroll
      for (int t = 0; t < 8; ++t) acc[a][t] = 0.f;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      float av[RPL > 0 ? RPL : 1];
      if (RPL == 4) {
        const float4 a4 = ld4(A + k * RRS + r0);
        av[0] = a4.x, av[RPL > 1 ? 1 : 0] = a4.y, av[RPL > 2 ? 2 : 0] = a4.z, av[RPL > 3 ? 3 : 0] = a4.w;
      } else if (RPL == 2) {
        const float2 a2 = *reinterpret_cast<const float2*>(A + k * RRS + r0);
        av[0] = a2.x, av[RPL > 1 ? 1 : 0] = a2.y;
      } else {
        av[0] = A[k * RRS + r0];
      }
      const float4 w0 = ld4(Wk + k * wld + j0), w1 = ld4(Wk + k * wld + j0 + 4);
      const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int x = 0; x < RPL; ++x)
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[x][t] = fmaf(av[x], w[t], acc[x][t]);
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float b = bias[j0 + t];
#pragma unroll
      for (int x = 0; x < RPL; ++x) {
        const float z = acc[x][t] + b;
        OUT[(j0 + t) * RRS + r0 + x] = ACT == ACT_TANH ? tanh_fast(z) : fmaxf(z, 0.f);
      }
    }
  }
}

template <int RPL>
__global__ void __launch_bounds__(RT, 1) k_rollout(const RolloutArgs A, const DiscLaunch L,
                                                   const float* __restrict__ env_params, float* __restrict__ env_obs,
                                                   const float* __restrict__ pol_params,
                                                   const float* __restrict__ pol_norm,
                                                   const float* __restrict__ disc_params, float* __restrict__ rollout,
                                                   float* __restrict__ ring, float* __restrict__ flat_out,
                                                   float* __restrict__ aux, const float* __restrict__ noise,
                                                   const int64_t* __restrict__ state) {
  constexpr int RR = rows_of(RPL), RRS = RR + TILE_PAD;
  extern __shared__ __align__(128) float smem[];
  const int tid = threadIdx.x;
  const int rt = tid < RR ? tid : RR - 1;  // tile row of this thread in the thread-per-env parts (threads >= RR idle there)
  const bool rowthread = tid < RR;
  const int Do = A.env.d_obs, Da = A.env.d_act, h = A.pol.hidden, HP = A.HP, JP = A.JP, IP = A.IP, KU = A.KU;
  const PolImg S(Do, Da, HP);
  float* psm = smem + A.pol_off;
  float* esm = smem + A.env_off;   // ABt[KU][IP] | c[IP] | w[IP]
  float* OBSU = smem + A.obsu_off; // [KU][RRS]: obs rows, then the control (clipped action / one-hot) rows
  float* XN = smem + A.xn_off;     // [max(KP)][RRS]: normalised MLP inputs
  float* H1 = smem + A.h1_off;
  float* H2 = smem + A.h2_off;
  float* NOBS = smem + A.nobs_off; // [IP][RRS]
  float* vec = smem + A.vec_off;   // lg[RT]
  float* lg = vec;

  // ---- one-time loads ------------------------------------------------------------------------------
  if (A.reward_mode != 0)
    for (int p = 0; p < L.npass; ++p)
      load_timg(smem + A.img_off + p * A.img_sz, L.pass[p], JP, disc_params,
                L.pass[p].has_norm ? L.pass[p].norm : nullptr, L.pass[p].eps);
  load_policy_img(psm, S, A.pol, HP, pol_params, pol_norm);
  for (int i = tid; i < (KU + 2) * IP; i += RT) esm[i] = 0.f;
  __syncthreads();
  {
    const float* eA = env_params;
    const float* eB = eA + Do * Do;
    const float* eC = eB + Do * Da;
    const float* eW = eC + Do;
    for (int i = tid; i < Do * Do; i += RT) {
      const int r = i / Do, c = i - r * Do;  // A[r][c] -> ABt[c][r]
      esm[c * IP + r] = eA[i];
    }
    for (int i = tid; i < Do * Da; i += RT) {
      const int r = i / Da, c = i - r * Da;  // B[r][c] -> ABt[Do + c][r]
      esm[(Do + c) * IP + r] = eB[i];
    }
    for (int i = tid; i < Do; i += RT) {
      esm[KU * IP + i] = eC[i];
      esm[(KU + 1) * IP + i] = eW[i];
    }
  }
  const float* ecv = esm + KU * IP;
  const float* ewv = esm + (KU + 1) * IP;
  const int64_t E = A.E, T = A.T, H = A.env.horizon;
  const int64_t e0 = (int64_t)blockIdx.x * RR;
  const int64_t e = e0 + rt;
  const bool live = rowthread && e < E;
  // The tile's environment state is SoA [d_obs][E] in HBM: one contiguous RR-float run per feature.  Full, 16-byte
  // aligned tiles are staged by the TMA unit (one cp.async.bulk per feature row, completion on an mbarrier: SASS
  // UBLKCP); ragged tail tiles take the plain coalesced loads.
  __shared__ __align__(8) uint64_t obs_bar;
  const bool bulk_tile = (e0 + RR <= E) && ((E & 3) == 0);
  if (tid == 0) {
    mbar_init(&obs_bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (bulk_tile) {
    if (tid == 0) {
      mbar_expect_tx(&obs_bar, (uint32_t)(Do * RR * 4));
      for (int k = 0; k < Do; ++k) bulk_g2s(OBSU + k * RRS, env_obs + (int64_t)k * E + e0, (uint32_t)(RR * 4), &obs_bar);
    }
    if (rowthread)
      for (int a = 0; a < Da; ++a) OBSU[(Do + a) * RRS + tid] = 0.f;
    mbar_wait(&obs_bar, 0);
  } else if (rowthread) {
    for (int k = 0; k < Do; ++k) OBSU[k * RRS + tid] = live ? env_obs[(int64_t)k * E + e] : 0.f;  // coalesced
    for (int a = 0; a < Da; ++a) OBSU[(Do + a) * RRS + tid] = 0.f;
  }
  __syncthreads();

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

  // value head on the latent tile H2 (thread per env)
  auto value_row = [&]() {
    float v0 = 0.f, v1 = 0.f;
    int j = 0;
    for (; j + 2 <= h; j += 2) {
      v0 = fmaf(psm[S.wv + j], H2[j * RRS + rt], v0);
      v1 = fmaf(psm[S.wv + j + 1], H2[(j + 1) * RRS + rt], v1);
    }
    if (j < h) v0 = fmaf(psm[S.wv + j], H2[j * RRS + rt], v0);
    return psm[S.bv] + (v0 + v1);
  };
  // XN <- feature-normalised copy of a [Do][RRS] observation tile
  auto norm_obs = [&](const float* __restrict__ SRC) {
    for (int i = tid; i < Do * (RR / 4); i += RT) {
      const int k = i / (RR / 4), r4 = (i - k * (RR / 4)) * 4;
      const float4 x = ld4(SRC + k * RRS + r4);
      const float m = psm[S.mean + k], is = psm[S.istd + k];
      st4(XN + k * RRS + r4, make_float4((x.x - m) * is, (x.y - m) * is, (x.z - m) * is, (x.w - m) * is));
    }
    __syncthreads();
  };
  auto value_of_obs = [&](const float* __restrict__ SRC) {
    norm_obs(SRC);
    tile_layer<ACT_TANH, RPL>(XN, Do, psm + S.w1v, HP, psm + S.b1v, H1, HP);
    __syncthreads();
    tile_layer<ACT_TANH, RPL>(H1, h, psm + S.w2v, HP, psm + S.b2v, H2, HP);
    __syncthreads();
    const float v = value_row();
    __syncthreads();
    return v;
  };

  bool done = false;
  for (int64_t t = 0; t < T; ++t) {
    float* row = rollout + (e * T + t) * rw;
    // ---- policy: value tower, then pi tower (H2 ends up holding the pi latent) ------------------------------
    const float value = value_of_obs(OBSU);
    tile_layer<ACT_TANH, RPL>(XN, Do, psm + S.w1p, HP, psm + S.b1p, H1, HP);
    __syncthreads();
    tile_layer<ACT_TANH, RPL>(H1, h, psm + S.w2p, HP, psm + S.b2p, H2, HP);
    __syncthreads();
    // ---- action head + sampling, thread per env ----------------------------------------------------------------
    float logp = 0.f;
    if (!rowthread) {
      // (threads beyond the tile's rows only take part in the tiled layers)
    } else if (!A.pol.discrete) {
      float z4[4] = {0.f, 0.f, 0.f, 0.f};
      for (int a = 0; a < Da; ++a) {
        float m0 = 0.f, m1 = 0.f;
        const float* wa = psm + S.wa + a * HP;
        int j = 0;
        for (; j + 2 <= h; j += 2) {
          m0 = fmaf(wa[j], H2[j * RRS + rt], m0);
          m1 = fmaf(wa[j + 1], H2[(j + 1) * RRS + rt], m1);
        }
        if (j < h) m0 = fmaf(wa[j], H2[j * RRS + rt], m0);
        const float m = psm[S.ba + a] + (m0 + m1);
        if (!A.deterministic && !noise && (a & 3) == 0)
          philox_normal4(A.env.seed, IMB_STREAM_ACT_NOISE, egid, (uint32_t)(gstep0 + t), a >> 2, z4);
        const float z = A.deterministic ? 0.f : noise ? (live ? noise[(t * E + e) * Da + a] : 0.f) : ((a & 3) == 0 ? z4[0] : (a & 3) == 1 ? z4[1] : (a & 3) == 2 ? z4[2] : z4[3]);
        const float ls = psm[S.lstd + a];
        const float sd = expf(ls);
        const float act = fmaf(sd, z, m);
        const float diff = act - m;
        logp += -(diff * diff) / (2.0f * sd * sd) - ls - 0.9189385332046727f;
        if (live) row[Do + a] = act;                                   // SB3 stores the UNCLIPPED action
        OBSU[(Do + a) * RRS + tid] = fminf(fmaxf(act, -1.0f), 1.0f);   // env and wrappers see the clipped one
      }
    } else {
      float mx = -INFINITY;
      for (int a = 0; a < Da; ++a) {
        float m0 = 0.f;
        const float* wa = psm + S.wa + a * HP;
        for (int j = 0; j < h; ++j) m0 = fmaf(wa[j], H2[j * RRS + rt], m0);
        const float m = psm[S.ba + a] + m0;
        OBSU[(Do + a) * RRS + tid] = m;  // logits, overwritten by the one-hot below
        mx = fmaxf(mx, m);
      }
      float se = 0.f;
      for (int a = 0; a < Da; ++a) se += expf(OBSU[(Do + a) * RRS + tid] - mx);
      const float lse = mx + logf(se);
      float u;
      if (noise) {
        u = live ? noise[t * E + e] : 0.f;
      } else {
        uint32_t k0, k1;
        philox_key(A.env.seed, IMB_STREAM_ACT_NOISE, k0, k1);
        u = u01(philox4x32(egid, (uint32_t)(gstep0 + t), 0u, 0u, k0, k1).x);
      }
      int chosen = Da - 1;
      float cdf = 0.f;
      bool found = false;
      for (int a = 0; a < Da; ++a) {
        cdf += expf(OBSU[(Do + a) * RRS + tid] - lse);
        if (!found && !(u >= cdf)) {
          chosen = a;
          found = true;
        }
      }
      if (A.deterministic) {
        chosen = 0;
        for (int a = 1; a < Da; ++a)
          if (OBSU[(Do + a) * RRS + tid] > OBSU[(Do + chosen) * RRS + tid]) chosen = a;
      }
      logp = OBSU[(Do + chosen) * RRS + tid] - lse;
      for (int a = 0; a < Da; ++a) OBSU[(Do + a) * RRS + tid] = (a == chosen) ? 1.f : 0.f;
      if (live) row[Do] = (float)chosen;
    }
    if (live) {
      for (int k = 0; k < Do; ++k) row[k] = OBSU[k * RRS + tid];
      row[col_logp] = logp;
      row[col_val] = value;
    }
    __syncthreads();

    // ---- environment step: NOBS = tanh([obs | u] . [A | B]^T + c) --------------------------------------------------
    tile_layer<ACT_TANH, RPL>(OBSU, KU, esm, IP, ecv, NOBS, IP);
    __syncthreads();
    float rew_env = 0.f;
    for (int i = 0; i < Do; ++i) rew_env = fmaf(ewv[i], NOBS[i * RRS + rt], rew_env);
    if (!A.env.discrete) {
      float pen = 0.f;
      for (int a = 0; a < Da; ++a) {
        const float uu = OBSU[(Do + a) * RRS + rt];
        pen = fmaf(uu, uu, pen);
      }
      rew_env -= 0.1f * pen;
    }
    done = ((t0 + t + 1) % H) == 0;
    const float donef = done ? 1.f : 0.f;

    // ---- learned reward on (obs, clipped act, terminal-fixed next obs, done) ----------------------------------------
    float reward = rew_env;
    if (A.reward_mode != 0) {
      for (int p = 0; p < L.npass; ++p) {
        const PassDesc& Pd = L.pass[p];
        const float* img = smem + A.img_off + p * A.img_sz;
        const int din = Pd.din;
        const float* mean = img + TImg::mean(din, JP);
        const float* istd = img + TImg::istd(din, JP);
        for (int i = tid; i < din * (RR / 4); i += RT) {
          const int k = i / (RR / 4), r4 = (i - k * (RR / 4)) * 4;
          const int fr = L.stage_row[Pd.in_slot[k]];  // batch feature row -> source tile row
          float4 x;
          if (fr < Do + Da) x = ld4(OBSU + fr * RRS + r4);
          else if (fr < 2 * Do + Da) x = ld4(NOBS + (fr - Do - Da) * RRS + r4);
          else x = make_float4(donef, donef, donef, donef);
          const float m = mean[k], is = istd[k];
          st4(XN + k * RRS + r4, make_float4((x.x - m) * is, (x.y - m) * is, (x.z - m) * is, (x.w - m) * is));
        }
        __syncthreads();
        const float* HL = XN;
        int hl = din;
        if (Pd.n_hidden >= 1) {
          tile_layer<ACT_RELU, RPL>(XN, din, img + TImg::w1t(din, JP), JP, img + TImg::b1(din, JP), H1, JP);
          __syncthreads();
          HL = H1;
          hl = Pd.h1;
        }
        if (Pd.n_hidden >= 2) {
          tile_layer<ACT_RELU, RPL>(H1, Pd.h1, img + TImg::w2t(din, JP), JP, img + TImg::b2(din, JP), H2, JP);
          __syncthreads();
          HL = H2;
          hl = Pd.h2;
        }
        const float* wf = img + TImg::wf(din, JP);
        float o0 = 0.f, o1 = 0.f;
        int j = 0;
        for (; j + 2 <= hl; j += 2) {
          o0 = fmaf(wf[j], HL[j * RRS + rt], o0);
          o1 = fmaf(wf[j + 1], HL[(j + 1) * RRS + rt], o1);
        }
        if (j < hl) o0 = fmaf(wf[j], HL[j * RRS + rt], o0);
        const float o = img[TImg::bf(din, JP)] + (o0 + o1);
        const float c = pass_coef(Pd.coef_kind, L.gamma, donef);
        lg[tid] = (p == 0) ? c * o : fmaf(c, o, lg[tid]);
        __syncthreads();
      }
      reward = (A.reward_mode == 1) ? softplus_f(lg[tid]) : lg[tid];
    }
    if (live) row[col_rew] = reward;

    // ---- time-limit bootstrap term gamma * V(terminal obs) (added after reward normalisation) -------------------
    float boot = 0.f;
    if (done) boot = A.hp.gamma * value_of_obs(NOBS);  // block-uniform branch (lock-step envs)
    if (live) {
      aux[2 * E + e * T + t] = boot;
      aux[2 * E + E * T + e * T + t] = rew_env;  // ground-truth env reward (BufferingWrapper records it)
      // ---- flattened transition row (reference order) -> ring / flat_out ---------------------------------------
      const int64_t f = flat_index(e, t, E, T, t0, H);
      float* dst0 = flat_out ? flat_out + f * tw : nullptr;
      float* dst1 = nullptr;
      if (ring && f >= skip) dst1 = ring + ((ring_idx0 + (f - skip)) % A.ring_capacity) * tw;
#pragma unroll 1
      for (int q = 0; q < 2; ++q) {
        float* dst = q == 0 ? dst0 : dst1;
        if (!dst) continue;
        for (int k = 0; k < Do + Da; ++k) dst[k] = OBSU[k * RRS + tid];
        for (int k = 0; k < Do; ++k) dst[Do + Da + k] = NOBS[k * RRS + tid];
        dst[2 * Do + Da] = donef;
      }
    }
    // ---- advance: on done the next observation is the reset observation ------------------------------------------
    if (done) {
      ++episode;
      if (rowthread)
        for (int k = 0; k < Do; ++k)
          OBSU[k * RRS + tid] = 0.1f * philox_normal(A.env.seed, IMB_STREAM_ENV_RESET, egid, (uint32_t)episode, k);
    } else if (rowthread) {
      for (int k = 0; k < Do; ++k) OBSU[k * RRS + tid] = NOBS[k * RRS + tid];
    }
    __syncthreads();
  }
  // ---- tail: state back to HBM, V(last obs) for GAE ------------------------------------------------------
  const float vlast = value_of_obs(OBSU);
  if (bulk_tile) {  // state tile back to HBM through the TMA unit as well (shared -> global bulk copies)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      for (int k = 0; k < Do; ++k)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(env_obs + (int64_t)k * E + e0),
                     "r"(smem_u32(OBSU + k * RRS)), "r"((uint32_t)(RR * 4))
                     : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  } else if (live) {
    for (int k = 0; k < Do; ++k) env_obs[(int64_t)k * E + e] = OBSU[k * RRS + tid];
  }
  if (live) {
    aux[e] = vlast;
    aux[E + e] = done ? 1.f : 0.f;
  }
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

template <int RPL>
static int launch_rollout_t(RolloutArgs A, const DiscLaunch& L, const float* env_params, float* env_obs,
                            const float* pol_params, const float* pol_norm, const float* disc_params, float* rollout,
                            float* ring, float* flat_out, float* aux, const float* noise, const int64_t* state,
                            cudaStream_t st) {
  constexpr int RR = rows_of(RPL), RRS = RR + TILE_PAD;
  auto al = [](int x) { return (x + 31) / 32 * 32; };
  const int Do = A.env.d_obs, Da = A.env.d_act;
  A.HP = A.pol.hidden <= 32 ? 32 : 64;
  A.IP = Do <= 32 ? 32 : 64;
  A.KU = Do + Da;
  int jp = 32, dmax = Do;
  if (A.reward_mode != 0) {
    for (int p = 0; p < L.npass; ++p) {
      if (L.pass[p].n_hidden >= 1 && L.pass[p].h1 > jp) jp = 64;
      if (L.pass[p].n_hidden >= 2 && L.pass[p].h2 > jp) jp = 64;
      if (L.pass[p].din > dmax) dmax = L.pass[p].din;
    }
  }
  A.JP = jp;
  const int wmax = A.HP > A.JP ? A.HP : A.JP;
  int o = 0;
  A.pol_off = o;
  o += al(PolImg(Do, Da, A.HP).total);
  A.env_off = o;
  o += al((A.KU + 2) * A.IP);
  A.img_off = o;
  A.img_sz = al(TImg::size(dmax, A.JP));
  if (A.reward_mode != 0) o += L.npass * A.img_sz;
  A.obsu_off = o;
  o += al(A.KU * RRS);
  A.xn_off = o;
  o += al(dmax * RRS);
  A.h1_off = o;
  o += al(wmax * RRS);
  A.h2_off = o;
  o += al(wmax * RRS);
  A.nobs_off = o;
  o += al(A.IP * RRS);
  A.vec_off = o;
  o += al(RT);
  A.total = o;
  const size_t bytes = (size_t)o * 4;
  IMB_REQUIRE(bytes <= IMB_SMEM_MAX, "rollout kernel needs %zu B of shared memory", bytes);
  static size_t attr_bytes = 0;
  if (bytes > attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(k_rollout<RPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) IMB_FAIL(-2, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_bytes = bytes;
  }
  const int blocks = (int)((A.E + RR - 1) / RR);
  k_rollout<RPL><<<blocks, RT, bytes, st>>>(A, L, env_params, env_obs, pol_params, pol_norm, disc_params, rollout, ring,
                                            flat_out, aux, noise, state);
  IMB_CHECK_LAUNCH("k_rollout");
  return 0;
}

static int launch_rollout(const RolloutArgs& A, const DiscLaunch& L, const float* env_params, float* env_obs,
                          const float* pol_params, const float* pol_norm, const float* disc_params, float* rollout,
                          float* ring, float* flat_out, float* aux, const float* noise, const int64_t* state,
                          cudaStream_t st) {
  // smallest tile that still covers the SMs: the kernel's duration is one CTA's latency
  const int64_t sms = imb_num_sms();
#define IMB_RL(R) \
  return launch_rollout_t<R>(A, L, env_params, env_obs, pol_params, pol_norm, disc_params, rollout, ring, flat_out, aux, \
                             noise, state, st)
  if (A.E <= sms * 8 * 2) IMB_RL(0);
  if (A.E <= sms * 32) IMB_RL(1);
  if (A.E <= sms * 64 * 2) IMB_RL(2);
  IMB_RL(4);
#undef IMB_RL
}
