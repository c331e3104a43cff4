// imb_ppo.cu -- the generator update (PPO) as ONE persistent single-CTA launch per round.
//
// The reference delegates this to stable-baselines3 (algorithms/adversarial/common.py:414,
// gen_algo.learn -> PPO.train); its arithmetic is restated in oracle/ppo_port.py (parity
// unpinned by the reference).  PPO.train is n_epochs x (N / batch_size)
// strictly sequential optimiser steps on 64-row minibatches: each step is ~1.3 MFLOP, far below
// launch latency, so the whole loop runs inside one kernel with parameters, gradients and Adam
// moments resident in shared memory:
//   gather minibatch rows by permutation -> [feature RunningNorm update] -> advantage
//   normalisation -> phase A (thread per row per tower: forward, loss, backward to dL/dz) ->
//   phase B (warps split output units: weight gradients) -> clip_grad_norm_ -> Adam.
// Also: imb_policy_logp = ActorCriticPolicy.evaluate_actions()[1] for the AIRL discriminator
// batch (common.py:476-519).
#include "imb_common.cuh"

namespace {

constexpr int PT = 256;  // threads: warps 0-3 = policy tower, warps 4-7 = value tower

template <int HP>
struct PpoCfg {
  static constexpr int MBR = (HP == 32) ? 128 : 64;  // max minibatch rows
  static constexpr int LD = HP + 1;
};

struct PpoArgs {
  imb_policy_desc pol;
  imb_ppo_hparams hp;
  int64_t n_rows;
  int rw;
  uint64_t seed;
  int moments_in_smem;
  int mbr;  // minibatch rows the shared-memory tiles are sized for (batch_size rounded up to 32)
};

__device__ __forceinline__ float block_sum(float v, float* red) {
  // red: >= 16 floats of shared memory; all PT threads must call
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < PT / 32; ++w) t += red[w];
  return t;
}

// transposed working images rebuilt from the master parameters after every optimiser step
template <int HP>
__device__ void build_images(const imb_policy_desc& pd, const float* __restrict__ Pm, float* __restrict__ w1t_pi,
                             float* __restrict__ w2t_pi, float* __restrict__ w1t_vf, float* __restrict__ w2t_vf) {
  const int Do = pd.d_obs, h = pd.hidden;
  for (int i = threadIdx.x; i < h * Do; i += PT) {
    const int j = i / Do, k = i - j * Do;
    w1t_pi[k * HP + j] = Pm[pd.off_pi_w1 + i];
    w1t_vf[k * HP + j] = Pm[pd.off_vf_w1 + i];
  }
  for (int i = threadIdx.x; i < h * h; i += PT) {
    const int j = i / h, ii = i - j * h;
    w2t_pi[ii * HP + j] = Pm[pd.off_pi_w2 + i];
    w2t_vf[ii * HP + j] = Pm[pd.off_vf_w2 + i];
  }
}

template <int HP>
__global__ void __launch_bounds__(PT, 1) k_ppo_update(const PpoArgs A, float* __restrict__ g_params,
                                                      float* __restrict__ g_norm, int32_t* __restrict__ g_norm_count,
                                                      float* __restrict__ g_m, float* __restrict__ g_v,
                                                      const float* __restrict__ rollout,
                                                      const int64_t* __restrict__ perm_in,
                                                      float* __restrict__ loss_log, int64_t* __restrict__ state) {
  constexpr int LD = PpoCfg<HP>::LD;
  const int MBR = A.mbr;
  extern __shared__ __align__(128) float smem[];
  __shared__ float red[32];
  __shared__ float bc[8];
  const imb_policy_desc& pd = A.pol;
  const int Do = pd.d_obs, Da = pd.d_act, h = pd.hidden, NP = pd.n_params;
  const int da_store = pd.discrete ? 1 : Da;
  const int col_act = Do, col_logp = Do + da_store, col_adv = col_logp + 3, col_ret = col_logp + 4;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  auto al = [](int x) { return (x + 31) / 32 * 32; };

  // ---- shared-memory carve-up ---------------------------------------------------------------------------
  int o = 0;
  float* Pm = smem + o; o += al(NP);
  float* G = smem + o; o += al(NP);
  float* Mm = g_m;
  float* Vm = g_v;
  if (A.moments_in_smem) {
    Mm = smem + o; o += al(NP);
    Vm = smem + o; o += al(NP);
  }
  float* w1t_pi = smem + o; o += al(Do * HP);
  float* w2t_pi = smem + o; o += al(HP * HP);
  float* w1t_vf = smem + o; o += al(Do * HP);
  float* w2t_vf = smem + o; o += al(HP * HP);
  const int xn_ld = Do | 1;
  float* XN = smem + o; o += al(MBR * xn_ld);
  const int mb_ld = (da_store + 3) | 1;  // act | logp_old | adv | ret
  float* MB = smem + o; o += al(MBR * mb_ld);
  float* T_H1[2], *T_LAT[2], *T_DZ2[2], *T_DZ1[2];
  for (int n = 0; n < 2; ++n) {
    T_H1[n] = smem + o; o += MBR * LD;
    T_LAT[n] = smem + o; o += MBR * LD;
    T_DZ2[n] = smem + o; o += MBR * LD;
    T_DZ1[n] = smem + o; o += MBR * LD;
  }
  const int dm_ld = Da | 1;
  float* DMEAN = smem + o; o += al(MBR * dm_ld);  // dL/d(action mean or logits) per row
  float* DLS = smem + o; o += al(MBR * dm_ld);    // dL/d(log_std) per row (Box)
  float* DVAL = smem + o; o += al(MBR);           // dL/d(value) per row
  float* nmean = smem + o; o += al(Do);
  float* nistd = smem + o; o += al(Do);
  int* s_idx = reinterpret_cast<int*>(smem + o); o += al(MBR);

  for (int i = tid; i < NP; i += PT) {
    Pm[i] = g_params[i];
    if (A.moments_in_smem) {
      Mm[i] = g_m[i];
      Vm[i] = g_v[i];
    }
  }
  for (int i = tid; i < al(Do * HP); i += PT) w1t_pi[i] = w1t_vf[i] = 0.f;
  for (int i = tid; i < al(HP * HP); i += PT) w2t_pi[i] = w2t_vf[i] = 0.f;
  float run_mean = 0.f, run_var = 1.f;  // feature norm state: thread k < Do owns feature k
  int32_t run_count = 0;
  if (pd.has_norm) {
    if (tid < Do) {
      run_mean = g_norm[tid];
      run_var = g_norm[Do + tid];
    }
    run_count = *g_norm_count;
  }
  __syncthreads();
  build_images<HP>(pd, Pm, w1t_pi, w2t_pi, w1t_vf, w2t_vf);
  __syncthreads();

  const int64_t N = A.n_rows;
  const int mb = A.hp.batch_size;
  const int64_t steps_per_epoch = (N + mb - 1) / mb;
  int64_t adam_step = state[IMB_ST_PPO_STEP];
  const int64_t perm_draw0 = state[IMB_ST_PPO_EPOCH];
  int64_t log_i = 0;

  for (int ep = 0; ep < A.hp.n_epochs; ++ep) {
    const FeistelKey fk = feistel_key(A.seed, IMB_STREAM_PPO_PERM, (uint64_t)(perm_draw0 + ep), (uint64_t)N);
    for (int64_t sidx = 0; sidx < steps_per_epoch; ++sidx) {
      const int64_t start = sidx * mb;
      const int nb = (int)min((int64_t)mb, N - start);
      // ---- 1. gather the minibatch rows (warp per row, lanes walk the row) ---------------------------------
      if (tid < nb) {
        const int64_t q = start + tid;
        s_idx[tid] = perm_in ? (int)perm_in[(int64_t)ep * N + q] : (int)feistel_perm(fk, (uint64_t)q, (uint64_t)N);
      }
      __syncthreads();
      for (int r = warp; r < nb; r += PT / 32) {
        const float* src = rollout + (int64_t)s_idx[r] * A.rw;
        for (int c = lane; c < Do; c += 32) XN[r * xn_ld + c] = src[c];
        for (int c = lane; c < da_store; c += 32) MB[r * mb_ld + c] = src[col_act + c];
        if (lane == 0) {
          MB[r * mb_ld + da_store + 0] = src[col_logp];
          MB[r * mb_ld + da_store + 1] = src[col_adv];
          MB[r * mb_ld + da_store + 2] = src[col_ret];
        }
      }
      __syncthreads();
      // ---- 2. feature RunningNorm (train mode: update with this minibatch, then normalise) -----------------
      if (pd.has_norm) {
        if (tid < Do) {
          float s = 0.f;
          for (int r = 0; r < nb; ++r) s += XN[r * xn_ld + tid];
          const float bmean = s / (float)nb;
          float m2 = 0.f;
          for (int r = 0; r < nb; ++r) {
            const float d = XN[r * xn_ld + tid] - bmean;
            m2 = fmaf(d, d, m2);
          }
          const float bvar = m2 / (float)nb, bn = (float)nb, c = (float)run_count, tot = c + bn;
          const float delta = bmean - run_mean;
          run_mean += delta * bn / tot;
          run_var *= c;
          run_var += bvar * bn;
          run_var += delta * delta * c * bn / tot;
          run_var /= tot;
          nmean[tid] = run_mean;
          nistd[tid] = 1.0f / sqrtf(run_var + pd.norm_eps);
        }
        run_count += nb;
        __syncthreads();
        for (int i = tid; i < nb * Do; i += PT) {
          const int r = i / Do, k = i - r * Do;
          XN[r * xn_ld + k] = (XN[r * xn_ld + k] - nmean[k]) * nistd[k];
        }
      }
      // ---- 3. advantage normalisation: (A - mean) / (std_unbiased + 1e-8) ---------------------------------
      float adv_mean = 0.f, adv_istd = 1.f;
      if (A.hp.normalize_advantage && nb > 1) {
        const float a = (tid < nb) ? MB[tid * mb_ld + da_store + 1] : 0.f;
        const float s = block_sum(a, red);
        adv_mean = s / (float)nb;
        const float d = (tid < nb) ? a - adv_mean : 0.f;
        const float m2 = block_sum(d * d, red);
        adv_istd = 1.0f / (sqrtf(m2 / (float)(nb - 1)) + 1e-8f);
      }
      __syncthreads();

      // ---- 4. phase A: thread per row per tower ---------------------------------------------------------------
      float l_pg = 0.f, l_v = 0.f, l_ent = 0.f;
      const int net = tid >> 7;          // 0 = policy tower (threads 0..127), 1 = value tower (128..255)
      const int r = tid & 127;
      const float inv_nb = 1.0f / (float)nb;
      if (r < nb && r < MBR) {
        const float* W1t = net ? w1t_vf : w1t_pi;
        const float* W2t = net ? w2t_vf : w2t_pi;
        const float* b1 = Pm + (net ? pd.off_vf_b1 : pd.off_pi_b1);
        const float* b2 = Pm + (net ? pd.off_vf_b2 : pd.off_pi_b2);
        const float* W2 = Pm + (net ? pd.off_vf_w2 : pd.off_pi_w2);  // torch layout [j][i]
        const float* x = XN + r * xn_ld;
        float h1[HP], lat[HP];
#pragma unroll
        for (int j = 0; j < HP; ++j) h1[j] = (j < h) ? b1[j] : 0.f;
        for (int k = 0; k < Do; ++k) {
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
          lat[j] = (j < h) ? b2[j] : 0.f;
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

        float dlat[HP];
#pragma unroll
        for (int j = 0; j < HP; ++j) dlat[j] = 0.f;
        const float* mbr = MB + r * mb_ld;
        if (net == 1) {
          // value head + MSE
          const float* wv = Pm + pd.off_val_w;
          float v = Pm[pd.off_val_b];
#pragma unroll
          for (int j = 0; j < HP; ++j) v = (j < h) ? fmaf(wv[j], lat[j], v) : v;
          const float ret = mbr[da_store + 2];
          const float dv = v - ret;
          l_v = dv * dv;
          const float g = A.hp.vf_coef * 2.0f * dv * inv_nb;
          DVAL[r] = g;
#pragma unroll
          for (int j = 0; j < HP; ++j) dlat[j] = (j < h) ? g * wv[j] : 0.f;
        } else {
          const float* Wa = Pm + pd.off_act_w;  // [Da][h]
          const float* ba = Pm + pd.off_act_b;
          const float adv = (mbr[da_store + 1] - adv_mean) * adv_istd;
          const float logp_old = mbr[da_store + 0];
          float logp = 0.f, ent = 0.f;
          float* dmean = DMEAN + r * dm_ld;
          float* dls = DLS + r * dm_ld;
          if (!pd.discrete) {
            const float* lstd = Pm + pd.off_log_std;
            for (int a = 0; a < Da; ++a) {
              float m = ba[a];
#pragma unroll
              for (int j = 0; j < HP; ++j) m = (j < h) ? fmaf(Wa[a * h + j], lat[j], m) : m;
              const float ls = lstd[a], sd = expf(ls), var = sd * sd;
              const float diff = mbr[a] - m;
              logp += -(diff * diff) / (2.0f * var) - ls - 0.9189385332046727f;
              ent += 1.4189385332046727f + ls;
              dmean[a] = diff / var;               // d logp / d mean
              dls[a] = diff * diff / var - 1.0f;   // d logp / d log_std
            }
          } else {
            float mx = -INFINITY;
            for (int a = 0; a < Da; ++a) {
              float m = ba[a];
#pragma unroll
              for (int j = 0; j < HP; ++j) m = (j < h) ? fmaf(Wa[a * h + j], lat[j], m) : m;
              dmean[a] = m;
              mx = fmaxf(mx, m);
            }
            float se = 0.f;
            for (int a = 0; a < Da; ++a) se += expf(dmean[a] - mx);
            const float lse = mx + logf(se);
            const int act = (int)mbr[0];
            logp = dmean[act] - lse;
            for (int a = 0; a < Da; ++a) {
              const float lp = dmean[a] - lse, p = expf(lp);
              ent -= p * lp;
              dls[a] = lp;  // temporarily: log p_a
            }
          }
          const float ratio = expf(logp - logp_old);
          const float lo = 1.0f - A.hp.clip_range, hi = 1.0f + A.hp.clip_range;
          const float pl1 = adv * ratio, pl2 = adv * fminf(fmaxf(ratio, lo), hi);
          l_pg = -fminf(pl1, pl2);
          l_ent = -ent;
          const bool inside = (ratio >= lo) && (ratio <= hi);
          const float dl_dlogp = (inside || pl1 < pl2) ? -adv * ratio * inv_nb : 0.f;
          const float dent = -A.hp.ent_coef * inv_nb;  // d(ent_coef * ent_loss)/d(entropy)
          if (!pd.discrete) {
            for (int a = 0; a < Da; ++a) {
              dls[a] = dl_dlogp * dls[a] + dent;   // dH/dlog_std = 1
              dmean[a] = dl_dlogp * dmean[a];
            }
          } else {
            const int act = (int)mbr[0];
            for (int a = 0; a < Da; ++a) {
              const float lp = dls[a], p = expf(lp);
              const float dlogp_dl = ((a == act) ? 1.f : 0.f) - p;
              const float dent_dl = -p * (lp + ent);
              dmean[a] = dl_dlogp * dlogp_dl + dent * dent_dl;
              dls[a] = 0.f;
            }
          }
          for (int a = 0; a < Da; ++a) {
            const float g = dmean[a];
#pragma unroll
            for (int j = 0; j < HP; ++j) dlat[j] = (j < h) ? fmaf(g, Wa[a * h + j], dlat[j]) : 0.f;
          }
        }
        // backward through the tower (tanh)
        float* tH1 = T_H1[net] + r * LD;
        float* tLAT = T_LAT[net] + r * LD;
        float* tDZ2 = T_DZ2[net] + r * LD;
        float* tDZ1 = T_DZ1[net] + r * LD;
        float dh1[HP];
#pragma unroll
        for (int i = 0; i < HP; ++i) dh1[i] = 0.f;
#pragma unroll
        for (int j = 0; j < HP; ++j) {
          const float dz2 = dlat[j] * (1.0f - lat[j] * lat[j]);
          tDZ2[j] = dz2;
          tLAT[j] = lat[j];
          if (j < h) {
            const float* w = W2 + j * h;
#pragma unroll
            for (int i = 0; i < HP; ++i) dh1[i] = (i < h) ? fmaf(w[i], dz2, dh1[i]) : 0.f;
          }
        }
#pragma unroll
        for (int i = 0; i < HP; ++i) {
          tH1[i] = h1[i];
          tDZ1[i] = dh1[i] * (1.0f - h1[i] * h1[i]);
        }
      }
      // loss terms for logging / parity
      {
        const float s_pg = block_sum(l_pg, red);
        const float s_v = block_sum(l_v, red);
        const float s_ent = block_sum(l_ent, red);
        if (tid == 0 && loss_log) {
          const float pg = s_pg * inv_nb, vl = s_v * inv_nb, el = s_ent * inv_nb;
          loss_log[log_i * 4 + 0] = pg;
          loss_log[log_i * 4 + 1] = vl;
          loss_log[log_i * 4 + 2] = el;
          loss_log[log_i * 4 + 3] = pg + A.hp.ent_coef * el + A.hp.vf_coef * vl;
        }
        ++log_i;
      }
      __syncthreads();

      // ---- 5. phase B: weight gradients (warps 0-3 policy tower, 4-7 value tower) ---------------------------------
      {
        const int bnet = warp >> 2, bw = warp & 3;
        constexpr int JW = HP / 4;
        const int j0 = bw * JW;
        const float* tH1 = T_H1[bnet];
        const float* tLAT = T_LAT[bnet];
        const float* tDZ2 = T_DZ2[bnet];
        const float* tDZ1 = T_DZ1[bnet];
        const int off_w1 = bnet ? pd.off_vf_w1 : pd.off_pi_w1, off_b1 = bnet ? pd.off_vf_b1 : pd.off_pi_b1;
        const int off_w2 = bnet ? pd.off_vf_w2 : pd.off_pi_w2, off_b2 = bnet ? pd.off_vf_b2 : pd.off_pi_b2;
        // dW1 / db1
        {
          float acc[JW][2], bs[JW];
#pragma unroll
          for (int jj = 0; jj < JW; ++jj) acc[jj][0] = acc[jj][1] = bs[jj] = 0.f;
          const bool k0 = lane < Do, k1 = lane + 32 < Do;
          for (int rr = 0; rr < nb; ++rr) {
            const float a0 = k0 ? XN[rr * xn_ld + lane] : 0.f;
            const float a1 = k1 ? XN[rr * xn_ld + lane + 32] : 0.f;
#pragma unroll
            for (int jj = 0; jj < JW; ++jj) {
              const float dz = tDZ1[rr * LD + j0 + jj];
              acc[jj][0] = fmaf(dz, a0, acc[jj][0]);
              acc[jj][1] = fmaf(dz, a1, acc[jj][1]);
              bs[jj] += dz;
            }
          }
#pragma unroll
          for (int jj = 0; jj < JW; ++jj) {
            const int j = j0 + jj;
            if (j < h) {
              if (k0) G[off_w1 + j * Do + lane] = acc[jj][0];
              if (k1) G[off_w1 + j * Do + lane + 32] = acc[jj][1];
              if (lane == 0) G[off_b1 + j] = bs[jj];
            }
          }
        }
        // dW2 / db2
        {
          float acc[JW][HP / 32], bs[JW];
#pragma unroll
          for (int jj = 0; jj < JW; ++jj) {
#pragma unroll
            for (int ii = 0; ii < HP / 32; ++ii) acc[jj][ii] = 0.f;
            bs[jj] = 0.f;
          }
          for (int rr = 0; rr < nb; ++rr) {
            float a[HP / 32];
#pragma unroll
            for (int ii = 0; ii < HP / 32; ++ii) a[ii] = tH1[rr * LD + lane + 32 * ii];
#pragma unroll
            for (int jj = 0; jj < JW; ++jj) {
              const float dz = tDZ2[rr * LD + j0 + jj];
#pragma unroll
              for (int ii = 0; ii < HP / 32; ++ii) acc[jj][ii] = fmaf(dz, a[ii], acc[jj][ii]);
              bs[jj] += dz;
            }
          }
#pragma unroll
          for (int jj = 0; jj < JW; ++jj) {
            const int j = j0 + jj;
            if (j < h) {
#pragma unroll
              for (int ii = 0; ii < HP / 32; ++ii) {
                const int i = lane + 32 * ii;
                if (i < h) G[off_w2 + j * h + i] = acc[jj][ii];
              }
              if (lane == 0) G[off_b2 + j] = bs[jj];
            }
          }
        }
        // heads
        if (bnet == 0) {
          // dWa[a][i], dba[a], dlog_std[a]: work items (a, i) over the 128 policy-side threads
          const int t4 = tid;  // 0..127
          for (int w = t4; w < Da * h; w += 128) {
            const int a = w / h, i = w - a * h;
            float acc = 0.f;
            for (int rr = 0; rr < nb; ++rr) acc = fmaf(DMEAN[rr * dm_ld + a], tLAT[rr * LD + i], acc);
            G[pd.off_act_w + w] = acc;
          }
          for (int a = t4; a < Da; a += 128) {
            float acc = 0.f, accl = 0.f;
            for (int rr = 0; rr < nb; ++rr) {
              acc += DMEAN[rr * dm_ld + a];
              accl += DLS[rr * dm_ld + a];
            }
            G[pd.off_act_b + a] = acc;
            if (!pd.discrete) G[pd.off_log_std + a] = accl;
          }
        } else {
          const int t4 = tid - 128;
          for (int i = t4; i <= h; i += 128) {
            float acc = 0.f;
            if (i < h) {
              for (int rr = 0; rr < nb; ++rr) acc = fmaf(DVAL[rr], tLAT[rr * LD + i], acc);
              G[pd.off_val_w + i] = acc;
            } else {
              for (int rr = 0; rr < nb; ++rr) acc += DVAL[rr];
              G[pd.off_val_b] = acc;
            }
          }
        }
      }
      __syncthreads();

      // ---- 6. clip_grad_norm_ + Adam (SB3: eps 1e-5) ---------------------------------------------------------------
      float ss = 0.f;
      for (int i = tid; i < NP; i += PT) ss = fmaf(G[i], G[i], ss);
      const float total = sqrtf(block_sum(ss, red));
      float clip = A.hp.max_grad_norm / (total + 1e-6f);
      clip = clip > 1.0f ? 1.0f : clip;
      ++adam_step;
      if (tid == 0) {
        const double b1c = 1.0 - pow(0.9, (double)adam_step), b2c = 1.0 - pow(0.999, (double)adam_step);
        bc[0] = (float)((double)A.hp.lr / b1c);
        bc[1] = (float)sqrt(b2c);
      }
      __syncthreads();
      const float step_size = bc[0], bc2s = bc[1];
      for (int i = tid; i < NP; i += PT) {
        const float g = G[i] * clip;
        const float mi = Mm[i] + (g - Mm[i]) * (1.0f - 0.9f);
        const float vi = Vm[i] * 0.999f + (1.0f - 0.999f) * g * g;
        Mm[i] = mi;
        Vm[i] = vi;
        Pm[i] -= step_size * (mi / (sqrtf(vi) / bc2s + A.hp.adam_eps));
      }
      __syncthreads();
      build_images<HP>(pd, Pm, w1t_pi, w2t_pi, w1t_vf, w2t_vf);
      __syncthreads();
    }
  }

  // ---- write back ------------------------------------------------------------------------------------------------
  for (int i = tid; i < NP; i += PT) {
    g_params[i] = Pm[i];
    if (A.moments_in_smem) {
      g_m[i] = Mm[i];
      g_v[i] = Vm[i];
    }
  }
  if (pd.has_norm) {
    if (tid < Do) {
      g_norm[tid] = run_mean;
      g_norm[Do + tid] = run_var;
    }
    if (tid == 0) *g_norm_count = run_count;
  }
  if (tid == 0) {
    state[IMB_ST_PPO_STEP] = adam_step;
    state[IMB_ST_PPO_EPOCH] = perm_draw0 + A.hp.n_epochs;
  }
}

// ---- log pi(a|s) for the AIRL discriminator batch --------------------------------------------------------------
// thread per batch column; obs rows [0,Do), act rows [Do, Do+Da_onehot) of the feature-major batch.
template <int HP>
__global__ void __launch_bounds__(128) k_policy_logp(const imb_policy_desc pd, const float* __restrict__ params,
                                                    const float* __restrict__ norm, float* __restrict__ batch,
                                                    int64_t ld, int64_t n, int row_logp, int w1t_off, int w2t_off,
                                                    int xn_off, int xn_ld) {
  extern __shared__ __align__(128) float smem[];
  const int Do = pd.d_obs, Da = pd.d_act, h = pd.hidden;
  float* Pm = smem;
  float* w1t = smem + w1t_off;
  float* w2t = smem + w2t_off;
  float* XNs = smem + xn_off;
  const int tid = threadIdx.x;
  for (int i = tid; i < pd.n_params; i += blockDim.x) Pm[i] = params[i];
  for (int i = tid; i < Do * HP + HP * HP; i += blockDim.x) w1t[i] = 0.f;
  __syncthreads();
  for (int i = tid; i < h * Do; i += blockDim.x) {
    const int j = i / Do, k = i - j * Do;
    w1t[k * HP + j] = Pm[pd.off_pi_w1 + i];
  }
  for (int i = tid; i < h * h; i += blockDim.x) {
    const int j = i / h, ii = i - j * h;
    w2t[ii * HP + j] = Pm[pd.off_pi_w2 + i];
  }
  __syncthreads();
  float* x = XNs + tid * xn_ld;
  for (int64_t col = (int64_t)blockIdx.x * blockDim.x + tid; col < n; col += (int64_t)gridDim.x * blockDim.x) {
    for (int k = 0; k < Do; ++k) {
      float v = batch[(int64_t)k * ld + col];
      if (pd.has_norm) v = (v - norm[k]) / sqrtf(norm[Do + k] + pd.norm_eps);
      x[k] = v;
    }
    float h1[HP], lat[HP];
#pragma unroll
    for (int j = 0; j < HP; ++j) h1[j] = (j < h) ? Pm[pd.off_pi_b1 + j] : 0.f;
    for (int k = 0; k < Do; ++k) {
      const float xv = x[k];
#pragma unroll
      for (int j = 0; j < HP; ++j) h1[j] = fmaf(w1t[k * HP + j], xv, h1[j]);
    }
#pragma unroll
    for (int j = 0; j < HP; ++j) {
      h1[j] = tanhf(h1[j]);
      lat[j] = (j < h) ? Pm[pd.off_pi_b2 + j] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < HP; ++i) {
      const float hv = h1[i];
#pragma unroll
      for (int j = 0; j < HP; ++j) lat[j] = fmaf(w2t[i * HP + j], hv, lat[j]);
    }
#pragma unroll
    for (int j = 0; j < HP; ++j) lat[j] = tanhf(lat[j]);
    float logp = 0.f;
    if (!pd.discrete) {
      for (int a = 0; a < Da; ++a) {
        float m = Pm[pd.off_act_b + a];
#pragma unroll
        for (int j = 0; j < HP; ++j) m = (j < h) ? fmaf(Pm[pd.off_act_w + a * h + j], lat[j], m) : m;
        const float ls = Pm[pd.off_log_std + a], sd = expf(ls);
        const float diff = batch[(int64_t)(Do + a) * ld + col] - m;
        logp += -(diff * diff) / (2.0f * sd * sd) - ls - 0.9189385332046727f;
      }
    } else {
      float mx = -INFINITY, chosen = 0.f;
      float lg[IMB_MAX_DIN];
      for (int a = 0; a < Da; ++a) {
        float m = Pm[pd.off_act_b + a];
#pragma unroll
        for (int j = 0; j < HP; ++j) m = (j < h) ? fmaf(Pm[pd.off_act_w + a * h + j], lat[j], m) : m;
        lg[a] = m;
        mx = fmaxf(mx, m);
        if (batch[(int64_t)(Do + a) * ld + col] > 0.5f) chosen = m;  // one-hot action rows
      }
      float se = 0.f;
      for (int a = 0; a < Da; ++a) se += expf(lg[a] - mx);
      logp = chosen - (mx + logf(se));
    }
    batch[(int64_t)row_logp * ld + col] = logp;
  }
}

}  // namespace

template <int HP>
static size_t ppo_smem_floats(const imb_policy_desc& pd, int moments_in_smem, int MBR) {
  constexpr int LD = PpoCfg<HP>::LD;
  auto al = [](int x) { return (x + 31) / 32 * 32; };
  const int Do = pd.d_obs, Da = pd.d_act, NP = pd.n_params;
  const int da_store = pd.discrete ? 1 : Da;
  size_t o = 0;
  o += al(NP) * (size_t)(moments_in_smem ? 4 : 2);
  o += 2 * (size_t)(al(Do * HP) + al(HP * HP));
  o += al(MBR * (Do | 1));
  o += al(MBR * ((da_store + 3) | 1));
  o += (size_t)8 * MBR * LD;
  o += 2 * (size_t)al(MBR * (Da | 1));
  o += al(MBR);
  o += 2 * (size_t)al(Do);
  o += al(MBR);
  return o;
}

template <int HP>
static int launch_ppo(const PpoArgs& A0, float* params, float* norm, int32_t* norm_count, float* m, float* v,
                      const float* rollout, const int64_t* perm, float* loss_log, int64_t* state, cudaStream_t st) {
  PpoArgs A = A0;
  IMB_REQUIRE(A.hp.batch_size >= 1 && A.hp.batch_size <= PpoCfg<HP>::MBR,
              "PPO minibatch size must be in [1, %d] for tower width %d", PpoCfg<HP>::MBR, HP);
  A.mbr = (A.hp.batch_size + 31) / 32 * 32;
  A.moments_in_smem = 1;
  size_t fl = ppo_smem_floats<HP>(A.pol, 1, A.mbr);
  if (fl * 4 > IMB_SMEM_MAX) {
    A.moments_in_smem = 0;
    fl = ppo_smem_floats<HP>(A.pol, 0, A.mbr);
  }
  IMB_REQUIRE(fl * 4 <= IMB_SMEM_MAX, "PPO kernel needs %zu B of shared memory", fl * 4);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_ppo_update<HP>, cudaFuncAttributeMaxDynamicSharedMemorySize, IMB_SMEM_MAX);
    if (e != cudaSuccess) IMB_FAIL(-2, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  k_ppo_update<HP><<<1, PT, fl * 4, st>>>(A, params, norm, norm_count, m, v, rollout, perm, loss_log, state);
  IMB_CHECK_LAUNCH("k_ppo_update");
  return 0;
}

extern "C" int imb_rollout_row_width(const imb_policy_desc* pol);

extern "C" int imb_ppo_update(const imb_policy_desc* pol, float* pol_params, float* pol_norm,
                              int32_t* pol_norm_count, float* exp_avg, float* exp_avg_sq, const float* rollout,
                              int64_t n_rows, const imb_ppo_hparams* hp, const int64_t* perm, uint64_t seed,
                              float* loss_log, int64_t* state, void* stream) {
  IMB_REQUIRE(pol->hidden >= 1 && pol->hidden <= 64, "policy tower width must be <= 64");
  IMB_REQUIRE(pol->d_obs <= IMB_MAX_DIN && pol->d_act <= IMB_MAX_DIN, "d_obs/d_act must be <= %d", IMB_MAX_DIN);
  IMB_REQUIRE(n_rows >= 1 && n_rows < (1ll << 31), "bad n_rows");
  PpoArgs A;
  A.pol = *pol;
  A.hp = *hp;
  A.n_rows = n_rows;
  A.rw = imb_rollout_row_width(pol);
  A.seed = seed;
  A.moments_in_smem = 1;
  cudaStream_t st = (cudaStream_t)stream;
  if (pol->hidden <= 32)
    return launch_ppo<32>(A, pol_params, pol_norm, pol_norm_count, exp_avg, exp_avg_sq, rollout, perm, loss_log,
                          state, st);
  return launch_ppo<64>(A, pol_params, pol_norm, pol_norm_count, exp_avg, exp_avg_sq, rollout, perm, loss_log, state,
                        st);
}

template <int HP>
static int launch_logp(const imb_policy_desc* pol, const float* params, const float* norm, float* batch, int64_t ld,
                       int64_t n, int row_logp, cudaStream_t st) {
  auto al = [](int x) { return (x + 31) / 32 * 32; };
  const int Do = pol->d_obs;
  int o = al(pol->n_params);
  const int w1t_off = o;
  o += Do * HP;
  const int w2t_off = o;
  o += al(HP * HP);
  const int xn_ld = Do | 1;
  const int xn_off = al(o);
  o = xn_off + al(128 * xn_ld);
  const size_t bytes = (size_t)o * 4;
  IMB_REQUIRE(bytes <= IMB_SMEM_MAX, "policy too large");
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_policy_logp<HP>, cudaFuncAttributeMaxDynamicSharedMemorySize, IMB_SMEM_MAX);
    attr_set = true;
  }
  int64_t blocks = (n + 127) / 128;
  const int64_t cap = (int64_t)imb_num_sms() * 2;
  if (blocks > cap) blocks = cap;
  k_policy_logp<HP><<<(int)blocks, 128, bytes, st>>>(*pol, params, norm, batch, ld, n, row_logp, w1t_off, w2t_off,
                                                      xn_off, xn_ld);
  IMB_CHECK_LAUNCH("k_policy_logp");
  return 0;
}

extern "C" int imb_policy_logp(const imb_policy_desc* pol, const float* pol_params, const float* pol_norm,
                               float* batch, int64_t ld, int64_t n, int32_t row_logp, void* stream) {
  if (n <= 0) return 0;
  IMB_REQUIRE(pol->hidden >= 1 && pol->hidden <= 64, "policy tower width must be <= 64");
  if (pol->hidden <= 32) return launch_logp<32>(pol, pol_params, pol_norm, batch, ld, n, row_logp, (cudaStream_t)stream);
  return launch_logp<64>(pol, pol_params, pol_norm, batch, ld, n, row_logp, (cudaStream_t)stream);
}
