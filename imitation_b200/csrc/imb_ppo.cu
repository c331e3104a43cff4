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
#include "imb_tile.cuh"

namespace {

constexpr int PT = 256;   // threads: 0..127 = policy tower, 128..255 = value tower
constexpr int PR = 64;    // minibatch rows per step (SB3 batch_size <= 64)
constexpr int PRS = PR + TILE_PAD;

struct PpoArgs {
  imb_policy_desc pol;
  imb_ppo_hparams hp;
  int64_t n_rows;
  int rw;
  uint64_t seed;
  int moments_in_smem;
  int HP, KP, slices;  // tower width / obs width padded to 32; row slices of the weight-gradient phase
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

// working images rebuilt from the master parameters after every optimiser step; per tower:
//   W1t[KP][HP] (k-major for layer 1), W2t[HP][HP] (k = input unit), W2p[HP][HP] (torch layout, padded)
__device__ void build_images(const imb_policy_desc& pd, const float* __restrict__ Pm, float* __restrict__ img,
                             int HP, int KP) {
  const int Do = pd.d_obs, h = pd.hidden;
  const int tsz = KP * HP + 2 * HP * HP;
  for (int i = threadIdx.x; i < h * Do; i += PT) {
    const int j = i / Do, k = i - j * Do;
    img[k * HP + j] = Pm[pd.off_pi_w1 + i];
    img[tsz + k * HP + j] = Pm[pd.off_vf_w1 + i];
  }
  for (int i = threadIdx.x; i < h * h; i += PT) {
    const int j = i / h, ii = i - j * h;
    const float a = Pm[pd.off_pi_w2 + i], b = Pm[pd.off_vf_w2 + i];
    img[KP * HP + ii * HP + j] = a;
    img[KP * HP + HP * HP + j * HP + ii] = a;
    img[tsz + KP * HP + ii * HP + j] = b;
    img[tsz + KP * HP + HP * HP + j * HP + ii] = b;
  }
}

__global__ void __launch_bounds__(PT, 1) k_ppo_update(const PpoArgs A, float* __restrict__ g_params,
                                                      float* __restrict__ g_norm, int32_t* __restrict__ g_norm_count,
                                                      float* __restrict__ g_m, float* __restrict__ g_v,
                                                      const float* __restrict__ rollout,
                                                      const int64_t* __restrict__ perm_in,
                                                      float* __restrict__ loss_log, int64_t* __restrict__ state) {
  extern __shared__ __align__(128) float smem[];
  __shared__ float red[32];
  __shared__ float bc[8];
  const imb_policy_desc& pd = A.pol;
  const int Do = pd.d_obs, Da = pd.d_act, h = pd.hidden, NP = pd.n_params, HP = A.HP, KP = A.KP;
  const int da_store = pd.discrete ? 1 : Da;
  const int col_logp = Do + da_store, col_adv = col_logp + 3;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int net = tid >> 7, tt = tid & 127;  // tower, thread within tower
  auto al = [](int x) { return (x + 31) / 32 * 32; };

  // ---- shared-memory carve-up ---------------------------------------------------------------------------
  int o = 0;
  float* Pm = smem + o; o += al(NP);
  float* G = smem + o; o += al(NP);
  float* GS = smem + o; o += A.slices * al(NP);  // per-slice weight gradients (deterministic sum)
  float* Mm = g_m;
  float* Vm = g_v;
  if (A.moments_in_smem) {
    Mm = smem + o; o += al(NP);
    Vm = smem + o; o += al(NP);
  }
  const int tsz = KP * HP + 2 * HP * HP;
  float* img = smem + o; o += al(2 * tsz);
  float* XN = smem + o; o += al(KP * PRS);
  float* TH1 = smem + o; o += 2 * HP * PRS;   // [tower][HP][PRS]
  float* TLAT = smem + o; o += 2 * HP * PRS;
  float* TDZ2 = smem + o; o += 2 * HP * PRS;
  float* TDZ1 = smem + o; o += 2 * HP * PRS;
  const int DAP = (Da + 3) / 4 * 4;
  float* DM = smem + o; o += al(DAP * PRS);    // dL/d(mean or logits), feature-major [a][row]
  float* DLS = smem + o; o += al(DAP * PRS);   // dL/d(log_std) per row
  float* MBv = smem + o; o += al((DAP + 3) * PRS);  // act[DAP rows] | logp_old | adv | ret, feature-major
  float* DVAL = smem + o; o += al(PRS);
  float* rstat = smem + o; o += al(2 * 64 + 4);  // running mean[64] | var[64] of the feature norm
  int* s_idx = reinterpret_cast<int*>(smem + o); o += al(PR);
  float* H1 = TH1 + net * HP * PRS;
  float* LAT = TLAT + net * HP * PRS;
  float* DZ2 = TDZ2 + net * HP * PRS;
  float* DZ1 = TDZ1 + net * HP * PRS;
  const float* W1t = img + net * tsz;
  const float* W2t = W1t + KP * HP;
  const float* W2p = W2t + HP * HP;

  for (int i = tid; i < NP; i += PT) {
    Pm[i] = g_params[i];
    if (A.moments_in_smem) {
      Mm[i] = g_m[i];
      Vm[i] = g_v[i];
    }
  }
  for (int i = tid; i < al(2 * tsz); i += PT) img[i] = 0.f;
  for (int i = tid; i < KP * PRS; i += PT) XN[i] = 0.f;
  for (int i = tid; i < 8 * HP * PRS; i += PT) TH1[i] = 0.f;  // TH1..TDZ1 are contiguous
  for (int i = tid; i < (DAP + 3) * PRS; i += PT) MBv[i] = 0.f;
  for (int i = tid; i < 2 * DAP * PRS; i += PT) DM[i] = 0.f;  // DM, DLS contiguous (both al() sized)
  if (tid < 64) {
    rstat[tid] = (pd.has_norm && tid < Do) ? g_norm[tid] : 0.f;
    rstat[64 + tid] = (pd.has_norm && tid < Do) ? g_norm[Do + tid] : 1.f;
  }
  int32_t run_count = pd.has_norm ? *g_norm_count : 0;
  __syncthreads();
  build_images(pd, Pm, img, HP, KP);
  __syncthreads();

  const int64_t N = A.n_rows;
  const int mb = A.hp.batch_size;
  const int64_t steps_per_epoch = (N + mb - 1) / mb;
  int64_t adam_step = state[IMB_ST_PPO_STEP];
  const int64_t perm_draw0 = state[IMB_ST_PPO_EPOCH];
  int64_t log_i = 0;
  const int rwg = Do + da_store + 3;      // gathered columns per row: obs | act | logp_old | adv | ret
  const int r0 = (tt & 15) * 4;           // this thread's 4 tile rows
  const int cgq = tt >> 4;                // column quad within a 32-wide block (0..7)
  const int off_w1 = net ? pd.off_vf_w1 : pd.off_pi_w1, off_b1 = net ? pd.off_vf_b1 : pd.off_pi_b1;
  const int off_w2 = net ? pd.off_vf_w2 : pd.off_pi_w2, off_b2 = net ? pd.off_vf_b2 : pd.off_pi_b2;

  for (int ep = 0; ep < A.hp.n_epochs; ++ep) {
    const FeistelKey fk = feistel_key(A.seed, IMB_STREAM_PPO_PERM, (uint64_t)(perm_draw0 + ep), (uint64_t)N);
    for (int64_t sidx = 0; sidx < steps_per_epoch; ++sidx) {
      const int64_t start = sidx * mb;
      const int nb = (int)min((int64_t)mb, N - start);
      const float inv_nb = 1.0f / (float)nb;
      // ---- 1. gather: every thread fetches independent elements (one L2 latency for the whole tile) ------
      if (tid < PR)
        s_idx[tid] = tid < nb ? (perm_in ? (int)perm_in[(int64_t)ep * N + start + tid]
                                         : (int)feistel_perm(fk, (uint64_t)(start + tid), (uint64_t)N))
                              : 0;
      __syncthreads();
      for (int e = tid; e < PR * rwg; e += PT) {
        const int r = e / rwg, c = e - r * rwg;
        float v = 0.f;
        if (r < nb) {
          const int sc = c < col_logp ? c : (c == col_logp ? col_logp : col_adv + (c - col_logp - 1));
          v = rollout[(int64_t)s_idx[r] * A.rw + sc];
        }
        if (c < Do) XN[c * PRS + r] = v;
        else MBv[(c - Do + (c >= col_logp ? DAP - da_store : 0)) * PRS + r] = v;
      }
      __syncthreads();
      // ---- 2. feature RunningNorm: update with this minibatch, then normalise (train mode) ------------------
      if (pd.has_norm) {
        for (int k = warp; k < Do; k += PT / 32) {
          const float x0 = lane < nb ? XN[k * PRS + lane] : 0.f, x1 = lane + 32 < nb ? XN[k * PRS + lane + 32] : 0.f;
          const float bmean = warp_sum(x0 + x1) * inv_nb;
          const float d0 = lane < nb ? x0 - bmean : 0.f, d1 = lane + 32 < nb ? x1 - bmean : 0.f;
          const float bvar = warp_sum(d0 * d0 + d1 * d1) * inv_nb;
          if (lane == 0) {
            float mean = rstat[k], var = rstat[64 + k];
            const float bn = (float)nb, c = (float)run_count, tot = c + bn, delta = bmean - mean;
            mean += delta * bn / tot;
            var *= c;
            var += bvar * bn;
            var += delta * delta * c * bn / tot;
            var /= tot;
            rstat[k] = mean;
            rstat[64 + k] = var;
          }
        }
        run_count += nb;
        __syncthreads();
        for (int e = tid; e < Do * PR; e += PT) {
          const int k = e / PR, r = e - k * PR;
          XN[k * PRS + r] = r < nb ? (XN[k * PRS + r] - rstat[k]) / sqrtf(rstat[64 + k] + pd.norm_eps) : 0.f;
        }
      }
      // ---- 3. advantage normalisation (warp 0): (A - mean) / (std_unbiased + 1e-8) ----------------------------
      if (warp == 0) {
        float* adv = MBv + (DAP + 1) * PRS;
        float am = 0.f, ais = 1.f;
        if (A.hp.normalize_advantage && nb > 1) {
          const float a0 = lane < nb ? adv[lane] : 0.f, a1 = lane + 32 < nb ? adv[lane + 32] : 0.f;
          am = warp_sum(a0 + a1) * inv_nb;
          const float d0 = lane < nb ? a0 - am : 0.f, d1 = lane + 32 < nb ? a1 - am : 0.f;
          ais = 1.0f / (sqrtf(warp_sum(d0 * d0 + d1 * d1) / (float)(nb - 1)) + 1e-8f);
        }
        if (lane < nb) adv[lane] = (adv[lane] - am) * ais;
        if (lane + 32 < nb) adv[lane + 32] = (adv[lane + 32] - am) * ais;
      }
      __syncthreads();

      // ---- 4. forward: two tanh layers per tower as tiled GEMMs ---------------------------------------------
      {
        const float* b1 = Pm + off_b1;
        for (int jh = 0; jh < HP / 32; ++jh) {
          const int j0 = jh * 32 + cgq * 4;
          float acc[4][4] = {};
          gemm_acc44<false>(acc, XN, PRS, r0, W1t, HP, j0, Do);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float b = (j0 + t < h) ? b1[j0 + t] : 0.f;
            st4(H1 + (j0 + t) * PRS + r0, make_float4(tanhf(acc[0][t] + b), tanhf(acc[1][t] + b),
                                                      tanhf(acc[2][t] + b), tanhf(acc[3][t] + b)));
          }
        }
      }
      __syncthreads();
      {
        const float* b2 = Pm + off_b2;
        for (int jh = 0; jh < HP / 32; ++jh) {
          const int j0 = jh * 32 + cgq * 4;
          float acc[4][4] = {};
          gemm_acc44<false>(acc, H1, PRS, r0, W2t, HP, j0, h);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float b = (j0 + t < h) ? b2[j0 + t] : 0.f;
            st4(LAT + (j0 + t) * PRS + r0, make_float4(tanhf(acc[0][t] + b), tanhf(acc[1][t] + b),
                                                       tanhf(acc[2][t] + b), tanhf(acc[3][t] + b)));
          }
        }
      }
      __syncthreads();

      // ---- 5. heads + losses + dL/dlatent, thread per row (64 threads per tower) -----------------------------------
      float l_pg = 0.f, l_v = 0.f, l_ent = 0.f;
      if (tt < PR) {
        const int r = tt;
        const bool live = r < nb;
        if (net == 1) {
          const float* wv = Pm + pd.off_val_w;
          float v = Pm[pd.off_val_b];
          for (int j = 0; j < h; ++j) v = fmaf(wv[j], LAT[j * PRS + r], v);
          const float dv = v - MBv[(DAP + 2) * PRS + r];
          float g = 0.f;
          if (live) {
            l_v = dv * dv;
            g = A.hp.vf_coef * 2.0f * dv * inv_nb;
          }
          DVAL[r] = g;
          for (int j = 0; j < h; ++j) {
            const float l = LAT[j * PRS + r];
            DZ2[j * PRS + r] = g * wv[j] * (1.0f - l * l);
          }
        } else {
          const float* Wa = Pm + pd.off_act_w;  // [Da][h]
          const float* ba = Pm + pd.off_act_b;
          const float adv = MBv[(DAP + 1) * PRS + r], logp_old = MBv[DAP * PRS + r];
          float logp = 0.f, ent = 0.f;
          if (!pd.discrete) {
            const float* lstd = Pm + pd.off_log_std;
            for (int a = 0; a < Da; ++a) {
              float m = ba[a];
              for (int j = 0; j < h; ++j) m = fmaf(Wa[a * h + j], LAT[j * PRS + r], m);
              const float ls = lstd[a], sd = expf(ls), var = sd * sd;
              const float diff = MBv[a * PRS + r] - m;
              logp += -(diff * diff) / (2.0f * var) - ls - 0.9189385332046727f;
              ent += 1.4189385332046727f + ls;
              DM[a * PRS + r] = diff / var;              // d logp / d mean
              DLS[a * PRS + r] = diff * diff / var - 1.0f;  // d logp / d log_std
            }
          } else {
            float mx = -INFINITY;
            for (int a = 0; a < Da; ++a) {
              float m = ba[a];
              for (int j = 0; j < h; ++j) m = fmaf(Wa[a * h + j], LAT[j * PRS + r], m);
              DM[a * PRS + r] = m;
              mx = fmaxf(mx, m);
            }
            float se = 0.f;
            for (int a = 0; a < Da; ++a) se += expf(DM[a * PRS + r] - mx);
            const float lse = mx + logf(se);
            const int act = (int)MBv[r];
            for (int a = 0; a < Da; ++a) {
              const float lp = DM[a * PRS + r] - lse;
              if (a == act) logp = lp;
              ent -= expf(lp) * lp;
              DLS[a * PRS + r] = lp;  // temporarily: log p_a
            }
          }
          const float ratio = expf(logp - logp_old);
          const float lo = 1.0f - A.hp.clip_range, hi = 1.0f + A.hp.clip_range;
          const float pl1 = adv * ratio, pl2 = adv * fminf(fmaxf(ratio, lo), hi);
          const bool inside = (ratio >= lo) && (ratio <= hi);
          float dl_dlogp = (inside || pl1 < pl2) ? -adv * ratio * inv_nb : 0.f;
          float dent = -A.hp.ent_coef * inv_nb;  // d(ent_coef * ent_loss) / d(entropy)
          if (live) {
            l_pg = -fminf(pl1, pl2);
            l_ent = -ent;
          } else {
            dl_dlogp = 0.f;
            dent = 0.f;
          }
          if (!pd.discrete) {
            for (int a = 0; a < Da; ++a) {
              DLS[a * PRS + r] = dl_dlogp * DLS[a * PRS + r] + dent;  // dH/dlog_std = 1
              DM[a * PRS + r] = dl_dlogp * DM[a * PRS + r];
            }
          } else {
            const int act = (int)MBv[r];
            for (int a = 0; a < Da; ++a) {
              const float lp = DLS[a * PRS + r], p = expf(lp);
              DM[a * PRS + r] = dl_dlogp * (((a == act) ? 1.f : 0.f) - p) + dent * (-p * (lp + ent));
              DLS[a * PRS + r] = 0.f;
            }
          }
          for (int j = 0; j < h; ++j) {
            float dl = 0.f;
            for (int a = 0; a < Da; ++a) dl = fmaf(DM[a * PRS + r], Wa[a * h + j], dl);
            const float l = LAT[j * PRS + r];
            DZ2[j * PRS + r] = dl * (1.0f - l * l);
          }
        }
      }
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

      // ---- 6. backward through layer 2: DZ1 = (DZ2 . W2) * (1 - H1^2) -----------------------------------------
      for (int jh = 0; jh < HP / 32; ++jh) {
        const int i0 = jh * 32 + cgq * 4;
        float acc[4][4] = {};
        gemm_acc44<false>(acc, DZ2, PRS, r0, W2p, HP, i0, h);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float4 hh = ld4(H1 + (i0 + t) * PRS + r0);
          st4(DZ1 + (i0 + t) * PRS + r0, make_float4(acc[0][t] * (1.f - hh.x * hh.x), acc[1][t] * (1.f - hh.y * hh.y),
                                                     acc[2][t] * (1.f - hh.z * hh.z), acc[3][t] * (1.f - hh.w * hh.w)));
        }
      }
      __syncthreads();

      // ---- 7. weight gradients per tower (128 threads): per-slice buffers, summed in fixed order ----------------
      {
        const int jl = lane & 7, il = lane >> 3;
        const float sj0[4] = {0.f, 0.f, 0.f, 0.f};
        {  // dW2 / db2
          const int nblk = (HP / 32) * (HP / 32), ntl = nblk * 32;
          const int lt = tt % ntl, sl = tt / ntl, blk = lt >> 5;
          const int jb = (blk % (HP / 32)) * 32, ib = (blk / (HP / 32)) * 32;
          const int rows = PR / A.slices;
          float acc[4][8] = {}, bacc[4] = {};
          wgrad_acc<false>(acc, bacc, DZ2, H1, PRS, jb + jl, ib + il, sl * rows, (sl + 1) * rows, nullptr, sj0);
          float* Gs = GS + sl * al(NP);
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int j = jb + jl + 8 * jj;
            if (j < h) {
#pragma unroll
              for (int ii = 0; ii < 8; ++ii) {
                const int i = ib + il + 4 * ii;
                if (i < h) Gs[off_w2 + j * h + i] = acc[jj][ii];
              }
              if (il == 0 && ib == 0) Gs[off_b2 + j] = bacc[jj];
            }
          }
        }
        {  // dW1 / db1  (KP x HP blocks; with KP*HP/1024 blocks per tower the slice count may be smaller)
          const int nblk = (HP / 32) * (KP / 32), ntl = nblk * 32;
          const int sl_n = 128 / ntl < A.slices ? 128 / ntl : A.slices;
          const int lt = tt % ntl, sl = tt / ntl, blk = lt >> 5;
          const int jb = (blk % (HP / 32)) * 32, ib = (blk / (HP / 32)) * 32;
          if (sl < sl_n) {
            const int rows = PR / sl_n;
            float acc[4][8] = {}, bacc[4] = {};
            wgrad_acc<false>(acc, bacc, DZ1, XN, PRS, jb + jl, ib + il, sl * rows, (sl + 1) * rows, nullptr, sj0);
            float* Gs = GS + sl * al(NP);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const int j = jb + jl + 8 * jj;
              if (j < h) {
#pragma unroll
                for (int ii = 0; ii < 8; ++ii) {
                  const int k = ib + il + 4 * ii;
                  if (k < Do) Gs[off_w1 + j * Do + k] = acc[jj][ii];
                }
                if (il == 0 && ib == 0) Gs[off_b1 + j] = bacc[jj];
              }
            }
          }
          // slices that did not run this contraction must contribute zeros
          for (int s2 = sl_n; s2 < A.slices; ++s2) {
            float* Gz = GS + s2 * al(NP);
            for (int i = tt; i < h * Do + h; i += 128) Gz[(i < h * Do ? off_w1 + i : off_b1 + (i - h * Do))] = 0.f;
          }
        }
        // heads (written straight to G: one owner per element)
        if (net == 0) {
          for (int w = tt; w < Da * h; w += 128) {
            const int a = w / h, j = w - a * h;
            float acc = 0.f;
            for (int r = 0; r < PR; r += 4) {
              const float4 d = ld4(DM + a * PRS + r), l = ld4(LAT + j * PRS + r);
              acc = fmaf(d.x, l.x, acc);
              acc = fmaf(d.y, l.y, acc);
              acc = fmaf(d.z, l.z, acc);
              acc = fmaf(d.w, l.w, acc);
            }
            G[pd.off_act_w + w] = acc;
          }
          for (int a = tt; a < Da; a += 128) {
            float acc = 0.f, accl = 0.f;
            for (int r = 0; r < PR; ++r) {
              acc += DM[a * PRS + r];
              accl += DLS[a * PRS + r];
            }
            G[pd.off_act_b + a] = acc;
            if (!pd.discrete) G[pd.off_log_std + a] = accl;
          }
        } else {
          for (int j = tt; j <= h; j += 128) {
            float acc = 0.f;
            if (j < h) {
              for (int r = 0; r < PR; r += 4) {
                const float4 d = ld4(DVAL + r), l = ld4(LAT + j * PRS + r);
                acc = fmaf(d.x, l.x, acc);
                acc = fmaf(d.y, l.y, acc);
                acc = fmaf(d.z, l.z, acc);
                acc = fmaf(d.w, l.w, acc);
              }
              G[pd.off_val_w + j] = acc;
            } else {
              for (int r = 0; r < PR; ++r) acc += DVAL[r];
              G[pd.off_val_b] = acc;
            }
          }
        }
      }
      __syncthreads();

      // ---- 8. slice sum -> clip_grad_norm_ -> Adam (SB3: eps 1e-5) --------------------------------------------------
      const int n_tower = 2 * (h * Do + h + h * h + h);  // towers' W1,b1,W2,b2 occupy the head of the vector
      float ss = 0.f;
      for (int i = tid; i < NP; i += PT) {
        float g;
        if (i < n_tower) {
          g = GS[i];
          for (int s2 = 1; s2 < A.slices; ++s2) g += GS[s2 * al(NP) + i];
          G[i] = g;
        } else {
          g = G[i];
        }
        ss = fmaf(g, g, ss);
      }
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
      build_images(pd, Pm, img, HP, KP);
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
      g_norm[tid] = rstat[tid];
      g_norm[Do + tid] = rstat[64 + tid];
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

static size_t ppo_smem_floats(const PpoArgs& A) {
  auto al = [](int x) { return (x + 31) / 32 * 32; };
  const int NP = A.pol.n_params, HP = A.HP, KP = A.KP, Da = A.pol.d_act;
  const int DAP = (Da + 3) / 4 * 4;
  size_t o = 0;
  o += (size_t)al(NP) * (2 + A.slices + (A.moments_in_smem ? 2 : 0));
  o += al(2 * (KP * HP + 2 * HP * HP));
  o += al(KP * PRS);
  o += (size_t)8 * HP * PRS;
  o += 2 * (size_t)al(DAP * PRS);
  o += al((DAP + 3) * PRS);
  o += al(PRS);
  o += al(2 * 64 + 4);
  o += al(PR);
  return o;
}

static int launch_ppo(const PpoArgs& A0, float* params, float* norm, int32_t* norm_count, float* m, float* v,
                      const float* rollout, const int64_t* perm, float* loss_log, int64_t* state, cudaStream_t st) {
  PpoArgs A = A0;
  IMB_REQUIRE(A.hp.batch_size >= 1 && A.hp.batch_size <= PR, "PPO minibatch size must be in [1, %d]", PR);
  A.HP = A.pol.hidden <= 32 ? 32 : 64;
  A.KP = A.pol.d_obs <= 32 ? 32 : 64;
  const int ntl = (A.HP / 32) * (A.HP / 32) * 32;
  A.slices = 128 / ntl;  // 4 for 32-wide towers, 1 for 64-wide
  A.moments_in_smem = 1;
  size_t fl = ppo_smem_floats(A);
  if (fl * 4 > IMB_SMEM_MAX) {
    A.moments_in_smem = 0;
    fl = ppo_smem_floats(A);
  }
  IMB_REQUIRE(fl * 4 <= IMB_SMEM_MAX, "PPO kernel needs %zu B of shared memory", fl * 4);
  static size_t attr_bytes = 0;
  if (fl * 4 > attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(k_ppo_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(fl * 4));
    if (e != cudaSuccess) IMB_FAIL(-2, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_bytes = fl * 4;
  }
  k_ppo_update<<<1, PT, fl * 4, st>>>(A, params, norm, norm_count, m, v, rollout, perm, loss_log, state);
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
  return launch_ppo(A, pol_params, pol_norm, pol_norm_count, exp_avg, exp_avg_sq, rollout, perm, loss_log, state,
                    (cudaStream_t)stream);
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
