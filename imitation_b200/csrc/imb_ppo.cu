// imb_ppo.cu -- the generator update (PPO) as ONE persistent thread-block-cluster launch per round.
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
#include <cooperative_groups.h>

#include "imb_common.cuh"
#include "imb_tile.cuh"

namespace {

constexpr int PT = 256;   // threads per CTA: 0..127 = policy tower, 128..255 = value tower
constexpr int CL = 8;     // CTAs per cluster: each owns RL rows of every minibatch and 1/CL of the parameters
constexpr int RL = 8;     // minibatch rows per CTA  (CL * RL = 64 >= SB3 batch_size)
constexpr int PR = CL * RL;
constexpr int PRS = PR + TILE_PAD;

struct PpoArgs {
  imb_policy_desc pol;
  imb_ppo_hparams hp;
  int64_t n_rows;
  int rw;
  uint64_t seed;
  int HP, KP, S;  // tower width / obs width padded to 32; parameters per slice (multiple of 4)
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

// working images rebuilt from the parameters after every optimiser step; per tower:
//   W1t[KP][HP] (k-major for layer 1), W2t[HP][HP] (k = input unit); the backward pass reads W2 in its
//   torch layout straight from the parameter vector (row stride h, consecutive lanes = consecutive floats)
__device__ void build_images(const imb_policy_desc& pd, const float* __restrict__ Pm, float* __restrict__ img,
                             int HP, int KP) {
  const int Do = pd.d_obs, h = pd.hidden;
  const int tsz = KP * HP + HP * HP;
  for (int i = threadIdx.x; i < h * Do; i += PT) {
    const int j = i / Do, k = i - j * Do;
    img[k * HP + j] = Pm[pd.off_pi_w1 + i];
    img[tsz + k * HP + j] = Pm[pd.off_vf_w1 + i];
  }
  for (int i = threadIdx.x; i < h * h; i += PT) {
    const int j = i / h, ii = i - j * h;
    img[KP * HP + ii * HP + j] = Pm[pd.off_pi_w2 + i];
    img[tsz + KP * HP + ii * HP + j] = Pm[pd.off_vf_w2 + i];
  }
}

__device__ __forceinline__ float dot8(const float* __restrict__ a, const float* __restrict__ b) {
  const float4 a0 = ld4(a), a1 = ld4(a + 4), b0 = ld4(b), b1 = ld4(b + 4);
  float s0 = a0.x * b0.x, s1 = a0.y * b0.y;
  s0 = fmaf(a0.z, b0.z, s0);
  s1 = fmaf(a0.w, b0.w, s1);
  s0 = fmaf(a1.x, b1.x, s0);
  s1 = fmaf(a1.y, b1.y, s1);
  s0 = fmaf(a1.z, b1.z, s0);
  s1 = fmaf(a1.w, b1.w, s1);
  return s0 + s1;
}

// PPO.train for one rollout: n_epochs x ceil(N / batch) optimiser steps, ONE cluster of CL CTAs.
// Every CTA gathers the whole minibatch (64 x ~26 floats) so that the feature RunningNorm and the
// advantage normalisation are computed redundantly and identically everywhere; forward/backward
// run on the CTA's own RL rows; the per-CTA partial gradients are exchanged through distributed
// shared memory: slice owners sum the CL partials in fixed order, the squared norms of the slices
// are exchanged for clip_grad_norm_, owners run Adam on their slice and push the new parameters to
// all CTAs.  Three cluster barriers per optimiser step, no global-memory traffic inside a step
// except the minibatch gather.
__global__ void __launch_bounds__(PT, 1) k_ppo_update(const PpoArgs A, float* __restrict__ g_params,
                                                      float* __restrict__ g_norm, int32_t* __restrict__ g_norm_count,
                                                      float* __restrict__ g_m, float* __restrict__ g_v,
                                                      const float* __restrict__ rollout,
                                                      const int64_t* __restrict__ perm_in,
                                                      float* __restrict__ loss_log, int64_t* __restrict__ state) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  extern __shared__ __align__(128) float smem[];
  __shared__ float red[32];
  __shared__ float bc[8];
  const imb_policy_desc& pd = A.pol;
  const int Do = pd.d_obs, Da = pd.d_act, h = pd.hidden, NP = pd.n_params, HP = A.HP, KP = A.KP, S = A.S;
  const int da_store = pd.discrete ? 1 : Da;
  const int col_logp = Do + da_store, col_adv = col_logp + 3;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int net = tid >> 7, tt = tid & 127;  // tower, thread within tower
  auto al = [](int x) { return (x + 31) / 32 * 32; };

  // ---- shared-memory carve-up (identical in every CTA: DSMEM addresses are rank + offset) ---------------
  int o = 0;
  float* Pm = smem + o; o += al(CL * S);        // full parameter vector (padded to CL slices)
  float* Ms = smem + o; o += al(S);             // Adam moments of the owned slice
  float* Vs = smem + o; o += al(S);
  float* GSL = smem + o; o += al(S);            // summed gradient of the owned slice
  float* RECV = smem + o; o += al(CL * S);      // [source CTA][S]: partial gradients pushed by every CTA
  float* SSQ = smem + o; o += 32;               // [CL] squared gradient norms of the slices
  float* LOSS = smem + o; o += 32;              // [CL][3] partial loss sums (read by CTA 0)
  const int tsz = KP * HP + HP * HP;
  float* img = smem + o; o += al(2 * tsz);
  float* XNf = smem + o; o += al(KP * PRS);     // full minibatch, feature-major
  const int DAP = (Da + 3) / 4 * 4;
  float* MBf = smem + o; o += al((DAP + 3) * PRS);  // act[DAP] | logp_old | adv | ret, full minibatch
  // own-row tiles, feature-major [feature][RL]
  float* XNo = smem + o; o += al(KP * RL);
  float* TH1 = smem + o; o += 2 * HP * RL;
  float* TLAT = smem + o; o += 2 * HP * RL;
  float* TDZ2 = smem + o; o += 2 * HP * RL;
  float* TDZ1 = smem + o; o += 2 * HP * RL;
  float* DM = smem + o; o += DAP * RL;          // dL/d(mean|logits) [a][RL]
  float* DLS = smem + o; o += DAP * RL;         // dL/d(log_std) per row [a][RL]
  float* MEAN = smem + o; o += DAP * RL;        // action means / logits [a][RL]
  float* DVAL = smem + o; o += 32;              // [RL] dL/dvalue ; VALS at +8 ; ONES at +16
  float* VALS = DVAL + 8;
  float* ONES = DVAL + 16;
  float* rstat = smem + o; o += al(2 * 64 + 4);
  int* s_idx = reinterpret_cast<int*>(smem + o); o += al(PR);
  unsigned short* offA = reinterpret_cast<unsigned short*>(smem + o); o += al((CL * S + 1) / 2);
  unsigned short* offB = reinterpret_cast<unsigned short*>(smem + o); o += al((CL * S + 1) / 2);
  unsigned short* imgpos = reinterpret_cast<unsigned short*>(smem + o); o += al((CL * S + 1) / 2);  // image slot of p
  unsigned int* glut = reinterpret_cast<unsigned int*>(smem + o); o += al(PR * (Do + da_store + 3));  // gather LUT
  float* H1 = TH1 + net * HP * RL;
  float* LAT = TLAT + net * HP * RL;
  float* DZ2 = TDZ2 + net * HP * RL;
  float* DZ1 = TDZ1 + net * HP * RL;
  const float* W1t = img + net * tsz;
  const float* W2t = W1t + KP * HP;

  for (int i = tid; i < CL * S; i += PT) {
    Pm[i] = i < NP ? g_params[i] : 0.f;
    RECV[i] = 0.f;
  }
  for (int i = tid; i < S; i += PT) {
    const int p = crank * S + i;
    Ms[i] = p < NP ? g_m[p] : 0.f;
    Vs[i] = p < NP ? g_v[p] : 0.f;
  }
  for (int i = tid; i < al(2 * tsz); i += PT) img[i] = 0.f;
  for (int i = tid; i < KP * PRS; i += PT) XNf[i] = 0.f;
  for (int i = tid; i < (DAP + 3) * PRS; i += PT) MBf[i] = 0.f;
  for (int i = tid; i < KP * RL; i += PT) XNo[i] = 0.f;
  for (int i = tid; i < 8 * HP * RL; i += PT) TH1[i] = 0.f;   // TH1..TDZ1 contiguous
  for (int i = tid; i < 3 * DAP * RL + 32; i += PT) DM[i] = 0.f;  // DM, DLS, MEAN, DVAL.. contiguous
  if (tid < 64) {
    rstat[tid] = (pd.has_norm && tid < Do) ? g_norm[tid] : 0.f;
    rstat[64 + tid] = (pd.has_norm && tid < Do) ? g_norm[Do + tid] : 1.f;
    SSQ[tid & 31] = 0.f;
    LOSS[tid & 31] = 0.f;
  }
  __syncthreads();
  if (tid < RL) ONES[tid] = 1.f;
  // gradient of parameter p = dot over the CTA's RL rows of two feature rows: offA[p], offB[p]
  // (float offsets into this CTA's shared memory); biases / log_std pair with ONES.
  for (int p = tid; p < CL * S; p += PT) {
    int a = (int)(ONES - smem), b2 = a;  // default: harmless
    if (p < NP) {
      auto tower = [&](int q, int tw) -> bool {  // q relative to the tower's first parameter
        const int w1 = h * Do, b1 = w1 + h, w2 = b1 + h * h, bb2 = w2 + h;
        const float* tDZ1 = TDZ1 + tw * HP * RL, *tDZ2 = TDZ2 + tw * HP * RL, *tH1 = TH1 + tw * HP * RL;
        if (q < w1) { a = (int)(tDZ1 - smem) + (q / Do) * RL; b2 = (int)(XNo - smem) + (q % Do) * RL; return true; }
        if (q < b1) { a = (int)(tDZ1 - smem) + (q - w1) * RL; b2 = (int)(ONES - smem); return true; }
        if (q < w2) { const int r = q - b1; a = (int)(tDZ2 - smem) + (r / h) * RL; b2 = (int)(tH1 - smem) + (r % h) * RL; return true; }
        if (q < bb2) { a = (int)(tDZ2 - smem) + (q - w2) * RL; b2 = (int)(ONES - smem); return true; }
        return false;
      };
      const int tp = h * Do + h + h * h + h;
      if (p < tp) tower(p, 0);
      else if (p < 2 * tp) tower(p - tp, 1);
      else if (p < pd.off_act_b) { const int r = p - pd.off_act_w; a = (int)(DM - smem) + (r / h) * RL; b2 = (int)(TLAT - smem) + (r % h) * RL; }
      else if (p < pd.off_val_w) { a = (int)(DM - smem) + (p - pd.off_act_b) * RL; b2 = (int)(ONES - smem); }
      else if (p < pd.off_val_b) { a = (int)(DVAL - smem); b2 = (int)(TLAT + HP * RL - smem) + (p - pd.off_val_w) * RL; }
      else if (p == pd.off_val_b) { a = (int)(DVAL - smem); b2 = (int)(ONES - smem); }
      else { a = (int)(DLS - smem) + (p - pd.off_log_std) * RL; b2 = (int)(ONES - smem); }
    }
    offA[p] = (unsigned short)(a / 4);   // all rows are 16-byte aligned: store offset / 4
    offB[p] = (unsigned short)(b2 / 4);
  }
  // image slot (float offset into img, 0xFFFF = none) of every parameter: W1 -> W1t[k][j], W2 -> W2t[i][j]
  for (int p = tid; p < CL * S; p += PT) {
    int pos = 0xFFFF;
    const int tp = h * Do + h + h * h + h;
    if (p < 2 * tp) {
      const int tw = p / tp, q = p - tw * tp, w1 = h * Do, b1 = w1 + h, w2 = b1 + h * h;
      if (q < w1) pos = tw * tsz + (q % Do) * HP + q / Do;
      else if (q >= b1 && q < w2) { const int r = q - b1; pos = tw * tsz + KP * HP + (r % h) * HP + r / h; }
    }
    imgpos[p] = (unsigned short)pos;
  }
  // gather LUT: element e of the minibatch tile -> (row r | source column sc << 8 | dst float offset << 16)
  const int rwg0 = Do + da_store + 3;
  for (int e = tid; e < PR * rwg0; e += PT) {
    const int r = e / rwg0, c = e - r * rwg0;
    const int sc = c < col_logp ? c : (c == col_logp ? col_logp : col_adv + (c - col_logp - 1));
    const int dst = c < Do ? (int)(XNf - smem) + c * PRS + r
                           : (int)(MBf - smem) + (c - Do + (c >= col_logp ? DAP - da_store : 0)) * PRS + r;
    glut[e] = (unsigned)r | ((unsigned)sc << 8) | ((unsigned)dst << 16);
  }
  int32_t run_count = pd.has_norm ? *g_norm_count : 0;
  __syncthreads();
  build_images(pd, Pm, img, HP, KP);
  cluster.sync();

  const int64_t N = A.n_rows;
  const int mb = A.hp.batch_size;
  const int64_t steps_per_epoch = (N + mb - 1) / mb;
  int64_t adam_step = state[IMB_ST_PPO_STEP];
  const int64_t perm_draw0 = state[IMB_ST_PPO_EPOCH];
  double b1pow = pow(0.9, (double)adam_step), b2pow = pow(0.999, (double)adam_step);  // beta^t, kept incrementally
  int64_t log_i = 0;
  const int rwg = Do + da_store + 3;  // gathered columns per row: obs | act | logp_old | adv | ret
  const int row0 = crank * RL;        // first minibatch row owned by this CTA
  // own-row GEMM mapping: thread -> column j = tt % HP and RPT = HP / 16 consecutive rows
  const int gj = tt % HP, RPT = HP / 16, gr0 = (tt / HP) * RPT;

  for (int ep = 0; ep < A.hp.n_epochs; ++ep) {
    const FeistelKey fk = feistel_key(A.seed, IMB_STREAM_PPO_PERM, (uint64_t)(perm_draw0 + ep), (uint64_t)N);
    for (int64_t sidx = 0; sidx < steps_per_epoch; ++sidx) {
      const int64_t start = sidx * mb;
      const int nb = (int)min((int64_t)mb, N - start);
      const float inv_nb = 1.0f / (float)nb;
      // ---- 1. gather the whole minibatch (independent loads, 4 in flight per thread) ------------------------
      if (tid < PR)
        s_idx[tid] = tid < nb ? (perm_in ? (int)perm_in[(int64_t)ep * N + start + tid]
                                         : (int)feistel_perm(fk, (uint64_t)(start + tid), (uint64_t)N))
                              : 0;
      __syncthreads();
      for (int e0 = tid; e0 < PR * rwg; e0 += 8 * PT) {  // up to 8 independent loads in flight per thread
        float v[8];
        unsigned lut[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int e = e0 + u * PT;
          lut[u] = e < PR * rwg ? glut[e] : 0xFFFFFFFFu;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          v[u] = 0.f;
          if (lut[u] != 0xFFFFFFFFu) {
            const int r = lut[u] & 0xFF;
            if (r < nb) v[u] = rollout[(int64_t)s_idx[r] * A.rw + ((lut[u] >> 8) & 0xFF)];
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (lut[u] != 0xFFFFFFFFu) smem[lut[u] >> 16] = v[u];
      }
      __syncthreads();
      // ---- 2. feature RunningNorm over the whole minibatch (identical in every CTA) -----------------------------
      if (pd.has_norm) {
        for (int k = warp; k < Do; k += PT / 32) {
          const float x0 = lane < nb ? XNf[k * PRS + lane] : 0.f, x1 = lane + 32 < nb ? XNf[k * PRS + lane + 32] : 0.f;
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
      }
      // own rows, normalised, feature-major [k][RL]
      for (int e = tid; e < Do * RL; e += PT) {
        const int k = e / RL, r = e - k * RL;
        const float x = XNf[k * PRS + row0 + r];
        XNo[e] = (row0 + r < nb) ? (pd.has_norm ? (x - rstat[k]) / sqrtf(rstat[64 + k] + pd.norm_eps) : x) : 0.f;
      }
      // ---- 3. advantage normalisation over the whole minibatch (warp 0, identical in every CTA) ----------------
      if (warp == 0) {
        float* adv = MBf + (DAP + 1) * PRS;
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

      // ---- 4. forward on the own rows: thread = (tower, column gj, RPT rows) -----------------------------------------
      auto own_gemm = [&](const float* __restrict__ Ain, const float* __restrict__ Wk, int wld, int K,
                          float (&acc)[4]) {
        acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
          const float w = Wk[k * wld + gj];
          const float* ar = Ain + k * RL + gr0;
          if (RPT == 2) {
            const float2 a = *reinterpret_cast<const float2*>(ar);
            acc[0] = fmaf(a.x, w, acc[0]);
            acc[1] = fmaf(a.y, w, acc[1]);
          } else {
            const float4 a = ld4(ar);
            acc[0] = fmaf(a.x, w, acc[0]);
            acc[1] = fmaf(a.y, w, acc[1]);
            acc[2] = fmaf(a.z, w, acc[2]);
            acc[3] = fmaf(a.w, w, acc[3]);
          }
        }
      };
      {
        float acc[4];
        own_gemm(XNo, W1t, HP, Do, acc);
        const float b = gj < h ? Pm[(net ? pd.off_vf_b1 : pd.off_pi_b1) + gj] : 0.f;
        for (int x = 0; x < RPT; ++x) H1[gj * RL + gr0 + x] = tanhf(acc[x] + b);
      }
      __syncthreads();
      {
        float acc[4];
        own_gemm(H1, W2t, HP, h, acc);
        const float b = gj < h ? Pm[(net ? pd.off_vf_b2 : pd.off_pi_b2) + gj] : 0.f;
        for (int x = 0; x < RPT; ++x) LAT[gj * RL + gr0 + x] = tanhf(acc[x] + b);
      }
      __syncthreads();

      // ---- 5. heads -----------------------------------------------------------------------------------------------------
      // (i) action means / logits: thread (a, r) ; value: 32 threads = (r, quarter of the latent)
      if (net == 0) {
        const float* Wa = Pm + pd.off_act_w;
        for (int w = tt; w < Da * RL; w += 128) {
          const int a = w / RL, r = w - a * RL;
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
          int j = 0;
          for (; j + 4 <= h; j += 4) {
            s0 = fmaf(Wa[a * h + j + 0], LAT[(j + 0) * RL + r], s0);
            s1 = fmaf(Wa[a * h + j + 1], LAT[(j + 1) * RL + r], s1);
            s2 = fmaf(Wa[a * h + j + 2], LAT[(j + 2) * RL + r], s2);
            s3 = fmaf(Wa[a * h + j + 3], LAT[(j + 3) * RL + r], s3);
          }
          for (; j < h; ++j) s0 = fmaf(Wa[a * h + j], LAT[j * RL + r], s0);
          MEAN[a * RL + r] = Pm[pd.off_act_b + a] + ((s0 + s1) + (s2 + s3));
        }
      } else if (tt < 4 * RL) {
        const float* wv = Pm + pd.off_val_w;
        const int r = tt >> 2, q = tt & 3;
        float s = 0.f;
        for (int j = q; j < h; j += 4) s = fmaf(wv[j], LAT[j * RL + r], s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (q == 0) VALS[r] = s + Pm[pd.off_val_b];
      }
      __syncthreads();
      // (ii) per-row loss terms and dL/d(head outputs): RL threads per tower
      float l_pg = 0.f, l_v = 0.f, l_ent = 0.f;
      if (tt < RL) {
        const int r = tt, gr = row0 + r;
        const bool live = gr < nb;
        if (net == 1) {
          const float dv = VALS[r] - MBf[(DAP + 2) * PRS + gr];
          float g = 0.f;
          if (live) {
            l_v = dv * dv;
            g = A.hp.vf_coef * 2.0f * dv * inv_nb;
          }
          DVAL[r] = g;
        } else {
          const float adv = MBf[(DAP + 1) * PRS + gr], logp_old = MBf[DAP * PRS + gr];
          float logp = 0.f, ent = 0.f;
          if (!pd.discrete) {
            const float* lstd = Pm + pd.off_log_std;
            for (int a = 0; a < Da; ++a) {
              const float ls = lstd[a], sd = expf(ls), var = sd * sd;
              const float diff = MBf[a * PRS + gr] - MEAN[a * RL + r];
              logp += -(diff * diff) / (2.0f * var) - ls - 0.9189385332046727f;
              ent += 1.4189385332046727f + ls;
              DM[a * RL + r] = diff / var;                 // d logp / d mean
              DLS[a * RL + r] = diff * diff / var - 1.0f;  // d logp / d log_std
            }
          } else {
            float mx = -INFINITY;
            for (int a = 0; a < Da; ++a) mx = fmaxf(mx, MEAN[a * RL + r]);
            float se = 0.f;
            for (int a = 0; a < Da; ++a) se += expf(MEAN[a * RL + r] - mx);
            const float lse = mx + logf(se);
            const int act = (int)MBf[gr];
            for (int a = 0; a < Da; ++a) {
              const float lp = MEAN[a * RL + r] - lse;
              if (a == act) logp = lp;
              ent -= expf(lp) * lp;
              DLS[a * RL + r] = lp;  // temporarily: log p_a
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
              DLS[a * RL + r] = dl_dlogp * DLS[a * RL + r] + dent;  // dH/dlog_std = 1
              DM[a * RL + r] = dl_dlogp * DM[a * RL + r];
            }
          } else {
            const int act = (int)MBf[gr];
            for (int a = 0; a < Da; ++a) {
              const float lp = DLS[a * RL + r], pp = expf(lp);
              DM[a * RL + r] = dl_dlogp * (((a == act) ? 1.f : 0.f) - pp) + dent * (-pp * (lp + ent));
              DLS[a * RL + r] = 0.f;
            }
          }
        }
      }
      // partial loss sums of this CTA -> CTA 0 (distributed shared memory)
      if (loss_log) {  // (uniform) loss terms are only reduced when the caller asked for the log
        const float s_pg = block_sum(l_pg, red);
        const float s_v = block_sum(l_v, red);
        const float s_ent = block_sum(l_ent, red);
        if (tid == 0) {
          float* L0 = cluster.map_shared_rank(LOSS, 0);
          L0[crank * 3 + 0] = s_pg;
          L0[crank * 3 + 1] = s_v;
          L0[crank * 3 + 2] = s_ent;
        }
      }
      __syncthreads();
      // (iii) dL/dz2 = (dL/dlatent) * (1 - lat^2): thread = (tower, column gj, RPT rows)
      for (int x = 0; x < RPT; ++x) {
        const int r = gr0 + x;
        float dl = 0.f;
        if (gj < h) {
          if (net == 1) {
            dl = DVAL[r] * Pm[pd.off_val_w + gj];
          } else {
            const float* Wa = Pm + pd.off_act_w;
            for (int a = 0; a < Da; ++a) dl = fmaf(DM[a * RL + r], Wa[a * h + gj], dl);
          }
        }
        const float l = LAT[gj * RL + r];
        DZ2[gj * RL + r] = dl * (1.0f - l * l);
      }
      __syncthreads();
      // ---- 6. backward through layer 2 --------------------------------------------------------------------------------------
      {
        float acc[4];
        own_gemm(DZ2, Pm + (net ? pd.off_vf_w2 : pd.off_pi_w2), h, h, acc);  // W2[j][i]: k = j, column = i
        for (int x = 0; x < RPT; ++x) {
          const float hh = H1[gj * RL + gr0 + x];
          DZ1[gj * RL + gr0 + x] = gj < h ? acc[x] * (1.f - hh * hh) : 0.f;
        }
      }
      __syncthreads();
      // ---- 7. partial gradient of every parameter: one dot product over the RL own rows ----------------------------
      // ... pushed straight into the owner's receive buffer RECV[this CTA][i]: one 16-byte DSMEM store per
      // parameter quad (scalar remote accesses are transaction-bound: ~10 k of them per step cost more
      // than all the arithmetic; see profiles/r01_summary.md)
      // CTA c starts with the quads owned by CTA c+1, so at any time the 8 senders target 8 different
      // receivers (all of them hitting owner 0 first made the receiving SM the bottleneck: ~40 % of the step)
      for (int q = tid; q < CL * S / 4; q += PT) {
        int qq = q + ((crank + 1) & (CL - 1)) * (S / 4);
        if (qq >= CL * S / 4) qq -= CL * S / 4;
        const int p0 = 4 * qq;
        if (p0 >= NP) continue;
        float4 g;
        g.x = dot8(smem + 4 * (int)offA[p0 + 0], smem + 4 * (int)offB[p0 + 0]);
        g.y = dot8(smem + 4 * (int)offA[p0 + 1], smem + 4 * (int)offB[p0 + 1]);
        g.z = dot8(smem + 4 * (int)offA[p0 + 2], smem + 4 * (int)offB[p0 + 2]);
        g.w = dot8(smem + 4 * (int)offA[p0 + 3], smem + 4 * (int)offB[p0 + 3]);
        const int owner = p0 / S;
        st4(cluster.map_shared_rank(RECV, owner) + crank * S + (p0 - owner * S), g);
      }
      cluster.sync();  // (a) all partial gradients (and partial losses) have landed at their owners

      // ---- 8. slice owners: sum the CL partials in fixed order, exchange squared norms ----------------------------------
      float ss = 0.f;
      {
        for (int i = tid; i < S; i += PT) {
          const int p = crank * S + i;
          float g = 0.f;
          if (p < NP) {
#pragma unroll
            for (int c = 0; c < CL; ++c) g += RECV[c * S + i];  // fixed order: deterministic
          }
          GSL[i] = g;
          ss = fmaf(g, g, ss);
        }
      }
      const float my_ssq = block_sum(ss, red);
      if (tid < CL) cluster.map_shared_rank(SSQ, tid)[crank] = my_ssq;
      if (crank == 0 && tid == 0 && loss_log) {
        float pg = 0.f, vl = 0.f, el = 0.f;
        for (int c = 0; c < CL; ++c) {
          pg += LOSS[c * 3 + 0];
          vl += LOSS[c * 3 + 1];
          el += LOSS[c * 3 + 2];
        }
        pg *= inv_nb, vl *= inv_nb, el *= inv_nb;
        loss_log[log_i * 4 + 0] = pg;
        loss_log[log_i * 4 + 1] = vl;
        loss_log[log_i * 4 + 2] = el;
        loss_log[log_i * 4 + 3] = pg + A.hp.ent_coef * el + A.hp.vf_coef * vl;
      }
      ++log_i;
      cluster.sync();  // (b) all slice norms are in every CTA's SSQ

      // ---- 9. clip_grad_norm_ + Adam on the owned slice; push the new parameters to every CTA ----------------------------
      float total = 0.f;
#pragma unroll
      for (int c = 0; c < CL; ++c) total += SSQ[c];
      total = sqrtf(total);
      float clip = A.hp.max_grad_norm / (total + 1e-6f);
      clip = clip > 1.0f ? 1.0f : clip;
      ++adam_step;
      b1pow *= 0.9;
      b2pow *= 0.999;
      const float step_size = (float)((double)A.hp.lr / (1.0 - b1pow)), bc2s = (float)sqrt(1.0 - b2pow);
      for (int i0 = 4 * tid; i0 < S; i0 += 4 * PT) {
        const int p0 = crank * S + i0;
        float np4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u;
          const float g = GSL[i] * clip;
          const float mi = Ms[i] + (g - Ms[i]) * (1.0f - 0.9f);
          const float vi = Vs[i] * 0.999f + (1.0f - 0.999f) * g * g;
          Ms[i] = mi;
          Vs[i] = vi;
          np4[u] = (p0 + u < NP) ? Pm[p0 + u] - step_size * (mi / (sqrtf(vi) / bc2s + A.hp.adam_eps)) : 0.f;
        }
        const float4 v4 = make_float4(np4[0], np4[1], np4[2], np4[3]);
#pragma unroll
        for (int c = 0; c < CL; ++c)  // rotated start: the 8 owners write to 8 different CTAs at a time
          st4(cluster.map_shared_rank(Pm, (crank + c) & (CL - 1)) + p0, v4);
      }
      cluster.sync();  // (c) every CTA has the new parameters
      // transposed working copies, rebuilt locally through the position look-up table (no divisions)
      for (int p = tid; p < 2 * (h * Do + h + h * h + h); p += PT) {
        const int ip = imgpos[p];
        if (ip != 0xFFFF) img[ip] = Pm[p];
      }
      __syncthreads();
    }
  }

  // ---- write back: slice owners store parameters and moments; CTA 0 stores norm state and counters -------------------
  for (int i = tid; i < S; i += PT) {
    const int p = crank * S + i;
    if (p < NP) {
      g_params[p] = Pm[p];
      g_m[p] = Ms[i];
      g_v[p] = Vs[i];
    }
  }
  if (crank == 0) {
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
  cluster.sync();  // no CTA may exit while peers can still address its shared memory
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
  const int HP = A.HP, KP = A.KP, Da = A.pol.d_act, S = A.S;
  const int DAP = (Da + 3) / 4 * 4;
  size_t o = 0;
  o += 2 * (size_t)al(CL * S) + 3 * (size_t)al(S) + 64;
  o += al(2 * (KP * HP + HP * HP));
  o += al(KP * PRS) + al((DAP + 3) * PRS);
  o += al(KP * RL) + (size_t)8 * HP * RL + (size_t)3 * DAP * RL + 32;
  o += al(2 * 64 + 4) + al(PR);
  o += 3 * (size_t)al((CL * S + 1) / 2);
  o += al(PR * (A.pol.d_obs + (A.pol.discrete ? 1 : Da) + 3));
  return o;
}

static int launch_ppo(const PpoArgs& A0, float* params, float* norm, int32_t* norm_count, float* m, float* v,
                      const float* rollout, const int64_t* perm, float* loss_log, int64_t* state, cudaStream_t st) {
  PpoArgs A = A0;
  IMB_REQUIRE(A.hp.batch_size >= 1 && A.hp.batch_size <= PR, "PPO minibatch size must be in [1, %d]", PR);
  A.HP = A.pol.hidden <= 32 ? 32 : 64;
  A.KP = A.pol.d_obs <= 32 ? 32 : 64;
  A.S = ((A.pol.n_params + CL - 1) / CL + 3) / 4 * 4;
  const size_t fl = ppo_smem_floats(A);
  IMB_REQUIRE(fl * 4 <= IMB_SMEM_MAX, "PPO kernel needs %zu B of shared memory per CTA", fl * 4);
  IMB_REQUIRE(fl < 65536, "PPO kernel: shared-memory float offsets must fit 16 bits (policy too large)");
  static size_t attr_bytes = 0;
  if (fl * 4 > attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(k_ppo_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(fl * 4));
    if (e != cudaSuccess) IMB_FAIL(-2, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_bytes = fl * 4;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL);
  cfg.blockDim = dim3(PT);
  cfg.dynamicSmemBytes = fl * 4;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k_ppo_update, A, params, norm, norm_count, m, v, rollout, perm, loss_log,
                                     state);
  if (e != cudaSuccess) IMB_FAIL(-2, "k_ppo_update (cluster launch): %s", cudaGetErrorString(e));
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
