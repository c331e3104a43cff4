// imb_ppo.cu -- the generator update (PPO) as ONE persistent thread-block-cluster launch per round.
//
// The reference delegates this to stable-baselines3 (algorithms/adversarial/common.py:414,
// gen_algo.learn -> PPO.train); its arithmetic is restated in oracle/ppo_port.py (parity
// unpinned by the reference).  PPO.train is n_epochs x (N / batch_size)
// strictly sequential optimiser steps on 64-row minibatches: each step is ~1.3 MFLOP, far below
// launch latency, so the whole loop runs inside one kernel with parameters, gradients and Adam
// moments resident in the shared memory of an 8-CTA thread-block cluster (see k_ppo_update):
//   prefetch minibatch rows by permutation -> [feature RunningNorm update] -> advantage
//   normalisation -> warp-autonomous forward / loss / backward chain -> weight gradients ->
//   DSMEM exchange (st.async + mbarriers) -> clip_grad_norm_ -> Adam on the owned slice -> parameter all-gather.
// Also: imb_policy_logp = ActorCriticPolicy.evaluate_actions()[1] for the AIRL discriminator
// batch (common.py:476-519).
#include <cooperative_groups.h>
#include <stdlib.h>

#include "imb_common.cuh"
#include "imb_tile.cuh"

#ifdef IMB_PPO_TIMING
// phase timing for profiles/: CTA 0 / thread 0 accumulates clock64() deltas per phase of the optimiser step
__device__ long long g_ppo_clk[24];
__device__ long long g_ppo_wclk[80];  // [slot][warp]: cycles since the top barrier at points of the warp chain (CTA 0)
#define PPO_TICK(i)                                  \
  do {                                               \
    if (tid == 0) {                                  \
      const long long now_ = clock64();              \
      clk_acc[i] += now_ - clk_last;                 \
      clk_last = now_;                               \
    }                                                \
  } while (0)
#define PPO_WCLK(slot)                                   \
  do {                                                   \
    if (lane == 0) wacc[slot] += clock64() - wclk0;      \
  } while (0)
#else
#define PPO_TICK(i) do {} while (0)
#define PPO_WCLK(slot) do {} while (0)
#endif
#ifndef PPO_TANH
#define PPO_TANH(x) tanh_fast(x)
#endif

namespace {

constexpr int PT = 256;   // threads per CTA: warps 0, 1 = policy chain, 2, 3 = value chain, 4-7 = statistics / prefetch
constexpr int CL = 8;     // CTAs per cluster: each owns RL rows of every minibatch and 1/CL of the gradient reduction
constexpr int RL = 8;     // minibatch rows per CTA  (CL * RL = 64 >= SB3 batch_size)
constexpr int PR = CL * RL;

struct PpoArgs {
  imb_policy_desc pol;
  imb_ppo_hparams hp;
  int64_t n_rows;
  int rw;
  uint64_t seed;
  int HP, KP, S;  // tower width / obs width padded to 32; padded-layout parameters per slice (multiple of 4)
  int RS2;        // staged row stride in shared memory: >= rw, = 4 (mod 8)
};

// Padded parameter layout used inside the kernel ("P-layout"): W1 rows have stride ldo = d_obs|1 and W2 rows
// stride ldh = hidden|1 (odd), so that BOTH access patterns of the step -- lane = output unit (forward,
// weight gradients) and lane = input unit (backward) -- are shared-memory bank-conflict free straight from the
// parameter vector; no transposed working copies have to be rebuilt after every optimiser step.  The pad
// elements have zero value and zero gradient for ever (Adam leaves them at 0).
struct PLay {
  int w1[2], b1[2], w2[2], b2[2], wa, ba, wv, bv, ls, ldo, ldh, total;
};
__host__ __device__ inline PLay make_play(const imb_policy_desc& pd) {
  PLay L;
  const int Do = pd.d_obs, Da = pd.d_act, h = pd.hidden;
  L.ldo = Do | 1;
  L.ldh = h | 1;
  int o = 0;
  for (int t = 0; t < 2; ++t) {
    L.w1[t] = o; o += h * L.ldo;
    L.b1[t] = o; o += h;
    L.w2[t] = o; o += h * L.ldh;
    L.b2[t] = o; o += h;
  }
  L.wa = o; o += Da * L.ldh;
  L.ba = o; o += Da;
  L.wv = o; o += h;
  L.bv = o; o += 1;
  L.ls = o; o += pd.discrete ? 0 : Da;
  L.total = o;
  return L;
}
// torch-flat parameter index -> P-layout index
__device__ inline int flat_to_play(const imb_policy_desc& pd, const PLay& L, int p) {
  const int Do = pd.d_obs, Da = pd.d_act, h = pd.hidden;
  auto in = [&](int off, int len) { return p >= off && p < off + len; };
  if (in(pd.off_pi_w1, h * Do)) { const int i = p - pd.off_pi_w1; return L.w1[0] + (i / Do) * L.ldo + i % Do; }
  if (in(pd.off_vf_w1, h * Do)) { const int i = p - pd.off_vf_w1; return L.w1[1] + (i / Do) * L.ldo + i % Do; }
  if (in(pd.off_pi_w2, h * h)) { const int i = p - pd.off_pi_w2; return L.w2[0] + (i / h) * L.ldh + i % h; }
  if (in(pd.off_vf_w2, h * h)) { const int i = p - pd.off_vf_w2; return L.w2[1] + (i / h) * L.ldh + i % h; }
  if (in(pd.off_pi_b1, h)) return L.b1[0] + p - pd.off_pi_b1;
  if (in(pd.off_vf_b1, h)) return L.b1[1] + p - pd.off_vf_b1;
  if (in(pd.off_pi_b2, h)) return L.b2[0] + p - pd.off_pi_b2;
  if (in(pd.off_vf_b2, h)) return L.b2[1] + p - pd.off_vf_b2;
  if (in(pd.off_act_w, Da * h)) { const int i = p - pd.off_act_w; return L.wa + (i / h) * L.ldh + i % h; }
  if (in(pd.off_act_b, Da)) return L.ba + p - pd.off_act_b;
  if (in(pd.off_val_w, h)) return L.wv + p - pd.off_val_w;
  if (p == pd.off_val_b) return L.bv;
  return L.ls + p - pd.off_log_std;
}

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

// sum over the 8 lanes of an aligned lane group; ALL 32 lanes must call it convergently (a per-group member
// mask makes the compiler serialise the groups through MATCH.ANY: 40x more instructions, measured)
__device__ __forceinline__ float group8_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}
__device__ __forceinline__ float rcp_fast(float x) { return __fdividef(1.0f, x); }
__device__ __forceinline__ float sqrt_fast(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ float dot8r(const float (&d)[8], const float* __restrict__ b) {
  const float4 b0 = ld4(b), b1 = ld4(b + 4);
  float s0 = d[0] * b0.x, s1 = d[1] * b0.y;
  s0 = fmaf(d[2], b0.z, s0);
  s1 = fmaf(d[3], b0.w, s1);
  s0 = fmaf(d[4], b1.x, s0);
  s1 = fmaf(d[5], b1.y, s1);
  s0 = fmaf(d[6], b1.z, s0);
  s1 = fmaf(d[7], b1.w, s1);
  return s0 + s1;
}
__device__ __forceinline__ void load8(float (&d)[8], const float* __restrict__ p) {
  const float4 a = ld4(p), b = ld4(p + 4);
  d[0] = a.x, d[1] = a.y, d[2] = a.z, d[3] = a.w, d[4] = b.x, d[5] = b.y, d[6] = b.z, d[7] = b.w;
}
__device__ __forceinline__ float sum8(const float (&d)[8]) {
  return ((d[0] + d[1]) + (d[2] + d[3])) + ((d[4] + d[5]) + (d[6] + d[7]));
}

// Weight gradients of ONE tower over the CTA's RL = 8 rows -> GP (P-layout, local shared memory), by NQ warps: thread =
// (unit gj = lane, every NQ-th input); the unit's dL/dz rows live in registers, the input rows are warp-uniform
// broadcasts, the scattered GP stores have odd lane strides (conflict free).  The head biases / log_std (plain row sums)
// ride with the last input group.  The dot products of a group are all computed into registers BEFORE the group's
// stores: the compiler cannot prove that GP and the activation tiles do not alias, and with a store between two dots it
// serialises them (load latency + FMA chain per dot, ~55 cycles each; the phase is latency bound).
struct WgradOff {
  int w1, b1, w2, b2, wa, ba, wv, bv, ls, ldo, ldh;
};
template <int NQ, int HPx>
__device__ __forceinline__ void tower_wgrad(const int tnet, const int gj, const int wq, const int h, const int Do, const int Da,
                                            const bool discrete, const WgradOff o, float* __restrict__ GP,
                                            const float* __restrict__ tH1, const float* __restrict__ tLAT,
                                            const float* __restrict__ tDZ2, const float* __restrict__ tDZ1,
                                            const float* __restrict__ XNc, const float* __restrict__ DM,
                                            const float* __restrict__ DLS, const float* __restrict__ DVAL) {
  constexpr int RLc = 8;
  constexpr int T2 = (HPx + NQ - 1) / NQ;  // layer-2 inputs per thread
  float dz2[8], dz1[8], lt[8];
  if (gj < h) {
    load8(dz2, tDZ2 + gj * RLc);
    load8(dz1, tDZ1 + gj * RLc);
    load8(lt, tLAT + gj * RLc);
    {  // dW2[gj][i], i = wq + NQ t
      float r[T2];
#pragma unroll
      for (int t = 0; t < T2; ++t) {
        const int i = wq + NQ * t;
        r[t] = dot8r(dz2, tH1 + (i < HPx ? i : 0) * RLc);
      }
#pragma unroll
      for (int t = 0; t < T2; ++t) {
        const int i = wq + NQ * t;
        if (i < h) GP[o.w2 + gj * o.ldh + i] = r[t];
      }
    }
    for (int k0 = wq; k0 < Do; k0 += 4 * NQ) {  // dW1[gj][k], four at a time
      float r[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int k = k0 + NQ * t;
        r[t] = dot8r(dz1, XNc + (k < Do ? k : 0) * RLc);
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int k = k0 + NQ * t;
        if (k < Do) GP[o.w1 + gj * o.ldo + k] = r[t];
      }
    }
    if (tnet == 0) {
      for (int a0 = wq; a0 < Da; a0 += 2 * NQ) {
        const int a1 = a0 + NQ;
        const float ra = dot8r(lt, DM + a0 * RLc), rb = dot8r(lt, DM + (a1 < Da ? a1 : a0) * RLc);
        GP[o.wa + a0 * o.ldh + gj] = ra;
        if (a1 < Da) GP[o.wa + a1 * o.ldh + gj] = rb;
      }
    } else if (wq == 0) {
      GP[o.wv + gj] = dot8r(lt, DVAL);
    }
    if (wq == 0) GP[o.b2 + gj] = sum8(dz2);
    if (wq == NQ - 1) GP[o.b1 + gj] = sum8(dz1);
  }
  if (wq == NQ - 1) {
    if (tnet == 0) {  // ba, log_std
      for (int t = gj; t < 2 * Da; t += HPx) {
        if (t < Da) {
          load8(dz2, DM + t * RLc);
          GP[o.ba + t] = sum8(dz2);
        } else if (!discrete) {
          load8(dz2, DLS + (t - Da) * RLc);
          GP[o.ls + t - Da] = sum8(dz2);
        }
      }
    } else if (gj == 0) {  // bv
      load8(dz2, DVAL);
      GP[o.bv] = sum8(dz2);
    }
  }
}

// PPO.train for one rollout: n_epochs x ceil(N / batch) optimiser steps, ONE cluster of CL CTAs.
// Data flow of one optimiser step (64-row minibatch, CTA c owns rows 8c..8c+7, every CTA holds all parameters):
//   * the minibatch rows are staged row-major in shared memory by asynchronous 16-byte row copies issued TWO
//     steps ahead by warps 4-7 (indices drawn on the fly, three buffers), completion on an mbarrier;
//   * minibatch statistics (feature RunningNorm update, advantage normalisation) over all 64 rows are computed
//     redundantly and identically by every CTA -- one step ahead, by warps 4-7, BESIDE the chain of warps 0-3;
//   * forward + loss + backward to dL/dz run WARP-AUTONOMOUSLY on warps 0-3: warp = (tower, 4 own rows), lane =
//     hidden unit; only __syncwarp() between layers (the two towers interact through the summed loss only), so
//     the ~8 dependent stages of the chain cost no CTA barrier, and every weight read from shared memory
//     feeds four FMAs (with 2 rows per warp on all 8 warps the chain was bound by shared-memory wavefronts);
//   * weight gradients over the CTA's 8 rows: thread = (unit, input subset) of one tower, staged in local shared
//     memory in the parameter layout.  The VALUE tower's are computed early by warps 2-7 (its chain ends ~1 k cycles
//     before the policy tower's: no action head, no loss terms) while warps 0, 1 finish the policy chain; the policy
//     tower's by all eight warps after the step's second barrier.  The staged vector is pushed to the slice owners
//     through distributed shared memory with 16-byte stores; owners sum the CL partials in fixed order, exchange the
//     squared slice norms, run clip_grad_norm_ + Adam on their slice (the moments never leave their owner) and
//     all-gather the new parameters into every CTA's copy.
// No cluster barrier inside the step loop (the exchanged data signals mbarriers at the receivers), three CTA
// barriers per optimiser step; the only global-memory traffic inside a step is the asynchronous minibatch prefetch.
template <int HP>
__global__ void __launch_bounds__(PT, 1) k_ppo_update(const PpoArgs A, float* __restrict__ g_params,
                                                      float* __restrict__ g_norm, int32_t* __restrict__ g_norm_count,
                                                      float* __restrict__ g_m, float* __restrict__ g_v,
                                                      const float* __restrict__ rollout,
                                                      const int64_t* __restrict__ perm_in,
                                                      float* __restrict__ loss_log, int64_t* __restrict__ state) {
  static_assert(HP == 32, "the warp-autonomous chain maps one lane to one hidden unit");
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  extern __shared__ __align__(128) float smem[];
  __shared__ float red[32];
  __shared__ float bc[8];
  __shared__ __align__(8) uint64_t mbar[3];  // one per staged-minibatch buffer (step s lives in buffer s % 3)
  __shared__ __align__(8) uint64_t xbar[3];  // [0]: partial gradients of the owned slice arrived; [1]: all new parameter slices; [2]: all slice norms
  __shared__ float SSQ[CL];                  // squared gradient norms of the 8 slices (each written by its owner)
  const imb_policy_desc& pd = A.pol;
  const int Do = pd.d_obs, Da = pd.d_act, h = pd.hidden, NP = pd.n_params, KP = A.KP, S = A.S;
  const PLay PL = make_play(pd);
  const int ldo = PL.ldo, ldh = PL.ldh;
  const int da_store = pd.discrete ? 1 : Da;
  const int col_logp = Do + da_store, col_adv = col_logp + 3, col_ret = col_logp + 4;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rw = A.rw, RS2 = A.RS2;          // rollout row width (multiple of 4) / staged row stride (= 4 mod 8)
  auto al = [](int x) { return (x + 31) / 32 * 32; };

  // ---- shared-memory carve-up (identical in every CTA: DSMEM addresses are rank + offset) ---------------
  int o = 0;
  float* Pm = smem + o; o += al(CL * S);        // parameters, P-layout, padded to CL slices
  float* Ms = smem + o; o += al(CL * S);        // Adam moments (P-layout indexing; only the owned slice is live)
  float* Vs = smem + o; o += al(CL * S);
  float* GP = smem + o; o += al(CL * S);        // own partial gradient (staging for the push)
  float* RECV = smem + o; o += al(CL * S);      // [source CTA][S]: partial gradients of the owned slice
  float* LOSS = smem + o; o += 32;              // [CL][3] partial loss sums (read by CTA 0)
  const int DAP = (Da + 3) / 4 * 4;
  const int rsz = al(PR * RS2);
  float* ROWS = smem + o; o += 3 * rsz;         // triple-buffered minibatch, row-major [64][RS2]
  // own-row tiles, feature-major [feature][RL]
  const int xsz = al(KP * RL);
  float* XNo = smem + o; o += 2 * xsz;          // normalised observations, double-buffered by step parity
  float* TH1 = smem + o; o += 2 * HP * RL;
  float* TLAT = smem + o; o += 2 * HP * RL;
  float* TDZ2 = smem + o; o += 2 * HP * RL;
  float* TDZ1 = smem + o; o += 2 * HP * RL;
  float* DM = smem + o; o += DAP * RL;          // dL/d(mean|logits) [a][RL]
  float* DLS = smem + o; o += DAP * RL;         // dL/d(log_std) per row [a][RL]
  float* MEAN = smem + o; o += DAP * RL;        // action means / logits [a][RL]
  float* DVAL = smem + o; o += 32;              // [RL] dL/dvalue
  float* rstat = smem + o; o += al(2 * 64 + 4); // running mean | var of the policy's feature RunningNorm

  for (int i = tid; i < CL * S; i += PT) Pm[i] = Ms[i] = Vs[i] = GP[i] = RECV[i] = 0.f;
  for (int i = tid; i < 3 * rsz; i += PT) ROWS[i] = 0.f;
  for (int i = tid; i < 2 * xsz + 8 * HP * RL + 3 * DAP * RL + 32; i += PT) XNo[i] = 0.f;  // XNo .. DVAL contiguous
  if (tid < 64) {
    rstat[tid] = (pd.has_norm && tid < Do) ? g_norm[tid] : 0.f;
    rstat[64 + tid] = (pd.has_norm && tid < Do) ? g_norm[Do + tid] : 1.f;
    LOSS[tid & 31] = 0.f;
  }
  if (tid == 0) {
    mbar_init(&mbar[0], 128);
    mbar_init(&mbar[1], 128);
    mbar_init(&mbar[2], 128);
    mbar_init(&xbar[0], 1);
    mbar_init(&xbar[1], 1);
    mbar_init(&xbar[2], 1);
    mbar_fence_init();
  }
  __syncthreads();
  for (int p = tid; p < NP; p += PT) {
    const int q = flat_to_play(pd, PL, p);
    Pm[q] = g_params[p];
    Ms[q] = g_m[p];
    Vs[q] = g_v[p];
  }
  int32_t run_count = pd.has_norm ? *g_norm_count : 0;

  const int64_t N = A.n_rows;
  const int Ni = (int)N;
  const int mb = A.hp.batch_size;
  const int64_t steps_per_epoch = (N + mb - 1) / mb;
  const int64_t n_steps = steps_per_epoch * A.hp.n_epochs;
  int64_t adam_step = state[IMB_ST_PPO_STEP];
  const int64_t perm_draw0 = state[IMB_ST_PPO_EPOCH];
  double b1pow = pow(0.9, (double)adam_step), b2pow = pow(0.999, (double)adam_step);  // beta^t, kept incrementally
  const int row0 = crank * RL;        // first minibatch row owned by this CTA
  const WgradOff wo_p = {PL.w1[0], PL.b1[0], PL.w2[0], PL.b2[0], PL.wa, PL.ba, PL.wv, PL.bv, PL.ls, ldo, ldh};
  const WgradOff wo_v = {PL.w1[1], PL.b1[1], PL.w2[1], PL.b2[1], PL.wa, PL.ba, PL.wv, PL.bv, PL.ls, ldo, ldh};

  // Asynchronous row gather of one minibatch (epoch ep, first row start) into buffer `buf` by warps 4-7: two
  // threads per row draw the row index and copy half of the 16-byte aligned rollout row each with 16-byte
  // cp.async (LDGSTS); completion is tracked by the buffer's mbarrier (one deferred arrival per thread).  (One
  // bulk-async copy per row and lane was tried first: the 32 per-lane UBLKCP issues serialise, ~2500 cycles per
  // warp.)  Rows are fetched TWO optimiser steps ahead (three buffers), so their latency is never waited for.
  auto issue_gather = [&](int ep, int start, int buf) {
    const int t = tid - 128;
    if (t < 0) return;
    const int r = t >> 1, half = t & 1;
    const int nbx = min(mb, Ni - start);
    if (r < nbx) {
      int64_t idx;
      if (perm_in) {
        idx = perm_in[(int64_t)ep * N + start + r];
      } else {
        const FeistelKey fk = feistel_key(A.seed, IMB_STREAM_PPO_PERM, (uint64_t)(perm_draw0 + ep), (uint64_t)N);
        idx = (int64_t)feistel_perm(fk, (uint64_t)(start + r), (uint64_t)N);
      }
      const float* src = rollout + idx * rw;
      float* dst = ROWS + buf * rsz + r * RS2;
      const int nq = rw >> 2, q0 = half ? (nq + 1) >> 1 : 0, q1 = half ? nq : (nq + 1) >> 1;
      for (int q = q0; q < q1; ++q) cp_async16(dst + 4 * q, src + 4 * q);
    }
    cp_async_mbar_arrive(&mbar[buf]);
  };
  // Statistics of one minibatch (step gs2, staged in buffer gs2 % 3): feature RunningNorm update + advantage
  // normalisation over all 64 rows, one 8-lane group per statistic, identically in every CTA; the group of
  // feature k also writes this CTA's own rows of it, normalised, into XNo[gs2 & 1][k][RL] (lane = row).  Run by
  // `ngrp` 8-lane groups (whole warps; group index gidx); all lanes run the same code (full-mask shuffles); idle
  // groups chew on the advantage column and discard the result.  The staged row stride RS2 = 4 (mod 8) makes
  // the 32 lanes of a warp (4 features x 8 rows) hit 32 banks.
  auto minibatch_stats = [&](int64_t gs2, int nbx, int gidx, int ngrp) {
    const int buf = (int)(gs2 % 3);
    float* R = ROWS + buf * rsz;
    float* XN = XNo + (int)(gs2 & 1) * xsz;
    const int gl = tid & 7;
    const float inv_nbx = 1.0f / (float)nbx;
    mbar_wait(&mbar[buf], (uint32_t)((gs2 / 3) & 1));  // all 64 row copies have landed
    for (int task0 = 0; task0 <= Do; task0 += ngrp) {
      const int task = task0 + gidx;
      const bool is_feat = task < Do, is_adv = task == Do;
      float* x = R + (is_feat ? task : col_adv);
      float v[8], s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[i] = (gl + 8 * i < nbx) ? x[(gl + 8 * i) * RS2] : 0.f;
        s += v[i];
      }
      const float bmean = group8_sum(s) * inv_nbx;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = (gl + 8 * i < nbx) ? v[i] - bmean : 0.f;
        q = fmaf(d, d, q);
      }
      const float ssd = group8_sum(q);
      if (is_feat) {
        float mean = 0.f, istd = 1.f;
        if (pd.has_norm) {  // every lane of the group computes the update; lane 0 stores it
          mean = rstat[task];
          float var = rstat[64 + task];
          const float bvar = ssd * inv_nbx;
          const float bn = (float)nbx, c = (float)run_count, itot = rcp_fast(c + bn), delta = bmean - mean;
          mean += delta * bn * itot;
          var *= c;
          var += bvar * bn;
          var += delta * delta * c * bn * itot;
          var *= itot;
          istd = rsqrtf(var + pd.norm_eps);
          if (gl == 0) {
            rstat[task] = mean;
            rstat[64 + task] = var;
          }
        }
        XN[task * RL + gl] = (row0 + gl < nbx) ? (x[(row0 + gl) * RS2] - mean) * istd : 0.f;
      } else if (is_adv) {
        float am = 0.f, ais = 1.f;
        if (A.hp.normalize_advantage && nbx > 1) {
          am = bmean;
          ais = rcp_fast(sqrt_fast(ssd / (float)(nbx - 1)) + 1e-8f);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (gl + 8 * i < nbx) x[(gl + 8 * i) * RS2] = (v[i] - am) * ais;
      }
    }
  };
  __syncthreads();
  issue_gather(0, 0, 0);
  if (n_steps > 1) issue_gather(mb >= Ni ? 1 : 0, mb >= Ni ? 0 : mb, 1);
  minibatch_stats(0, min(mb, Ni), tid >> 3, PT / 8);
  if (pd.has_norm) run_count += min(mb, Ni);  // (every thread keeps the count; the statistics' owners use it)
  // The gradient exchange is synchronised by the data itself: every 16-byte DSMEM store (st.async) completes
  // bytes on an mbarrier of the RECEIVING CTA, which waits until the expected byte count of the phase has
  // landed -- no cluster-wide barrier inside the step loop.  Each barrier is re-armed (one arrival + expected
  // bytes) by its owner right after the previous phase completed, which is always before a peer can send for
  // the next phase (a peer's next-phase data depends on data this CTA sends later).
  const uint32_t xbytes = (uint32_t)(CL * S * 4);
  const unsigned qmagic = (unsigned)(0x100000000ull / (unsigned)(S / 4)) + 1u;  // exact quotient for the < 2^16 quads here
  const uint32_t recv_sa = smem_u32(RECV), pm_sa = smem_u32(Pm), ssq_sa = smem_u32(SSQ);
  const uint32_t xbar0_sa = smem_u32(&xbar[0]), xbar1_sa = smem_u32(&xbar[1]), xbar2_sa = smem_u32(&xbar[2]);
  if (tid == 0) {
    mbar_expect_tx(&xbar[0], xbytes);
    mbar_expect_tx(&xbar[1], xbytes);
    mbar_expect_tx(&xbar[2], (uint32_t)(CL * 4));
  }
  cluster.sync();

#ifdef IMB_PPO_TIMING
  long long clk_acc[16] = {0}, clk_last = clock64();
  long long wacc[10] = {0};  // per-warp chain clocks, kept in registers (a global RMW per sample stalls the chain)
#endif
  int ep_now = 0, start = 0;  // epoch and first row of the current step
  for (int64_t gs = 0; gs < n_steps; ++gs) {
    const int cur = (int)(gs & 1);
    const float* Rc = ROWS + (int)(gs % 3) * rsz;
    const float* XNc = XNo + cur * xsz;
    const int nb = min(mb, Ni - start);
    const float inv_nb = 1.0f / (float)nb;
    int ep_next = ep_now, start_next = start + mb;
    if (start_next >= Ni) {
      start_next = 0;
      ++ep_next;
    }
    ++adam_step;
    if (tid == PT - 1) {  // Adam bias corrections in double, off the critical path (double-buffered by step parity)
      b1pow *= 0.9;
      b2pow *= 0.999;
      bc[2 * cur] = (float)((double)A.hp.lr / (1.0 - b1pow));
      bc[2 * cur + 1] = (float)sqrt(1.0 - b2pow);
    }
    __syncthreads();  // the parameters written by the previous step's Adam (and XNo, the statistics) are visible
    PPO_TICK(0);
#ifdef IMB_PPO_TIMING
    const long long wclk0 = clock64();
#endif
    float l_pg = 0.f, l_v = 0.f, l_ent = 0.f;
    if (warp >= 4) {
      // ---- 1b. warps 4-7, beside the chain: the NEXT step's minibatch statistics and own-row tile (they do not
      //          depend on the parameters), then the row prefetch for the step after it ------------------------------
      if (gs + 1 < n_steps) minibatch_stats(gs + 1, min(mb, Ni - start_next), (tid - 128) >> 3, (PT - 128) / 8);
      PPO_WCLK(1);
      if (gs + 2 < n_steps) {
        int ep2 = ep_next, start2 = start_next + mb;
        if (start2 >= Ni) {
          start2 = 0;
          ++ep2;
        }
        issue_gather(ep2, start2, (int)((gs + 2) % 3));
      }
      PPO_WCLK(3);
    } else {
      // ---- 1a. warp-autonomous chain on warps 0-3: forward, loss terms, backward to dL/dz for (tower, 4 own rows).
      //          Four rows per warp: every weight fetched from shared memory feeds four FMAs (the chain was bound
      //          by shared-memory wavefronts with two rows per warp and all eight warps fetching) ---------------------
      const int cnet = warp >> 1;                // 0: policy tower (warps 0, 1), 1: value tower (warps 2, 3)
      const int j = lane, jc = lane < h ? lane : 0, r0 = 4 * (warp & 1);
      const bool jl = lane < h;
      const int rr = lane >> 3, la = lane & 7;   // per-row parts: lane octet rr handles row r0 + rr
      const int gr = row0 + r0 + rr;
      const bool live = gr < nb;
      const float* row = Rc + gr * RS2;
      float* cH1 = TH1 + cnet * HP * RL;
      float* cLAT = TLAT + cnet * HP * RL;
      float* cDZ2 = TDZ2 + cnet * HP * RL;
      float* cDZ1 = TDZ1 + cnet * HP * RL;
      const float* cW1 = Pm + (cnet ? PL.w1[1] : PL.w1[0]);
      const float* cW2 = Pm + (cnet ? PL.w2[1] : PL.w2[0]);
      const int c_b1 = cnet ? PL.b1[1] : PL.b1[0], c_b2 = cnet ? PL.b2[1] : PL.b2[0];
      // policy warps: everything the loss needs that does not depend on the forward pass is fetched now, so its
      // latency (shared-memory loads, the exponential) hides behind the layers
      const bool gfast = cnet == 0 && !pd.discrete && Da <= 8;  // one action per lane of the octet
      float pre_adv = 0.f, pre_lpo = 0.f, pre_act = 0.f, pre_ls = 0.f, pre_ivar = 0.f;
      if (cnet == 0) {
        pre_adv = row[col_adv];
        pre_lpo = row[col_logp];
        if (gfast && la < Da) {
          pre_act = row[Do + la];
          pre_ls = Pm[PL.ls + la];
          pre_ivar = __expf(-2.0f * pre_ls);
        }
      }
      // layer 1
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      {
        const float* wp = cW1 + jc * ldo;
#pragma unroll 8
        for (int k = 0; k < Do; ++k) {
          const float w = wp[k];
          const float4 x = ld4(XNc + k * RL + r0);
          a0 = fmaf(x.x, w, a0);
          a1 = fmaf(x.y, w, a1);
          a2 = fmaf(x.z, w, a2);
          a3 = fmaf(x.w, w, a3);
        }
      }
      float b = Pm[c_b1 + jc];
      const float h10 = jl ? PPO_TANH(a0 + b) : 0.f, h11 = jl ? PPO_TANH(a1 + b) : 0.f;
      const float h12 = jl ? PPO_TANH(a2 + b) : 0.f, h13 = jl ? PPO_TANH(a3 + b) : 0.f;
      st4(cH1 + j * RL + r0, make_float4(h10, h11, h12, h13));
      __syncwarp();
      PPO_WCLK(3);
      // layer 2
      a0 = a1 = a2 = a3 = 0.f;
      {
        const float* wp = cW2 + jc * ldh;
#pragma unroll 8
        for (int i = 0; i < h; ++i) {
          const float w = wp[i];
          const float4 x = ld4(cH1 + i * RL + r0);
          a0 = fmaf(x.x, w, a0);
          a1 = fmaf(x.y, w, a1);
          a2 = fmaf(x.z, w, a2);
          a3 = fmaf(x.w, w, a3);
        }
      }
      b = Pm[c_b2 + jc];
      const float lat0 = jl ? PPO_TANH(a0 + b) : 0.f, lat1 = jl ? PPO_TANH(a1 + b) : 0.f;
      const float lat2 = jl ? PPO_TANH(a2 + b) : 0.f, lat3 = jl ? PPO_TANH(a3 + b) : 0.f;
      st4(cLAT + j * RL + r0, make_float4(lat0, lat1, lat2, lat3));
      PPO_WCLK(4);
      // heads + loss terms; dl0..dl3 = dL/dlatent of this unit for the four rows
      float dl0 = 0.f, dl1 = 0.f, dl2 = 0.f, dl3 = 0.f;
      auto oct_sum = [&](float v) {
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        return v;
      };
      if (cnet == 1) {
        // value head: four sums over the 32 units, transposed on the way so that octet rr ends with row r0 + rr's
        const float wvj = jl ? Pm[PL.wv + j] : 0.f;
        const float p0 = lat0 * wvj, p1 = lat1 * wvj, p2 = lat2 * wvj, p3 = lat3 * wvj;
        const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
        float k0 = up16 ? p2 : p0, k1 = up16 ? p3 : p1;
        k0 += __shfl_xor_sync(0xffffffffu, up16 ? p0 : p2, 16);
        k1 += __shfl_xor_sync(0xffffffffu, up16 ? p1 : p3, 16);
        float kk = up8 ? k1 : k0;
        kk += __shfl_xor_sync(0xffffffffu, up8 ? k0 : k1, 8);
        const float val = oct_sum(kk) + Pm[PL.bv];
        const float dv = val - row[col_ret];
        const float dval = live ? A.hp.vf_coef * 2.0f * dv * inv_nb : 0.f;
        if (la == 0) {
          if (live) l_v = dv * dv;
          DVAL[r0 + rr] = dval;
        }
        dl0 = __shfl_sync(0xffffffffu, dval, 0) * wvj;
        dl1 = __shfl_sync(0xffffffffu, dval, 8) * wvj;
        dl2 = __shfl_sync(0xffffffffu, dval, 16) * wvj;
        dl3 = __shfl_sync(0xffffffffu, dval, 24) * wvj;
      } else {
        const float* Wa = Pm + PL.wa;
        // action means / logits from the latent tile in shared memory: lane = (action ab + lane / 4, units q + 4 i of
        // quarter q = lane % 4): one 16-byte load brings a unit's four rows, so a weight and a latent load feed four FMAs
        // (one lane per (action, row) cost two loads per FMA: 64 loads per lane, ~700 cycles of the chain); the four
        // row sums are then reduced over the quad by a transposing butterfly (3 shuffles) that leaves row r0 + q in
        // lane q.  Pad units hold zeros; Wa rows have the odd stride ldh.
        __syncwarp();
        {
          const int asub = lane >> 2, q = lane & 3;
          const bool b0 = (lane & 1) != 0, b1 = (lane & 2) != 0;
          for (int ab = 0; ab < Da; ab += 8) {
            const int a = ab + asub, ac = a < Da ? a : 0;
            const float* wr = Wa + ac * ldh + q;
            const float* lr = cLAT + q * RL + r0;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int i = 0; i < HP / 4; ++i) {
              const float w = wr[4 * i];
              const float4 l4 = ld4(lr + 4 * i * RL);
              s0 = fmaf(w, l4.x, s0);
              s1 = fmaf(w, l4.y, s1);
              s2 = fmaf(w, l4.z, s2);
              s3 = fmaf(w, l4.w, s3);
            }
            float ka = b0 ? s1 : s0, kb = b0 ? s3 : s2;
            ka += __shfl_xor_sync(0xffffffffu, b0 ? s0 : s1, 1);
            kb += __shfl_xor_sync(0xffffffffu, b0 ? s2 : s3, 1);
            float kk = b1 ? kb : ka;
            kk += __shfl_xor_sync(0xffffffffu, b1 ? ka : kb, 2);
            if (a < Da) MEAN[a * RL + r0 + q] = kk + Pm[PL.ba + a];
          }
        }
        __syncwarp();
        PPO_WCLK(5);
        const float adv = pre_adv, logp_old = pre_lpo;
        float logp = 0.f, ent = 0.f;
        float r_dm = 0.f, r_dls = 0.f;  // fast path: this lane's d logp / d mean, d logp / d log_std
        int act = 0;
        if (gfast) {
          if (la < Da) {
            const float diff = pre_act - MEAN[la * RL + r0 + rr];
            const float d2 = diff * diff * pre_ivar;
            logp = -0.5f * d2 - pre_ls - 0.9189385332046727f;
            ent = 1.4189385332046727f + pre_ls;
            r_dm = diff * pre_ivar;
            r_dls = d2 - 1.0f;
          }
        } else if (!pd.discrete) {
          const float* lstd = Pm + PL.ls;
          for (int a = la; a < Da; a += 8) {
            const float ls = lstd[a], ivar = __expf(-2.0f * ls);
            const float diff = row[Do + a] - MEAN[a * RL + r0 + rr];
            const float d2 = diff * diff * ivar;
            logp += -0.5f * d2 - ls - 0.9189385332046727f;
            ent += 1.4189385332046727f + ls;
            DM[a * RL + r0 + rr] = diff * ivar;   // d logp / d mean
            DLS[a * RL + r0 + rr] = d2 - 1.0f;    // d logp / d log_std
          }
        } else {
          float mx = -INFINITY;
          for (int a = la; a < Da; a += 8) mx = fmaxf(mx, MEAN[a * RL + r0 + rr]);
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
          float se = 0.f;
          for (int a = la; a < Da; a += 8) se += expf(MEAN[a * RL + r0 + rr] - mx);
          const float lse = mx + logf(oct_sum(se));
          act = (int)row[Do];
          for (int a = la; a < Da; a += 8) {
            const float lp = MEAN[a * RL + r0 + rr] - lse;
            if (a == act) logp = lp;
            ent -= expf(lp) * lp;
            DLS[a * RL + r0 + rr] = lp;  // temporarily: log p_a
          }
        }
        logp = oct_sum(logp);
        PPO_WCLK(6);
        if (pd.discrete || loss_log) ent = oct_sum(ent);  // (Gaussian: the entropy only feeds the loss log)
        const float ratio = __expf(logp - logp_old);
        const float lo = 1.0f - A.hp.clip_range, hi = 1.0f + A.hp.clip_range;
        const float pl1 = adv * ratio, pl2 = adv * fminf(fmaxf(ratio, lo), hi);
        const bool inside = (ratio >= lo) && (ratio <= hi);
        float dl_dlogp = (inside || pl1 < pl2) ? -adv * ratio * inv_nb : 0.f;
        float dent = -A.hp.ent_coef * inv_nb;  // d(ent_coef * ent_loss) / d(entropy)
        if (live) {
          if (la == 0) {
            l_pg = -fminf(pl1, pl2);
            l_ent = -ent;
          }
        } else {
          dl_dlogp = 0.f;
          dent = 0.f;
        }
        if (gfast) {
          if (la < Da) {
            DLS[la * RL + r0 + rr] = dl_dlogp * r_dls + dent;  // dH/dlog_std = 1
            DM[la * RL + r0 + rr] = dl_dlogp * r_dm;
          }
        } else if (!pd.discrete) {
          for (int a = la; a < Da; a += 8) {
            DLS[a * RL + r0 + rr] = dl_dlogp * DLS[a * RL + r0 + rr] + dent;  // dH/dlog_std = 1
            DM[a * RL + r0 + rr] = dl_dlogp * DM[a * RL + r0 + rr];
          }
        } else {
          for (int a = la; a < Da; a += 8) {
            const float lp = DLS[a * RL + r0 + rr], pp = expf(lp);
            DM[a * RL + r0 + rr] = dl_dlogp * (((a == act) ? 1.f : 0.f) - pp) + dent * (-pp * (lp + ent));
            DLS[a * RL + r0 + rr] = 0.f;
          }
        }
        __syncwarp();
        PPO_WCLK(7);
#pragma unroll 4
        for (int a = 0; a < Da; ++a) {
          const float waj = Wa[a * ldh + jc];
          const float4 d = ld4(DM + a * RL + r0);
          dl0 = fmaf(d.x, waj, dl0);
          dl1 = fmaf(d.y, waj, dl1);
          dl2 = fmaf(d.z, waj, dl2);
          dl3 = fmaf(d.w, waj, dl3);
        }
      }
      PPO_WCLK(2);
      // dL/dz2, backward through layer 2 (lane = input unit i: dH1[i] = sum_j DZ2[j] W2[j][i]), dL/dz1
      st4(cDZ2 + j * RL + r0,
          make_float4(jl ? dl0 * (1.0f - lat0 * lat0) : 0.f, jl ? dl1 * (1.0f - lat1 * lat1) : 0.f,
                      jl ? dl2 * (1.0f - lat2 * lat2) : 0.f, jl ? dl3 * (1.0f - lat3 * lat3) : 0.f));
      __syncwarp();
      a0 = a1 = a2 = a3 = 0.f;
      {
        const float* wp = cW2 + jc;
#pragma unroll 8
        for (int jj = 0; jj < h; ++jj) {
          const float w = wp[jj * ldh];
          const float4 d = ld4(cDZ2 + jj * RL + r0);
          a0 = fmaf(d.x, w, a0);
          a1 = fmaf(d.y, w, a1);
          a2 = fmaf(d.z, w, a2);
          a3 = fmaf(d.w, w, a3);
        }
      }
      st4(cDZ1 + j * RL + r0,
          make_float4(jl ? a0 * (1.f - h10 * h10) : 0.f, jl ? a1 * (1.f - h11 * h11) : 0.f,
                      jl ? a2 * (1.f - h12 * h12) : 0.f, jl ? a3 * (1.f - h13 * h13) : 0.f));
    }
    if (pd.has_norm && gs + 1 < n_steps) run_count += min(mb, Ni - start_next);
    PPO_WCLK(0);
    // ---- 1c. the VALUE tower's weight gradients, early: its chain (warps 2, 3) finishes ~1 k cycles before the policy
    //          tower's (no action head, no loss terms) and the statistics / prefetch warps are done by then as well, so
    //          warps 2-7 meet on a named barrier and compute them while warps 0, 1 are still in the policy chain ------
    if (warp >= 2) {
      asm volatile("bar.sync 1, 192;" ::: "memory");
      tower_wgrad<6, HP>(1, lane, warp - 2, h, Do, Da, pd.discrete != 0, wo_v, GP, TH1 + HP * RL, TLAT + HP * RL,
                         TDZ2 + HP * RL, TDZ1 + HP * RL, XNc, DM, DLS, DVAL);
    }
    PPO_WCLK(8);
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
    PPO_TICK(3);
    // ---- 2. the POLICY tower's weight gradients (and the action head's) by all eight warps ------------------------------
    tower_wgrad<PT / 32, HP>(0, lane, warp, h, Do, Da, pd.discrete != 0, wo_p, GP, TH1, TLAT, TDZ2, TDZ1, XNc, DM, DLS, DVAL);
    __syncthreads();
    PPO_TICK(7);
    // ---- 3. push the partials to the slice owners: RECV[this CTA][i], one 16-byte DSMEM store per quad ----------
    // CTA c starts with the quads owned by CTA c+1, so at any time the 8 senders target 8 different receivers
    for (int q = tid; q < CL * S / 4; q += PT) {
      int qq = q + ((crank + 1) & (CL - 1)) * (S / 4);
      if (qq >= CL * S / 4) qq -= CL * S / 4;
      const int p0 = 4 * qq, owner = (int)__umulhi((unsigned)qq, qmagic);  // = qq / (S / 4)
      st_async_v4(mapa_u32(recv_sa + (uint32_t)(crank * S + (p0 - owner * S)) * 4u, owner), ld4(GP + p0),
                  mapa_u32(xbar0_sa, owner));
    }
    if (loss_log) cluster.sync();  // (test / logging path only) the partial losses have landed in CTA 0
    if (crank == 0 && tid == 0 && loss_log) {
      float pg = 0.f, vl = 0.f, el = 0.f;
      for (int c = 0; c < CL; ++c) {
        pg += LOSS[c * 3 + 0];
        vl += LOSS[c * 3 + 1];
        el += LOSS[c * 3 + 2];
      }
      pg *= inv_nb, vl *= inv_nb, el *= inv_nb;
      loss_log[gs * 4 + 0] = pg;
      loss_log[gs * 4 + 1] = vl;
      loss_log[gs * 4 + 2] = el;
      loss_log[gs * 4 + 3] = pg + A.hp.ent_coef * el + A.hp.vf_coef * vl;
    }
    PPO_TICK(8);
    // wait until the 8 partials of the owned slice have landed
    mbar_wait(&xbar[0], (uint32_t)(gs & 1));
    if (tid == 0) mbar_expect_tx(&xbar[0], xbytes);  // re-arm for the next step
    PPO_TICK(9);

    // ---- 4. slice owners: sum the CL partials in fixed order (one quad per thread), exchange the squared slice
    //         norms for clip_grad_norm_ (4 bytes to every CTA, same st.async + mbarrier mechanism) ------------------------
    const int i0 = 4 * tid;
    const bool own = i0 < S;  // S / 4 <= PT is checked by the launcher
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    float ss = 0.f;
    if (own) {
      g = ld4(RECV + i0);
#pragma unroll
      for (int c = 1; c < CL; ++c) {  // fixed order: deterministic
        const float4 t = ld4(RECV + c * S + i0);
        g.x += t.x, g.y += t.y, g.z += t.z, g.w += t.w;
      }
      ss = (g.x * g.x + g.y * g.y) + (g.z * g.z + g.w * g.w);
    }
    const float my_ssq = block_sum(ss, red);
    if (tid < CL) st_async_f32(mapa_u32(ssq_sa + (uint32_t)crank * 4u, tid), my_ssq, mapa_u32(xbar2_sa, tid));
    PPO_TICK(10);
    mbar_wait(&xbar[2], (uint32_t)(gs & 1));
    if (tid == 0) mbar_expect_tx(&xbar[2], (uint32_t)(CL * 4));  // re-arm for the next step
    PPO_TICK(11);

    // ---- 5. clip_grad_norm_ + Adam on the OWNED slice (moments never leave their owner); the new parameters are
    //         all-gathered into every CTA's parameter vector ----------------------------------------------------------
    float total = 0.f;
#pragma unroll
    for (int c = 0; c < CL; ++c) total += SSQ[c];  // same order everywhere: the replicas' clip factors agree bit for bit
    total = sqrtf(total);
    float clip = A.hp.max_grad_norm / (total + 1e-6f);
    clip = clip > 1.0f ? 1.0f : clip;
    if (own) {
      const float step_size = bc[2 * cur], inv_bc2s = rcp_fast(bc[2 * cur + 1]);
      const int q0 = crank * S + i0;
      const float4 m4 = ld4(Ms + q0), v4 = ld4(Vs + q0), p4 = ld4(Pm + q0);
      // approximate sqrt / division (~1e-7 relative on an update that is itself ~lr relative to the weights)
      auto adam1 = [&](float gg, float& m, float& v, float& pw) {
        gg *= clip;
        m = m + (gg - m) * (1.0f - 0.9f);
        v = v * 0.999f + (1.0f - 0.999f) * gg * gg;
        pw -= step_size * __fdividef(m, fmaf(sqrt_fast(v), inv_bc2s, A.hp.adam_eps));
      };
      float4 m = m4, v = v4, pw = p4;
      adam1(g.x, m.x, v.x, pw.x);
      adam1(g.y, m.y, v.y, pw.y);
      adam1(g.z, m.z, v.z, pw.z);
      adam1(g.w, m.w, v.w, pw.w);
      st4(Ms + q0, m);
      st4(Vs + q0, v);
#pragma unroll
      for (int c = 0; c < CL; ++c) {  // rotated start: the 8 owners write to 8 different CTAs at a time
        const int dstc = (crank + c) & (CL - 1);
        st_async_v4(mapa_u32(pm_sa + (uint32_t)q0 * 4u, dstc), pw, mapa_u32(xbar1_sa, dstc));
      }
    }
    mbar_wait(&xbar[1], (uint32_t)(gs & 1));  // every CTA's new parameter slice has landed in Pm
    if (tid == 0) mbar_expect_tx(&xbar[1], xbytes);  // re-arm for the next step
    PPO_TICK(12);
    ep_now = ep_next;
    start = start_next;
    // (the barrier at the top of the next step orders these parameter writes before their first use)
  }
  __syncthreads();
#ifdef IMB_PPO_TIMING
  if (crank == 0 && tid == 0)
    for (int i = 0; i < 16; ++i) g_ppo_clk[i] = clk_acc[i];
  if (crank == 0 && lane == 0)
    for (int i = 0; i < 10; ++i) g_ppo_wclk[i * 8 + warp] = wacc[i];
#endif

  // ---- write back (CTA 0): parameters and moments in torch order, norm state, counters ------------------------------
  for (int p = tid; p < NP; p += PT) {  // moments live with their slice owner
    const int q = flat_to_play(pd, PL, p);
    if (q / S == crank) {
      g_m[p] = Ms[q];
      g_v[p] = Vs[q];
    }
  }
  if (crank == 0) {
    for (int p = tid; p < NP; p += PT) g_params[p] = Pm[flat_to_play(pd, PL, p)];
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

#include "imb_ppo_gen.cuh"

}  // namespace

static size_t ppo_smem_floats(const PpoArgs& A) {
  auto al = [](int x) { return (x + 31) / 32 * 32; };
  const int HP = A.HP, KP = A.KP, Da = A.pol.d_act, S = A.S;
  const int DAP = (Da + 3) / 4 * 4;
  size_t o = 0;
  o += 5 * (size_t)al(CL * S) + 32;
  o += 3 * (size_t)al(PR * A.RS2);
  o += 2 * (size_t)al(KP * RL) + (size_t)8 * HP * RL + (size_t)3 * DAP * RL + 32;
  o += al(2 * 64 + 4);
  return o;
}

// cluster launch of one of the two PPO kernels (CL CTAs of PT threads, one cluster)
template <typename K>
static int launch_cluster(K kernel, const char* name, size_t smem_bytes, size_t* attr_bytes, cudaStream_t st, const PpoArgs& A,
                          float* params, float* norm, int32_t* norm_count, float* m, float* v, const float* rollout,
                          const int64_t* perm, float* loss_log, int64_t* state) {
  IMB_REQUIRE(smem_bytes <= IMB_SMEM_MAX, "%s needs %zu B of shared memory per CTA (policy / minibatch too large)", name,
              smem_bytes);
  if (smem_bytes > *attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) IMB_FAIL(-2, "cudaFuncSetAttribute(%s): %s", name, cudaGetErrorString(e));
    *attr_bytes = smem_bytes;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL);
  cfg.blockDim = dim3(PT);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, A, params, norm, norm_count, m, v, rollout, perm, loss_log, state);
  if (e != cudaSuccess) IMB_FAIL(-2, "%s (cluster launch): %s", name, cudaGetErrorString(e));
  return 0;
}

static int launch_ppo(const PpoArgs& A0, float* params, float* norm, int32_t* norm_count, float* m, float* v,
                      const float* rollout, const int64_t* perm, float* loss_log, int64_t* state, cudaStream_t st) {
  PpoArgs A = A0;
  IMB_REQUIRE(A.hp.batch_size >= 1 && A.hp.batch_size <= GEN_MAX_MB, "PPO minibatch size must be in [1, %d]", GEN_MAX_MB);
  IMB_REQUIRE(A.rw % 4 == 0, "rollout row width must be a multiple of 4 floats (bulk row copies)");
  A.KP = A.pol.d_obs <= 32 ? 32 : 64;
  A.S = ((make_play(A.pol).total + CL - 1) / CL + 3) / 4 * 4;
  A.RS2 = ((A.rw + 4) % 8 == 4) ? A.rw + 4 : A.rw + 8;
  // IMB_PPO_FORCE_GENERAL=1 (tests): run the general kernel on shapes the specialised one covers
  const char* force = getenv("IMB_PPO_FORCE_GENERAL");
  const bool general = A.pol.hidden > 32 || A.hp.batch_size > PR || (force && force[0] == '1');
  if (!general) {
    // 64-row minibatch resident in shared memory, one lane per hidden unit (k_ppo_update)
    A.HP = 32;
    IMB_REQUIRE(A.S / 4 <= PT, "policy too large for the PPO update kernel (%d parameters per slice)", A.S);
    static size_t attr_bytes = 0;
    return launch_cluster(k_ppo_update<32>, "k_ppo_update", ppo_smem_floats(A) * 4, &attr_bytes, st, A, params, norm,
                          norm_count, m, v, rollout, perm, loss_log, state);
  }
  // tower width up to 64 and / or minibatches of more than 64 rows (k_ppo_update_gen)
  A.HP = A.pol.hidden <= 32 ? 32 : 64;
  const size_t bytes = (size_t)gen_layout(A.S, A.HP, A.KP, A.pol.d_act, A.hp.batch_size).total * 4;
  if (A.HP == 32) {
    static size_t attr_bytes = 0;
    return launch_cluster(k_ppo_update_gen<1>, "k_ppo_update_gen<1>", bytes, &attr_bytes, st, A, params, norm, norm_count, m,
                          v, rollout, perm, loss_log, state);
  }
  static size_t attr_bytes2 = 0;
  return launch_cluster(k_ppo_update_gen<2>, "k_ppo_update_gen<2>", bytes, &attr_bytes2, st, A, params, norm, norm_count, m,
                        v, rollout, perm, loss_log, state);
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

#ifdef IMB_PPO_TIMING
extern "C" __attribute__((visibility("default"))) int imb_debug_ppo_clocks(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_ppo_clk, 16 * sizeof(long long));
}
extern "C" __attribute__((visibility("default"))) int imb_debug_ppo_warp_clocks(long long* out, int reset) {
  static const long long zero[80] = {0};
  if (reset) return (int)cudaMemcpyToSymbol(g_ppo_wclk, zero, sizeof(zero));
  return (int)cudaMemcpyFromSymbol(out, g_ppo_wclk, 80 * sizeof(long long));
}
#endif
