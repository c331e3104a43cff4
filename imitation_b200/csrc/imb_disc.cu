// imb_disc.cu -- stage 3 of the GAIL/AIRL round: the discriminator update, fused.
//
// Replaces (reference, /root/reference/src/imitation): rewards/reward_nets.py:441-457
// (BasicRewardNet.forward), :701-736 (ShapedRewardNet.forward), util/networks.py:79-134
// (RunningNorm), algorithms/adversarial/common.py:353-372 (BCE-with-logits, backward, Adam)
// and :27-92 (compute_train_stats).
//
// Kernel plan (no spinning grid barriers -- every dependency is a kernel boundary or the
// "last block done" ticket, so a bug cannot hang the GPU):
//   k_norm_stats   per-chunk (n, mean, M2) per input feature; last CTA Chan-merges the chunks in
//                  fixed order into the running stats (RunningNorm.update_stats).
//   k_disc_fwdbwd  persistent CTAs over 128-row tiles of the feature-major batch.  The tile is
//                  staged [feature][row] into shared memory by cp.async.bulk (TMA unit) with a
//                  2-stage mbarrier pipeline.  Phase A (thread per row): normalise, MLP forward,
//                  BCE-with-logits, backward to dL/dz per layer, activations to smem tiles.
//                  Phase B (warps split output columns): the three weight-gradient contractions
//                  dW = D^T . Act over the tile, accumulated in shared memory across tiles.
//                  Weights (<34 KB) stay in shared memory; activations never touch HBM.
//   k_disc_reduce  warp-per-parameter deterministic sum of the per-CTA partials.
//   k_disc_adam    torch.optim.Adam step + the 9 train statistics.
#include "imb_common.cuh"
#include "imb_mlp.cuh"

thread_local char g_imb_err[512] = {0};

extern "C" int imb_version(void) { return 1; }
extern "C" const char* imb_last_error(void) { return g_imb_err; }

namespace {

constexpr int NORM_CHUNK = 2048;      // rows per CTA in k_norm_stats
constexpr int MAXG = 296;             // max CTAs of k_disc_fwdbwd (2 per SM)

// ---- workspace layout (floats) -----------------------------------------------------------------
struct WsLayout {
  int64_t gacc;      // [P] accumulated gradient
  int64_t stats;     // [16] reduced sums of the last minibatch
  int64_t meta;      // [16] ints: grid of the last fwdbwd launch, n rows, n_expert
  int64_t snap;      // [2*IMB_MAX_DIN] potential-norm stats after the first (next_obs) update
  int64_t ticket;    // [16] uint tickets
  int64_t normpart;  // [MAXCHUNKS][2*IMB_MAX_DIN + 4]
  int64_t partial;   // [MAXG][P + 16]
  int64_t total;
};
constexpr int MAXCHUNKS = 4096;  // up to 8M rows per norm launch
__host__ __device__ inline int64_t part_stride(int P) { return (int64_t)((P + 16 + 31) / 32) * 32; }
inline WsLayout ws_layout(int P) {
  WsLayout w;
  int64_t o = 0;
  w.gacc = o;
  o += (P + 31) / 32 * 32;
  w.stats = o;
  o += 32;
  w.meta = o;
  o += 32;
  w.snap = o;
  o += 2 * IMB_MAX_DIN;
  w.ticket = o;
  o += 32;
  w.normpart = o;
  o += (int64_t)MAXCHUNKS * (2 * IMB_MAX_DIN + 4);
  w.partial = o;
  o += (int64_t)MAXG * part_stride(P);
  w.total = o;
  return w;
}

// ---- RunningNorm statistics ----------------------------------------------------------------------
struct NormLaunch {
  int din;
  short row[IMB_MAX_DIN];  // batch feature rows
};

// One CTA per NORM_CHUNK rows; warp w handles features w, w+nw, ...: exact two-pass (mean, M2)
// inside the chunk, then the last CTA to finish merges all chunks in index order (Chan et al.)
// and folds the batch into the running statistics exactly as util/networks.py:111-134 does.
__global__ void __launch_bounds__(256) k_norm_stats(NormLaunch L, const float* __restrict__ batch, int64_t ld,
                                                    int64_t n, float* __restrict__ run_mean_var,
                                                    int32_t* __restrict__ count, float* __restrict__ snap_out,
                                                    float* __restrict__ part, unsigned int* __restrict__ ticket) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * NORM_CHUNK;
  const int64_t r1 = min(n, r0 + (int64_t)NORM_CHUNK);
  const int cn = (int)(r1 - r0);
  const int PS = 2 * IMB_MAX_DIN + 4;
  float* my = part + (int64_t)blockIdx.x * PS;
  for (int k = warp; k < L.din; k += nw) {
    const float* src = batch + (int64_t)L.row[k] * ld + r0;
    float s = 0.f;
    for (int i = lane; i < cn; i += 32) s += src[i];
    s = warp_sum(s);
    const float mean = s / (float)cn;
    float m2 = 0.f;
    for (int i = lane; i < cn; i += 32) {
      float dlt = src[i] - mean;
      m2 = fmaf(dlt, dlt, m2);
    }
    m2 = warp_sum(m2);
    if (lane == 0) {
      my[k] = mean;
      my[IMB_MAX_DIN + k] = m2;
    }
  }
  if (threadIdx.x == 0) my[2 * IMB_MAX_DIN] = (float)cn;
  __threadfence();
  __shared__ bool is_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const int32_t old_count = *count;
  for (int k = threadIdx.x; k < L.din; k += blockDim.x) {
    float na = 0.f, ma = 0.f, m2a = 0.f;
    for (unsigned int c = 0; c < gridDim.x; ++c) {
      const float* p = part + (int64_t)c * PS;
      const float nb = __ldcg(p + 2 * IMB_MAX_DIN), mb = __ldcg(p + k), m2b = __ldcg(p + IMB_MAX_DIN + k);
      const float nt = na + nb;
      const float dlt = mb - ma;
      ma = ma + dlt * (nb / nt);
      m2a = m2a + m2b + dlt * dlt * (na * nb / nt);
      na = nt;
    }
    const float b_mean = ma, b_var = m2a / na, b_n = na;
    float mean = run_mean_var[k], var = run_mean_var[L.din + k];
    const float cnt = (float)old_count;
    const float tot = cnt + b_n;
    const float delta = b_mean - mean;
    mean += delta * b_n / tot;
    var *= cnt;
    var += b_var * b_n;
    var += delta * delta * cnt * b_n / tot;
    var /= tot;
    run_mean_var[k] = mean;
    run_mean_var[L.din + k] = var;
    if (snap_out) {
      snap_out[k] = mean;
      snap_out[L.din + k] = var;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    *count = old_count + (int32_t)n;
    *ticket = 0u;  // re-arm for the next launch
  }
}

template <int H>
struct TileCfg {
  static constexpr int R = (H == 32) ? 128 : 64;  // rows per tile
  static constexpr int LD = H + 1;                // activation tile row stride (odd -> conflict-free)
};

// ---- the fused forward / BCE / backward kernel ---------------------------------------------------
// Dynamic shared memory carve-up (floats):
//   [mlp images for each distinct MLP][AW: P accumulators][xs: 2 stages x nstage x 128]
//   [XN: R x xn_ld][H1: R x LD][H2: R x LD][DZ1: R x LD][gv: R][red: 32]
template <int H>
__global__ void __launch_bounds__(NT) k_disc_fwdbwd(const DiscLaunch L, const float* __restrict__ params,
                                                   const float* __restrict__ batch, int64_t ld, int64_t n,
                                                   int64_t n_expert, float loss_scale,
                                                   const float* __restrict__ grad_out,
                                                   float* __restrict__ logits_out, float* __restrict__ partial,
                                                   int n_mlp_images, int img1_off, int aw_off, int xs_off,
                                                   int xn_off, int xn_ld, int tiles_off) {
  constexpr int R = TileCfg<H>::R;
  constexpr int LD = TileCfg<H>::LD;
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bars[2];
  float* img[MAX_PASS];
  img[0] = smem;
  img[1] = smem + img1_off;  // potential image using pass-1 norm (Phi(s'))
  img[2] = smem + img1_off;  // pass 2 shares weights; its norm constants live right after (see below)
  float* AW = smem + aw_off;
  float* xs = smem + xs_off;
  float* XN = smem + xn_off;
  float* H1 = smem + tiles_off;
  float* H2 = H1 + R * LD;
  float* DZ1 = H2 + R * LD;
  float* gv = DZ1 + R * LD;
  float* red = gv + R;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // -- one-time: weights -> smem, zero accumulators, init barriers -------------------------------
  load_mlp<H>(img[0], L.pass[0], params);
  // second normalisation table for Phi(s) (pass 2): stored after image 1
  float* mean2 = nullptr;
  float* istd2 = nullptr;
  if (L.npass == 3) {
    load_mlp<H>(img[1], L.pass[1], params);
    const int din = L.pass[2].din;
    mean2 = img[1] + MlpSm<H>::size(din);
    istd2 = mean2 + IMB_MAX_DIN;
    for (int i = tid; i < din; i += NT) {
      if (L.pass[2].has_norm) {
        mean2[i] = L.pass[2].norm[i];
        istd2[i] = 1.0f / sqrtf(L.pass[2].norm[din + i] + L.pass[2].eps);
      } else {
        mean2[i] = 0.f;
        istd2[i] = 1.f;
      }
    }
  }
  for (int i = tid; i < L.P; i += NT) AW[i] = 0.f;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();

  const int64_t ntiles = (n + R - 1) / R;
  const uint32_t stage_bytes = (uint32_t)L.nstage * R * 4u;
  auto issue = [&](int64_t tile, int stage) {
    // one elected thread arms the barrier and issues one bulk copy per staged feature row
    mbar_expect_tx(&bars[stage], stage_bytes);
    float* dst = xs + (size_t)stage * L.nstage * XS_LD;
    for (int s = 0; s < L.nstage; ++s)
      bulk_g2s(dst + s * XS_LD, batch + (int64_t)L.stage_row[s] * ld + tile * R, R * 4u, &bars[stage]);
  };
  uint32_t phase[2] = {0u, 0u};
  if (tid == 0 && (int64_t)blockIdx.x < ntiles) issue(blockIdx.x, 0);

  float s_loss = 0.f, s_ent = 0.f;
  int c_exp = 0, c_gen = 0, c_pred_exp = 0;

  int it = 0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int stage = it & 1;
    const int64_t next = tile + gridDim.x;
    if (tid == 0 && next < ntiles) issue(next, stage ^ 1);  // prefetch while this tile computes
    mbar_wait(&bars[stage], phase[stage]);
    phase[stage] ^= 1u;
    const float* x = xs + (size_t)stage * L.nstage * XS_LD;
    const int64_t row = tile * R + tid;
    const bool active = (tid < R) && (row < n);
    const float done = (L.done_slot >= 0 && active) ? x[L.done_slot * XS_LD + tid] : 0.f;

    // ---- logit: forward sweep over the passes (single pass: folded into phase A below) --------
    float g = 0.f;  // dL/dlogit for this row
    float h1[H], h2[H];
    if (L.npass > 1) {
      float logit = 0.f;
      if (active) {
        for (int p = 0; p < L.npass; ++p) {
          const PassDesc& P = L.pass[p];
          const float* mean = (p == 2) ? mean2 : img[p] + MlpSm<H>::mean_off(P.din);
          const float* istd = (p == 2) ? istd2 : img[p] + MlpSm<H>::istd_off(P.din);
          float* xn = XN + tid * xn_ld;
          for (int k = 0; k < P.din; ++k) xn[k] = (x[P.in_slot[k] * XS_LD + tid] - mean[k]) * istd[k];
          const float out = mlp_forward_row<H, false>(img[p], P, xn, h1, h2);
          logit = fmaf(pass_coef(P.coef_kind, L.gamma, done), out, logit);
        }
        if (L.logp_slot >= 0) logit -= x[L.logp_slot * XS_LD + tid];
      }
      if (active) {
        if (logits_out) logits_out[row] = logit;
        if (grad_out) {
          g = grad_out[row];
        } else {
          const float y = (row < n_expert) ? 1.f : 0.f;
          const float sg = sigmoid_f(logit);
          const float sp = fmaxf(logit, 0.f) + log1pf(expf(-fabsf(logit)));
          s_loss += sp - logit * y;
          s_ent += sp - logit * sg;
          const bool pred_exp = !(logit < 0.f);
          c_pred_exp += pred_exp;
          if (y > 0.5f) c_exp += pred_exp; else c_gen += !pred_exp;
          g = (sg - y) * loss_scale;
        }
      }
    }

    for (int p = 0; p < L.npass; ++p) {
      const PassDesc& P = L.pass[p];
      const float* sm = img[p];
      const float* mean = (p == 2) ? mean2 : sm + MlpSm<H>::mean_off(P.din);
      const float* istd = (p == 2) ? istd2 : sm + MlpSm<H>::istd_off(P.din);
      // ---------------- phase A: thread per row ------------------------------------------------
      float gp = 0.f;
      if (active) {
        float* xn = XN + tid * xn_ld;
        for (int k = 0; k < P.din; ++k) xn[k] = (x[P.in_slot[k] * XS_LD + tid] - mean[k]) * istd[k];
        const float out = mlp_forward_row<H, true>(sm, P, xn, h1, h2);
        if (L.npass == 1) {
          float logit = out;
          if (L.logp_slot >= 0) logit -= x[L.logp_slot * XS_LD + tid];
          if (logits_out) logits_out[row] = logit;
          if (grad_out) {
            g = grad_out[row];
          } else {
            const float y = (row < n_expert) ? 1.f : 0.f;
            const float sg = sigmoid_f(logit);
            const float sp = fmaxf(logit, 0.f) + log1pf(expf(-fabsf(logit)));
            s_loss += sp - logit * y;
            s_ent += sp - logit * sg;
            const bool pred_exp = !(logit < 0.f);
            c_pred_exp += pred_exp;
            if (y > 0.5f) c_exp += pred_exp; else c_gen += !pred_exp;
            g = (sg - y) * loss_scale;
          }
        }
        gp = g * pass_coef(P.coef_kind, L.gamma, done);
        // backward to dL/dz1 (and keep h1, h2 for the weight gradients)
        const float* wf = sm + MlpSm<H>::wf_off(P.din);
        if (P.n_hidden == 2) {
          const float* W2 = sm + MlpSm<H>::w2_off(P.din);
          float dz1[H];
#pragma unroll
          for (int i = 0; i < H; ++i) dz1[i] = 0.f;
#pragma unroll
          for (int j = 0; j < H; ++j) {
            const float dz2 = (h2[j] > 0.f) ? gp * wf[j] : 0.f;
            const float4* w = reinterpret_cast<const float4*>(W2 + j * H);
#pragma unroll
            for (int i4 = 0; i4 < H / 4; ++i4) {
              const float4 ww = w[i4];
              dz1[4 * i4 + 0] = fmaf(ww.x, dz2, dz1[4 * i4 + 0]);
              dz1[4 * i4 + 1] = fmaf(ww.y, dz2, dz1[4 * i4 + 1]);
              dz1[4 * i4 + 2] = fmaf(ww.z, dz2, dz1[4 * i4 + 2]);
              dz1[4 * i4 + 3] = fmaf(ww.w, dz2, dz1[4 * i4 + 3]);
            }
          }
#pragma unroll
          for (int i = 0; i < H; ++i) {
            H1[tid * LD + i] = h1[i];
            H2[tid * LD + i] = h2[i];
            DZ1[tid * LD + i] = (h1[i] > 0.f) ? dz1[i] : 0.f;
          }
        } else if (P.n_hidden == 1) {
#pragma unroll
          for (int i = 0; i < H; ++i) {
            H1[tid * LD + i] = h1[i];
            DZ1[tid * LD + i] = (h1[i] > 0.f) ? gp * wf[i] : 0.f;
          }
        }
      } else if (tid < R) {
        // inactive (padding) rows contribute zeros
        float* xn = XN + tid * xn_ld;
        for (int k = 0; k < P.din; ++k) xn[k] = 0.f;
#pragma unroll
        for (int i = 0; i < H; ++i) {
          H1[tid * LD + i] = 0.f;
          H2[tid * LD + i] = 0.f;
          DZ1[tid * LD + i] = 0.f;
        }
      }
      if (tid < R) gv[tid] = gp;
      __syncthreads();

      // ---------------- phase B: weight gradients, warps split the output columns ----------------
      float* A = AW + P.param_off;
      const int din = P.din;
      constexpr int JW = H / 4;  // output units per warp
      const int j0 = warp * JW;
      if (P.n_hidden == 0) {
        // dwf[k] = sum_r g_r xn[r][k]; dbf = sum_r g_r
        for (int k = tid; k <= din; k += NT) {
          float acc = 0.f;
          if (k < din) {
            for (int r = 0; r < R; ++r) acc = fmaf(gv[r], XN[r * xn_ld + k], acc);
          } else {
            for (int r = 0; r < R; ++r) acc += gv[r];
          }
          A[k] += acc;
        }
      } else {
        const int h1w = P.h1;
        int off_w1 = 0, off_b1 = h1w * din, off_w2 = off_b1 + h1w;
        int off_b2 = off_w2 + ((P.n_hidden == 2) ? P.h2 * h1w : 0);
        int off_wf = (P.n_hidden == 2) ? off_b2 + P.h2 : off_w2;
        const int hl = (P.n_hidden == 2) ? P.h2 : h1w;
        // (1) dW1[j][k] = sum_r dz1[r][j] xn[r][k];  db1[j] = sum_r dz1[r][j]
        {
          float acc[JW][2];
          float bs[JW];
#pragma unroll
          for (int jj = 0; jj < JW; ++jj) {
            acc[jj][0] = acc[jj][1] = 0.f;
            bs[jj] = 0.f;
          }
          const bool k0ok = lane < din, k1ok = (lane + 32) < din;
          for (int r = 0; r < R; ++r) {
            const float a0 = k0ok ? XN[r * xn_ld + lane] : 0.f;
            const float a1 = k1ok ? XN[r * xn_ld + lane + 32] : 0.f;
#pragma unroll
            for (int jj = 0; jj < JW; ++jj) {
              const float dz = DZ1[r * LD + j0 + jj];
              acc[jj][0] = fmaf(dz, a0, acc[jj][0]);
              acc[jj][1] = fmaf(dz, a1, acc[jj][1]);
              bs[jj] += dz;
            }
          }
#pragma unroll
          for (int jj = 0; jj < JW; ++jj) {
            const int j = j0 + jj;
            if (j < h1w) {
              if (k0ok) A[off_w1 + j * din + lane] += acc[jj][0];
              if (k1ok) A[off_w1 + j * din + lane + 32] += acc[jj][1];
              if (lane == 0) A[off_b1 + j] += bs[jj];
            }
          }
        }
        // (2) dW2[j][i] = sum_r dz2[r][j] h1[r][i];  db2[j] = sum_r dz2[r][j]
        if (P.n_hidden == 2) {
          const float* wf = sm + MlpSm<H>::wf_off(din);
          float wfr[JW];
#pragma unroll
          for (int jj = 0; jj < JW; ++jj) wfr[jj] = wf[j0 + jj];
          float acc[JW][H / 32];
          float bs[JW];
#pragma unroll
          for (int jj = 0; jj < JW; ++jj) {
#pragma unroll
            for (int ii = 0; ii < H / 32; ++ii) acc[jj][ii] = 0.f;
            bs[jj] = 0.f;
          }
          for (int r = 0; r < R; ++r) {
            const float gr = gv[r];
            float a[H / 32];
#pragma unroll
            for (int ii = 0; ii < H / 32; ++ii) a[ii] = H1[r * LD + lane + 32 * ii];
#pragma unroll
            for (int jj = 0; jj < JW; ++jj) {
              const float dz2 = (H2[r * LD + j0 + jj] > 0.f) ? gr * wfr[jj] : 0.f;
#pragma unroll
              for (int ii = 0; ii < H / 32; ++ii) acc[jj][ii] = fmaf(dz2, a[ii], acc[jj][ii]);
              bs[jj] += dz2;
            }
          }
#pragma unroll
          for (int jj = 0; jj < JW; ++jj) {
            const int j = j0 + jj;
            if (j < P.h2) {
#pragma unroll
              for (int ii = 0; ii < H / 32; ++ii) {
                const int i = lane + 32 * ii;
                if (i < h1w) A[off_w2 + j * h1w + i] += acc[jj][ii];
              }
              if (lane == 0) A[off_b2 + j] += bs[jj];
            }
          }
        }
        // (3) dwf[j] = sum_r g_r hlast[r][j];  dbf = sum_r g_r
        {
          const float* HL = (P.n_hidden == 2) ? H2 : H1;
          for (int j = tid; j <= hl; j += NT) {
            float acc = 0.f;
            if (j < hl) {
              for (int r = 0; r < R; ++r) acc = fmaf(gv[r], HL[r * LD + j], acc);
            } else {
              for (int r = 0; r < R; ++r) acc += gv[r];
            }
            A[off_wf + j] += acc;
          }
        }
      }
      __syncthreads();
    }
  }

  // ---- per-CTA partials: gradients + statistics -------------------------------------------------
  float* my = partial + (int64_t)blockIdx.x * part_stride(L.P);
  for (int i = tid; i < L.P; i += NT) my[i] = AW[i];
  s_loss = warp_sum(s_loss);
  s_ent = warp_sum(s_ent);
  c_exp = warp_sum_i(c_exp);
  c_gen = warp_sum_i(c_gen);
  c_pred_exp = warp_sum_i(c_pred_exp);
  if (lane == 0) {
    red[warp * 5 + 0] = s_loss;
    red[warp * 5 + 1] = s_ent;
    red[warp * 5 + 2] = (float)c_exp;
    red[warp * 5 + 3] = (float)c_gen;
    red[warp * 5 + 4] = (float)c_pred_exp;
  }
  __syncthreads();
  if (tid < 5) {
    float v = 0.f;
    for (int w = 0; w < NT / 32; ++w) v += red[w * 5 + tid];
    my[L.P + tid] = v;
  }
}

// ---- deterministic reduction of the per-CTA partials -----------------------------------------
// warp per parameter (and per statistic): lanes stride over the G partial rows, shuffle-reduce.
__global__ void __launch_bounds__(256) k_disc_reduce(int P, int G, const float* __restrict__ partial,
                                                    float* __restrict__ gacc, float* __restrict__ stats,
                                                    float* __restrict__ grad_out_flat) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int64_t ps = part_stride(P);
  for (int p = gw; p < P + 5; p += nwarps) {
    float acc = 0.f;
    for (int c = lane; c < G; c += 32) acc += partial[(int64_t)c * ps + p];
    acc = warp_sum(acc);
    if (lane == 0) {
      if (p < P) {
        const float v = gacc[p] + acc;
        gacc[p] = v;
        if (grad_out_flat) grad_out_flat[p] = v;
      } else {
        stats[p - P] = acc;  // statistics of the LAST minibatch only (common.py:376-381)
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_disc_adam(int P, imb_adam opt, float* __restrict__ params,
                                                  float* __restrict__ m, float* __restrict__ v,
                                                  const float* __restrict__ grad, float grad_div,
                                                  const float* __restrict__ stats, const int* __restrict__ meta,
                                                  const int64_t* __restrict__ state,
                                                  float* __restrict__ stats_out) {
  // bias corrections in double like torch's Python-scalar arithmetic (torch/optim/adam.py)
  const int64_t step = state[IMB_ST_DISC_STEP] + 1;
  const double bc1d = 1.0 - pow((double)opt.beta1, (double)step);
  const double bc2d = 1.0 - pow((double)opt.beta2, (double)step);
  const float step_size = (float)((double)opt.lr / bc1d);
  const float bc2_sqrt = (float)sqrt(bc2d);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P) {
    const float g = grad[i] / grad_div;
    const float mi = m[i] + (g - m[i]) * (1.0f - opt.beta1);        // torch: exp_avg.lerp_(grad, 1-beta1)
    const float vi = v[i] * opt.beta2 + (1.0f - opt.beta2) * g * g;  // exp_avg_sq.mul_(b2).addcmul_(g,g,1-b2)
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + opt.eps;
    params[i] -= step_size * (mi / denom);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && stats_out) {
    const float n = (float)meta[1], n_exp = (float)meta[2], n_gen = n - n_exp;
    const float loss_sum = stats[0], ent_sum = stats[1], c_exp = stats[2], c_gen = stats[3], c_pred = stats[4];
    const float nanv = __int_as_float(0x7fc00000);
    stats_out[0] = loss_sum * reinterpret_cast<const float*>(meta)[3];                 // disc_loss (scaled minibatch mean)
    stats_out[1] = n > 0 ? (c_exp + c_gen) / n : nanv;         // disc_acc
    stats_out[2] = n_exp >= 1 ? c_exp / n_exp : nanv;          // disc_acc_expert
    stats_out[3] = c_gen / fmaxf(1.f, n_gen);                  // disc_acc_gen
    stats_out[4] = n > 0 ? ent_sum / n : nanv;                 // disc_entropy
    stats_out[5] = n > 0 ? n_exp / n : nanv;                   // disc_proportion_expert_true
    stats_out[6] = n > 0 ? c_pred / n : nanv;                  // disc_proportion_expert_pred
    stats_out[7] = n_exp;
    stats_out[8] = n_gen;
  }
}

__global__ void k_state_add(int64_t* state, int idx, int64_t v) { state[idx] += v; }
__global__ void k_set_meta(int* meta, int G, int64_t n, int64_t n_expert, float loss_scale) {
  meta[0] = G;
  meta[1] = (int)n;
  meta[2] = (int)n_expert;
  reinterpret_cast<float*>(meta)[3] = loss_scale;
}

// ---- forward only (reward relabel / predict) ---------------------------------------------------
// thread per row straight from global memory (coalesced over rows for each feature).
template <int H>
__global__ void __launch_bounds__(NT) k_reward_fwd(const DiscLaunch L, const float* __restrict__ params,
                                                  const float* __restrict__ batch, int64_t ld, int64_t n,
                                                  int out_mode, float* __restrict__ out, int img1_off, int xn_off,
                                                  int xn_ld) {
  extern __shared__ __align__(128) float smem[];
  float* img[MAX_PASS] = {smem, smem + img1_off, smem + img1_off};
  float* XN = smem + xn_off;
  const int tid = threadIdx.x;
  load_mlp<H>(img[0], L.pass[0], params);
  float* mean2 = nullptr;
  float* istd2 = nullptr;
  if (L.npass == 3) {
    load_mlp<H>(img[1], L.pass[1], params);
    const int din = L.pass[2].din;
    mean2 = img[1] + MlpSm<H>::size(din);
    istd2 = mean2 + IMB_MAX_DIN;
    for (int i = tid; i < din; i += NT) {
      if (L.pass[2].has_norm) {
        mean2[i] = L.pass[2].norm[i];
        istd2[i] = 1.0f / sqrtf(L.pass[2].norm[din + i] + L.pass[2].eps);
      } else {
        mean2[i] = 0.f;
        istd2[i] = 1.f;
      }
    }
  }
  __syncthreads();
  float h1[H], h2[H];
  for (int64_t row = (int64_t)blockIdx.x * NT + tid; row < n; row += (int64_t)gridDim.x * NT) {
    const float done = (L.done_slot >= 0) ? batch[(int64_t)L.stage_row[L.done_slot] * ld + row] : 0.f;
    float logit = 0.f;
    for (int p = 0; p < L.npass; ++p) {
      const PassDesc& P = L.pass[p];
      const float* mean = (p == 2) ? mean2 : img[p] + MlpSm<H>::mean_off(P.din);
      const float* istd = (p == 2) ? istd2 : img[p] + MlpSm<H>::istd_off(P.din);
      float* xn = XN + tid * xn_ld;
      for (int k = 0; k < P.din; ++k)
        xn[k] = (batch[(int64_t)L.stage_row[P.in_slot[k]] * ld + row] - mean[k]) * istd[k];
      const float o = mlp_forward_row<H, false>(img[p], P, xn, h1, h2);
      logit = fmaf(pass_coef(P.coef_kind, L.gamma, done), o, logit);
    }
    if (out_mode >= 1 && L.logp_slot >= 0) logit -= batch[(int64_t)L.stage_row[L.logp_slot] * ld + row];
    out[row] = (out_mode == 2) ? softplus_f(logit) : logit;
  }
}

// ---- NormalizedRewardNet.predict_processed over consecutive env steps ---------------------------
// single CTA: for t in steps: normalise the E rewards of step t with the running stats, then merge
// step t's raw rewards into the stats (reward_nets.py:637-671 + networks.py:111-134).
__global__ void __launch_bounds__(1024) k_reward_norm_scan(float* __restrict__ rews, int64_t E, int64_t T,
                                                          int64_t step_stride, int64_t env_stride,
                                                          float* __restrict__ mv, int32_t* __restrict__ count,
                                                          float eps, int update) {
  __shared__ float red[64];
  __shared__ float bc[2];
  float mean = mv[0], var = mv[1];
  int32_t cnt = *count;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  for (int64_t t = 0; t < T; ++t) {
    float* r = rews + t * step_stride;
    const float istd = 1.0f / sqrtf(var + eps);
    float s = 0.f;
    for (int64_t e = tid; e < E; e += blockDim.x) s += r[e * env_stride];
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (tid == 0) {
      float a = 0.f;
      for (int w = 0; w < nw; ++w) a += red[w];
      bc[0] = a / (float)E;
    }
    __syncthreads();
    const float bmean = bc[0];
    float m2 = 0.f;
    for (int64_t e = tid; e < E; e += blockDim.x) {
      const float v = r[e * env_stride];
      const float dlt = v - bmean;
      m2 = fmaf(dlt, dlt, m2);
      r[e * env_stride] = (v - mean) * istd;  // normalise with the stats BEFORE this step's update
    }
    m2 = warp_sum(m2);
    if (lane == 0) red[32 + warp] = m2;
    __syncthreads();
    if (tid == 0) {
      float a = 0.f;
      for (int w = 0; w < nw; ++w) a += red[32 + w];
      bc[1] = a / (float)E;
    }
    __syncthreads();
    if (update) {
      const float bvar = bc[1], bn = (float)E, c = (float)cnt, tot = c + bn;
      const float delta = bmean - mean;
      mean += delta * bn / tot;
      var *= c;
      var += bvar * bn;
      var += delta * delta * c * bn / tot;
      var /= tot;
      cnt += (int32_t)E;
    }
    __syncthreads();
  }
  if (tid == 0 && update) {
    mv[0] = mean;
    mv[1] = var;
    *count = cnt;
  }
}

// ---- host-side launch helpers -----------------------------------------------------------------
template <int H>
struct SmemPlan {
  int img1_off, aw_off, xs_off, xn_off, xn_ld, tiles_off, total_floats;
};
template <int H>
SmemPlan<H> plan_smem(const DiscLaunch& L, bool train) {
  SmemPlan<H> s;
  constexpr int R = TileCfg<H>::R;
  constexpr int LD = TileCfg<H>::LD;
  auto al = [](int x) { return (x + 31) / 32 * 32; };
  int o = al(MlpSm<H>::size(L.pass[0].din));
  s.img1_off = o;
  if (L.npass == 3) o += al(MlpSm<H>::size(L.pass[1].din) + 2 * IMB_MAX_DIN);
  int maxdin = 0;
  for (int p = 0; p < L.npass; ++p) maxdin = L.pass[p].din > maxdin ? L.pass[p].din : maxdin;
  s.xn_ld = maxdin | 1;
  if (train) {
    s.aw_off = o;
    o += al(L.P);
    s.xs_off = o;
    o += al(2 * L.nstage * XS_LD);
    s.xn_off = o;
    o += al(R * s.xn_ld);
    s.tiles_off = o;
    o += al(3 * R * LD + R + 32);
  } else {
    s.aw_off = s.xs_off = s.tiles_off = 0;
    s.xn_off = o;
    o += al(NT * s.xn_ld);
  }
  s.total_floats = o;
  return s;
}

inline int pick_H(const DiscLaunch& L) {
  int h = 0;
  for (int p = 0; p < L.npass; ++p) {
    if (L.pass[p].n_hidden >= 1 && L.pass[p].h1 > h) h = L.pass[p].h1;
    if (L.pass[p].n_hidden >= 2 && L.pass[p].h2 > h) h = L.pass[p].h2;
  }
  return h <= 32 ? 32 : 64;
}

}  // namespace

extern "C" int64_t imb_disc_workspace_floats(const imb_disc_desc* d) { return ws_layout(d->n_params).total; }

static int norm_launch(const imb_mlp& m, const short* rows, const float* batch, int64_t ld, int64_t n,
                       float* norm_state, int32_t* norm_count, float* snap, float* ws, const WsLayout& w,
                       cudaStream_t st) {
  NormLaunch NL;
  NL.din = m.din;
  for (int k = 0; k < m.din; ++k) NL.row[k] = rows[k];
  const int chunks = (int)((n + NORM_CHUNK - 1) / NORM_CHUNK);
  IMB_REQUIRE(chunks >= 1 && chunks <= MAXCHUNKS, "norm update: n=%lld out of range", (long long)n);
  k_norm_stats<<<chunks, 256, 0, st>>>(NL, batch, ld, n, norm_state + m.norm_off, norm_count + m.count_idx, snap,
                                       ws + w.normpart, reinterpret_cast<unsigned int*>(ws + w.ticket));
  IMB_CHECK_LAUNCH("k_norm_stats");
  return 0;
}

extern "C" int imb_disc_norm_update(const imb_disc_desc* d, const float* batch, int64_t ld, int64_t n,
                                    float* norm_state, int32_t* norm_count, float* ws, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  IMB_REQUIRE(n >= 1, "norm update needs n >= 1");
  const WsLayout w = ws_layout(d->n_params);
  DiscLaunch L;
  if (int rc = build_launch(d, norm_state, nullptr, L)) return rc;
  short rows[IMB_MAX_DIN];
  if (d->base.has_norm) {
    for (int k = 0; k < d->base.din; ++k) rows[k] = L.stage_row[L.pass[0].in_slot[k]];
    if (int rc = norm_launch(d->base, rows, batch, ld, n, norm_state, norm_count, nullptr, ws, w, st)) return rc;
  }
  if (d->shaped && d->potential.has_norm) {
    // reference order: Phi(next_state) first, then Phi(state); both update the same RunningNorm
    for (int k = 0; k < d->potential.din; ++k) rows[k] = L.stage_row[L.pass[1].in_slot[k]];
    if (int rc = norm_launch(d->potential, rows, batch, ld, n, norm_state, norm_count, ws + w.snap, ws, w, st))
      return rc;
    for (int k = 0; k < d->potential.din; ++k) rows[k] = L.stage_row[L.pass[2].in_slot[k]];
    if (int rc = norm_launch(d->potential, rows, batch, ld, n, norm_state, norm_count, nullptr, ws, w, st)) return rc;
  }
  return 0;
}

template <int H>
static int launch_fwdbwd(const DiscLaunch& L, const float* params, const float* batch, int64_t ld, int64_t n,
                         int64_t n_expert, float loss_scale, const float* grad_out, float* logits_out, float* ws,
                         const WsLayout& w, cudaStream_t st) {
  constexpr int R = TileCfg<H>::R;
  const SmemPlan<H> s = plan_smem<H>(L, true);
  const size_t bytes = (size_t)s.total_floats * 4;
  IMB_REQUIRE(bytes <= IMB_SMEM_MAX, "discriminator too large for the fused kernel (%zu B smem)", bytes);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_disc_fwdbwd<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, IMB_SMEM_MAX);
    if (e != cudaSuccess) IMB_FAIL(-2, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int64_t ntiles = (n + R - 1) / R;
  int per_sm = (int)((IMB_SMEM_MAX) / (bytes + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 2) per_sm = 2;
  int64_t G = (int64_t)imb_num_sms() * per_sm;
  if (G > MAXG) G = MAXG;
  if (G > ntiles) G = ntiles;
  k_set_meta<<<1, 1, 0, st>>>(reinterpret_cast<int*>(ws + w.meta), (int)G, n, n_expert, loss_scale);
  k_disc_fwdbwd<H><<<(int)G, NT, bytes, st>>>(L, params, batch, ld, n, n_expert, loss_scale, grad_out, logits_out,
                                               ws + w.partial, L.npass, s.img1_off, s.aw_off, s.xs_off, s.xn_off,
                                               s.xn_ld, s.tiles_off);
  IMB_CHECK_LAUNCH("k_disc_fwdbwd");
  return (int)G;
}

// host mirror of the grid chosen by the last fwd/bwd launch (stream-ordered use only)
static thread_local int g_last_grid = 0;

extern "C" int imb_disc_fwd_bwd(const imb_disc_desc* d, const float* params, const float* norm_state,
                                const float* batch, int64_t ld, int64_t n, int64_t n_expert, float loss_scale,
                                const float* grad_out, float* logits_out, int flags, float* ws, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  IMB_REQUIRE(n >= 1, "fwd_bwd needs n >= 1");
  IMB_REQUIRE(ld % 4 == 0 && ld >= (n + IMB_TILE_ROWS - 1) / IMB_TILE_ROWS * IMB_TILE_ROWS,
              "batch leading dimension must cover n rounded up to %d rows", IMB_TILE_ROWS);
  const WsLayout w = ws_layout(d->n_params);
  DiscLaunch L;
  // in training mode the Phi(s') pass uses the stats snapshot taken between the two norm updates
  const bool snap = d->shaped && d->potential.has_norm && (flags & IMB_F_TRAIN_NORM);
  if (int rc = build_launch(d, norm_state, snap ? ws + w.snap : nullptr, L)) return rc;
  if (flags & IMB_F_ZERO_GRAD) {
    cudaError_t e = cudaMemsetAsync(ws + w.gacc, 0, sizeof(float) * d->n_params, st);
    if (e != cudaSuccess) IMB_FAIL(-2, "memset: %s", cudaGetErrorString(e));
  }
  int G = (pick_H(L) == 32) ? launch_fwdbwd<32>(L, params, batch, ld, n, n_expert, loss_scale, grad_out, logits_out,
                                                ws, w, st)
                            : launch_fwdbwd<64>(L, params, batch, ld, n, n_expert, loss_scale, grad_out, logits_out,
                                                ws, w, st);
  if (G < 0) return G;
  g_last_grid = G;
  return 0;
}

extern "C" int imb_disc_reduce(const imb_disc_desc* d, float* ws, float* grad_out_flat, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const WsLayout w = ws_layout(d->n_params);
  IMB_REQUIRE(g_last_grid > 0, "imb_disc_reduce called before imb_disc_fwd_bwd");
  const int P = d->n_params;
  const int warps = P + 5;
  int blocks = (warps * 32 + 255) / 256;
  if (blocks > 2 * imb_num_sms()) blocks = 2 * imb_num_sms();
  k_disc_reduce<<<blocks, 256, 0, st>>>(P, g_last_grid, ws + w.partial, ws + w.gacc, ws + w.stats, grad_out_flat);
  IMB_CHECK_LAUNCH("k_disc_reduce");
  return 0;
}

extern "C" int imb_disc_adam(const imb_disc_desc* d, const imb_adam* opt, float* params, float* exp_avg,
                             float* exp_avg_sq, const float* grad_flat_or_null, float grad_div, float* ws,
                             int64_t* state, float* stats_out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const WsLayout w = ws_layout(d->n_params);
  const int P = d->n_params;
  const float* grad = grad_flat_or_null ? grad_flat_or_null : ws + w.gacc;
  // loss statistic: sum * loss_scale recorded by the last fwd/bwd launch (meta[3])
  k_disc_adam<<<(P + 255) / 256, 256, 0, st>>>(P, *opt, params, exp_avg, exp_avg_sq, grad, grad_div, ws + w.stats,
                                               reinterpret_cast<const int*>(ws + w.meta), state, stats_out);
  IMB_CHECK_LAUNCH("k_disc_adam");
  k_state_add<<<1, 1, 0, st>>>(state, IMB_ST_DISC_STEP, 1);
  IMB_CHECK_LAUNCH("k_state_add");
  return 0;
}

template <int H>
static int launch_fwd(const DiscLaunch& L, const float* params, const float* batch, int64_t ld, int64_t n,
                      int out_mode, float* out, cudaStream_t st) {
  const SmemPlan<H> s = plan_smem<H>(L, false);
  const size_t bytes = (size_t)s.total_floats * 4;
  IMB_REQUIRE(bytes <= IMB_SMEM_MAX, "reward net too large for the fused kernel (%zu B smem)", bytes);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_reward_fwd<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, IMB_SMEM_MAX);
    if (e != cudaSuccess) IMB_FAIL(-2, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  int64_t blocks = (n + NT - 1) / NT;
  const int64_t cap = (int64_t)imb_num_sms() * 4;
  if (blocks > cap) blocks = cap;
  k_reward_fwd<H><<<(int)blocks, NT, bytes, st>>>(L, params, batch, ld, n, out_mode, out, s.img1_off, s.xn_off,
                                                   s.xn_ld);
  IMB_CHECK_LAUNCH("k_reward_fwd");
  return 0;
}

extern "C" int imb_reward_forward(const imb_disc_desc* d, const float* params, const float* norm_state,
                                  const float* batch, int64_t ld, int64_t n, int out_mode, float* out,
                                  void* stream) {
  if (n <= 0) return 0;
  DiscLaunch L;
  imb_disc_desc dd = *d;
  if (out_mode == 0) dd.subtract_logp = 0;
  if (int rc = build_launch(&dd, norm_state, nullptr, L)) return rc;
  return (pick_H(L) == 32) ? launch_fwd<32>(L, params, batch, ld, n, out_mode, out, (cudaStream_t)stream)
                           : launch_fwd<64>(L, params, batch, ld, n, out_mode, out, (cudaStream_t)stream);
}

extern "C" int imb_reward_norm_scan(float* rews, int64_t n_envs, int64_t n_steps, int64_t step_stride,
                                    int64_t env_stride, float* norm_state2, int32_t* norm_count, float eps,
                                    int update_stats, void* stream) {
  IMB_REQUIRE(n_envs >= 1 && n_steps >= 0, "bad sizes");
  if (n_steps == 0) return 0;
  int threads = 1024;
  while (threads > 32 && threads / 2 >= n_envs) threads /= 2;
  k_reward_norm_scan<<<1, threads, 0, (cudaStream_t)stream>>>(rews, n_envs, n_steps, step_stride, env_stride,
                                                              norm_state2, norm_count, eps, update_stats);
  IMB_CHECK_LAUNCH("k_reward_norm_scan");
  return 0;
}

extern "C" int imb_state_init(int64_t* state, void* stream) {
  cudaError_t e = cudaMemsetAsync(state, 0, sizeof(int64_t) * IMB_ST_WORDS, (cudaStream_t)stream);
  if (e != cudaSuccess) IMB_FAIL(-2, "memset: %s", cudaGetErrorString(e));
  return 0;
}
