// imb_mlp.cuh -- shared-memory MLP images and the per-row forward used by the discriminator
// kernels (imb_disc.cu) and by the reward relabel inside the rollout kernel (imb_rollout.cu).
#pragma once
#include "imb_common.cuh"

namespace {

constexpr int NT = 128;               // threads per CTA in the disc fwd/bwd kernel
constexpr int XS_LD = IMB_TILE_ROWS;  // staged tile: [slot][128 rows]
constexpr int MAX_STAGE_ROWS = 200;   // 2*64 + 64 + 2 = 194
constexpr int MAX_PASS = 3;

// ---- launch descriptor passed by value to the kernels --------------------------------------------
struct PassDesc {
  int din, n_hidden, h1, h2;
  int has_norm;
  float eps;
  int param_off;     // into flat params / accumulators
  int coef_kind;     // 0: +1, 1: +gamma*(1-done), 2: -1
  const float* norm; // [mean(din) | var(din)] to use for this pass (nullptr if none)
  unsigned char in_slot[IMB_MAX_DIN];  // staged slot of MLP input k
};
struct DiscLaunch {
  int npass;
  int nstage;        // staged feature rows
  int done_slot;     // staged slot of the done row (-1 if unused)
  int logp_slot;     // staged slot of the log pi row (-1 if unused)
  int P;             // total params
  float gamma;
  short stage_row[MAX_STAGE_ROWS];  // batch feature row of each staged slot
  PassDesc pass[MAX_PASS];
};

__host__ __attribute__((unused)) inline int mlp_params(const imb_mlp& m) {
  int hl = m.n_hidden == 0 ? m.din : (m.n_hidden == 1 ? m.h1 : m.h2);
  int p = 0;
  if (m.n_hidden >= 1) p += m.h1 * m.din + m.h1;
  if (m.n_hidden >= 2) p += m.h2 * m.h1 + m.h2;
  p += hl * m.n_out + m.n_out;
  return p;
}

// Build the launch descriptor: which batch feature rows are staged and how each pass maps its
// MLP inputs onto them.  Batch feature rows: obs [0,Do), act [Do,Do+Da), next_obs [Do+Da,2Do+Da),
// done 2Do+Da, logp 2Do+Da+1.
__host__ int build_launch(const imb_disc_desc* d, const float* norm_state, const float* snapA, DiscLaunch& L) {
  memset(&L, 0, sizeof(L));
  const int Do = d->d_obs, Da = d->d_act;
  const int row_obs = 0, row_act = Do, row_nobs = Do + Da, row_done = 2 * Do + Da, row_logp = row_done + 1;
  short slot_of[2 * IMB_MAX_DIN + IMB_MAX_DIN + 2 + 64];
  const int nrows = row_logp + 1;
  IMB_REQUIRE(nrows <= (int)(sizeof(slot_of) / sizeof(short)), "d_obs/d_act too large");
  for (int i = 0; i < nrows; ++i) slot_of[i] = -1;
  int ns = 0;
  auto need = [&](int row) {
    if (slot_of[row] < 0) {
      slot_of[row] = (short)ns;
      L.stage_row[ns++] = (short)row;
    }
    return (int)slot_of[row];
  };
  // pass 0: base net on the selected concat
  PassDesc& b = L.pass[0];
  int k = 0;
  auto push = [&](PassDesc& p, int row) -> int {
    if (k >= IMB_MAX_DIN) return -1;
    p.in_slot[k++] = (unsigned char)need(row);
    return 0;
  };
  if (d->use_state)
    for (int i = 0; i < Do; ++i)
      if (push(b, row_obs + i)) IMB_FAIL(-1, "base MLP input wider than %d", IMB_MAX_DIN);
  if (d->use_action)
    for (int i = 0; i < Da; ++i)
      if (push(b, row_act + i)) IMB_FAIL(-1, "base MLP input wider than %d", IMB_MAX_DIN);
  if (d->use_next_state)
    for (int i = 0; i < Do; ++i)
      if (push(b, row_nobs + i)) IMB_FAIL(-1, "base MLP input wider than %d", IMB_MAX_DIN);
  if (d->use_done)
    if (push(b, row_done)) IMB_FAIL(-1, "base MLP input wider than %d", IMB_MAX_DIN);
  IMB_REQUIRE(k == d->base.din, "base.din=%d does not match the selected inputs (%d)", d->base.din, k);
  auto fill = [&](PassDesc& p, const imb_mlp& m, int coef, const float* norm) {
    p.din = m.din;
    p.n_hidden = m.n_hidden;
    p.h1 = m.h1;
    p.h2 = m.h2;
    p.has_norm = m.has_norm;
    p.eps = m.norm_eps;
    p.param_off = m.param_off;
    p.coef_kind = coef;
    p.norm = m.has_norm ? norm : nullptr;
  };
  fill(b, d->base, 0, norm_state + d->base.norm_off);
  L.npass = 1;
  L.done_slot = -1;
  L.logp_slot = -1;
  if (d->shaped) {
    IMB_REQUIRE(d->potential.din == Do, "potential.din must equal d_obs");
    PassDesc& p1 = L.pass[1];  // Phi(s')   (evaluated first by the reference)
    k = 0;
    for (int i = 0; i < Do; ++i) push(p1, row_nobs + i);
    fill(p1, d->potential, 1, snapA ? snapA : norm_state + d->potential.norm_off);
    PassDesc& p2 = L.pass[2];  // Phi(s)
    k = 0;
    for (int i = 0; i < Do; ++i) push(p2, row_obs + i);
    fill(p2, d->potential, 2, norm_state + d->potential.norm_off);
    L.npass = 3;
    L.done_slot = need(row_done);
  }
  if (d->subtract_logp) L.logp_slot = need(row_logp);
  L.nstage = ns;
  L.P = d->n_params;
  L.gamma = d->gamma;
  IMB_REQUIRE(ns <= MAX_STAGE_ROWS, "too many staged rows");
  for (int p = 0; p < L.npass; ++p) {
    const PassDesc& q = L.pass[p];
    IMB_REQUIRE(q.n_hidden >= 0 && q.n_hidden <= 2, "n_hidden must be 0..2");
    IMB_REQUIRE(q.din >= 1 && q.din <= IMB_MAX_DIN, "din out of range");
    if (q.n_hidden >= 1) IMB_REQUIRE(q.h1 >= 1 && q.h1 <= IMB_MAX_HIDDEN, "h1 out of range");
    if (q.n_hidden >= 2) IMB_REQUIRE(q.h2 >= 1 && q.h2 <= IMB_MAX_HIDDEN, "h2 out of range");
  }
  return 0;
}

// ---- shared-memory image of one MLP -----------------------------------------------------------
// W1t[k][j] (din x H, zero padded), b1[H], W2[j][i] and W2t[i][j] (H x H), b2[H], wf[H], bf.
template <int H>
struct MlpSm {
  static constexpr int W1T = 0;
  __host__ __device__ static constexpr int b1_off(int din) { return din * H; }
  __host__ __device__ static constexpr int w2_off(int din) { return din * H + H; }
  __host__ __device__ static constexpr int w2t_off(int din) { return din * H + H + H * H; }
  __host__ __device__ static constexpr int b2_off(int din) { return din * H + H + 2 * H * H; }
  __host__ __device__ static constexpr int wf_off(int din) { return din * H + 2 * H + 2 * H * H; }
  __host__ __device__ static constexpr int mean_off(int din) { return din * H + 3 * H + 2 * H * H + 4; }
  __host__ __device__ static constexpr int istd_off(int din) { return mean_off(din) + IMB_MAX_DIN; }
  __host__ __device__ static constexpr int size(int din) { return istd_off(din) + IMB_MAX_DIN; }
};

// Load one MLP's parameters (torch layout) into its shared-memory image.  wf/bf: for n_hidden==0
// the "final" weights act on the (normalised) inputs and are stored in the W1t column 0.
template <int H>
__device__ void load_mlp(float* sm, const PassDesc& p, const float* __restrict__ params) {
  const int din = p.din;
  const float* q = params + p.param_off;
  const int tid = threadIdx.x, nt = blockDim.x;
  float* W1t = sm;
  float* b1 = sm + MlpSm<H>::b1_off(din);
  float* W2 = sm + MlpSm<H>::w2_off(din);
  float* W2t = sm + MlpSm<H>::w2t_off(din);
  float* b2 = sm + MlpSm<H>::b2_off(din);
  float* wf = sm + MlpSm<H>::wf_off(din);
  for (int i = tid; i < MlpSm<H>::mean_off(din); i += nt) sm[i] = 0.f;
  __syncthreads();
  int off = 0;
  int hl = din;
  if (p.n_hidden >= 1) {
    for (int i = tid; i < p.h1 * din; i += nt) {
      int j = i / din, k = i - j * din;
      W1t[k * H + j] = q[off + i];
    }
    off += p.h1 * din;
    for (int i = tid; i < p.h1; i += nt) b1[i] = q[off + i];
    off += p.h1;
    hl = p.h1;
  }
  if (p.n_hidden >= 2) {
    for (int i = tid; i < p.h2 * p.h1; i += nt) {
      int j = i / p.h1, ii = i - j * p.h1;
      float v = q[off + i];
      W2[j * H + ii] = v;
      W2t[ii * H + j] = v;
    }
    off += p.h2 * p.h1;
    for (int i = tid; i < p.h2; i += nt) b2[i] = q[off + i];
    off += p.h2;
    hl = p.h2;
  }
  if (p.n_hidden == 0) {
    for (int i = tid; i < din; i += nt) W1t[i * H] = q[off + i];  // column 0 holds wf over inputs
  } else {
    for (int i = tid; i < hl; i += nt) wf[i] = q[off + i];
  }
  off += hl;
  if (tid == 0) wf[H] = q[off];  // bf
  float* mean = sm + MlpSm<H>::mean_off(din);
  float* istd = sm + MlpSm<H>::istd_off(din);
  for (int i = tid; i < din; i += nt) {
    if (p.has_norm) {
      mean[i] = p.norm[i];
      istd[i] = 1.0f / sqrtf(p.norm[din + i] + p.eps);
    } else {
      mean[i] = 0.f;
      istd[i] = 1.f;
    }
  }
}

// Forward for one row held by this thread.  xn: this row's normalised inputs in shared memory
// (stride 1).  KEEP: also return h1/h2 (post-ReLU) for the backward.
template <int H, bool KEEP>
__device__ __forceinline__ float mlp_forward_row(const float* __restrict__ sm, const PassDesc& p,
                                                 const float* __restrict__ xn, float (&h1)[H], float (&h2)[H]) {
  const int din = p.din;
  const float* W1t = sm;
  const float* wf = sm + MlpSm<H>::wf_off(din);
  if (p.n_hidden == 0) {
    float acc = wf[H];
    for (int k = 0; k < din; ++k) acc = fmaf(W1t[k * H], xn[k], acc);
    return acc;
  }
  const float* b1 = sm + MlpSm<H>::b1_off(din);
#pragma unroll
  for (int j = 0; j < H; ++j) h1[j] = b1[j];
  for (int k = 0; k < din; ++k) {
    const float xv = xn[k];
    const float4* w = reinterpret_cast<const float4*>(W1t + k * H);
#pragma unroll
    for (int j4 = 0; j4 < H / 4; ++j4) {
      const float4 ww = w[j4];
      h1[4 * j4 + 0] = fmaf(ww.x, xv, h1[4 * j4 + 0]);
      h1[4 * j4 + 1] = fmaf(ww.y, xv, h1[4 * j4 + 1]);
      h1[4 * j4 + 2] = fmaf(ww.z, xv, h1[4 * j4 + 2]);
      h1[4 * j4 + 3] = fmaf(ww.w, xv, h1[4 * j4 + 3]);
    }
  }
#pragma unroll
  for (int j = 0; j < H; ++j) h1[j] = fmaxf(h1[j], 0.f);
  if (p.n_hidden == 1) {
    float acc = wf[H];
#pragma unroll
    for (int j = 0; j < H; ++j) acc = fmaf(wf[j], h1[j], acc);
    return acc;
  }
  const float* W2t = sm + MlpSm<H>::w2t_off(din);
  const float* b2 = sm + MlpSm<H>::b2_off(din);
#pragma unroll
  for (int j = 0; j < H; ++j) h2[j] = b2[j];
#pragma unroll
  for (int i = 0; i < H; ++i) {
    const float hv = h1[i];
    const float4* w = reinterpret_cast<const float4*>(W2t + i * H);
#pragma unroll
    for (int j4 = 0; j4 < H / 4; ++j4) {
      const float4 ww = w[j4];
      h2[4 * j4 + 0] = fmaf(ww.x, hv, h2[4 * j4 + 0]);
      h2[4 * j4 + 1] = fmaf(ww.y, hv, h2[4 * j4 + 1]);
      h2[4 * j4 + 2] = fmaf(ww.z, hv, h2[4 * j4 + 2]);
      h2[4 * j4 + 3] = fmaf(ww.w, hv, h2[4 * j4 + 3]);
    }
  }
  float acc = wf[H];
#pragma unroll
  for (int j = 0; j < H; ++j) {
    h2[j] = fmaxf(h2[j], 0.f);
    acc = fmaf(wf[j], h2[j], acc);
  }
  return acc;
}

// coefficient of a pass's output in the logit: r + gamma*(1-done)*Phi(s') - Phi(s)
__device__ __forceinline__ float pass_coef(int kind, float gamma, float done) {
  return kind == 0 ? 1.0f : (kind == 1 ? gamma * (1.0f - done) : -1.0f);
}


}  // namespace
