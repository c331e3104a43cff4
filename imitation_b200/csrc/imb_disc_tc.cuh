// imb_disc_tc.cuh -- the fused discriminator forward / BCE / backward kernel on the 5th-generation tensor cores.
//
// Same contract as k_disc_fwdbwd (imb_disc.cu): reads the feature-major disc batch, writes logits and one partial
// [gradient | statistics] vector per CTA.  Replaces (reference): rewards/reward_nets.py:441-457 + util/networks.py:264-281
// (three addmm of BasicRewardNet.forward) and the five mm of its backward (algorithms/adversarial/common.py:360-369).
//
// fp32 parity (north star: logits within 1e-5) on tf32 tensor cores: every operand is split x = hi + lo with hi exact in
// tf32, and every contraction is the three-term sum  hi*hi + hi*lo + lo*hi  accumulated in fp32 in TMEM (error ~2^-22 per
// product, the dropped lo*lo term is 2^-24).
//
// One CTA = 128-row tiles, persistent.  Warps 0-7: epilogue (thread = (row, 16-column half)); warp 8: the single thread
// that issues tcgen05.mma; warp 9: the single thread that issues the cp.async.bulk tile loads (two stages).
//
// Scheduling.  The chain E0 M1 E1 M2 E2 M3 E3 of a tile is strictly sequential and every hand-off between the epilogue
// warps and the tensor core costs ~300 cycles, while M4 is 1.5 k cycles of tensor time that only needs the operand atoms
// to stay put.  So the MMA thread is a small scheduler: chain MMAs are issued the moment their operand is ready, the
// weight-gradient MMAs of the PREVIOUS tile are issued in chunks of one k-step (3 MMAs, ~96 cycles; at most two chunks in
// flight) whenever no chain step is ready -- they run under the epilogue phases of the next tile.  What makes that legal
// with one set of operand atoms (two sets do not fit in 227 KB): a1 is double-buffered, the x atoms are written late
// (recomputed from the still-resident stage in E2, after M4 of the previous tile has drained), dz2 / dz1 likewise.
// Because [a1 | x] must be adjacent atoms, odd tiles use [x | a1'] and accumulate into a second TMEM accumulator with the
// column blocks swapped; the two are added in the final read-out.
//
//   E0  stage -> normalise -> [x | 1] split -> TMEM operand columns + smem x atom
//   M1  z1 = [x|1] . [W1|b1]^T                 TS: A from TMEM (M=128 rows), B = weight image in smem, 16 cycles / MMA
//   E1  a1 = relu(z1), mask bits               -> TMEM operand + smem a1 atom
//   M2  z2 = a1 . W2^T
//   E2  a2 = relu(z2 + b2); logit; BCE; g; dz2 = g * w3 * [z2 > 0]; dW3 += g * a2   -> TMEM operand + smem dz2 atom
//   M3  dz1' = dz2 . W2                        (B = transposed weight image)
//   E3  dz1 = dz1' * [z1 > 0]                  -> smem dz1 atom
//   M4  [dW2 db2 ; dW1 db1] += [dz2 | dz1]^T . [a1 | x 1]     SS, M=64, N=32+K1, K=128 rows, both operands MN-major:
//       tf32 MN-major operands exist only in the SWIZZLE_128B_BASE32B layout, i.e. plain [row][32 floats] atoms whose
//       32-byte chunk c of row r sits at chunk c ^ (r & 3) -- which is exactly what a thread that owns a row writes.
//       The accumulator stays in TMEM across all tiles of the CTA.
//
// Measured MMA costs on B200 (profiles/micro/tc_probe_r02b.log): TS M128 N32 16.0 cycles (tensor floor), SS M64 N56 30.0
// (shared-memory operand bandwidth: 3.75 KB / 128 B per cycle), SS M128 N32 40 (why the forward operands go through TMEM).
#pragma once
#include "imb_common.cuh"
#include "imb_mlp.cuh"
#include "imb_tc.cuh"

namespace {

// epilogue thread = (row, CW-column group): CW = 16 -> 8 epilogue warps, CW = 8 -> 16 (the phases are latency-bound:
// twice the warps per scheduler nearly halves them; needs <= 112 registers per thread)
template <int CW> struct TcCfg {
  static constexpr int NH = 32 / CW;             // column groups
  static constexpr int EPI_WARPS = 4 * NH;
  static constexpr int EPI_THREADS = EPI_WARPS * 32;
  static constexpr int THREADS = EPI_THREADS + 64;  // + MMA warp + loader warp
};
// background weight-gradient chunks (of 16 per tile) issued after: the job's creation | M1 | M2 of the next tile
constexpr int TC_JOB0 = 2, TC_JOB1 = 5;
constexpr int TC_ATOM_F = 128 * 32;             // floats of one [128 rows][32] operand atom (16 KB)
constexpr int TC_STAGE_F = 32 * 128;            // floats of one input stage: 32 slots x 128 rows (slot 31 = zeros)
constexpr int TC_TMEM_COLS = 256;
// TMEM columns
constexpr int TC_ACC = 0, TC_OPA_HI = 32, TC_OPA_LO = 64, TC_WG0 = 96, TC_WG1 = 160;

struct TcPlan {  // float offsets into the 1024-byte aligned dynamic shared memory
  int a_hi, a_lo;                                 // [dz2 | dz1]
  int b_hi, b_lo;                                 // [a1 (even tiles) | x | a1 (odd tiles)]
  int stage;                                      // 2 x [32 slots][128] staged batch rows
  int w1_hi, w1_lo, w2_hi, w2_lo, w2t_hi, w2t_lo;  // weight images, K-major no-swizzle: [(k/4)][n][4]
  int vec;                                        // b2[32] w3[32] b3,pad[4] mean[32] istd[32] red[512] fin[4*40]
  int total;
  int K1, KC;                                     // input columns incl. the ones column, rounded up to 8; K1 / 8
};

inline bool tc_applicable(const DiscLaunch& L) {
  const PassDesc& p = L.pass[0];
  return L.npass == 1 && p.n_hidden == 2 && p.h1 == 32 && p.h2 == 32 && p.din + 1 <= 32 && L.logp_slot < 0 &&
         L.done_slot < 0 && L.nstage <= 31;
}

inline TcPlan tc_plan(const DiscLaunch& L) {
  TcPlan t;
  t.K1 = (L.pass[0].din + 1 + 7) / 8 * 8;
  t.KC = t.K1 / 8;
  int o = 0;
  t.a_hi = o; o += 2 * TC_ATOM_F;
  t.a_lo = o; o += 2 * TC_ATOM_F;
  t.b_hi = o; o += 3 * TC_ATOM_F;
  t.b_lo = o; o += 3 * TC_ATOM_F;
  t.stage = o; o += 2 * TC_STAGE_F;
  t.w1_hi = o; o += 32 * 32;
  t.w1_lo = o; o += 32 * 32;
  t.w2_hi = o; o += 32 * 32;
  t.w2_lo = o; o += 32 * 32;
  t.w2t_hi = o; o += 32 * 32;
  t.w2t_lo = o; o += 32 * 32;
  t.vec = o; o += 32 + 32 + 4 + 32 + 32 + 512 + 4 * 40;
  t.total = o;
  return t;
}

#ifdef IMB_TC_TIMING
#define TCK(i)                                  \
  do {                                          \
    const long long t__ = clock64();            \
    tclk[i] += (float)(t__ - tlast);            \
    tlast = t__;                                \
  } while (0)
#else
#define TCK(i) do { } while (0)
#endif

struct TcBars {
  uint64_t full[2], empty[2], opa, acc, wg;
};

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// epilogue -> MMA hand-off, one arrival per warp: every lane completes its TMEM stores (and, with SMEM, makes its
// generic-proxy shared-memory stores of the whole tile visible to the async proxy the MMA reads operands through),
// the warp converges, lane 0 arrives (release; cumulative over the lanes it synchronised with)
template <bool SMEM>
__device__ __forceinline__ void epi_signal(uint64_t* bar, int lane) {
  tc::st_wait();
  tc::fence_before_sync();
  if (SMEM) tc::fence_async_smem();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}
// CW values of one row -> the row's 32-byte chunks of a swizzled atom, conflict-free across the warp: rows r and
// r + 4 share a chunk position, so lanes with bit 2 of the row set store their 16-byte pieces in swapped order
template <int CW>
__device__ __forceinline__ void store_tile(float* atom, const int (&poff)[CW / 4], bool sw, const float (&v)[CW]) {
#pragma unroll
  for (int p = 0; p < CW / 4; ++p) {
    const int a = p * 4, b = (p ^ 1) * 4;
    const float4 val = make_float4(sw ? v[b] : v[a], sw ? v[b + 1] : v[a + 1], sw ? v[b + 2] : v[a + 2],
                                   sw ? v[b + 3] : v[a + 3]);
    *reinterpret_cast<float4*>(atom + poff[p]) = val;
  }
}
template <int CW>
__device__ __forceinline__ void lds_vec(const float* p, float (&v)[CW]) {  // CW consecutive floats, 16-byte aligned
#pragma unroll
  for (int i = 0; i < CW / 4; ++i) {
    const float4 t = *reinterpret_cast<const float4*>(p + 4 * i);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}

template <int CW> __device__ __forceinline__ void tmem_ld(uint32_t a, float (&v)[CW]);
template <> __device__ __forceinline__ void tmem_ld<16>(uint32_t a, float (&v)[16]) { tc::ld16(a, v); }
template <> __device__ __forceinline__ void tmem_ld<8>(uint32_t a, float (&v)[8]) { tc::ld8(a, v); }
template <int CW> __device__ __forceinline__ void tmem_st(uint32_t a, const float (&v)[CW]);
template <> __device__ __forceinline__ void tmem_st<16>(uint32_t a, const float (&v)[16]) { tc::st16(a, v); }
template <> __device__ __forceinline__ void tmem_st<8>(uint32_t a, const float (&v)[8]) { tc::st8(a, v); }

template <int CW>
__global__ void __launch_bounds__(TcCfg<CW>::THREADS, 1)
    k_disc_fwdbwd_tc(const DiscLaunch L, const TcPlan T, const float* __restrict__ params,
                     const float* __restrict__ batch, int64_t ld, int64_t n, int64_t n_expert, float loss_scale,
                     const float* __restrict__ grad_out, float* __restrict__ logits_out, float* __restrict__ partial,
                     int* __restrict__ meta, int64_t pstride) {
  constexpr int TC_EPI_WARPS = TcCfg<CW>::EPI_WARPS, TC_EPI_THREADS = TcCfg<CW>::EPI_THREADS,
                TC_THREADS = TcCfg<CW>::THREADS, NH = TcCfg<CW>::NH, NP = CW / 4;
  // (the swizzled operand atoms need a 1024-byte aligned base; another kernel of this translation unit declares the
  //  dynamic shared array with a smaller alignment, so align by hand -- the launch adds 1 KB of slack)
  extern __shared__ float tc_smem_raw[];
  float* smem = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u) / 4;
  __shared__ __align__(8) TcBars bars;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const PassDesc& Pd = L.pass[0];
  const int din = Pd.din, K1 = T.K1, KC = T.KC;
  const int P = L.P;
  float* vec = smem + T.vec;
  float* b2s = vec;
  float* w3s = vec + 32;
  float* b3s = vec + 64;
  float* means = vec + 68;   // per input column: (mean, 1/std); the ones column is (-1, 1) over a zero input, padding (0, 0)
  float* istds = vec + 100;
  float* red = vec + 132;    // [NH][128] head partial sums
  float* fin = vec + 644;    // [4][40] final reductions

  // ---- prologue: weight images (hi / lo), vectors, barriers, TMEM ----------------------------------------------
  {
    const float* q = params + Pd.param_off;
    const int off_b1 = 32 * din, off_w2 = off_b1 + 32, off_b2 = off_w2 + 1024, off_wf = off_b2 + 32, off_bf = off_wf + 32;
    for (int i = tid; i < 32 * K1; i += TC_THREADS) {  // W1 image: n = j, k
      const int j = i / K1, k = i - j * K1;
      const float v = k < din ? q[j * din + k] : (k == din ? q[off_b1 + j] : 0.f);
      float hi, lo;
      tc::split_tf32(v, hi, lo);
      const int o = (k >> 2) * 128 + j * 4 + (k & 3);
      smem[T.w1_hi + o] = hi;
      smem[T.w1_lo + o] = lo;
    }
    for (int i = tid; i < 1024; i += TC_THREADS) {
      const int j2 = i >> 5, i1 = i & 31;
      float hi, lo;
      tc::split_tf32(q[off_w2 + i], hi, lo);
      const int o = (i1 >> 2) * 128 + j2 * 4 + (i1 & 3);   // forward: n = j2, k = i1
      smem[T.w2_hi + o] = hi;
      smem[T.w2_lo + o] = lo;
      const int ot = (j2 >> 2) * 128 + i1 * 4 + (j2 & 3);  // backward: n = i1, k = j2
      smem[T.w2t_hi + ot] = hi;
      smem[T.w2t_lo + ot] = lo;
    }
    if (tid < 32) {
      b2s[tid] = q[off_b2 + tid];
      w3s[tid] = q[off_wf + tid];
      float m = 0.f, is = 0.f;
      if (tid < din) {
        m = Pd.has_norm ? Pd.norm[tid] : 0.f;
        is = Pd.has_norm ? 1.0f / sqrtf(Pd.norm[din + tid] + Pd.eps) : 1.f;
      } else if (tid == din) {
        m = -1.f;
        is = 1.f;
      }
      means[tid] = m;
      istds[tid] = is;
      if (tid == 0) b3s[0] = q[off_bf];
    }
    if (tid >= 64 && tid < 64 + 256) {  // the all-zero slot 31 of both stages
      const int i = tid - 64;
      smem[T.stage + (i >> 7) * TC_STAGE_F + 31 * 128 + (i & 127)] = 0.f;
    }
  }
  if (tid == 0) {
    mbar_init(&bars.full[0], 1);
    mbar_init(&bars.full[1], 1);
    mbar_init(&bars.empty[0], TC_EPI_WARPS);
    mbar_init(&bars.empty[1], TC_EPI_WARPS);
    mbar_init(&bars.opa, TC_EPI_WARPS);
    mbar_init(&bars.acc, 1);
    mbar_init(&bars.wg, 1);
    mbar_fence_init();
    if (blockIdx.x == 0) {  // launch record for k_disc_reduce / k_disc_adam (statistics of the LAST minibatch)
      meta[0] = (int)gridDim.x;
      meta[1] = (int)n;
      meta[2] = (int)n_expert;
      reinterpret_cast<float*>(meta)[3] = loss_scale;
    }
  }
  if (warp == TC_EPI_WARPS) tc::tmem_alloc(&tmem_base_s, TC_TMEM_COLS);
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tb = tmem_base_s;
  const int64_t ntiles = (n + 127) / 128;
  // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const int nt = (int)((ntiles - (int64_t)blockIdx.x + (int64_t)gridDim.x - 1) / (int64_t)gridDim.x);

  if (warp == TC_EPI_WARPS + 1) {
    // ================================ loader: two stages ==============================================
    if (elect_one()) {
      for (int it = 0; it < nt; ++it) {
        const int64_t tile = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
        const int s = it & 1;
        if (it >= 2) mbar_wait(&bars.empty[s], (uint32_t)(((it >> 1) - 1) & 1));
        float* xs = smem + T.stage + s * TC_STAGE_F;
        int64_t cnt = ld - tile * 128;
        if (cnt > 128) cnt = 128;
        mbar_expect_tx(&bars.full[s], (uint32_t)(L.nstage * cnt * 4));
        for (int k = 0; k < L.nstage; ++k)
          bulk_g2s(xs + k * 128, batch + (int64_t)L.stage_row[k] * ld + tile * 128, (uint32_t)(cnt * 4), &bars.full[s]);
      }
    }
  } else if (warp == TC_EPI_WARPS) {
    // ================================ MMA issuer =======================================================
    if (elect_one()) {
      const uint32_t s0 = smem_u32(smem);
      // weight images: K-major, no swizzle: LBO = 32 * 16 B (next chunk of four k), SBO = 128 B (next eight n);
      // one MMA consumes two k chunks = 1024 B
      const uint64_t dW1h = tc::smem_desc(s0 + T.w1_hi * 4, 512, 128), dW1l = tc::smem_desc(s0 + T.w1_lo * 4, 512, 128);
      const uint64_t dW2h = tc::smem_desc(s0 + T.w2_hi * 4, 512, 128), dW2l = tc::smem_desc(s0 + T.w2_lo * 4, 512, 128);
      const uint64_t dWth = tc::smem_desc(s0 + T.w2t_hi * 4, 512, 128), dWtl = tc::smem_desc(s0 + T.w2t_lo * 4, 512, 128);
      // wgrad operands: MN-major SWIZZLE_128B_BASE32B: LBO = atom stride (16 KB), SBO = 512 B (next four k rows);
      // one MMA consumes eight rows = 1024 B.  Even tiles: B = [a1 | x] from atom 0, odd tiles: B = [x | a1'] from atom 1.
      const uint64_t SW = (uint64_t)1 << 61;
      const uint64_t dAh = tc::smem_desc(s0 + T.a_hi * 4, 16384, 512) | SW, dAl = tc::smem_desc(s0 + T.a_lo * 4, 16384, 512) | SW;
      const uint64_t dBh0 = tc::smem_desc(s0 + T.b_hi * 4, 16384, 512) | SW, dBl0 = tc::smem_desc(s0 + T.b_lo * 4, 16384, 512) | SW;
      const uint32_t id_f = tc::idesc_tf32(128, 32, 0, 0);
      const uint32_t id_w = tc::idesc_tf32(64, 64, 1, 1);
      const uint32_t ACC = tb + TC_ACC, OH = tb + TC_OPA_HI, OL = tb + TC_OPA_LO;
      uint32_t nopa = 0;
      // weight-gradient k-steps [k0, k1) of tile jt (its operand atoms stay untouched until bars.wg completes)
      auto wgrad = [&](int jt, int k0, int k1) {
        const int p = jt & 1;
        const uint32_t WG = tb + (p ? TC_WG1 : TC_WG0);
        const uint64_t bh0 = dBh0 + (p ? 1024 : 0), bl0 = dBl0 + (p ? 1024 : 0);  // + one atom = 16 KB >> 4
        for (int ks = k0; ks < k1; ++ks) {
          tc::mma_ss(WG, dAh + 64 * ks, bh0 + 64 * ks, id_w, (jt >= 2 || ks > 0) ? 1u : 0u);
          tc::mma_ss(WG, dAh + 64 * ks, bl0 + 64 * ks, id_w, 1);
          tc::mma_ss(WG, dAl + 64 * ks, bh0 + 64 * ks, id_w, 1);
        }
        if (k1 == 16) tc::commit(&bars.wg);
      };
      for (int it = 0; it < nt; ++it) {
        mbar_wait(&bars.opa, nopa++ & 1);  // x operand in TMEM
        tc::fence_after_sync();
        for (int ks = 0; ks < KC; ++ks) {
          tc::mma_ts(ACC, OH + 8 * ks, dW1h + 64 * ks, id_f, ks > 0);
          tc::mma_ts(ACC, OH + 8 * ks, dW1l + 64 * ks, id_f, 1);
          tc::mma_ts(ACC, OL + 8 * ks, dW1h + 64 * ks, id_f, 1);
        }
        tc::commit(&bars.acc);
        if (it > 0) wgrad(it - 1, TC_JOB0, TC_JOB0 + TC_JOB1);  // runs under E1
        mbar_wait(&bars.opa, nopa++ & 1);  // a1
        tc::fence_after_sync();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          tc::mma_ts(ACC, OH + 8 * ks, dW2h + 64 * ks, id_f, ks > 0);
          tc::mma_ts(ACC, OH + 8 * ks, dW2l + 64 * ks, id_f, 1);
          tc::mma_ts(ACC, OL + 8 * ks, dW2h + 64 * ks, id_f, 1);
        }
        tc::commit(&bars.acc);
        if (it > 0) wgrad(it - 1, TC_JOB0 + TC_JOB1, 16);  // runs under E2; completes before E2 stores dz2 / x
        mbar_wait(&bars.opa, nopa++ & 1);  // dz2
        tc::fence_after_sync();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          tc::mma_ts(ACC, OH + 8 * ks, dWth + 64 * ks, id_f, ks > 0);
          tc::mma_ts(ACC, OH + 8 * ks, dWtl + 64 * ks, id_f, 1);
          tc::mma_ts(ACC, OL + 8 * ks, dWth + 64 * ks, id_f, 1);
        }
        tc::commit(&bars.acc);
        mbar_wait(&bars.opa, nopa++ & 1);  // the tile's operand atoms are complete
        tc::fence_after_sync();
        wgrad(it, 0, it + 1 < nt ? TC_JOB0 : 16);  // runs under E0 of the next tile (all of it after the last tile)
      }
    }
  } else {
    // ================================ epilogue warps ================================================
    const int q = warp & 3, h = warp >> 2;
    const int row = 32 * q + lane;
    const int c0 = CW * h;
    const uint32_t tl = tb + ((uint32_t)(32 * q) << 16);
    const bool sw = (row >> 2) & 1;
    int poff[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const int pq = p ^ (int)sw;
      poff[p] = row * 32 + ((((c0 >> 3) + (pq >> 1)) ^ (row & 3)) << 3) + (pq & 1) * 4;
    }
    float* A_hi = smem + T.a_hi;
    float* A_lo = smem + T.a_lo;
    float* B_hi = smem + T.b_hi;
    float* B_lo = smem + T.b_lo;
    int xoff[CW];  // staged word of input column c0 + c (slot 31 = zeros for the ones column and the padding)
#pragma unroll
    for (int c = 0; c < CW; ++c) xoff[c] = ((c0 + c < din) ? (int)Pd.in_slot[c0 + c] : 31) * 128 + row;
    float dw3[CW];
#pragma unroll
    for (int c = 0; c < CW; ++c) dw3[c] = 0.f;
    float db3 = 0.f, s_loss = 0.f, s_ent = 0.f;
    int c_exp = 0, c_gen = 0, c_pred_exp = 0;
    uint32_t nacc = 0;
    float v[CW], hi[CW], lo[CW];
#ifdef IMB_TC_TIMING
    float tclk[11];
#pragma unroll
    for (int i = 0; i < 11; ++i) tclk[i] = 0.f;
    long long tlast = clock64();
#endif
    // normalised inputs of this thread's 16 columns (branch-free: all loads first)
    auto load_x = [&](const float* xs, bool valid) {
      float mn[CW], is[CW];
      lds_vec<CW>(means + c0, mn);
      lds_vec<CW>(istds + c0, is);
#pragma unroll
      for (int c = 0; c < CW; ++c) v[c] = xs[xoff[c]];
#pragma unroll
      for (int c = 0; c < CW; ++c) v[c] = valid ? (v[c] - mn[c]) * is[c] : 0.f;
    };
    for (int it = 0; it < nt; ++it) {
      const int64_t tile = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
      const int nv = (int)min((int64_t)128, n - tile * 128);
      const bool valid = row < nv;
      const int s = it & 1;
      const float* xs = smem + T.stage + s * TC_STAGE_F;
      // ---- E0: inputs -> TMEM operand ----------------------------------------------------------------
      mbar_wait(&bars.full[s], (uint32_t)((it >> 1) & 1));
      TCK(0);
      if (c0 < K1) {
        load_x(xs, valid);
#pragma unroll
        for (int c = 0; c < CW; ++c) tc::split_tf32(v[c], hi[c], lo[c]);
        tmem_st<CW>(tl + TC_OPA_HI + c0, hi);
        tmem_st<CW>(tl + TC_OPA_LO + c0, lo);
      }
      epi_signal<false>(&bars.opa, lane);
      TCK(1);
      // ---- E1: a1 = relu(z1) ------------------------------------------------------------------------
      mbar_wait(&bars.acc, nacc++ & 1);
      TCK(2);
      tc::fence_after_sync();
      tmem_ld<CW>(tl + TC_ACC + c0, v);
      tc::ld_wait();
      uint32_t m1 = 0;
#pragma unroll
      for (int c = 0; c < CW; ++c) {
        m1 |= (v[c] > 0.f ? 1u : 0u) << c;
        tc::split_tf32(fmaxf(v[c], 0.f), hi[c], lo[c]);
      }
      tmem_st<CW>(tl + TC_OPA_HI + c0, hi);
      tmem_st<CW>(tl + TC_OPA_LO + c0, lo);
      // (the a1 atoms of this parity were last read by the weight-gradient MMAs of tile it - 2, which completed before
      //  tile it - 1 stored its dz2)
      store_tile<CW>(B_hi + (s ? 2 * TC_ATOM_F : 0), poff, sw, hi);
      store_tile<CW>(B_lo + (s ? 2 * TC_ATOM_F : 0), poff, sw, lo);
      epi_signal<false>(&bars.opa, lane);
      TCK(3);
      // ---- E2: a2, logit, loss, dz2 ---------------------------------------------------------------------
      mbar_wait(&bars.acc, nacc++ & 1);
      TCK(4);
      tc::fence_after_sync();
      tmem_ld<CW>(tl + TC_ACC + c0, v);
      lds_vec<CW>(b2s + c0, hi);
      lds_vec<CW>(w3s + c0, lo);
      tc::ld_wait();
      float part = 0.f;
#pragma unroll
      for (int c = 0; c < CW; ++c) {
        v[c] = fmaxf(v[c] + hi[c], 0.f);  // a2 (post-ReLU; > 0 <=> z2 > 0)
        part = fmaf(lo[c], v[c], part);
      }
      red[h * 128 + row] = part;
      named_bar_sync(1 + q, 32 * NH);
      float logit;
      if (NH == 2) logit = (red[row] + red[128 + row]) + b3s[0];
      else logit = ((red[row] + red[128 + row]) + (red[256 + row] + red[384 + row])) + b3s[0];
      float g = 0.f;
      if (valid) {
        const int64_t grow = tile * 128 + row;
        if (h == 0 && logits_out) logits_out[grow] = logit;
        if (grad_out) {
          g = grad_out[grow];
        } else {
          const float y = (grow < n_expert) ? 1.f : 0.f;
          // sigmoid and softplus from ONE exponential: e = exp(-|x|) is what sigmoid_f / log1pf(expf(-|x|)) evaluate
          const float e = expf(-fabsf(logit));
          const float sg = logit >= 0.f ? 1.0f / (1.0f + e) : e / (1.0f + e);
          if (h == 0) {
            const float sp = fmaxf(logit, 0.f) + log1pf(e);
            s_loss += sp - logit * y;
            s_ent += sp - logit * sg;
            const bool pred_exp = !(logit < 0.f);
            c_pred_exp += pred_exp;
            if (y > 0.5f) c_exp += pred_exp; else c_gen += !pred_exp;
          }
          g = (sg - y) * loss_scale;
        }
      }
      if (h == 0) db3 += g;
#pragma unroll
      for (int c = 0; c < CW; ++c) {
        dw3[c] = fmaf(g, v[c], dw3[c]);
        const float dz = v[c] > 0.f ? g * lo[c] : 0.f;
        tc::split_tf32(dz, hi[c], v[c]);  // hi[] / v[] = dz2 hi / lo
      }
      tmem_st<CW>(tl + TC_OPA_HI + c0, hi);
      tmem_st<CW>(tl + TC_OPA_LO + c0, v);
      epi_signal<false>(&bars.opa, lane);
      TCK(5);
      // the atoms written from here on are still being read by the previous tile's weight-gradient MMAs
      if (it > 0) mbar_wait(&bars.wg, (uint32_t)((it - 1) & 1));
      TCK(6);
      store_tile<CW>(A_hi, poff, sw, hi);
      store_tile<CW>(A_lo, poff, sw, v);
      if (c0 < K1) {  // x atoms: recomputed from the stage (cheaper than 32 live registers since E0)
        load_x(xs, valid);
#pragma unroll
        for (int c = 0; c < CW; ++c) tc::split_tf32(v[c], hi[c], lo[c]);
        store_tile<CW>(B_hi + TC_ATOM_F, poff, sw, hi);
        store_tile<CW>(B_lo + TC_ATOM_F, poff, sw, lo);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.empty[s]);
      TCK(7);
      // ---- E3: dz1 ----------------------------------------------------------------------------------------
      mbar_wait(&bars.acc, nacc++ & 1);
      TCK(8);
      tc::fence_after_sync();
      tmem_ld<CW>(tl + TC_ACC + c0, v);
      tc::ld_wait();
#pragma unroll
      for (int c = 0; c < CW; ++c) tc::split_tf32(((m1 >> c) & 1u) ? v[c] : 0.f, hi[c], lo[c]);
      store_tile<CW>(A_hi + TC_ATOM_F, poff, sw, hi);
      store_tile<CW>(A_lo + TC_ATOM_F, poff, sw, lo);
      epi_signal<true>(&bars.opa, lane);
      TCK(9);
    }
    // ---- the CTA's partial vector -------------------------------------------------------------------------
    float* my = partial + (int64_t)blockIdx.x * pstride;
    const int off_b1 = Pd.param_off + 32 * din, off_w2 = off_b1 + 32, off_b2 = off_w2 + 1024, off_wf = off_b2 + 32,
              off_bf = off_wf + 32;
    if (nt > 0) {
      mbar_wait(&bars.wg, (uint32_t)((nt - 1) & 1));
      tc::fence_after_sync();
      // accumulator row m (0..31: dz2 unit m; 32..63: dz1 unit m - 32) lives in lane (m % 16) + 32 * (m / 16).
      // WG0 (even tiles): columns 0..31 a1 units, 32 + k input column k; WG1 (odd tiles): the two blocks swapped.
      // column k == din is the ones column -> bias gradients.  Column half h = 0 reads the a1 block, h = 1 the x block.
      const int m = 16 * q + (lane & 15), j = m & 31;
      if (h < 2) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float w0[16], w1[16];
          tc::ld16(tl + TC_WG0 + 32 * h + 16 * half, w0);
          if (nt > 1) tc::ld16(tl + TC_WG1 + 32 * (1 - h) + 16 * half, w1);
          tc::ld_wait();
          if (lane < 16) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int c = 16 * half + i;
              const float val = nt > 1 ? w0[i] + w1[i] : w0[i];
              if (h == 0) {
                if (m < 32) my[off_w2 + j * 32 + c] = val;
              } else {
                if (c < din) {
                  if (m >= 32) my[Pd.param_off + j * din + c] = val;
                } else if (c == din) {
                  my[(m < 32 ? off_b2 : off_b1) + j] = val;
                }
              }
            }
          }
        }
      }
    } else {  // (a CTA without tiles: cannot happen with grid <= ntiles, kept for safety)
      for (int i = tid; i < P; i += TC_EPI_THREADS) my[i] = 0.f;
    }
    // dW3 / db3 / statistics: warp shuffles, then the four row quarters through shared memory
#pragma unroll
    for (int c = 0; c < CW; ++c) dw3[c] = warp_sum(dw3[c]);
    db3 = warp_sum(db3);
    s_loss = warp_sum(s_loss);
    s_ent = warp_sum(s_ent);
    c_exp = warp_sum_i(c_exp);
    c_gen = warp_sum_i(c_gen);
    c_pred_exp = warp_sum_i(c_pred_exp);
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < CW; ++c) fin[q * 40 + c0 + c] = dw3[c];
      if (h == 0) {
        fin[q * 40 + 32] = db3;
        fin[q * 40 + 33] = s_loss;
        fin[q * 40 + 34] = s_ent;
        fin[q * 40 + 35] = (float)c_exp;
        fin[q * 40 + 36] = (float)c_gen;
        fin[q * 40 + 37] = (float)c_pred_exp;
      }
    }
    named_bar_sync(5, TC_EPI_THREADS);
    if (tid < 38) {
      const float sum = (fin[tid] + fin[40 + tid]) + (fin[80 + tid] + fin[120 + tid]);
      if (tid < 32) my[off_wf + tid] = sum;
      else if (tid == 32) my[off_bf] = sum;
      else my[P + (tid - 33)] = sum;
    }
#ifdef IMB_TC_TIMING
    if (tid == 0) {  // phase clocks of (row 0, column half 0): cycles summed over this CTA's tiles; [10] = tiles
      tclk[10] = (float)nt;
      for (int i = 0; i < 11; ++i) my[P + 5 + i] = tclk[i];
    }
#endif
  }
  // ---- teardown -------------------------------------------------------------------------------------------
  __syncwarp();
  tc::fence_before_sync();
  __syncthreads();
  if (warp == TC_EPI_WARPS) tc::tmem_dealloc(tb, TC_TMEM_COLS);
}

}  // namespace
