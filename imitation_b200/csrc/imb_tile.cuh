// imb_tile.cuh -- shared-memory tiled fp32 GEMM building blocks for the small-MLP kernels.
//
// Why: a thread-per-row MLP reads one broadcast weight per FMA from shared memory; the LSU
// delivers 4 B/lane/cycle, i.e. one operand per lane per cycle per SM, while the FMA pipes want
// four -- so that form tops out at 25 % of fp32 peak (measured 10 TFLOP/s), and its fully
// unrolled code (200 KB) thrashes the instruction cache.  Here every operand fetched from shared
// memory is reused from registers: each thread owns a TR x TJ output tile (2.67-4 FMAs per
// fetched float) and the reduction loops stay rolled (a few KB of code).
//
// Tiles are FEATURE-MAJOR in shared memory: element (feature f, row r) at base[f * RS + r] with
// RS = R + 4; the 4-float pad staggers consecutive feature rows by 4 banks so that the 8 lanes of
// one LDS.128 phase never collide.
#pragma once
#include "imb_common.cuh"
#include "imb_mlp.cuh"

namespace {

constexpr int TILE_PAD = 4;

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }

// acc[NQ*4][8] += sum_k A[k][rq[q]+0..3] * Wk[k * wld + j0 + 0..7]
//   A: feature-major tile (row stride RS); Wk: k-major weights (row stride wld, 16-byte aligned).
// MASK: the A operand is generated on the fly as  (S[k][r] > 0) ? gq[r] * sk[k] : 0
//   (dL/dz of a ReLU layer: S = post-activation tile, gq = upstream scalar per row, sk = per-k scale).
template <int NQ, bool MASK>
__device__ __forceinline__ void gemm_acc(float (&acc)[NQ * 4][8], const float* __restrict__ A, int RS,
                                         const int (&rq)[NQ], const float* __restrict__ Wk, int wld, int j0,
                                         int K, const float4 (&gq)[NQ], const float* __restrict__ sk) {
#pragma unroll 2
  for (int k = 0; k < K; ++k) {
    float4 a[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) a[q] = ld4(A + k * RS + rq[q]);
    if (MASK) {
      const float s = sk[k];
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        a[q].x = a[q].x > 0.f ? gq[q].x * s : 0.f;
        a[q].y = a[q].y > 0.f ? gq[q].y * s : 0.f;
        a[q].z = a[q].z > 0.f ? gq[q].z * s : 0.f;
        a[q].w = a[q].w > 0.f ? gq[q].w * s : 0.f;
      }
    }
    const float4 w0 = ld4(Wk + k * wld + j0), w1 = ld4(Wk + k * wld + j0 + 4);
    const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const float av[4] = {a[q].x, a[q].y, a[q].z, a[q].w};
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[q * 4 + x][t] = fmaf(av[x], w[t], acc[q * 4 + x][t]);
    }
  }
}

// 4 x 4 variant (PPO: 64-row minibatch tiles, 16 outputs per thread)
template <bool TANHMASK>
__device__ __forceinline__ void gemm_acc44(float (&acc)[4][4], const float* __restrict__ A, int RS, int r0,
                                           const float* __restrict__ Wk, int wld, int j0, int K) {
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float4 a = ld4(A + k * RS + r0);
    const float4 w = ld4(Wk + k * wld + j0);
    const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int t = 0; t < 4; ++t) acc[x][t] = fmaf(av[x], wv[t], acc[x][t]);
  }
}

// Weight-gradient contraction over tile rows [ra, rb) (multiples of 4):
//   acc[jj][ii] += sum_r D[jb + jl + 8*jj][r] * ACT[ib + il + 4*ii][r]     jj < 4, ii < 8
//   bacc[jj]    += sum_r D[jb + jl + 8*jj][r]                             (bias gradient)
// One warp (jl = lane % 8, il = lane / 8) covers a 32 x 32 block of the weight gradient; the j set
// is strided by 8 so the 8 lanes of an LDS.128 phase read 8 consecutive feature rows (conflict-free
// with the 4-bank stagger) and the ACT loads are phase-uniform broadcasts.
// MASK: D is generated on the fly as (S[j][r] > 0) ? g[r] * sj[jj] : 0.
template <bool MASK>
__device__ __forceinline__ void wgrad_acc(float (&acc)[4][8], float (&bacc)[4], const float* __restrict__ D,
                                          const float* __restrict__ ACT, int RS, int jrow0, int irow0, int ra, int rb,
                                          const float* __restrict__ gvec, const float (&sj)[4]) {
  for (int r = ra; r < rb; r += 4) {
    float4 d[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) d[jj] = ld4(D + (jrow0 + 8 * jj) * RS + r);
    if (MASK) {
      const float4 g = ld4(gvec + r);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        d[jj].x = d[jj].x > 0.f ? g.x * sj[jj] : 0.f;
        d[jj].y = d[jj].y > 0.f ? g.y * sj[jj] : 0.f;
        d[jj].z = d[jj].z > 0.f ? g.z * sj[jj] : 0.f;
        d[jj].w = d[jj].w > 0.f ? g.w * sj[jj] : 0.f;
      }
    }
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) bacc[jj] += (d[jj].x + d[jj].y) + (d[jj].z + d[jj].w);
#pragma unroll
    for (int ii = 0; ii < 8; ++ii) {
      const float4 a = ld4(ACT + (irow0 + 4 * ii) * RS + r);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        float t = acc[jj][ii];
        t = fmaf(d[jj].x, a.x, t);
        t = fmaf(d[jj].y, a.y, t);
        t = fmaf(d[jj].z, a.z, t);
        t = fmaf(d[jj].w, a.w, t);
        acc[jj][ii] = t;
      }
    }
  }
}

// Shared-memory image of one MLP, hidden widths padded to JP (32 or 64), everything zero padded:
//   W1t[din][JP] b1[JP] W2t[JP][JP] W2[JP][JP] b2[JP] wf[64] bf,pad[4] mean[64] istd[64]
struct TImg {
  __host__ __device__ static int w1t(int, int) { return 0; }
  __host__ __device__ static int b1(int din, int JP) { return din * JP; }
  __host__ __device__ static int w2t(int din, int JP) { return din * JP + JP; }
  __host__ __device__ static int w2(int din, int JP) { return din * JP + JP + JP * JP; }
  __host__ __device__ static int b2(int din, int JP) { return din * JP + JP + 2 * JP * JP; }
  __host__ __device__ static int wf(int din, int JP) { return din * JP + 2 * JP + 2 * JP * JP; }
  __host__ __device__ static int bf(int din, int JP) { return wf(din, JP) + 64; }
  __host__ __device__ static int mean(int din, int JP) { return bf(din, JP) + 4; }
  __host__ __device__ static int istd(int din, int JP) { return mean(din, JP) + 64; }
  __host__ __device__ static int size(int din, int JP) { return istd(din, JP) + 64; }
};

__device__ void load_timg(float* sm, const PassDesc& p, int JP, const float* __restrict__ params,
                          const float* __restrict__ norm, float eps) {
  const int din = p.din, tid = threadIdx.x, nt = blockDim.x;
  const float* q = params + p.param_off;
  for (int i = tid; i < TImg::mean(din, JP); i += nt) sm[i] = 0.f;
  __syncthreads();
  int off = 0, hl = din;
  // (loops unrolled so that the global loads of 8 iterations are in flight together: the image build is the
  //  whole prologue of a CTA that processes a single tile)
  if (p.n_hidden >= 1) {
#pragma unroll 8
    for (int i = tid; i < p.h1 * din; i += nt) {
      const int j = i / din, k = i - j * din;
      sm[TImg::w1t(din, JP) + k * JP + j] = q[off + i];
    }
    off += p.h1 * din;
    for (int i = tid; i < p.h1; i += nt) sm[TImg::b1(din, JP) + i] = q[off + i];
    off += p.h1;
    hl = p.h1;
  }
  if (p.n_hidden >= 2) {
#pragma unroll 8
    for (int i = tid; i < p.h2 * p.h1; i += nt) {
      const int j = i / p.h1, ii = i - j * p.h1;
      const float v = q[off + i];
      sm[TImg::w2(din, JP) + j * JP + ii] = v;
      sm[TImg::w2t(din, JP) + ii * JP + j] = v;
    }
    off += p.h2 * p.h1;
    for (int i = tid; i < p.h2; i += nt) sm[TImg::b2(din, JP) + i] = q[off + i];
    off += p.h2;
    hl = p.h2;
  }
  for (int i = tid; i < hl; i += nt) sm[TImg::wf(din, JP) + i] = q[off + i];
  off += hl;
  if (tid == 0) sm[TImg::bf(din, JP)] = q[off];
  for (int i = tid; i < din; i += nt) {
    sm[TImg::mean(din, JP) + i] = norm ? norm[i] : 0.f;
    sm[TImg::istd(din, JP) + i] = norm ? 1.0f / sqrtf(norm[din + i] + eps) : 1.f;
  }
}


}  // namespace
