// imb_sync.cu -- pack / unpack of the per-round replica state around the ONE all-reduce of a multi-GPU
// round (north_star: "a single NCCL all-reduce ... per round"; imitation_b200/distributed.py).
// Parameters and Adam moments are averaged; RunningNorm statistics (util/networks.py:96-134) are merged
// EXACTLY through their additive sufficient statistics S0 = n, S1 = n*mean, S2 = n*(var + mean^2) relative
// to the common round-start state; staging is float64 so the merge is exact.
#include "imb_common.cuh"

namespace {

__device__ __forceinline__ void suff(const imb_sync_desc& d, int i, int e, double& out) {
  // e: 0 -> S0, 1..k -> S1[e-1], k+1..2k -> S2[e-k-1]
  const int k = d.k[i];
  const double n = (double)*d.count[i];
  if (e == 0) {
    out = n;
  } else if (e <= k) {
    out = n * (double)d.mean[i][e - 1];
  } else {
    const double m = (double)d.mean[i][e - k - 1];
    out = n * ((double)d.var[i][e - k - 1] + m * m);
  }
}

__global__ void k_sync_snapshot(const imb_sync_desc d, double* __restrict__ start) {
  int o = 0;
  for (int i = 0; i < d.n_norm; ++i) {
    const int len = 1 + 2 * d.k[i];
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < len; e += gridDim.x * blockDim.x) suff(d, i, e, start[o + e]);
    o += len;
  }
}

__global__ void k_sync_pack(const imb_sync_desc d, double* __restrict__ buf) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t o = 0;
  for (int i = 0; i < d.n_avg; ++i) {
    for (int64_t e = t0; e < d.avg_n[i]; e += stride) buf[o + e] = (double)d.avg[i][e];
    o += d.avg_n[i];
  }
  for (int i = 0; i < d.n_norm; ++i) {
    const int len = 1 + 2 * d.k[i];
    for (int64_t e = t0; e < len; e += stride) suff(d, i, (int)e, buf[o + e]);
    o += len;
  }
}

__global__ void k_sync_unpack(const imb_sync_desc d, const double* __restrict__ buf, const double* __restrict__ start,
                              int world) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t o = 0;
  const double inv = 1.0 / (double)world;
  for (int i = 0; i < d.n_avg; ++i) {
    for (int64_t e = t0; e < d.avg_n[i]; e += stride) d.avg[i][e] = (float)(buf[o + e] * inv);
    o += d.avg_n[i];
  }
  int so = 0;
  for (int i = 0; i < d.n_norm; ++i) {
    const int k = d.k[i], len = 1 + 2 * k;
    const double wm1 = (double)(world - 1);
    const double n = buf[o] - wm1 * start[so];
    for (int64_t e = t0; e < k; e += stride) {
      if (n > 0.0) {
        const double a = buf[o + 1 + e] - wm1 * start[so + 1 + e];
        const double b = buf[o + 1 + k + e] - wm1 * start[so + 1 + k + e];
        const double mean = a / n;
        double var = b / n - mean * mean;
        var = var > 0.0 ? var : 0.0;
        d.mean[i][e] = (float)mean;
        d.var[i][e] = (float)var;
      }
    }
    if (t0 == 0) *d.count[i] = (int32_t)llrint(n);
    o += len;
    so += len;
  }
}

int check(const imb_sync_desc* d) {
  IMB_REQUIRE(d->n_avg >= 0 && d->n_avg <= IMB_SYNC_MAX_AVG, "n_avg must be in [0, %d]", IMB_SYNC_MAX_AVG);
  IMB_REQUIRE(d->n_norm >= 0 && d->n_norm <= IMB_SYNC_MAX_NORM, "n_norm must be in [0, %d]", IMB_SYNC_MAX_NORM);
  return 0;
}

}  // namespace

extern "C" int64_t imb_sync_buffer_doubles(const imb_sync_desc* d) {
  int64_t n = 0;
  for (int i = 0; i < d->n_avg; ++i) n += d->avg_n[i];
  for (int i = 0; i < d->n_norm; ++i) n += 1 + 2 * d->k[i];
  return n;
}

extern "C" int imb_sync_snapshot(const imb_sync_desc* d, double* start, void* stream) {
  if (int rc = check(d)) return rc;
  if (d->n_norm == 0) return 0;
  k_sync_snapshot<<<1, 256, 0, (cudaStream_t)stream>>>(*d, start);
  IMB_CHECK_LAUNCH("k_sync_snapshot");
  return 0;
}

extern "C" int imb_sync_pack(const imb_sync_desc* d, double* buf, void* stream) {
  if (int rc = check(d)) return rc;
  k_sync_pack<<<32, 256, 0, (cudaStream_t)stream>>>(*d, buf);
  IMB_CHECK_LAUNCH("k_sync_pack");
  return 0;
}

extern "C" int imb_sync_unpack(const imb_sync_desc* d, const double* buf, const double* start, int32_t world,
                               void* stream) {
  if (int rc = check(d)) return rc;
  IMB_REQUIRE(world >= 1, "world must be >= 1");
  k_sync_unpack<<<32, 256, 0, (cudaStream_t)stream>>>(*d, buf, start, world);
  IMB_CHECK_LAUNCH("k_sync_unpack");
  return 0;
}
