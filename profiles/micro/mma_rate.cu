// Throughput / latency of legacy mma.sync.m16n8k8 TF32 and of FFMA on this GPU (profiles/r01_summary.md).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma(float (&d)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int NACC>
__global__ void k_mma(float* out, int iters, long long* cycles) {
  float d[NACC][4];
  for (int i = 0; i < NACC; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
  unsigned a = 0x3f800000u + threadIdx.x, b = 0x3f000000u + threadIdx.x;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) mma(d[i], a, a, a, a, b, b);
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < NACC; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

__global__ void k_ffma(float* out, int iters, long long* cycles) {
  float acc[16];
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3f;
  const float x = 1.0001f, y = 0.9999f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], x, y);
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <typename F>
static float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  f();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 8 * 1024 * 4);
  cudaMallocManaged(&cyc, 8);
  const int iters = 20000;
  // latency: one warp, one dependent accumulator
  k_mma<1><<<1, 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  printf("mma.sync m16n8k8 tf32 dependent-issue latency: %.1f cycles\n", (double)*cyc / iters);
  k_mma<8><<<1, 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  printf("one warp, 8 independent accumulators: %.1f cycles per mma\n", (double)*cyc / iters / 8);
  for (int warps : {4, 8, 16, 32}) {
    const int blocks = 148 * 2;
    float ms = time_ms([&] { k_mma<8><<<blocks, warps * 32>>>(out, iters, cyc); });
    const double flops = 2.0 * 16 * 8 * 8 * 8.0 * iters * warps * blocks;
    printf("%2d warps/CTA x %d CTAs: %.1f TFLOP/s dense tf32 (legacy mma.sync)\n", warps, blocks, flops / ms / 1e9);
  }
  {
    const int blocks = 148 * 2, threads = 1024;
    float ms = time_ms([&] { k_ffma<<<blocks, threads>>>(out, iters, cyc); });
    printf("FFMA: %.1f TFLOP/s fp32\n", 2.0 * 16 * iters * (double)threads * blocks / ms / 1e9);
  }
  return 0;
}
