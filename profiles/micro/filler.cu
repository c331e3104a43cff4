// filler.cu -- a do-nothing kernel that keeps N CTAs resident (one per SM through its shared-memory request) for a given
// number of cycles; used by profiles/ppo_low_grid_probe.py to ask whether the 8-CTA persistent PPO launch runs at a
// different speed when the rest of the chip is occupied (B300_MICROARCH.md, "I-cache": issue throttle at low grid).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC -o libfiller.so filler.cu
#include <cuda_runtime.h>
extern "C" __global__ void k_filler(long long cycles, int sleep_ns, int fma_work, float* sink) {
  extern __shared__ float sm[];
  const long long t0 = clock64();
  float acc = (float)threadIdx.x;
  while (clock64() - t0 < cycles) {
    if (sleep_ns) __nanosleep(sleep_ns);
    for (int i = 0; i < fma_work; ++i) acc = fmaf(acc, 1.0000001f, 0.5f);
  }
  if (acc == 12345.678f) sink[0] = acc + sm[0];
}
extern "C" int launch_filler(int blocks, int threads, int smem_bytes, long long cycles, int sleep_ns, int fma_work, float* sink,
                             void* stream) {
  cudaFuncSetAttribute(k_filler, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  k_filler<<<blocks, threads, smem_bytes, (cudaStream_t)stream>>>(cycles, sleep_ns, fma_work, sink);
  return (int)cudaGetLastError();
}
