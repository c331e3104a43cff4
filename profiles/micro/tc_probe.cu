// tc_probe.cu -- validates the tcgen05 descriptor / layout assumptions of csrc/imb_tc.cuh on a real B200 and measures
// the issue rate of the MMA shapes the discriminator kernel uses.  One experiment per process:  tc_probe <exp> [variant]
//   1  SS  A K-major [128 x 32] x B K-major [32 x 32]           -> D[128 x 32]          (forward layers)
//   2  SS  A MN-major (M=64, K=128) x B MN-major (N=56, K=128)  -> D[64 x 56]           (weight gradients; tf32 MN-major
//          operands exist only in the SWIZZLE_128B_BASE32B layout: [row][32 floats], 32-byte chunk c of row r at c ^ (r & 3))
//   3  TS  A in TMEM [128 x 32] x B K-major [32 x 32]           -> D[128 x 32]          (forward layers, A from TMEM)
//   4  SS  A K-major [128 x 32] x B MN-major (N=32, K=32)       -> D[128 x 32]          (backward dz1 = dz2 . W2)
//   5  timing of the shapes above (cycles per MMA at steady state)
// variant 1 swaps LBO and SBO (in case the field meaning is the other way round).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I imitation_b200/csrc -o profiles/micro/tc_probe profiles/micro/tc_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "imb_tc.cuh"

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e = (x);                                                             \
    if (e != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                       \
    }                                                                                \
  } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t* b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ bool bar_try(uint64_t* b, uint32_t par) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(s32(b)), "r"(par)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bar_wait(uint64_t* b, uint32_t par) {
  for (uint32_t s = 0; !bar_try(b, par); ++s)
    if (s > (1u << 24)) __trap();
}

// chunked operand tile: element (row, col) of a [ROWS x COLS] matrix at float index (col/4)*ROWS*4 + row*4 + col%4
__host__ __device__ inline int cidx(int row, int col, int ROWS) { return (col >> 2) * ROWS * 4 + row * 4 + (col & 3); }

// SW128_32B MN-major source tile: [ROWS][32 floats] per atom (atoms of 32 columns ROWS*32 floats apart); the 32-byte
// chunk (8 floats) c of row r sits at chunk position c ^ (r & 3)
__host__ __device__ inline int sidx(int row, int col, int ROWS) {
  const int atom = col >> 5, c = col & 31;
  return atom * ROWS * 32 + row * 32 + ((((c >> 3) ^ (row & 3))) << 3) + (c & 7);
}
__device__ __forceinline__ uint64_t desc_sw32b(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return tc::smem_desc(saddr, lbo, sbo) | ((uint64_t)1 << 61);
}

struct Args {
  int exp, variant;
  const float* a;  // chunked A-source tile
  const float* b;  // chunked B-source tile
  float* d;        // TMEM dump [128 lanes][64 cols]
  long long* cyc;
};

__global__ void __launch_bounds__(128) k_probe(Args g) {
  extern __shared__ __align__(1024) float sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  float* A = sm;              // up to 128 x 64 floats = 32 KB
  float* B = sm + 128 * 64;   // up to 128 x 64
  for (int i = tid; i < 128 * 64; i += 128) {
    A[i] = g.a[i];
    B[i] = g.b[i];
  }
  if (tid == 0) bar_init(&bar, 1);
  if (warp == 0) tc::tmem_alloc(&tbase_s, 256);
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tb = tbase_s;
  const uint32_t D = tb;         // accumulator columns [0,64)
  const uint32_t TA = tb + 64;   // A operand columns [64,128)
  // zero the accumulator region so untouched cells read as 0
  {
    float z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = 0.f;
    for (int c = 0; c < 64; c += 16) tc::st16(tc::tmem_addr(D, 32 * warp, c), z);
    tc::st_wait();
  }
  if (g.exp == 3) {  // A operand into TMEM: thread = row, 32 K values
    float v[16];
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = A[cidx(tid, h * 16 + i, 128)];
      tc::st16(tc::tmem_addr(TA, 32 * warp, h * 16), v);
    }
    tc::st_wait();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const bool swap = g.variant == 1;
  auto desc = [&](const float* p, uint32_t lbo, uint32_t sbo) {
    return swap ? tc::smem_desc(s32(p), sbo, lbo) : tc::smem_desc(s32(p), lbo, sbo);
  };
  if (tid == 0) {
    if (g.exp == 1 || g.exp == 3 || g.exp == 4) {
      // A: [128 rows x 32 K] K-major: LBO = 128*16 (next K chunk), SBO = 128 (next 8 rows)
      // B (exp 1, 3): W[n][k] as chunked [32 rows(n) x 32 cols(k)]: K-major, LBO = 32*16, SBO = 128
      // B (exp 4): W2[k][n] as chunked [32 rows(k) x 32 cols(n)]: MN-major, SBO = 32*16 (next 4 n), LBO = 128 (next 8 k)
      const uint32_t id = tc::idesc_tf32(128, 32, 0, g.exp == 4 ? 1 : 0);
      for (int ks = 0; ks < 4; ++ks) {
        uint64_t bd;
        if (g.exp == 4)
          bd = desc(B + ks * 8 * 4, 128, 32 * 16);            // advance 8 k rows: 8 * 16 B
        else
          bd = desc(B + ks * 2 * 32 * 4, 32 * 16, 128);       // advance 2 K chunks
        if (g.exp == 3) {
          tc::mma_ts(D, TA + ks * 8, bd, id, ks > 0);
        } else {
          const uint64_t ad = desc(A + ks * 2 * 128 * 4, 128 * 16, 128);
          tc::mma_ss(D, ad, bd, id, ks > 0);
        }
      }
    } else if (g.exp == 2) {
      // A-source [128 rows x 64 units] = two SW128_32B atoms (32 units each, 128*128 B apart); MN-major A: LBO = atom
      // stride, SBO = 512 (next group of four K rows).  B-source [128 rows x 56 feats] likewise.  One MMA = 8 K rows = 1 KB.
      const uint32_t id = tc::idesc_tf32(64, 56, 1, 1);
      for (int ks = 0; ks < 16; ++ks) {
        uint64_t ad = desc_sw32b(s32(A + ks * 8 * 32), 128 * 128, 512);
        uint64_t bd = desc_sw32b(s32(B + ks * 8 * 32), 128 * 128, 512);
        if (swap) {
          ad = desc_sw32b(s32(A + ks * 8 * 32), 512, 128 * 128);
          bd = desc_sw32b(s32(B + ks * 8 * 32), 512, 128 * 128);
        }
        tc::mma_ss(D, ad, bd, id, ks > 0);
      }
    }
    tc::commit(&bar);
  }
  bar_wait(&bar, 0);
  tc::fence_after_sync();
  for (int c = 0; c < 64; c += 16) {
    float v[16];
    tc::ld16(tc::tmem_addr(D, 32 * warp, c), v);
    tc::ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) g.d[tid * 64 + c + i] = v[i];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tb, 256);
}

// timing: `reps` back-to-back MMAs of one shape issued by the elected lane of warp 0 (tight loop: descriptors are
// precomputed, four MMAs per iteration), then commit + wait
template <int MODE>
__global__ void __launch_bounds__(128) k_time(int reps, long long* out) {
  extern __shared__ __align__(1024) float sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 2 * 128 * 64; i += 128) sm[i] = 0.f;
  if (tid == 0) bar_init(&bar, 1);
  if (warp == 0) tc::tmem_alloc(&tbase_s, 512);
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tb = tbase_s;
  if (warp == 0) {
    uint32_t pred = 0;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    if (pred) {
      const uint32_t a0 = s32(sm), b0 = s32(sm + 128 * 64);
      uint64_t ad[4], bd[4];
      uint32_t id = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (MODE == 0 || MODE == 1) {         // M128 N32, B K-major no swizzle
          ad[k] = tc::smem_desc(a0 + k * 4096, 2048, 128);
          bd[k] = tc::smem_desc(b0 + k * 1024, 512, 128);
          id = tc::idesc_tf32(128, 32, 0, 0);
        } else if (MODE == 2) {               // wgrad M64 N56, both MN-major SW128_32B
          ad[k] = desc_sw32b(a0 + k * 1024, 16384, 512);
          bd[k] = desc_sw32b(b0 + k * 1024, 16384, 512);
          id = tc::idesc_tf32(64, 56, 1, 1);
        } else if (MODE == 3) {               // wgrad M64 N64
          ad[k] = desc_sw32b(a0 + k * 1024, 16384, 512);
          bd[k] = desc_sw32b(b0 + k * 1024, 16384, 512);
          id = tc::idesc_tf32(64, 64, 1, 1);
        } else if (MODE == 4) {               // wgrad M128 N64 (all four A atoms)
          ad[k] = desc_sw32b(a0 + k * 1024, 4096, 512);
          bd[k] = desc_sw32b(b0 + k * 1024, 16384, 512);
          id = tc::idesc_tf32(128, 64, 1, 1);
        } else if (MODE == 5) {               // M128 N256 K-major A, MN-major B: the peak-rate reference shape
          ad[k] = tc::smem_desc(a0 + k * 4096, 2048, 128);
          bd[k] = desc_sw32b(b0 + k * 1024, 4096, 512);
          id = tc::idesc_tf32(128, 256, 0, 1);
        } else if (MODE == 6) {               // M128 N64 TS
          bd[k] = tc::smem_desc(b0 + k * 2048, 1024, 128);
          id = tc::idesc_tf32(128, 64, 0, 0);
        } else if (MODE == 7) {               // M128 N16 TS
          bd[k] = tc::smem_desc(b0 + k * 512, 256, 128);
          id = tc::idesc_tf32(128, 16, 0, 0);
        }
      }
      uint32_t par = 0;
      for (int round = 0; round < 2; ++round) {
        const long long t0 = clock64();
#pragma unroll 1
        for (int r = 0; r < reps; r += 4) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (MODE == 1 || MODE == 6 || MODE == 7)
              tc::mma_ts(tb, tb + 256 + k * 8, bd[k], id, 1);
            else
              tc::mma_ss(tb, ad[k], bd[k], id, 1);
          }
        }
        const long long t1 = clock64();
        tc::commit(&bar);
        bar_wait(&bar, par);
        par ^= 1;
        const long long t2 = clock64();
        out[2 * round] = t2 - t0;
        out[2 * round + 1] = t1 - t0;
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tb, 512);
}

template <int MODE>
static void run_time(const char* name, long long* d, size_t smem) {
  CK(cudaFuncSetAttribute(k_time<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int reps = 1024;
  k_time<MODE><<<1, 128, smem>>>(reps, d);
  CK(cudaDeviceSynchronize());
  long long h[4];
  CK(cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost));
  printf("timing %-34s : %.1f cycles / MMA complete, %.1f cycles / MMA issue (first round %.1f)\n", name,
         (double)h[2] / reps, (double)h[3] / reps, (double)h[0] / reps);
}

int main(int argc, char** argv) {
  const int exp = argc > 1 ? atoi(argv[1]) : 1;
  const int variant = argc > 2 ? atoi(argv[2]) : 0;
  CK(cudaSetDevice(0));
  const size_t smem = 2 * 128 * 64 * 4;
  if (exp == 5) {
    long long* d;
    CK(cudaMalloc(&d, 64));
    run_time<0>("SS M128 N32 K/K", d, smem);
    run_time<1>("TS M128 N32", d, smem);
    run_time<2>("SS M64 N56 MN/MN sw128_32B", d, smem);
    run_time<3>("SS M64 N64 MN/MN", d, smem);
    run_time<4>("SS M128 N64 MN/MN", d, smem);
    run_time<5>("SS M128 N256 K/MN (peak shape)", d, smem);
    run_time<6>("TS M128 N64", d, smem);
    run_time<7>("TS M128 N16", d, smem);
    return 0;
  }
  // small integers / 8: exact in tf32, every partial sum exact in fp32
  std::vector<float> a(128 * 64, 0.f), b(128 * 64, 0.f), dref(128 * 64, 0.f), dout(128 * 64, 0.f);
  srand(1234 + exp);
  auto rv = []() { return (float)((rand() % 17) - 8) / 8.0f; };
  int M = 128, N = 32, K = 32;
  std::vector<float> Am, Bm;  // logical matrices: Am[m][k], Bm[n][k]
  if (exp == 1 || exp == 3 || exp == 4) {
    Am.resize(128 * 32);
    Bm.resize(32 * 32);
    for (auto& v : Am) v = rv();
    for (auto& v : Bm) v = rv();
    for (int m = 0; m < 128; ++m)
      for (int k = 0; k < 32; ++k) a[cidx(m, k, 128)] = Am[m * 32 + k];
    for (int n = 0; n < 32; ++n)
      for (int k = 0; k < 32; ++k) {
        if (exp == 4)
          b[cidx(k, n, 32)] = Bm[n * 32 + k];  // stored [k rows][n cols] (W2[k][n]): MN-major B
        else
          b[cidx(n, k, 32)] = Bm[n * 32 + k];  // stored [n rows][k cols]: K-major B
      }
  } else if (exp == 2) {
    M = 64; N = 56; K = 128;
    Am.resize(64 * 128);
    Bm.resize(56 * 128);
    for (auto& v : Am) v = rv();
    for (auto& v : Bm) v = rv();
    for (int m = 0; m < 64; ++m)
      for (int k = 0; k < 128; ++k) a[sidx(k, m, 128)] = Am[m * 128 + k];  // source tile [row k][unit m]
    for (int n = 0; n < 56; ++n)
      for (int k = 0; k < 128; ++k) b[sidx(k, n, 128)] = Bm[n * 128 + k];
  } else {
    printf("unknown experiment\n");
    return 1;
  }
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0.f;
      for (int k = 0; k < K; ++k) s += Am[m * K + k] * Bm[n * K + k];
      dref[m * 64 + n] = s;
    }
  float *da, *db, *dd;
  CK(cudaMalloc(&da, a.size() * 4));
  CK(cudaMalloc(&db, b.size() * 4));
  CK(cudaMalloc(&dd, dout.size() * 4));
  CK(cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dd, 0, dout.size() * 4));
  CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  Args g{exp, variant, da, db, dd, nullptr};
  k_probe<<<1, 128, smem>>>(g);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(dout.data(), dd, dout.size() * 4, cudaMemcpyDeviceToHost));
  // compare under the expected lane map
  int bad = 0;
  float maxerr = 0.f;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      const int lane = (M == 64) ? (m % 16) + 32 * (m / 16) : m;
      const float e = fabsf(dout[lane * 64 + n] - dref[m * 64 + n]);
      if (e > 1e-6f) ++bad;
      if (e > maxerr) maxerr = e;
    }
  printf("exp %d variant %d: M=%d N=%d K=%d  mismatches %d / %d  max err %g  -> %s\n", exp, variant, M, N, K, bad, M * N,
         maxerr, bad == 0 ? "OK" : "WRONG");
  if (bad) {
    // help decode: where does D[0][0..3], D[1][0], D[16][0], D[32][0] land?  print a small TMEM corner + nonzero lane set
    printf("ref  D[0][0..3] = %g %g %g %g   D[1][0] = %g  D[16][0] = %g\n", dref[0], dref[1], dref[2], dref[3], dref[64],
           dref[16 * 64]);
    for (int l = 0; l < 4; ++l)
      printf("tmem lane %d: %g %g %g %g %g %g %g %g\n", l, dout[l * 64 + 0], dout[l * 64 + 1], dout[l * 64 + 2],
             dout[l * 64 + 3], dout[l * 64 + 4], dout[l * 64 + 5], dout[l * 64 + 6], dout[l * 64 + 7]);
    printf("lanes with non-zero data:");
    for (int l = 0; l < 128; ++l) {
      bool nz = false;
      for (int c = 0; c < 64; ++c) nz |= dout[l * 64 + c] != 0.f;
      if (nz) printf(" %d", l);
    }
    printf("\n");
  }
  return bad ? 1 : 0;
}
