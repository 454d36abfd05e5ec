// Micro-benchmark 2: what spaces out the UMMAs of a real GEMM mainloop?  One CTA per SM, M=128, N=160 / 256,
// K=16, SS operands, one accumulator chain.  Variants:
//   commit_every = c : a tcgen05.commit (to a rotating, never-awaited mbarrier) after every c UMMAs (0 = none)
//   ldtm = 1         : warps 4..7 stream tcgen05.ld.32x32b.x16 from the OTHER 256 accumulator columns meanwhile
//   wait = 1         : the issuing thread waits (mbarrier) for a "full" barrier that a producer thread arrives on
//                      right after it saw the matching "empty" commit (the two-barrier ring of the GEMM, no loads)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I instancediffusion_b200/csrc \
//          -o tools/micro/umma_bench2 tools/micro/umma_bench2.cu
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "common.cuh"

using namespace idiff;

constexpr int STAGES = 6;

template <int N_, int CE, bool RING>  // CE: commits per group of four UMMAs (0 / 1); compile-time so that the loop is lean
__global__ void __launch_bounds__(256, 1) k(int ldtm, int reps, long long* cyc) {
  constexpr int n = N_;
  constexpr bool ring = RING;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, cbar[STAGES], fbar[STAGES];
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&cbar[s], 1);
      mbar_init(&fbar[s], 1);
    }
    stop = 0;
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    const uint64_t adesc = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem + 16384), 16, 1024);
    const uint32_t idesc = make_idesc_f16(128, n, 0, 0, 0);
    const long long t0 = clock64();
    int s = 0, ph = 0;
    for (int g = 0; g < reps / 4; ++g) {
      if (RING) mbar_wait(&fbar[s], ph);  // "operands landed" (arrives once the stage was released)
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) umma_f16_ss(tmem, adesc, bdesc, idesc, 1u);
      if (CE) {
        umma_commit(&cbar[s]);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    cyc[blockIdx.x] = t1 - t0;
    stop = 1;
  } else if (threadIdx.x == 32 && ring) {
    // producer of the ring: the first STAGES "full" arrivals are free, then one per released stage
    int s = 0;
    for (int j = 0; j < reps / 4; ++j) {
      if (j >= STAGES) mbar_wait(&cbar[s], ((j / STAGES) - 1) & 1);  // the (j/STAGES - 1)-th completion of this stage's commit
      mbar_arrive(&fbar[s]);
      if (++s == STAGES) s = 0;
    }
  } else if (warp >= 4 && ldtm) {
    uint32_t v[16], acc = 0;
    const uint32_t base = tmem + 256 + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    while (!stop) {
#pragma unroll 1
      for (int c = 0; c < 256; c += 16) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(base + c) : "memory");
        tmem_ld_wait();
        acc += v[0] + v[15];
      }
    }
    if (acc == 0x12345678u) cyc[200] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

template <int N_, int CE, bool RING>
static void run(int ldtm, long long* cyc) {
  const int reps = 4096;
  const int smem_bytes = 65 * 1024 + 1024;
  cudaFuncSetAttribute(k<N_, CE, RING>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  for (int rep = 0; rep < 2; ++rep) {
    k<N_, CE, RING><<<148, 256, smem_bytes>>>(ldtm, reps, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
  }
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0;
  for (int i = 0; i < 148; ++i) c += h[i];
  printf("N=%3d commit per 4 UMMAs=%d ldtm=%d ring=%d: %7.1f clk per UMMA (math floor %5.1f)\n", N_, CE, ldtm, (int)RING,
         c / 148 / reps, N_ / 2.0);
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, 256 * 8);
  run<160, 0, false>(0, cyc);
  run<160, 1, false>(0, cyc);
  run<160, 0, false>(1, cyc);
  run<160, 1, false>(1, cyc);
  run<160, 1, true>(0, cyc);
  run<160, 1, true>(1, cyc);
  run<256, 0, false>(0, cyc);
  run<256, 1, false>(0, cyc);
  run<256, 1, false>(1, cyc);
  run<256, 1, true>(0, cyc);
  run<256, 1, true>(1, cyc);
  return 0;
}
