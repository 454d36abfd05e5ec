// Micro-benchmark: tcgen05.mma (kind::f16, cta_group::1, M=128, K=16) cost per instruction as a
// function of N, of the accumulator dependency (same TMEM accumulator vs four rotating ones) and of
// the A operand source (shared memory "SS" vs tensor memory "TS").  The d=40 attention kernel issues
// many small UMMAs (N = 64 and N = 16); this measures what one of them really costs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I instancediffusion_b200/csrc \
//          -o tools/micro/umma_bench tools/micro/umma_bench.cu
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "common.cuh"

using namespace idiff;

template <bool A_TMEM>
__global__ void __launch_bounds__(128, 1) k(int n, int rotate, int reps, long long* cyc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint64_t adesc = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem + 16384), 16, 1024);
    const uint32_t idesc = make_idesc_f16(128, n, 0, 0, 0);
    const uint32_t a_tmem = tmem + 496;  // 8 columns of packed fp16 (K = 16)
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      // rotate: N <= 64: four accumulators of 64 columns; larger N: two accumulators of 256 columns
      const uint32_t d = tmem + (rotate ? (n > 64 ? (r & 1) * 256 : (r & 3) * 64) : 0);
      if (A_TMEM) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
            "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(1u)
            : "memory");
      } else {
        umma_f16_ss(d, adesc, bdesc, idesc, 1u);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    cyc[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, 148 * 8);
  const int reps = 4096;
  const int smem_bytes = 49 * 1024 + 1024;
  cudaFuncSetAttribute(k<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaFuncSetAttribute(k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const int ns[] = {16, 64, 128, 160, 192, 256};
  for (int ts = 0; ts < 2; ++ts)
    for (int rotate = 0; rotate < 2; ++rotate)
      for (int n : ns) {
        if (rotate && n > 192 && ts) continue;  // (the TMEM A operand sits in columns 448..)
        for (int rep = 0; rep < 2; ++rep) {  // first pass warms up
          if (ts) k<true><<<148, 128, smem_bytes>>>(n, rotate, reps, cyc);
          else k<false><<<148, 128, smem_bytes>>>(n, rotate, reps, cyc);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("error: %s (ts=%d n=%d)\n", cudaGetErrorString(e), ts, n);
            return 1;
          }
        }
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double c = 0;
        for (int i = 0; i < 148; ++i) c += h[i];
        c /= 148;
        printf("A=%s accumulators=%s N=%3d: %7.1f clk per UMMA (math floor %5.1f)\n", ts ? "tmem" : "smem",
               rotate ? "4 rotating" : "1 (chain) ", n, c / reps, n / 2.0);
      }
  // fixed latency of a short batch: issue `r` UMMAs, commit, wait for the mbarrier
  for (int n : {16, 64, 128})
    for (int r : {1, 3, 16, 19, 38}) {
      for (int rep = 0; rep < 2; ++rep) {
        k<false><<<148, 128, smem_bytes>>>(n, 0, r, cyc);
        cudaDeviceSynchronize();
      }
      long long h[148];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double c = 0;
      for (int i = 0; i < 148; ++i) c += h[i];
      printf("batch of %2d UMMAs (N=%3d) + commit + mbarrier wait: %7.1f clk total\n", r, n, c / 148);
    }
  return 0;
}
