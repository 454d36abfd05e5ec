// Micro-benchmark: MUFU.EX2 (fp32) and FFMA warp-instruction throughput per SM on the target GPU.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/mufu_bench tools/micro/mufu_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = -0.001f * (threadIdx.x + i);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      else asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  for (int mode = 0; mode < 2; ++mode) {
    for (int warps = 4; warps <= 32; warps *= 2) {
      if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, cyc); else k<1><<<148, warps * 32>>>(out, iters, cyc);
      cudaDeviceSynchronize();
      long long h[148];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double c = 0;
      for (int i = 0; i < 148; ++i) c += h[i];
      c /= 148;
      double ops = (double)iters * 8 * warps * 32;
      printf("%s warps/SM=%2d: %.1f lane-ops/clk/SM\n", mode == 0 ? "MUFU.EX2" : "FFMA    ", warps, ops / c);
    }
  }
  return 0;
}
