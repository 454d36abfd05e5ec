// HBM-bound normalisation passes over fp16 token-major (NHWC) activations.
//   GroupNorm32 / Normalize : util.py:223-225 (eps 1e-5, fp32 statistics), attention.py:75-76 (eps 1e-6)
//   LayerNorm               : attention.py:294-295, 320-322
// Statistics are accumulated in fp32.  SiLU (openaimodel.py:184,208,462) is fused into the
// GroupNorm apply pass.
#include "../../include/idiff_b200.h"
#include "common.cuh"
#include "host.cuh"

namespace idiff {

constexpr int GN_THREADS = 256;
constexpr int GN_MAX_GROUPS = 32;

// grid (chunks, B). Each thread owns channel pairs cp = t, t+256, ... and walks the pixel chunk.
__global__ void __launch_bounds__(GN_THREADS)
gn_stats_kernel(const __half2* __restrict__ x, float* __restrict__ stats, int hw, int C, int groups,
                int pix_per_block) {
  __shared__ float sm[GN_MAX_GROUPS * 2];
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(hw, p0 + pix_per_block);
  const int CP = C >> 1;
  const int cpg = C / groups;
  if (threadIdx.x < groups * 2) sm[threadIdx.x] = 0.f;
  __syncthreads();
  const __half2* xb = x + (long)b * hw * CP;
  for (int cp = threadIdx.x; cp < CP; cp += GN_THREADS) {
    float s = 0.f, ss = 0.f;
    for (int pix = p0; pix < p1; ++pix) {
      const float2 v = __half22float2(xb[(long)pix * CP + cp]);
      s += v.x + v.y;
      ss += v.x * v.x + v.y * v.y;
    }
    const int g = (2 * cp) / cpg;
    atomicAdd(&sm[2 * g], s);
    atomicAdd(&sm[2 * g + 1], ss);
  }
  __syncthreads();
  if (threadIdx.x < groups * 2) atomicAdd(&stats[(long)b * groups * 2 + threadIdx.x], sm[threadIdx.x]);
}

// grid (chunks, B); dynamic smem: 2*C floats (per-channel scale / shift).
__global__ void __launch_bounds__(GN_THREADS)
gn_apply_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, const float* __restrict__ gamma,
                const float* __restrict__ beta, const float* __restrict__ stats, int hw, int C,
                int groups, float eps, int fuse_silu, int pix_per_block) {
  extern __shared__ float sm_ab[];
  float* sa = sm_ab;
  float* sb = sm_ab + C;
  const int b = blockIdx.y;
  const int cpg = C / groups;
  const float inv_n = 1.0f / (float)((long)cpg * hw);
  for (int c = threadIdx.x; c < C; c += GN_THREADS) {
    const int g = c / cpg;
    const float s = stats[((long)b * groups + g) * 2];
    const float ss = stats[((long)b * groups + g) * 2 + 1];
    const float mean = s * inv_n;
    const float var = fmaxf(ss * inv_n - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    const float a = rstd * gamma[c];
    sa[c] = a;
    sb[c] = beta[c] - mean * a;
  }
  __syncthreads();
  const int CV = C >> 3;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(hw, p0 + pix_per_block);
  const long base = (long)b * hw * CV;
  const int total = (p1 - p0) * CV;
  for (int i = threadIdx.x; i < total; i += GN_THREADS) {
    const int pix = p0 + i / CV;
    const int cv = i - (i / CV) * CV;
    const uint4 v = x[base + (long)pix * CV + cv];
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_half2(u[j]);
      const int c = cv * 8 + 2 * j;
      float r0 = f.x * sa[c] + sb[c];
      float r1 = f.y * sa[c + 1] + sb[c + 1];
      if (fuse_silu) {
        r0 = silu_f(r0);
        r1 = silu_f(r1);
      }
      o[j] = pack_half2(r0, r1);
    }
    y[base + (long)pix * CV + cv] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// One warp per row; C % 8 == 0, C <= 8 * 32 * LN_MAX_VEC.
constexpr int LN_MAX_VEC = 5;  // C <= 1280
__global__ void __launch_bounds__(256)
layernorm_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, const float* __restrict__ gamma,
                 const float* __restrict__ beta, int rows, int C, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int CV = C >> 3;
  float v[LN_MAX_VEC * 8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_VEC; ++i) {
    const int cv = lane + i * 32;
    if (cv < CV) {
      const uint4 u = x[(long)row * CV + cv];
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_half2(w[j]);
        v[i * 8 + 2 * j] = f.x;
        v[i * 8 + 2 * j + 1] = f.y;
        s += f.x + f.y;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_VEC; ++i) {
    const int cv = lane + i * 32;
    if (cv < CV) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i * 8 + j] - mean;
        ss += d * d;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss / (float)C + eps);
#pragma unroll
  for (int i = 0; i < LN_MAX_VEC; ++i) {
    const int cv = lane + i * 32;
    if (cv < CV) {
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = cv * 8 + 2 * j;
        const float r0 = (v[i * 8 + 2 * j] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
        const float r1 = (v[i * 8 + 2 * j + 1] - mean) * rstd * __ldg(gamma + c + 1) + __ldg(beta + c + 1);
        o[j] = pack_half2(r0, r1);
      }
      y[(long)row * CV + cv] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

}  // namespace idiff

extern "C" int idiff_groupnorm(const void* x, void* y, const float* gamma, const float* beta,
                               float* stats_ws, int batch, int hw, int channels, int groups,
                               float eps, int fuse_silu, void* stream) {
  using namespace idiff;
  IDIFF_REQUIRE(x && y && gamma && beta && stats_ws, "idiff_groupnorm: null pointer argument");
  IDIFF_REQUIRE(groups > 0 && groups <= GN_MAX_GROUPS && channels % groups == 0,
                "idiff_groupnorm: bad groups=%d for C=%d", groups, channels);
  IDIFF_REQUIRE(channels % 8 == 0 && (channels / groups) % 2 == 0,
                "idiff_groupnorm: C=%d must be a multiple of 8 with an even group size", channels);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  IDIFF_CHECK_CUDA(cudaMemsetAsync(stats_ws, 0, sizeof(float) * 2 * groups * batch, s));
  // ~64K elements per block
  int ppb = (65536 + channels - 1) / channels;
  if (ppb < 1) ppb = 1;
  const int chunks = (hw + ppb - 1) / ppb;
  dim3 grid(chunks, batch);
  gn_stats_kernel<<<grid, GN_THREADS, 0, s>>>(reinterpret_cast<const __half2*>(x), stats_ws, hw,
                                              channels, groups, ppb);
  gn_apply_kernel<<<grid, GN_THREADS, 2 * channels * sizeof(float), s>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), gamma, beta, stats_ws, hw,
      channels, groups, eps, fuse_silu, ppb);
  IDIFF_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int idiff_layernorm(const void* x, void* y, const float* gamma, const float* beta,
                               int rows, int channels, float eps, void* stream) {
  using namespace idiff;
  IDIFF_REQUIRE(x && y && gamma && beta, "idiff_layernorm: null pointer argument");
  IDIFF_REQUIRE(channels % 8 == 0 && channels <= 8 * 32 * LN_MAX_VEC,
                "idiff_layernorm: unsupported C=%d", channels);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int rows_per_block = 8;
  layernorm_kernel<<<(rows + rows_per_block - 1) / rows_per_block, 256, 0, s>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), gamma, beta, rows, channels, eps);
  IDIFF_CHECK_CUDA(cudaGetLastError());
  return 0;
}
