// tcgen05 GEMM / implicit-GEMM conv3x3 for sm_100a.
//
//   out[M, N] = epilogue( A[M, K] . W[N, K]^T )        fp16 operands, fp32 accumulation in TMEM
//
// One CTA computes a 128 x BN output tile.  Warp roles (192 threads):
//   warp 0      TMA producer: 128B-swizzled A/B tiles of 64 K-elements into a STAGES-deep ring
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16)
//   warps 2..5  epilogue: tcgen05.ld accumulator rows -> bias / SiLU / GEGLU / gate / residual
//               -> fp16 (or fp32 NCHW) global stores
// conv3x3 mode gathers the A tile tap by tap with a 4-D TMA box over the NHWC activation;
// out-of-image taps are zero-filled by the TMA unit (no im2col buffer, no halo copy).
//
// Replaces (reference file:line): attention.py:41,62,121-125,175-179,297,354,363;
// openaimodel.py:109,134,186,205,213,361-363,464; text_grounding_net.py:75-81.
#include "../../include/idiff_b200.h"
#include "common.cuh"
#include "host.cuh"

#include <stdlib.h>

namespace idiff {

namespace v2 {
int gemm_v2(const idiff_gemm_args* a, cudaStream_t stream);  // gemm2.cu
}

constexpr int BM = 128;
constexpr int BK = 64;                    // 64 fp16 = 128 B = one swizzle row
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int GEMM_THREADS = 192;

struct GemmKParams {
  int M, N, K, num_kb;
  // conv geometry (conv != 0)
  int conv, H, W, Bn, PW, PH, PB, tiles_w, tiles_h, kb_per_tap;
  // epilogue
  const float* bias;
  const __half* rowadd;
  const __half* residual;
  void* out;
  int ldo, ldr, ldra, rows_per_batch, flags;
  float gate;
};

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : ((BN == 128) ? 3 : 4);
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int MIN_CTAS = (BN == 256) ? 1 : 2;
};

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, GemmCfg<BN>::MIN_CTAS)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const GemmKParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full_bar = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x;
  const int m_tile = blockIdx.y;
  const int n0 = n_tile * BN;

  // tile origin
  int m0 = m_tile * BM;
  int b0 = 0, h0 = 0, w0 = 0;
  if (p.conv) {
    const int tw = m_tile % p.tiles_w;
    const int th = (m_tile / p.tiles_w) % p.tiles_h;
    const int tb = m_tile / (p.tiles_w * p.tiles_h);
    b0 = tb * p.PB;
    h0 = th * p.PH;
    w0 = tw * p.PW;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc<BN>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();  // the next kernel's prologue may overlap this kernel (host.cuh launch_pdl)
  pdl_wait();               // operands come from earlier kernels: nothing above touched global memory

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
        if (p.conv) {
          const int tap = kb / p.kb_per_tap;
          const int cb = kb - tap * p.kb_per_tap;
          const int ky = tap / 3, kx = tap - ky * 3;
          tma_load_4d(sA + s * A_STAGE_BYTES, &tmA, &full_bar[s], cb * BK, w0 + kx - 1,
                      h0 + ky - 1, b0);
        } else {
          tma_load_2d(sA + s * A_STAGE_BYTES, &tmA, &full_bar[s], kb * BK, m0);
        }
        tma_load_2d(sB + s * Cfg::B_STAGE_BYTES, &tmB, &full_bar[s], kb * BK, n0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(BM, BN, /*fp16*/ 0, 0, 0);
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_base = smem_u32(sA + s * A_STAGE_BYTES);
        const uint32_t b_base = smem_u32(sB + s * Cfg::B_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t adesc = make_smem_desc_sw128(a_base + k * 32, 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(b_base + k * 32, 16, 1024);
          umma_f16_ss(tmem_base, adesc, bdesc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs retire
      }
      umma_commit(tmem_full_bar);  // accumulator complete
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;
    long out_row;
    bool row_ok;
    int batch_idx;
    int pix = 0;
    if (p.conv) {
      const int pw = r % p.PW;
      const int ph_ = (r / p.PW) % p.PH;
      const int pb = r / (p.PW * p.PH);
      const int b = b0 + pb, h = h0 + ph_, w = w0 + pw;
      row_ok = (b < p.Bn) && (h < p.H) && (w < p.W);
      pix = h * p.W + w;
      out_row = (long)b * p.H * p.W + pix;
      batch_idx = b;
    } else {
      out_row = (long)m0 + r;
      row_ok = out_row < p.M;
      batch_idx = (int)(out_row / p.rows_per_batch);
      pix = (int)(out_row - (long)batch_idx * p.rows_per_batch);
    }
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const bool geglu = (p.flags & IDIFF_EPI_GEGLU) != 0;
    const bool do_silu = (p.flags & IDIFF_EPI_SILU) != 0;
    const bool nchw = (p.flags & IDIFF_OUT_F32_NCHW) != 0;
    const int n_out_cols = geglu ? BN / 2 : BN;
    const int out_col0 = geglu ? n_tile * (BN / 2) : n0;
    const int n_out_total = geglu ? p.N / 2 : p.N;

    for (int c0 = 0; c0 < n_out_cols; c0 += 32) {
      if (out_col0 + c0 >= n_out_total) break;  // warp-uniform
      uint32_t v[32];
      tmem_ld_32x32b_x32(trow + c0, v);
      float x[32];
      if (geglu) {
        uint32_t g[32];
        tmem_ld_32x32b_x32(trow + BN / 2 + c0, g);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float val = __uint_as_float(v[j]);
          float gat = __uint_as_float(g[j]);
          if (p.bias) {
            val += __ldg(p.bias + n0 + c0 + j);
            gat += __ldg(p.bias + n0 + BN / 2 + c0 + j);
          }
          x[j] = val * gelu_erf_f(gat);
        }
      } else {
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = n0 + c0 + j;
          float val = __uint_as_float(v[j]);
          if (col < p.N) {
            if (p.bias) val += __ldg(p.bias + col);
            if (p.rowadd && row_ok) val += __half2float(p.rowadd[(long)batch_idx * p.ldra + col]);
          }
          if (do_silu) val = silu_f(val);
          x[j] = val;
        }
      }
      if (!row_ok) continue;
      if (nchw) {
        float* o = reinterpret_cast<float*>(p.out);
        const long hw = p.conv ? (long)p.H * p.W : (long)p.rows_per_batch;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = out_col0 + c0 + j;
          if (col < n_out_total) o[((long)batch_idx * n_out_total + col) * hw + pix] = x[j];
        }
      } else {
        __half* o = reinterpret_cast<__half*>(p.out) + out_row * p.ldo + out_col0 + c0;
        const __half* res =
            p.residual ? p.residual + out_row * p.ldr + out_col0 + c0 : nullptr;
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          if (out_col0 + c0 + j8 * 8 >= n_out_total) break;
          float y[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) y[j] = x[j8 * 8 + j];
          if (res) {
            const uint4 rv = *reinterpret_cast<const uint4*>(res + j8 * 8);
            const uint32_t ru[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = unpack_half2(ru[j]);
              y[2 * j] = f.x + p.gate * y[2 * j];
              y[2 * j + 1] = f.y + p.gate * y[2 * j + 1];
            }
          }
          uint4 ov;
          ov.x = pack_half2(y[0], y[1]);
          ov.y = pack_half2(y[2], y[3]);
          ov.z = pack_half2(y[4], y[5]);
          ov.w = pack_half2(y[6], y[7]);
          *reinterpret_cast<uint4*>(o + j8 * 8) = ov;
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<BN>(tmem_base);
  }
}

static void choose_patch(int H, int W, int B, int* PW, int* PH, int* PB) {
  int pw = 1;
  while (pw * 2 <= 128 && (W % (pw * 2)) == 0) pw *= 2;
  int ph = 1;
  while (pw * ph * 2 <= 128 && ph < H) ph *= 2;
  int pb = 128 / (pw * ph);
  *PW = pw;
  *PH = ph;
  *PB = pb;
  (void)B;
}

template <int BN>
static int launch_gemm(const idiff_gemm_args* a, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  GemmKParams p;
  memset(&p, 0, sizeof(p));
  p.M = a->M;
  p.N = a->N;
  p.K = a->K;
  p.num_kb = (a->K + BK - 1) / BK;
  p.bias = a->bias;
  p.rowadd = reinterpret_cast<const __half*>(a->rowadd);
  p.residual = reinterpret_cast<const __half*>(a->residual);
  p.out = a->out;
  p.ldo = a->ldo;
  p.ldr = a->ldr;
  p.ldra = a->ldra > 0 ? a->ldra : a->N;
  p.rows_per_batch = a->rows_per_batch > 0 ? a->rows_per_batch : a->M;
  p.flags = a->flags;
  p.gate = a->gate;

  CUtensorMap tmA, tmB;
  int m_tiles;
  if (a->conv_h > 0) {
    const int H = a->conv_h, W = a->conv_w, B = a->conv_b, C = a->conv_cin;
    IDIFF_REQUIRE(C % BK == 0, "conv3x3: Cin=%d must be a multiple of %d", C, BK);
    IDIFF_REQUIRE(a->K == 9 * C, "conv3x3: K=%d must equal 9*Cin=%d", a->K, 9 * C);
    IDIFF_REQUIRE(a->M == B * H * W, "conv3x3: M=%d must equal B*H*W=%d", a->M, B * H * W);
    p.conv = 1;
    p.H = H;
    p.W = W;
    p.Bn = B;
    choose_patch(H, W, B, &p.PW, &p.PH, &p.PB);
    p.tiles_w = W / p.PW;
    p.tiles_h = (H + p.PH - 1) / p.PH;
    const int tiles_b = (B + p.PB - 1) / p.PB;
    p.kb_per_tap = C / BK;
    p.rows_per_batch = H * W;
    m_tiles = p.tiles_w * p.tiles_h * tiles_b;
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    const uint32_t box[4] = {(uint32_t)BK, (uint32_t)p.PW, (uint32_t)p.PH, (uint32_t)p.PB};
    if (encode_tmap_f16(&tmA, a->a, 4, dims, strides, box)) return -1;
  } else {
    m_tiles = (a->M + BM - 1) / BM;
    const uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->M};
    const uint64_t strides[1] = {(uint64_t)a->lda * 2};
    const uint32_t box[2] = {(uint32_t)BK, (uint32_t)BM};
    if (encode_tmap_f16(&tmA, a->a, 2, dims, strides, box)) return -1;
  }
  {
    const uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->N};
    const uint64_t strides[1] = {(uint64_t)a->ldw * 2};
    const uint32_t box[2] = {(uint32_t)BK, (uint32_t)BN};
    if (encode_tmap_f16(&tmB, a->w, 2, dims, strides, box)) return -1;
  }
  const int n_tiles = (a->N + BN - 1) / BN;
  static bool attr_set = false;
  if (!attr_set) {
    IDIFF_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel<BN>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          Cfg::SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid(n_tiles, m_tiles, 1);
  IDIFF_CHECK_CUDA(launch_pdl(gemm_kernel<BN>, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, tmA, tmB, p));
  IDIFF_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace idiff

extern "C" int idiff_gemm(const idiff_gemm_args* a, void* stream) {
  using namespace idiff;
  IDIFF_REQUIRE(a && a->a && a->w && a->out, "idiff_gemm: null pointer argument");
  IDIFF_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "idiff_gemm: bad shape M=%d N=%d K=%d", a->M, a->N, a->K);
  const bool geglu = (a->flags & IDIFF_EPI_GEGLU) != 0;
  const bool nchw = (a->flags & IDIFF_OUT_F32_NCHW) != 0;
  if (geglu) {
    IDIFF_REQUIRE(a->N % 256 == 0, "idiff_gemm: GEGLU needs N %% 256 == 0 (N=%d)", a->N);
    IDIFF_REQUIRE(!a->residual && !a->rowadd && !nchw, "idiff_gemm: GEGLU excludes residual/rowadd/NCHW");
  }
  if (!nchw) {
    IDIFF_REQUIRE(a->N % 8 == 0, "idiff_gemm: N=%d must be a multiple of 8 for fp16 output", a->N);
    IDIFF_REQUIRE(a->ldo % 8 == 0, "idiff_gemm: ldo=%d must be a multiple of 8", a->ldo);
    IDIFF_REQUIRE((reinterpret_cast<uintptr_t>(a->out) & 15) == 0, "idiff_gemm: out not 16B aligned");
    if (a->residual) {
      IDIFF_REQUIRE(a->ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(a->residual) & 15) == 0,
                    "idiff_gemm: residual must be 16B aligned with ldr %% 8 == 0");
    }
  } else {
    IDIFF_REQUIRE(!a->residual, "idiff_gemm: NCHW fp32 output excludes residual");
  }
  // gemm2.cu (persistent / stream-K / wide tiles) is the production kernel; the first-generation
  // kernel in this file stays selectable for A/B measurements (IDIFF_GEMM_V1=1).
  static const bool use_v1_env = []() {
    const char* e = getenv("IDIFF_GEMM_V1");
    return e && e[0] == '1';
  }();
  const bool use_v1 = use_v1_env && !geglu;  // the v1 kernel's 128-wide tiles predate the 128-row GEGLU groups
  if (use_v1) return launch_gemm<128>(a, reinterpret_cast<cudaStream_t>(stream));
  return v2::gemm_v2(a, reinterpret_cast<cudaStream_t>(stream));
}
