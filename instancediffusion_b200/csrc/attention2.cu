// Attention v2 for sm_100a (head_dim 40, the 64x64-resolution level that carries 83 % of the
// attention time): two 128-query tiles per CTA processed in ping-pong.
//
// What changed against attention.cu, and why (measured: profiles/r1_v0_*):
//   * 2 Q tiles / 8 softmax warps per CTA: the K/V tiles are fetched once for 256 queries and each
//     SM sub-partition has two softmax warps to hide latency; S_q(j+1) = Q_q K(j+1)^T is issued as
//     soon as softmax_q(j) has pulled S_q(j) into registers, so the tensor pipe works while the
//     exponentials run.
//   * single pass over S: all 128 scores of a row are read from TMEM once into registers.
//   * exp2 on packed halves (ex2.approx.ftz.f16x2): half the MUFU work, and the result is already
//     the fp16 pair the P.V product wants.  d=40 attention is exp-bound.
//   * the row sums l = sum_k P are computed by the tensor core as P . 1 (a 16-column block of
//     ones in shared memory, UMMA N=16) -- exactly the rounded P that multiplies V.
//   * lazy rescaling: the running maximum used for the exponent is only advanced when it grew by
//     more than 2^8 in the exp2 domain, so the O/L rescale in TMEM is rare (P <= 256 fits fp16).
// Warp roles (384 threads = 3 warpgroups): warp 0 TMA producer, warp 1 TMEM allocator + UMMA issuer
// (warps 2-3 idle; the group releases registers with setmaxnreg.dec), warps 4..7 softmax of Q tile 0,
// warps 8..11 softmax of Q tile 1 (one thread per query row, 232 registers via setmaxnreg.inc: a
// row's 128 scores and its 64 packed probabilities live in registers).
#include "../../include/idiff_b200.h"
#include "common.cuh"
#include "host.cuh"

namespace idiff {
namespace att2 {

constexpr int THREADS = 384;  // 3 warpgroups: {TMA, UMMA, 2 idle} + softmax(Q0) + softmax(Q1)
constexpr int BQ = 128;
constexpr int BKV = 128;

struct Params {
  int heads, nq, n0, n1, kv1_broadcast;
  float scale_log2e;
  __half* out;
  int out_ld;
};

template <int D>
struct Cfg {
  static_assert(D <= 64, "attention2 handles one 64-wide d chunk");
  static constexpr int KSTEPS = (D + 15) / 16;
  static constexpr int DV = 64;
  static constexpr int STAGES = 3;
  static constexpr int Q_BYTES = BQ * 128;         // per Q tile
  static constexpr int KV_BYTES = BKV * 128;       // one of K or V
  static constexpr int P_BYTES = 2 * BQ * 128;     // per Q tile: two 64-key chunks
  static constexpr int ONES_BYTES = 4096;          // [2 chunks][16 rows][128 B] of fp16 1.0
  static constexpr int SMEM_BYTES = 2 * Q_BYTES + STAGES * 2 * KV_BYTES + 2 * P_BYTES + ONES_BYTES + 1024 + 256;
  // TMEM columns
  static constexpr int S_COL0 = 0, S_COL1 = 128;
  static constexpr int O_COL0 = 256, L_COL0 = 320, O_COL1 = 336, L_COL1 = 400;
};

IDIFF_DEVICE void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
IDIFF_DEVICE void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

IDIFF_DEVICE uint32_t exp2_pack_h2(float x0, float x1) {
  // {2^x0, 2^x1} as packed fp16.  ex2.approx.f16x2 was measured here first: on sm_100 it lowers to
  // two scalar MUFU.EX2.F16 + a PRMT and ran at about half the fp32 MUFU rate (profiles/), so the
  // exponentials stay fp32 MUFU.EX2 followed by one F2FP pack.
  return pack_half2(exp2_approx(x0), exp2_approx(x1));
}

template <int D>
__global__ void __launch_bounds__(THREADS, 1)
attention2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK0,
                  const __grid_constant__ CUtensorMap tmV0, const __grid_constant__ CUtensorMap tmK1,
                  const __grid_constant__ CUtensorMap tmV1, const Params p) {
  using C = Cfg<D>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                                  // [2][16 KiB]
  uint8_t* sK = sQ + 2 * C::Q_BYTES;                   // [STAGES][16 KiB]
  uint8_t* sV = sK + STAGES * C::KV_BYTES;             // [STAGES][16 KiB]
  uint8_t* sP = sV + STAGES * C::KV_BYTES;             // [2][32 KiB]
  uint8_t* sOnes = sP + 2 * C::P_BYTES;                // 4 KiB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + C::ONES_BYTES);
  uint64_t* q_full = bars;               // 1
  uint64_t* k_full = bars + 1;           // STAGES
  uint64_t* v_full = k_full + STAGES;    // STAGES
  uint64_t* kv_empty = v_full + STAGES;  // STAGES
  uint64_t* s_full = kv_empty + STAGES;  // 2
  uint64_t* s_free = s_full + 2;         // 2 (128 arrivals)
  uint64_t* p_full = s_free + 2;         // 2 (128 arrivals)
  uint64_t* pv_done = p_full + 2;        // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 2 * BQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int T0 = (p.n0 + BKV - 1) / BKV;
  const int T1 = (p.n1 + BKV - 1) / BKV;
  const int T = T0 + T1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK0);
    tma_prefetch_desc(&tmV0);
    mbar_init(q_full, 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int q = 0; q < 2; ++q) {
      mbar_init(&s_full[q], 1);
      mbar_init(&s_free[q], 128);
      mbar_init(&p_full[q], 128);
      mbar_init(&pv_done[q], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  // the block of ones that turns the row sums into a UMMA (every element is 1.0, so any layout works)
  for (int i = threadIdx.x; i < C::ONES_BYTES / 4; i += THREADS)
    reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3C003C00u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Register re-partition between warpgroups (the setmaxnreg must sit at the head of each role
  // branch so that ptxas allocates the branch bodies against the new limits).
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;\n");
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * C::Q_BYTES);
      tma_load_4d(sQ, &tmQ, q_full, 0, h, q0, b);
      tma_load_4d(sQ + C::Q_BYTES, &tmQ, q_full, 0, h, q0 + BQ, b);
      for (int j = 0; j < T; ++j) {
        const int s = j % STAGES;
        mbar_wait(&kv_empty[s], ((j / STAGES) & 1) ^ 1);
        const bool seg1 = j >= T0;
        const int row = (seg1 ? (j - T0) : j) * BKV;
        const int bb = seg1 ? (p.kv1_broadcast ? 0 : b) : b;
        mbar_expect_tx(&k_full[s], C::KV_BYTES);
        tma_load_4d(sK + s * C::KV_BYTES, seg1 ? &tmK1 : &tmK0, &k_full[s], 0, h, row, bb);
        mbar_expect_tx(&v_full[s], C::KV_BYTES);
        tma_load_4d(sV + s * C::KV_BYTES, seg1 ? &tmV1 : &tmV0, &v_full[s], 0, h, row, bb);
      }
    }
  } else if (warp == 1) {
    // ===================== UMMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc_f16(BQ, BKV, 0, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_f16(BQ, C::DV, 0, 0, /*B MN-major*/ 1);
      constexpr uint32_t idesc_l = make_idesc_f16(BQ, 16, 0, 0, 0);
      // Every descriptor here shares its high word (SBO = 1024 B, version 1, SWIZZLE_128B); the low
      // word is (address >> 4) | (LBO >> 4) << 16, so stepping an operand by X bytes is lo += X >> 4.
      // The issuing thread therefore spends one integer add per UMMA (the tiles are small: N = 64 /
      // 16 UMMAs last 32 / 8 clocks, so descriptor arithmetic in the issue loop would dominate).
      constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
      constexpr uint32_t LBO_K = (16u >> 4) << 16;             // K-major operands (unused field)
      constexpr uint32_t LBO_V = ((BKV * 128u) >> 4) << 16;    // MN-major V: next 64-wide d chunk
      const uint32_t q_lo0 = (smem_u32(sQ) >> 4) | LBO_K;
      const uint32_t p_lo0 = (smem_u32(sP) >> 4) | LBO_K;
      const uint32_t ones_lo = (smem_u32(sOnes) >> 4) | LBO_K;
      const uint32_t k_lo0 = (smem_u32(sK) >> 4) | LBO_K;
      const uint32_t v_lo0 = (smem_u32(sV) >> 4) | LBO_V;
      auto umma_lo = [&](uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "mov.b64 da, {%1, %3};\n\t"
            "mov.b64 db, {%2, %3};\n\t"
            "setp.ne.b32 p, %5, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}\n" ::"r"(d_tmem),
            "r"(a_lo), "r"(b_lo), "r"(DESC_HI), "r"(idesc), "r"(acc)
            : "memory");
      };
      auto issue_qk = [&](int q, int j) {
        const uint32_t q_lo = q_lo0 + q * (C::Q_BYTES >> 4);
        const uint32_t k_lo = k_lo0 + (j % STAGES) * (C::KV_BYTES >> 4);
        const uint32_t d_tmem = tmem_base + (q ? C::S_COL1 : C::S_COL0);
#pragma unroll
        for (int kk = 0; kk < C::KSTEPS; ++kk)
          umma_lo(d_tmem, q_lo + kk * 2, k_lo + kk * 2, idesc_qk, kk > 0 ? 1u : 0u);
        umma_commit(&s_full[q]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(0, 0);
      issue_qk(1, 0);
      for (int j = 0; j < T; ++j) {
        const int st = j % STAGES;
        if (j + 1 < T) {
          mbar_wait(&k_full[(j + 1) % STAGES], ((j + 1) / STAGES) & 1);
          for (int q = 0; q < 2; ++q) {
            mbar_wait(&s_free[q], j & 1);  // softmax_q(j) holds S_q(j) in registers
            tc_fence_after();
            issue_qk(q, j + 1);
          }
        }
        mbar_wait(&v_full[st], (j / STAGES) & 1);
        const uint32_t v_lo = v_lo0 + st * (C::KV_BYTES >> 4);
        for (int q = 0; q < 2; ++q) {
          mbar_wait(&p_full[q], j & 1);
          tc_fence_after();
          const uint32_t p_lo = p_lo0 + q * (C::P_BYTES >> 4);
          const uint32_t o_tmem = tmem_base + (q ? C::O_COL1 : C::O_COL0);
          const uint32_t l_tmem = tmem_base + (q ? C::L_COL1 : C::L_COL0);
          const uint32_t acc0 = j > 0 ? 1u : 0u;
#pragma unroll
          for (int kk = 0; kk < BKV / 16; ++kk) {
            const uint32_t a_lo = p_lo + (kk >> 2) * ((BQ * 128) >> 4) + (kk & 3) * 2;
            umma_lo(o_tmem, a_lo, v_lo + kk * (2048 >> 4), idesc_pv, kk > 0 ? 1u : acc0);
            umma_lo(l_tmem, a_lo, ones_lo + (kk >> 2) * (2048 >> 4) + (kk & 3) * 2, idesc_l, kk > 0 ? 1u : acc0);
          }
          umma_commit(&pv_done[q]);
        }
        umma_commit(&kv_empty[st]);
      }
    }
    __syncwarp();
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;\n");
    // ===================== softmax / correction / epilogue =====================
    const int q = (warp - 4) >> 2;  // Q tile of this warp group
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_off + (q ? C::S_COL1 : C::S_COL0);
    const uint32_t o_addr = tmem_base + lane_off + (q ? C::O_COL1 : C::O_COL0);  // O (64) then L (16)
    const float c = p.scale_log2e;
    float m_used = -INFINITY;  // maximum the exponent is taken against (raw score units)
    uint8_t* p_row = sP + q * C::P_BYTES + r * 128;
    const int sw = r & 7;

    for (int j = 0; j < T; ++j) {
      const bool seg1 = j >= T0;
      const int row0 = (seg1 ? (j - T0) : j) * BKV;
      const int nvalid = min(BKV, (seg1 ? p.n1 : p.n0) - row0);
      mbar_wait(&s_full[q], j & 1);
      tc_fence_after();
      float s[BKV];
#pragma unroll
      for (int c0 = 0; c0 < BKV; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(s_addr + c0, v);
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) s[c0 + jj] = __uint_as_float(v[jj]);
      }
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_free[q]);  // S_q may be overwritten by Q_q K(j+1)^T

      if (nvalid < BKV) {
#pragma unroll
        for (int jj = 0; jj < BKV; ++jj)
          if (jj >= nvalid) s[jj] = -INFINITY;
      }
      float m_tile = s[0];
#pragma unroll
      for (int jj = 1; jj < BKV; ++jj) m_tile = fmaxf(m_tile, s[jj]);
      // lazy maximum: only move the reference when it grew by more than 8 in the exp2 domain
      float alpha = 1.0f;
      if ((m_tile - m_used) * c > 8.0f) {
        alpha = exp2_approx((m_used - m_tile) * c);  // first tile: 2^-inf = 0 (O is not read then)
        m_used = m_tile;
      }
      const float mc = m_used * c;
      uint32_t pk[BKV / 2];
#pragma unroll
      for (int jj = 0; jj < BKV; jj += 2)
        pk[jj >> 1] = exp2_pack_h2(fmaf(s[jj], c, -mc), fmaf(s[jj + 1], c, -mc));

      if (j > 0) {
        mbar_wait(&pv_done[q], (j - 1) & 1);  // P buffer free, O/L quiescent
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll
          for (int c0 = 0; c0 < 64; c0 += 32) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(o_addr + c0, o);
            tmem_ld_wait();
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) o[jj] = __float_as_uint(__uint_as_float(o[jj]) * alpha);
            tmem_st_32x32b_x32(o_addr + c0, o);
          }
          uint32_t l[16];
          tmem_ld_32x32b_x16(o_addr + 64, l);
          tmem_ld_wait();
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) l[jj] = __float_as_uint(__uint_as_float(l[jj]) * alpha);
          tmem_st_32x32b_x16(o_addr + 64, l);
          tmem_st_wait();
        }
      }
#pragma unroll
      for (int i = 0; i < BKV / 8; ++i) {
        uint4 val = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
        uint8_t* dst = p_row + (i >> 3) * (BQ * 128) + (((i & 7) ^ sw) << 4);
        *reinterpret_cast<uint4*>(dst) = val;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_full[q]);
    }

    // epilogue: O / l -> fp16
    mbar_wait(&pv_done[q], (T - 1) & 1);
    tc_fence_after();
    uint32_t l[16];
    tmem_ld_32x32b_x16(o_addr + 64, l);
    tmem_ld_wait();
    const float inv_l = 1.0f / __uint_as_float(l[0]);
    const int qrow = q0 + q * BQ + r;
    const bool row_ok = qrow < p.nq;
    __half* orow = p.out + ((long)b * p.nq + qrow) * p.out_ld + h * D;
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
      if (c0 >= D) break;
      uint32_t o[32];
      tmem_ld_32x32b_x32(o_addr + c0, o);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          if (c0 + j8 * 8 < D) {
            uint4 ov;
            ov.x = pack_half2(__uint_as_float(o[j8 * 8 + 0]) * inv_l, __uint_as_float(o[j8 * 8 + 1]) * inv_l);
            ov.y = pack_half2(__uint_as_float(o[j8 * 8 + 2]) * inv_l, __uint_as_float(o[j8 * 8 + 3]) * inv_l);
            ov.z = pack_half2(__uint_as_float(o[j8 * 8 + 4]) * inv_l, __uint_as_float(o[j8 * 8 + 5]) * inv_l);
            ov.w = pack_half2(__uint_as_float(o[j8 * 8 + 6]) * inv_l, __uint_as_float(o[j8 * 8 + 7]) * inv_l);
            *reinterpret_cast<uint4*>(orow + c0 + j8 * 8) = ov;
          }
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

static int make_head_tmap(CUtensorMap* m, const void* base, int d, int heads, int rows, int batch,
                          int ld, int box_rows) {
  const uint64_t dims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)rows, (uint64_t)batch};
  const uint64_t strides[3] = {(uint64_t)d * 2, (uint64_t)ld * 2, (uint64_t)rows * ld * 2};
  const uint32_t box[4] = {64u, 1u, (uint32_t)box_rows, 1u};
  return encode_tmap_f16(m, base, 4, dims, strides, box);
}

int attention_v2_d40(const idiff_attn_args* a, cudaStream_t stream) {
  constexpr int D = 40;
  using C = Cfg<D>;
  CUtensorMap tmQ, tmK0, tmV0, tmK1, tmV1;
  if (make_head_tmap(&tmQ, a->q, D, a->heads, a->nq, a->batch, a->q_ld, BQ)) return -1;
  if (make_head_tmap(&tmK0, a->k0, D, a->heads, a->n0, a->batch, a->k0_ld, BKV)) return -1;
  if (make_head_tmap(&tmV0, a->v0, D, a->heads, a->n0, a->batch, a->v0_ld, BKV)) return -1;
  if (a->n1 > 0) {
    const int b1 = a->kv1_batch == 1 ? 1 : a->batch;
    if (make_head_tmap(&tmK1, a->k1, D, a->heads, a->n1, b1, a->k1_ld, BKV)) return -1;
    if (make_head_tmap(&tmV1, a->v1, D, a->heads, a->n1, b1, a->v1_ld, BKV)) return -1;
  } else {
    tmK1 = tmK0;
    tmV1 = tmV0;
  }
  Params p;
  p.heads = a->heads;
  p.nq = a->nq;
  p.n0 = a->n0;
  p.n1 = a->n1;
  p.kv1_broadcast = (a->kv1_batch == 1) ? 1 : 0;
  p.scale_log2e = a->scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__half*>(a->out);
  p.out_ld = a->out_ld;
  static bool attr_set = false;
  if (!attr_set) {
    IDIFF_CHECK_CUDA(cudaFuncSetAttribute(attention2_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          C::SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid((a->nq + 2 * BQ - 1) / (2 * BQ), a->heads, a->batch);
  attention2_kernel<D><<<grid, THREADS, C::SMEM_BYTES, stream>>>(tmQ, tmK0, tmV0, tmK1, tmV1, p);
  IDIFF_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace att2
}  // namespace idiff
