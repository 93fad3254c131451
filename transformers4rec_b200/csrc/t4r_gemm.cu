// t4r_gemm.cu -- persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M, N] = epilogue( A[M, K] * B[N, K]^T )
//
// A and B arrive as split-bf16 planes (hi, lo; see include/t4r_b200.h), are staged
// into shared memory by TMA (128-byte swizzle, K-major), multiplied by
// tcgen05.mma (kind::f16, bf16 x bf16 -> fp32 in TMEM; three products per K step:
// hi*hi + hi*lo + lo*hi) and drained by four epilogue warps with tcgen05.ld.
//
// Roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-9 =
// epilogue (TMEM lane quadrant = warp_id % 4; warps 2-5 own the first half of a
// tile's columns, warps 6-9 the second half).  Two accumulator stages in TMEM let
// the epilogue of tile i overlap the main loop of tile i+1.  Tiles are visited
// m-fastest so CTAs running at the same time share the B tile in L2.
//
// Epilogues: dense (bias / ReLU / GELU / mask-replace / residual / LayerNorm, fp32
// and split-bf16 outputs) and head (online log-sum-exp partials + label rank).
#include <cuda.h>
#include <math.h>
#include <mutex>

#include "t4r_common.cuh"
#include "t4r_tmem_ld.cuh"
#include "t4r_internal.h"

namespace t4r {

// ----------------------------------------------------------------------------
// tensor maps (driver entry point resolved at run time: no link-time libcuda)
// ----------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

// 2-D bf16 tensor [rows, Kp] row-major, box = [box_rows, 64] with 128B swizzle.
static int make_tmap(CUtensorMap* map, const __nv_bfloat16* base, int64_t rows, int Kp, int box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    return T4R_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_error("operand planes must be 16-byte aligned");
    return T4R_ERR_INVALID;
  }
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(Kp), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(Kp) * 2};
  cuuint32_t box[2] = {64u, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) rows=%lld Kp=%d box_rows=%d", (int)r, (long long)rows, Kp,
              box_rows);
    return T4R_ERR_CUDA;
  }
  return 0;
}

// ----------------------------------------------------------------------------
// kernel
// ----------------------------------------------------------------------------
constexpr int BM = 128;
constexpr int A_PLANE_BYTES = BM * 128;  // 128 rows x 64 bf16

template <int BN>
struct GemmCfg {
  static constexpr int B_PLANE_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = 2 * A_PLANE_BYTES + 2 * B_PLANE_BYTES;
  static constexpr int STAGES = (BN == 256) ? 2 : ((BN == 128) ? 3 : 4);
  static constexpr int TMEM_COLS = 2 * BN;  // two accumulator stages (power of two)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 4096 /*LN exchange*/;
};

struct GemmDev {
  int M;
  int64_t N;
  int nkb;
  int nprod;
  const int32_t* m_dev;
  GemmEpilogue ep;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == T4R_ACT_RELU) return fmaxf(v, 0.f);
  if (act == T4R_ACT_GELU) return gelu_erf(v);
  return v;
}

// value of one 32-column chunk before LayerNorm: acc + bias -> act -> mask -> + residual
__device__ __forceinline__ void dense_chunk(float (&v)[32], const GemmEpilogue& ep, int64_t row, int64_t ncol0,
                                            int code) {
  if (ep.bias) {
    const float4* b4 = reinterpret_cast<const float4*>(ep.bias + ncol0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 b = __ldg(b4 + j);
      v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
    }
  }
  if (ep.act != T4R_ACT_NONE) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], ep.act);
  }
  if (code == 1) {
    const float4* m4 = reinterpret_cast<const float4*>(ep.mask_vec + ncol0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 b = __ldg(m4 + j);
      v[4 * j + 0] = b.x; v[4 * j + 1] = b.y; v[4 * j + 2] = b.z; v[4 * j + 3] = b.w;
    }
  } else if (code == 2) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
  }
  if (ep.residual) {
    const float4* r4 = reinterpret_cast<const float4*>(ep.residual + row * ep.ldr + ncol0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 b = r4[j];
      v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
    }
  }
}

__device__ __forceinline__ void store_f32_chunk(float* dst, const float (&v)[32], float scale) {
  float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    d4[j] = make_float4(v[4 * j + 0] * scale, v[4 * j + 1] * scale, v[4 * j + 2] * scale, v[4 * j + 3] * scale);
}

__device__ __forceinline__ void store_planes_chunk(__nv_bfloat16* hi_dst, __nv_bfloat16* lo_dst, const float (&v)[32]) {
  uint32_t h[16], l[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    __nv_bfloat16 h0, l0, h1, l1;
    split_bf16(v[2 * j], h0, l0);
    split_bf16(v[2 * j + 1], h1, l1);
    h[j] = pack_bf16x2(h0, h1);
    l[j] = pack_bf16x2(l0, l1);
  }
  uint4* hd = reinterpret_cast<uint4*>(hi_dst);
  uint4* ld = reinterpret_cast<uint4*>(lo_dst);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    hd[j] = make_uint4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
    ld[j] = make_uint4(l[4 * j], l[4 * j + 1], l[4 * j + 2], l[4 * j + 3]);
  }
}

// Each epilogue thread owns one output row (its TMEM lane) and COLS = BN/2 columns
// (warps 2-5 take the first half of the tile's columns, warps 6-9 the second half).
template <int BN, bool LN>
__device__ __forceinline__ void epilogue_dense(const GemmDev& p, uint32_t taddr, int64_t row, bool row_ok, int64_t n0,
                                               float2* xch_mine, const float2* xch_other) {
  constexpr int COLS = BN / 2;
  const GemmEpilogue& ep = p.ep;
  int code = 0;
  if (row_ok && ep.row_code) code = ep.row_code[row];
  if constexpr (LN) {
    // single pass: the thread's whole half-row lives in registers
    float v[COLS];
    tmem_ld<COLS>(taddr, v);
    float mean_h = 0.f, m2_h = 0.f;
    if (row_ok) {
#pragma unroll
      for (int c = 0; c < COLS / 32; ++c) {
        float (&vc)[32] = *reinterpret_cast<float (*)[32]>(&v[c * 32]);
        dense_chunk(vc, ep, row, n0 + c * 32, code);
      }
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < COLS; ++j) sum += v[j];
      mean_h = sum * (1.f / COLS);
#pragma unroll
      for (int j = 0; j < COLS; ++j) {
        const float d = v[j] - mean_h;
        m2_h = fmaf(d, d, m2_h);
      }
    }
    // combine the two halves of the row (Chan et al. pairwise update, equal counts)
    *xch_mine = make_float2(mean_h, m2_h);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float2 o = *xch_other;
    const float delta = o.x - mean_h;
    const float mean = 0.5f * (mean_h + o.x);
    const float m2 = m2_h + o.y + delta * delta * (0.5f * COLS);
    const float rstd = rsqrtf(m2 * (1.f / BN) + ep.ln_eps);
    if (row_ok) {
#pragma unroll
      for (int c = 0; c < COLS / 32; ++c) {
        float (&vc)[32] = *reinterpret_cast<float (*)[32]>(&v[c * 32]);
        const int64_t ncol0 = n0 + c * 32;
        if (ep.out_pre) store_f32_chunk(ep.out_pre + row * ep.ldp + ncol0, vc, 1.f);
        const float4* g4 = reinterpret_cast<const float4*>(ep.ln_gamma + ncol0);
        const float4* b4 = reinterpret_cast<const float4*>(ep.ln_beta + ncol0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 g = __ldg(g4 + j), b = __ldg(b4 + j);
          vc[4 * j + 0] = (vc[4 * j + 0] - mean) * rstd * g.x + b.x;
          vc[4 * j + 1] = (vc[4 * j + 1] - mean) * rstd * g.y + b.y;
          vc[4 * j + 2] = (vc[4 * j + 2] - mean) * rstd * g.z + b.z;
          vc[4 * j + 3] = (vc[4 * j + 3] - mean) * rstd * g.w + b.w;
        }
        if (ep.out_f32) store_f32_chunk(ep.out_f32 + row * ep.ldo + ncol0, vc, ep.out_scale);
        if (ep.out_planes) {
          __nv_bfloat16* hi = ep.out_planes + row * ep.ldpl + ncol0;
          store_planes_chunk(hi, hi + ep.plane_stride, vc);
        }
      }
    }
  } else {
#pragma unroll 1
    for (int c = 0; c < COLS / 32; ++c) {
      float v[32];
      tmem_ld<32>(taddr + c * 32, v);
      if (!row_ok) continue;
      const int64_t ncol0 = n0 + c * 32;
      if (ncol0 >= p.N) continue;
      dense_chunk(v, ep, row, ncol0, code);
      if (ep.out_f32) {
        float* dst = ep.out_f32 + row * ep.ldo + ncol0;
        if ((ep.ldo & 3) == 0 && ncol0 + 32 <= p.N) {
          store_f32_chunk(dst, v, ep.out_scale);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (ncol0 + j < p.N) dst[j] = v[j] * ep.out_scale;
        }
      }
      if (ep.out_planes) {
        __nv_bfloat16* hi = ep.out_planes + row * ep.ldpl + ncol0;
        store_planes_chunk(hi, hi + ep.plane_stride, v);
      }
    }
  }
}

// head epilogue: per row, online log-sum-exp (base 2) over this thread's COLS classes
// of the tile, optional logQ bias / accidental-hit removal (sampled softmax), rank count.
template <int BN>
__device__ __forceinline__ void epilogue_head(const GemmDev& p, uint32_t taddr, int64_t row, bool row_ok, int64_t n0,
                                              int part_idx) {
  constexpr int COLS = BN / 2;
  const GemmEpilogue& ep = p.ep;
  constexpr float kLog2e = 1.4426950408889634f;
  const float scale2 = ep.inv_tau * kLog2e;
  float m_run = -INFINITY, s_run = 0.f;
  int cnt = 0;
  int64_t label = -1;
  float tgt = 0.f;
  const bool want_rank = (ep.row_rank != nullptr);
  if (row_ok && ep.row_label) label = ep.row_label[row];
  if (row_ok && want_rank) tgt = ep.row_tgt[row];
  const bool full_tile = (n0 + COLS <= p.N);
#pragma unroll 1
  for (int c = 0; c < COLS / 32; ++c) {
    float v[32];
    tmem_ld<32>(taddr + c * 32, v);
    if (!row_ok) continue;
    const int64_t ncol0 = n0 + c * 32;
    if (ncol0 >= p.N) continue;
    if (ep.col_bias) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (full_tile || ncol0 + j < p.N) v[j] += __ldg(ep.col_bias + ncol0 + j);
    }
    if (ep.col_ids) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if ((full_tile || ncol0 + j < p.N) && __ldg(ep.col_ids + ncol0 + j) == label) v[j] = ep.hit_value;
    }
    if (!full_tile) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (ncol0 + j >= p.N) v[j] = -INFINITY;
    }
    if (want_rank) {
      // classes scoring above the label; ties resolved "lower id first" like a stable top-k
      const int64_t lab_col = label - ep.col_offset;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float x = v[j] * ep.inv_tau;
        const int64_t col = ncol0 + j;
        cnt += (col != lab_col) && ((x > tgt) || (x == tgt && col < lab_col));
      }
    }
    float cmax = v[0];
#pragma unroll
    for (int j = 1; j < 32; ++j) cmax = fmaxf(cmax, v[j]);
    cmax *= scale2;  // scale2 > 0
    const float m_new = fmaxf(m_run, cmax);
    if (m_new > -INFINITY) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) acc += exp2f(fmaf(v[j], scale2, -m_new));
      s_run = s_run * exp2f(m_run - m_new) + acc;
      m_run = m_new;
    }
  }
  if (row_ok) {
    ep.part_m[static_cast<int64_t>(part_idx) * ep.part_ld + row] = m_run;
    ep.part_s[static_cast<int64_t>(part_idx) * ep.part_ld + row] = s_run;
    if (want_rank && cnt) atomicAdd(ep.row_rank + row, cnt);
  }
}

template <int BN, bool LN, bool HEAD>
__global__ void __launch_bounds__(320, 1)
gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                   const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                   const GemmDev p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tfull_bar = empty_bar + Cfg::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float2* xch = reinterpret_cast<float2*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + 256);  // [2 parity][2 half][128]

  const int warp = warp_id();
  const int lane = lane_id();

  int M_eff = p.M;
  if (p.m_dev) M_eff = min(p.M, *p.m_dev);
  const int tiles_m = (M_eff + BM - 1) / BM;
  const int tiles_n = static_cast<int>((p.N + BN - 1) / BN);
  const int64_t num_tiles = static_cast<int64_t>(tiles_m) * tiles_n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmAh);
    tma_prefetch_desc(&tmAl);
    tma_prefetch_desc(&tmBh);
    tma_prefetch_desc(&tmBl);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t bytes = (p.nprod == 3) ? Cfg::STAGE_BYTES : (A_PLANE_BYTES + Cfg::B_PLANE_BYTES);
      for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = static_cast<int>(tile % tiles_m) * BM;
        const int n0 = static_cast<int>(tile / tiles_m) * BN;
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * Cfg::STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], bytes);
          tma_load_2d(st, &tmAh, &full_bar[stage], kb * 64, m0);
          tma_load_2d(st + 2 * A_PLANE_BYTES, &tmBh, &full_bar[stage], kb * 64, n0);
          if (p.nprod == 3) {
            tma_load_2d(st + A_PLANE_BYTES, &tmAl, &full_bar[stage], kb * 64, m0);
            tma_load_2d(st + 2 * A_PLANE_BYTES + Cfg::B_PLANE_BYTES, &tmBl, &full_bar[stage], kb * 64, n0);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[as], aph ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t a_hi = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t a_lo = a_hi + A_PLANE_BYTES;
          const uint32_t b_hi = a_hi + 2 * A_PLANE_BYTES;
          const uint32_t b_lo = b_hi + Cfg::B_PLANE_BYTES;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint64_t da_hi = umma_desc_sw128(a_hi + k4 * 32);
            const uint64_t db_hi = umma_desc_sw128(b_hi + k4 * 32);
            if (p.nprod == 3) {
              const uint64_t da_lo = umma_desc_sw128(a_lo + k4 * 32);
              const uint64_t db_lo = umma_desc_sw128(b_lo + k4 * 32);
              umma_bf16(d_tmem, da_lo, db_hi, idesc, (kb | k4) != 0);
              umma_bf16(d_tmem, da_hi, db_lo, idesc, 1u);
              umma_bf16(d_tmem, da_hi, db_hi, idesc, 1u);
            } else {
              umma_bf16(d_tmem, da_hi, db_hi, idesc, (kb | k4) != 0);
            }
          }
          umma_commit(&empty_bar[stage]);  // smem slot free once these MMAs retire
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[as]);  // accumulator ready for the epilogue
        as ^= 1;
        if (as == 0) aph ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps (2..9) =====================
    const int quad = warp & 3;          // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;   // which half of the tile's columns
    constexpr int COLS = BN / 2;
    int as = 0;
    uint32_t aph = 0;
    uint32_t tile_parity = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int tile_n = static_cast<int>(tile / tiles_m);
      const int64_t m0 = static_cast<int64_t>(tile % tiles_m) * BM;
      const int64_t n0 = static_cast<int64_t>(tile_n) * BN + half * COLS;
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                             static_cast<uint32_t>(as * BN + half * COLS);
      const int64_t row = m0 + quad * 32 + lane;
      const bool row_ok = row < M_eff;
      if (HEAD) {
        epilogue_head<BN>(p, taddr, row, row_ok, n0, tile_n * 2 + half);
      } else {
        float2* xm = xch + (tile_parity * 2 + half) * 128 + quad * 32 + lane;
        const float2* xo = xch + (tile_parity * 2 + (half ^ 1)) * 128 + quad * 32 + lane;
        epilogue_dense<BN, LN>(p, taddr, row, row_ok, n0, xm, xo);
      }
      tile_parity ^= 1;
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      as ^= 1;
      if (as == 0) aph ^= 1;
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ----------------------------------------------------------------------------
// host launcher
// ----------------------------------------------------------------------------
static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    g_num_sms = n;
  }
  return g_num_sms;
}

template <int BN, bool LN, bool HEAD>
static int launch_inst(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl,
                       const GemmDev& dp, int64_t max_tiles, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_bf16x3_kernel<BN, LN, HEAD>;
  static bool attr_set = false;
  if (!attr_set) {
    T4R_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  int grid = static_cast<int>(max_tiles < num_sms() ? max_tiles : num_sms());
  if (grid < 1) grid = 1;
  kern<<<grid, 320, Cfg::SMEM_BYTES, stream>>>(ah, al, bh, bl, dp);
  T4R_LAUNCH_CHECK("gemm_bf16x3_kernel");
  return 0;
}

int launch_gemm(const GemmProblem& pb, const GemmEpilogue& ep, cudaStream_t stream) {
  T4R_REQUIRE(pb.M > 0 && pb.N > 0 && pb.Kp > 0 && pb.Kp % 64 == 0, "gemm: bad shape M=%lld N=%lld Kp=%d",
              (long long)pb.M, (long long)pb.N, pb.Kp);
  T4R_REQUIRE(pb.M < (1ll << 31), "gemm: M too large");
  T4R_REQUIRE(pb.nprod == 1 || pb.nprod == 3, "gemm: nprod must be 1 or 3");
  const bool ln = ep.ln_gamma != nullptr;
  int bn = pb.bn;
  if (ln) {
    T4R_REQUIRE(pb.N == 64 || pb.N == 128 || pb.N == 256,
                "fused LayerNorm epilogue supports N in {64,128,256}, got %lld", (long long)pb.N);
    bn = static_cast<int>(pb.N);
  } else if (ep.head) {
    if (bn == 0) bn = 256;
  } else {
    if (bn == 0) bn = (pb.N % 256 == 0) ? 256 : ((pb.N % 128 == 0) ? 128 : 64);
    T4R_REQUIRE(ep.out_planes == nullptr || pb.N % 64 == 0, "gemm: planes output needs N %% 64 == 0");
    T4R_REQUIRE((ep.bias == nullptr && ep.residual == nullptr && ep.row_code == nullptr) || pb.N % 32 == 0,
                "gemm: dense epilogue needs N %% 32 == 0");
  }
  T4R_REQUIRE(bn == 64 || bn == 128 || bn == 256, "gemm: bad BN %d", bn);

  CUtensorMap ah, al, bh, bl;
  T4R_TRY(make_tmap(&ah, pb.a_planes, pb.M, pb.Kp, BM));
  T4R_TRY(make_tmap(&al, pb.a_planes + pb.a_rows * pb.Kp, pb.M, pb.Kp, BM));
  T4R_TRY(make_tmap(&bh, pb.b_planes, pb.N, pb.Kp, bn));
  T4R_TRY(make_tmap(&bl, pb.b_planes + pb.b_rows * pb.Kp, pb.N, pb.Kp, bn));

  GemmDev dp;
  dp.M = static_cast<int>(pb.M);
  dp.N = pb.N;
  dp.nkb = pb.Kp / 64;
  dp.nprod = pb.nprod;
  dp.m_dev = pb.m_dev;
  dp.ep = ep;
  const int64_t max_tiles = ((pb.M + BM - 1) / BM) * ((pb.N + bn - 1) / bn);

  if (ep.head) {
    if (bn == 256) return launch_inst<256, false, true>(ah, al, bh, bl, dp, max_tiles, stream);
    if (bn == 128) return launch_inst<128, false, true>(ah, al, bh, bl, dp, max_tiles, stream);
    return launch_inst<64, false, true>(ah, al, bh, bl, dp, max_tiles, stream);
  }
  if (ln) {
    if (bn == 256) return launch_inst<256, true, false>(ah, al, bh, bl, dp, max_tiles, stream);
    if (bn == 128) return launch_inst<128, true, false>(ah, al, bh, bl, dp, max_tiles, stream);
    return launch_inst<64, true, false>(ah, al, bh, bl, dp, max_tiles, stream);
  }
  if (bn == 256) return launch_inst<256, false, false>(ah, al, bh, bl, dp, max_tiles, stream);
  if (bn == 128) return launch_inst<128, false, false>(ah, al, bh, bl, dp, max_tiles, stream);
  return launch_inst<64, false, false>(ah, al, bh, bl, dp, max_tiles, stream);
}

}  // namespace t4r
