// t4r_attn_mma.cu -- K5 on the warp-level tensor path (mma.sync m16n8k16 bf16, fp32 accumulate).
//
// One warp owns one (session, head).  L <= 32 keys/queries are padded to a 32 x 32 problem:
//   S1 = Qaug K^T            (32 x 32, k = dh)       Qaug = [q_0..q_{L-1}; r_w_bias; r_r_bias; 0...]
//   S2 = Qaug R^T            (32 x 2L, k = dh)       XLNet only
//   s[i,j] = (S1[i,j] + S1[L,j] + S2[i,j+L-i] + S2[L+1,j+L-i]) / sqrt(dh)
//            = ((q_i + r_w_bias).k_j + (q_i + r_r_bias).R[j+L-i]) / sqrt(dh)      HF:xlnet:95-140, rel_shift :81-93
//   P = softmax_j(s)  (no mask for XLNet 'bi'; j <= i for GPT-2, HF:gpt2:54-72)
//   O = P V           (32 x dh, k = 32)
// Appending the two bias vectors as extra query rows turns the bias terms into two more rows of
// the same products, so every operand is a raw split-bf16 plane written by the QKV GEMM epilogue /
// the positional projection.  All products are issued three times (hi*hi + hi*lo + lo*hi) like
// the tcgen05 GEMMs, so the result stays fp32-grade.  The S1 accumulator fragments are reused in
// place as the A operand of P V (C-fragment layout == A-fragment layout of two 8-wide tiles);
// only S2 takes a shared-memory round trip (the relative shift moves data across lanes).
#include <math.h>
#include <stdlib.h>

#include "t4r_common.cuh"
#include "t4r_internal.h"

namespace t4r {

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x2(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// three split products: c += a_lo b_hi + a_hi b_lo + a_hi b_hi
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                     const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
  mma_bf16(c, al, bh);
  mma_bf16(c, ah, bl);
  mma_bf16(c, ah, bh);
}
__device__ __forceinline__ void split_pair(float x, float y, uint32_t& hi, uint32_t& lo) {
  split_bf16x2(x, y, hi, lo);
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// Shared-memory tiles hold only the L (+2) real rows; every ldmatrix row address beyond them
// points at one shared all-zero row, so the per-warp footprint stays small enough for 3-4
// blocks per SM (the kernel is latency-bound: one cp.async batch + ~2k instructions per session).
// MASKED (XLNet permutation language modeling, HF:xlnet two-stream attention): blockIdx.z is the stream -- 0 = content
// stream h, 1 = query stream g.  The planes then hold 2 B L rows (h rows, then g rows); queries come from the stream's
// own rows, keys / values always from the h rows, and score (i, j) of session b is replaced by -1e30 where
// plm_mask[b, i, j] != 0 (for h: except on the diagonal -- HF's non_tgt_mask).  -1e30, not -inf: a fully masked row
// then softmaxes to the uniform distribution exactly as HF's fp32 `score - 1e30 * mask` does.
template <int DH, bool REL, bool MASKED = false>
__global__ void __launch_bounds__(128)
attn_mma_kernel(const __nv_bfloat16* __restrict__ qkv, int64_t qkv_plane_stride, const __nv_bfloat16* __restrict__ rpl,
                int64_t r_plane_stride, const float* __restrict__ rw, const float* __restrict__ rr, int B, int L, int d,
                int sessions_per_block, __nv_bfloat16* __restrict__ out_planes, int64_t out_plane_stride,
                const uint8_t* __restrict__ plm_mask = nullptr) {
  const int stream = MASKED ? static_cast<int>(blockIdx.z) : 0;
  const int64_t stream_rows = MASKED ? static_cast<int64_t>(stream) * B * L : 0;
  constexpr int LDS = DH + 8;            // row stride (bf16): 16-byte rows, conflict-free ldmatrix
  constexpr int KT = DH / 16;            // k tiles of the q.k / q.R products
  constexpr int NTC = DH / 8;            // n tiles of the P V product
  constexpr int C8 = DH / 8;             // 16-byte chunks per row
  extern __shared__ __align__(16) uint8_t smem_a[];
  const int h = blockIdx.y;
  const int warp = warp_id(), lane = lane_id();
  const int nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int QR = REL ? L + 2 : L;                    // query rows incl. the two bias rows
  const int RR = 2 * L;                              // rows of R
  const int S2LD = ((2 * L + 7) / 8) * 8 + 1;        // fp32 stride of the shifted-term scratch
  // block-shared: zero row, R planes; per warp: Q, K, V planes and the S2 scratch
  __nv_bfloat16* zrow = reinterpret_cast<__nv_bfloat16*>(smem_a);                  // [LDS]
  __nv_bfloat16* Rs = zrow + LDS;                                                   // [2][RR][LDS]
  const int warp_elems = 2 * QR * LDS + 4 * L * LDS;
  const int s2_floats = REL ? QR * S2LD : 0;
  uint8_t* wbase = smem_a + (LDS + (REL ? 2 * RR * LDS : 0)) * 2;
  wbase += static_cast<size_t>(warp) * (((warp_elems * 2 + s2_floats * 4) + 15) / 16 * 16);
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(wbase);   // [2][QR][LDS]
  __nv_bfloat16* Ks = Qs + 2 * QR * LDS;                         // [2][L][LDS]
  __nv_bfloat16* Vs = Ks + 2 * L * LDS;                          // [2][L][LDS]
  float* S2s = reinterpret_cast<float*>(Vs + 2 * L * LDS);       // [QR][S2LD]
  const uint32_t zaddr = smem_u32(zrow);

  for (int c = threadIdx.x; c < LDS; c += blockDim.x) zrow[c] = __float2bfloat16_rn(0.f);
  if (REL) {
    for (int idx = threadIdx.x; idx < 2 * RR * C8; idx += blockDim.x) {
      const int pl = idx / (RR * C8), rem = idx % (RR * C8);
      const int m = rem / C8, c8 = rem % C8;
      *reinterpret_cast<uint4*>(Rs + (pl * RR + m) * LDS + 8 * c8) =
          __ldg(reinterpret_cast<const uint4*>(rpl + pl * r_plane_stride + static_cast<int64_t>(m) * d + h * DH + 8 * c8));
    }
  }
  __syncthreads();
  const float scale = rsqrtf(static_cast<float>(DH));
  const int b_begin = blockIdx.x * sessions_per_block;
  const int b_end = min(B, b_begin + sessions_per_block);

  for (int b = b_begin + warp; b < b_end; b += nwarps) {
    __syncwarp();
    // ---- stage q, k, v planes with one batch of 16-byte cp.async per lane
    {
      const __nv_bfloat16* base = qkv + static_cast<int64_t>(b) * L * 3 * d + h * DH;
      for (int idx = lane; idx < 2 * L * C8; idx += 32) {
        const int pl = idx / (L * C8), rem = idx % (L * C8);
        const int i = rem / C8, c8 = rem % C8;
        const __nv_bfloat16* rowp = base + pl * qkv_plane_stride + static_cast<int64_t>(i) * 3 * d + 8 * c8;
        cp_async16(smem_u32(Qs + (pl * QR + i) * LDS + 8 * c8), rowp + stream_rows * 3 * d);  // queries: this stream's rows
        cp_async16(smem_u32(Ks + (pl * L + i) * LDS + 8 * c8), rowp + d);
        cp_async16(smem_u32(Vs + (pl * L + i) * LDS + 8 * c8), rowp + 2 * d);
      }
      if (REL) {
        // query rows L and L+1 carry r_w_bias / r_r_bias (split on the fly)
        for (int c = 2 * lane; c < DH; c += 64) {   // pairs: cvt.rn.bf16x2 (the scalar conversion runs on the XU pipe)
          uint32_t hi, lo;
          const float2 w2 = __ldg(reinterpret_cast<const float2*>(rw + h * DH + c));
          split_bf16x2(w2.x, w2.y, hi, lo);
          *reinterpret_cast<uint32_t*>(Qs + (0 * QR + L) * LDS + c) = hi;
          *reinterpret_cast<uint32_t*>(Qs + (1 * QR + L) * LDS + c) = lo;
          const float2 r2 = __ldg(reinterpret_cast<const float2*>(rr + h * DH + c));
          split_bf16x2(r2.x, r2.y, hi, lo);
          *reinterpret_cast<uint32_t*>(Qs + (0 * QR + L + 1) * LDS + c) = hi;
          *reinterpret_cast<uint32_t*>(Qs + (1 * QR + L + 1) * LDS + c) = lo;
        }
      }
      cp_async_wait_all();
    }
    __syncwarp();

    // ---- A fragments of Qaug
    uint32_t aq[2][2][KT][4];  // [plane][m tile][k tile]
#pragma unroll
    for (int pl = 0; pl < 2; ++pl)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int kt = 0; kt < KT; ++kt)
        {
          const int row = mt * 16 + (lane & 15);
          const int col = kt * 16 + (lane >> 4) * 8;
          ldsm_x4(aq[pl][mt][kt], row < QR ? smem_u32(Qs + (pl * QR + row) * LDS + col) : zaddr);
        }

    // ---- S1 = Qaug K^T
    float s1[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) s1[mt][nt][e] = 0.f;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
        uint32_t bh[2], bl[2];
        const int row = nt * 8 + (lane & 7);
        const int col = kt * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x2(bh, row < L ? smem_u32(Ks + row * LDS + col) : zaddr);
        ldsm_x2(bl, row < L ? smem_u32(Ks + (L + row) * LDS + col) : zaddr);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) mma3(s1[mt][nt], aq[0][mt][kt], aq[1][mt][kt], bh, bl);
      }

    if (REL) {
      // ---- S2 = Qaug R^T -> scratch
      const int nt2_max = (2 * L + 7) / 8;
#pragma unroll 1
      for (int nt2 = 0; nt2 < nt2_max; ++nt2) {
        float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
          uint32_t bh[2], bl[2];
          const int row = nt2 * 8 + (lane & 7);
          const int col = kt * 16 + ((lane >> 3) & 1) * 8;
          ldsm_x2(bh, row < RR ? smem_u32(Rs + row * LDS + col) : zaddr);
          ldsm_x2(bl, row < RR ? smem_u32(Rs + (RR + row) * LDS + col) : zaddr);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) mma3(acc[mt], aq[0][mt][kt], aq[1][mt][kt], bh, bl);
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int col = nt2 * 8 + 2 * t;
          const int r0 = mt * 16 + g, r1 = r0 + 8;
          if (r0 < QR) { S2s[r0 * S2LD + col] = acc[mt][0]; S2s[r0 * S2LD + col + 1] = acc[mt][1]; }
          if (r1 < QR) { S2s[r1 * S2LD + col] = acc[mt][2]; S2s[r1 * S2LD + col + 1] = acc[mt][3]; }
        }
      }
      __syncwarp();
    }

    // ---- scores -> probabilities (in place in s1)
    const int mtL = L >> 4, rL = L & 15;             // where query row L (r_w_bias) sits in the fragments
    const int srcL = ((rL & 7) << 2) | t;
    float rmax[2][2], rsum[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) rmax[mt][hf] = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float wb0 = 0.f, wb1 = 0.f;
      if (REL) {
        // (r_w_bias . k_j) for this lane's two columns: row L of S1, held by lane 4*(rL%8)+t
        const float v0 = (mtL == 0) ? ((rL < 8) ? s1[0][nt][0] : s1[0][nt][2]) : ((rL < 8) ? s1[1][nt][0] : s1[1][nt][2]);
        const float v1 = (mtL == 0) ? ((rL < 8) ? s1[0][nt][1] : s1[0][nt][3]) : ((rL < 8) ? s1[1][nt][1] : s1[1][nt][3]);
        wb0 = __shfl_sync(0xffffffffu, v0, srcL);
        wb1 = __shfl_sync(0xffffffffu, v1, srcL);
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = mt * 16 + g + ((e >> 1) << 3);
          const int j = nt * 8 + 2 * t + (e & 1);
          float v = s1[mt][nt][e];
          if (REL) {
            const int ii = i < L ? i : 0;
            const int jj = j < L ? j : 0;
            const int m = jj + L - ii;
            v += ((e & 1) ? wb1 : wb0) + S2s[ii * S2LD + m] + S2s[(L + 1) * S2LD + m];
          }
          v *= scale;
          if (MASKED) {
            if (i < L && j < L && !(stream == 0 && i == j) && plm_mask[(static_cast<int64_t>(b) * L + i) * L + j]) v = -1e30f;
          }
          if (j >= L || (!REL && j > i)) v = -INFINITY;
          s1[mt][nt][e] = v;
          rmax[mt][e >> 1] = fmaxf(rmax[mt][e >> 1], v);
        }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float m = rmax[mt][hf];
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
        rmax[mt][hf] = m;
        rsum[mt][hf] = 0.f;
      }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float v = s1[mt][nt][e];
          // 2^((v - max) log2 e) on the SFU (ex2.approx, 2 ulp): expf's range reduction was 20 % of this kernel's samples
          const float p = (v == -INFINITY) ? 0.f : fast_exp2((v - rmax[mt][e >> 1]) * 1.4426950408889634f);
          s1[mt][nt][e] = p;
          rsum[mt][e >> 1] += p;
        }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float sm = rsum[mt][hf];
        sm += __shfl_xor_sync(0xffffffffu, sm, 1);
        sm += __shfl_xor_sync(0xffffffffu, sm, 2);
        rsum[mt][hf] = (sm > 0.f) ? 1.f / sm : 0.f;
      }

    // ---- O = P V : S1's C fragments are the A fragments of P (two 8-wide tiles per k tile)
    float o[2][NTC][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nc = 0; nc < NTC; ++nc)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[mt][nc][e] = 0.f;
#pragma unroll
    for (int kt = 0; kt < 2; ++kt) {
      uint32_t ph[2][4], plo[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        split_pair(s1[mt][2 * kt][0], s1[mt][2 * kt][1], ph[mt][0], plo[mt][0]);
        split_pair(s1[mt][2 * kt][2], s1[mt][2 * kt][3], ph[mt][1], plo[mt][1]);
        split_pair(s1[mt][2 * kt + 1][0], s1[mt][2 * kt + 1][1], ph[mt][2], plo[mt][2]);
        split_pair(s1[mt][2 * kt + 1][2], s1[mt][2 * kt + 1][3], ph[mt][3], plo[mt][3]);
      }
#pragma unroll
      for (int nc = 0; nc < NTC; ++nc) {
        uint32_t bh[2], bl[2];
        const int row = kt * 16 + (lane & 15);
        ldsm_x2_trans(bh, row < L ? smem_u32(Vs + row * LDS + nc * 8) : zaddr);
        ldsm_x2_trans(bl, row < L ? smem_u32(Vs + (L + row) * LDS + nc * 8) : zaddr);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) mma3(o[mt][nc], ph[mt], plo[mt], bh, bl);
      }
    }

    // ---- normalise, split, store (rows < L)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int i = mt * 16 + g + hf * 8;
        if (i < L) {
          const float inv = rsum[mt][hf];
          __nv_bfloat16* hi = out_planes + (stream_rows + static_cast<int64_t>(b) * L + i) * d + h * DH + 2 * t;
#pragma unroll
          for (int nc = 0; nc < NTC; ++nc) {
            uint32_t wh, wl;
            split_pair(o[mt][nc][2 * hf] * inv, o[mt][nc][2 * hf + 1] * inv, wh, wl);
            *reinterpret_cast<uint32_t*>(hi + nc * 8) = wh;
            *reinterpret_cast<uint32_t*>(hi + out_plane_stride + nc * 8) = wl;
          }
        }
      }
  }
}

template <int DH, bool REL, bool MASKED = false>
static int launch_attn_mma_inst(const __nv_bfloat16* qkv, int64_t qkv_ps, const __nv_bfloat16* rpl, int64_t r_ps,
                                const float* rw, const float* rr, int B, int L, int d, int H,
                                __nv_bfloat16* out_planes, int64_t out_ps, cudaStream_t s,
                                const uint8_t* plm_mask = nullptr) {
  constexpr int LDS = DH + 8;
  const int QR = REL ? L + 2 : L;
  const int S2LD = ((2 * L + 7) / 8) * 8 + 1;
  const size_t per_warp = ((static_cast<size_t>(2 * QR * LDS + 4 * L * LDS) * 2 + (REL ? QR * S2LD * 4 : 0)) + 15) / 16 * 16;
  const size_t shared_part = static_cast<size_t>(LDS + (REL ? 2 * 2 * L * LDS : 0)) * 2;
  const int warps = 4;
  const size_t smem = shared_part + warps * per_warp;
  T4R_REQUIRE(smem <= 200 * 1024, "attn_mma: shared memory %zu too large", smem);
  auto kern = attn_mma_kernel<DH, REL, MASKED>;
  static size_t attr = 0;
  if (smem > attr) {
    T4R_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr = smem;
  }
  int spb = 2 * warps;
  while (spb > warps && static_cast<int64_t>((B + spb - 1) / spb) * H < 148 * 8) spb >>= 1;
  dim3 grid((B + spb - 1) / spb, H, MASKED ? 2 : 1);
  kern<<<grid, warps * 32, smem, s>>>(qkv, qkv_ps, rpl, r_ps, rw, rr, B, L, d, spb, out_planes, out_ps, plm_mask);
  T4R_LAUNCH_CHECK("attn_mma_kernel");
  return 0;
}


// ============================================================================
// 32 < L <= 64 (BASELINE configs[4]: L = 50): the same scheme with TWO warps per (session, head).  The 64 x 64 score
// problem does not fit one warp's registers (S1 alone would be 128 accumulators), so the query rows are split:
// warp w of a pair owns m tiles 2w, 2w+1 (rows 32w .. 32w+31 of Qaug); K, V, R and the staged Q are shared by the
// pair through shared memory.  What crossed lanes by shuffle in the one-warp kernel crosses warps through shared
// memory here: the (r_w_bias . k_j) row of S1 goes to WB[64], the (r_r_bias . R) row of S2 is read from the pair's S2
// scratch like every other row; named barriers (one id per pair, 64 threads) order the stages.  A block holds one or
// two pairs (shared-memory footprint decides) which share the head's R planes.  Validated on a B200 in round 2
// (12 / 12 parity cases; XLNet layer at L = 50: 1.60 -> 1.05 ms) and the default since; T4R_ATTN_MMA64=0 keeps
// 32 < L <= 64 on the FFMA kernel (attn_kernel in t4r_kernels.cu).
// ============================================================================
__device__ __forceinline__ void pair_bar(int pair) { asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory"); }

template <int DH, bool REL, bool MASKED = false>
__global__ void __launch_bounds__(128)
attn_mma64_kernel(const __nv_bfloat16* __restrict__ qkv, int64_t qkv_plane_stride, const __nv_bfloat16* __restrict__ rpl,
                  int64_t r_plane_stride, const float* __restrict__ rw, const float* __restrict__ rr, int B, int L, int d,
                  int sessions_per_block, __nv_bfloat16* __restrict__ out_planes, int64_t out_plane_stride,
                  const uint8_t* __restrict__ plm_mask = nullptr) {
  const int stream = MASKED ? static_cast<int>(blockIdx.z) : 0;
  const int64_t stream_rows = MASKED ? static_cast<int64_t>(stream) * B * L : 0;
  constexpr int LDS = DH + 8;
  constexpr int KT = DH / 16;
  constexpr int NTC = DH / 8;
  constexpr int C8 = DH / 8;
  constexpr int NT = 8;                              // key tiles: 64 keys
  extern __shared__ __align__(16) uint8_t smem_a[];
  const int h = blockIdx.y;
  const int warp = warp_id(), lane = lane_id();
  const int npairs = blockDim.x >> 6;
  const int pair = warp >> 1, wp = warp & 1;        // warp pair, warp within the pair
  const int pt = threadIdx.x & 63;                   // thread within the pair
  const int g = lane >> 2, t = lane & 3;
  const int QR = REL ? L + 2 : L;
  const int RR = 2 * L;
  const int S2LD = ((2 * L + 7) / 8) * 8 + 1;
  __nv_bfloat16* zrow = reinterpret_cast<__nv_bfloat16*>(smem_a);                  // [LDS]
  __nv_bfloat16* Rs = zrow + LDS;                                                   // [2][RR][LDS]
  const int pair_elems = 2 * QR * LDS + 4 * L * LDS;
  const int pair_floats = REL ? QR * S2LD + 64 : 0;
  uint8_t* pbase = smem_a + (LDS + (REL ? 2 * RR * LDS : 0)) * 2;
  pbase = smem_a + (((pbase - smem_a) + 15) / 16 * 16);
  pbase += static_cast<size_t>(pair) * (((pair_elems * 2 + pair_floats * 4) + 15) / 16 * 16);
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(pbase);   // [2][QR][LDS]
  __nv_bfloat16* Ks = Qs + 2 * QR * LDS;                         // [2][L][LDS]
  __nv_bfloat16* Vs = Ks + 2 * L * LDS;                          // [2][L][LDS]
  float* S2s = reinterpret_cast<float*>(Vs + 2 * L * LDS);       // [QR][S2LD]
  float* WB = S2s + (REL ? QR * S2LD : 0);                       // [64]  (r_w_bias . k_j)
  const uint32_t zaddr = smem_u32(zrow);

  for (int c = threadIdx.x; c < LDS; c += blockDim.x) zrow[c] = __float2bfloat16_rn(0.f);
  if (REL) {
    for (int idx = threadIdx.x; idx < 2 * RR * C8; idx += blockDim.x) {
      const int pl = idx / (RR * C8), rem = idx % (RR * C8);
      const int m = rem / C8, c8 = rem % C8;
      *reinterpret_cast<uint4*>(Rs + (pl * RR + m) * LDS + 8 * c8) =
          __ldg(reinterpret_cast<const uint4*>(rpl + pl * r_plane_stride + static_cast<int64_t>(m) * d + h * DH + 8 * c8));
    }
  }
  __syncthreads();
  const float scale = rsqrtf(static_cast<float>(DH));
  const int b_begin = blockIdx.x * sessions_per_block;
  const int b_end = min(B, b_begin + sessions_per_block);
  const int mt0 = 2 * wp;                            // first of this warp's two m tiles

  for (int b = b_begin + pair; b < b_end; b += npairs) {   // same trip count for both warps of a pair
    pair_bar(pair);                                          // the previous session's tiles are no longer read
    {
      const __nv_bfloat16* base = qkv + static_cast<int64_t>(b) * L * 3 * d + h * DH;
      for (int idx = pt; idx < 2 * L * C8; idx += 64) {
        const int pl = idx / (L * C8), rem = idx % (L * C8);
        const int i = rem / C8, c8 = rem % C8;
        const __nv_bfloat16* rowp = base + pl * qkv_plane_stride + static_cast<int64_t>(i) * 3 * d + 8 * c8;
        cp_async16(smem_u32(Qs + (pl * QR + i) * LDS + 8 * c8), rowp + stream_rows * 3 * d);
        cp_async16(smem_u32(Ks + (pl * L + i) * LDS + 8 * c8), rowp + d);
        cp_async16(smem_u32(Vs + (pl * L + i) * LDS + 8 * c8), rowp + 2 * d);
      }
      if (REL) {
        for (int c = pt; c < DH; c += 64) {
          __nv_bfloat16 hi, lo;
          split_bf16(__ldg(rw + h * DH + c), hi, lo);
          Qs[(0 * QR + L) * LDS + c] = hi;
          Qs[(1 * QR + L) * LDS + c] = lo;
          split_bf16(__ldg(rr + h * DH + c), hi, lo);
          Qs[(0 * QR + L + 1) * LDS + c] = hi;
          Qs[(1 * QR + L + 1) * LDS + c] = lo;
        }
      }
      cp_async_wait_all();
    }
    pair_bar(pair);

    // ---- A fragments of this warp's 32 rows of Qaug
    uint32_t aq[2][2][KT][4];  // [plane][local m tile][k tile]
#pragma unroll
    for (int pl = 0; pl < 2; ++pl)
#pragma unroll
      for (int ml = 0; ml < 2; ++ml)
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
          const int row = (mt0 + ml) * 16 + (lane & 15);
          const int col = kt * 16 + (lane >> 4) * 8;
          ldsm_x4(aq[pl][ml][kt], row < QR ? smem_u32(Qs + (pl * QR + row) * LDS + col) : zaddr);
        }

    // ---- S1 = Qaug K^T (this warp's rows x 64 keys)
    float s1[2][NT][4];
#pragma unroll
    for (int ml = 0; ml < 2; ++ml)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) s1[ml][nt][e] = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
        uint32_t bh[2], bl[2];
        const int row = nt * 8 + (lane & 7);
        const int col = kt * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x2(bh, row < L ? smem_u32(Ks + row * LDS + col) : zaddr);
        ldsm_x2(bl, row < L ? smem_u32(Ks + (L + row) * LDS + col) : zaddr);
#pragma unroll
        for (int ml = 0; ml < 2; ++ml) mma3(s1[ml][nt], aq[0][ml][kt], aq[1][ml][kt], bh, bl);
      }

    if (REL) {
      // ---- S2 = Qaug R^T for this warp's rows -> the pair's scratch
      const int nt2_max = (2 * L + 7) / 8;
#pragma unroll 1
      for (int nt2 = 0; nt2 < nt2_max; ++nt2) {
        float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
          uint32_t bh[2], bl[2];
          const int row = nt2 * 8 + (lane & 7);
          const int col = kt * 16 + ((lane >> 3) & 1) * 8;
          ldsm_x2(bh, row < RR ? smem_u32(Rs + row * LDS + col) : zaddr);
          ldsm_x2(bl, row < RR ? smem_u32(Rs + (RR + row) * LDS + col) : zaddr);
#pragma unroll
          for (int ml = 0; ml < 2; ++ml) mma3(acc[ml], aq[0][ml][kt], aq[1][ml][kt], bh, bl);
        }
#pragma unroll
        for (int ml = 0; ml < 2; ++ml) {
          const int col = nt2 * 8 + 2 * t;
          const int r0 = (mt0 + ml) * 16 + g, r1 = r0 + 8;
          if (r0 < QR) { S2s[r0 * S2LD + col] = acc[ml][0]; S2s[r0 * S2LD + col + 1] = acc[ml][1]; }
          if (r1 < QR) { S2s[r1 * S2LD + col] = acc[ml][2]; S2s[r1 * S2LD + col + 1] = acc[ml][3]; }
        }
      }
      // ---- row L of S1 (r_w_bias . k_j) -> WB, written by the warp and lanes that hold it
      const int mtL = L >> 4, rL = L & 15;
      if ((mtL >> 1) == wp && g == (rL & 7)) {
        const int mlL = mtL & 1;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const float v0 = (mlL == 0) ? ((rL < 8) ? s1[0][nt][0] : s1[0][nt][2]) : ((rL < 8) ? s1[1][nt][0] : s1[1][nt][2]);
          const float v1 = (mlL == 0) ? ((rL < 8) ? s1[0][nt][1] : s1[0][nt][3]) : ((rL < 8) ? s1[1][nt][1] : s1[1][nt][3]);
          WB[nt * 8 + 2 * t] = v0;
          WB[nt * 8 + 2 * t + 1] = v1;
        }
      }
      pair_bar(pair);   // both warps' S2 rows (incl. row L + 1) and WB are visible
    }

    // ---- scores -> probabilities (in place in s1)
    float rmax[2][2], rsum[2][2];
#pragma unroll
    for (int ml = 0; ml < 2; ++ml)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) rmax[ml][hf] = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int ml = 0; ml < 2; ++ml)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = (mt0 + ml) * 16 + g + ((e >> 1) << 3);
          const int j = nt * 8 + 2 * t + (e & 1);
          float v = s1[ml][nt][e];
          if (REL) {
            const int ii = i < L ? i : 0;
            const int jj = j < L ? j : 0;
            const int m = jj + L - ii;
            v += WB[j] + S2s[ii * S2LD + m] + S2s[(L + 1) * S2LD + m];
          }
          v *= scale;
          if (MASKED) {
            if (i < L && j < L && !(stream == 0 && i == j) && plm_mask[(static_cast<int64_t>(b) * L + i) * L + j]) v = -1e30f;
          }
          if (j >= L || (!REL && j > i)) v = -INFINITY;
          s1[ml][nt][e] = v;
          rmax[ml][e >> 1] = fmaxf(rmax[ml][e >> 1], v);
        }
    }
#pragma unroll
    for (int ml = 0; ml < 2; ++ml)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float m = rmax[ml][hf];
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
        rmax[ml][hf] = m;
        rsum[ml][hf] = 0.f;
      }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int ml = 0; ml < 2; ++ml)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float v = s1[ml][nt][e];
          const float p = (v == -INFINITY) ? 0.f : fast_exp2((v - rmax[ml][e >> 1]) * 1.4426950408889634f);
          s1[ml][nt][e] = p;
          rsum[ml][e >> 1] += p;
        }
#pragma unroll
    for (int ml = 0; ml < 2; ++ml)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float sm = rsum[ml][hf];
        sm += __shfl_xor_sync(0xffffffffu, sm, 1);
        sm += __shfl_xor_sync(0xffffffffu, sm, 2);
        rsum[ml][hf] = (sm > 0.f) ? 1.f / sm : 0.f;
      }

    // ---- O = P V (k = 64 keys = 4 k tiles)
    float o[2][NTC][4];
#pragma unroll
    for (int ml = 0; ml < 2; ++ml)
#pragma unroll
      for (int nc = 0; nc < NTC; ++nc)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[ml][nc][e] = 0.f;
#pragma unroll
    for (int kt = 0; kt < NT / 2; ++kt) {
      uint32_t ph[2][4], plo[2][4];
#pragma unroll
      for (int ml = 0; ml < 2; ++ml) {
        split_pair(s1[ml][2 * kt][0], s1[ml][2 * kt][1], ph[ml][0], plo[ml][0]);
        split_pair(s1[ml][2 * kt][2], s1[ml][2 * kt][3], ph[ml][1], plo[ml][1]);
        split_pair(s1[ml][2 * kt + 1][0], s1[ml][2 * kt + 1][1], ph[ml][2], plo[ml][2]);
        split_pair(s1[ml][2 * kt + 1][2], s1[ml][2 * kt + 1][3], ph[ml][3], plo[ml][3]);
      }
#pragma unroll
      for (int nc = 0; nc < NTC; ++nc) {
        uint32_t bh[2], bl[2];
        const int row = kt * 16 + (lane & 15);
        ldsm_x2_trans(bh, row < L ? smem_u32(Vs + row * LDS + nc * 8) : zaddr);
        ldsm_x2_trans(bl, row < L ? smem_u32(Vs + (L + row) * LDS + nc * 8) : zaddr);
#pragma unroll
        for (int ml = 0; ml < 2; ++ml) mma3(o[ml][nc], ph[ml], plo[ml], bh, bl);
      }
    }

    // ---- normalise, split, store (rows < L)
#pragma unroll
    for (int ml = 0; ml < 2; ++ml)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int i = (mt0 + ml) * 16 + g + hf * 8;
        if (i < L) {
          const float inv = rsum[ml][hf];
          __nv_bfloat16* hi = out_planes + (stream_rows + static_cast<int64_t>(b) * L + i) * d + h * DH + 2 * t;
#pragma unroll
          for (int nc = 0; nc < NTC; ++nc) {
            uint32_t wh, wl;
            split_pair(o[ml][nc][2 * hf] * inv, o[ml][nc][2 * hf + 1] * inv, wh, wl);
            *reinterpret_cast<uint32_t*>(hi + nc * 8) = wh;
            *reinterpret_cast<uint32_t*>(hi + out_plane_stride + nc * 8) = wl;
          }
        }
      }
  }
}

template <int DH, bool REL, bool MASKED = false>
static int launch_attn_mma64_inst(const __nv_bfloat16* qkv, int64_t qkv_ps, const __nv_bfloat16* rpl, int64_t r_ps,
                                  const float* rw, const float* rr, int B, int L, int d, int H,
                                  __nv_bfloat16* out_planes, int64_t out_ps, cudaStream_t s,
                                  const uint8_t* plm_mask = nullptr) {
  constexpr int LDS = DH + 8;
  const int QR = REL ? L + 2 : L;
  const int S2LD = ((2 * L + 7) / 8) * 8 + 1;
  const size_t per_pair = ((static_cast<size_t>(2 * QR * LDS + 4 * L * LDS) * 2 + (REL ? (QR * S2LD + 64) * 4 : 0)) + 15) / 16 * 16;
  const size_t shared_part = (static_cast<size_t>(LDS + (REL ? 2 * 2 * L * LDS : 0)) * 2 + 15) / 16 * 16;
  int pairs = 2;
  if (shared_part + pairs * per_pair > 200 * 1024) pairs = 1;
  const size_t smem = shared_part + pairs * per_pair;
  T4R_REQUIRE(smem <= 200 * 1024, "attn_mma64: shared memory %zu too large", smem);
  auto kern = attn_mma64_kernel<DH, REL, MASKED>;
  static size_t attr = 0;
  if (smem > attr) {
    T4R_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr = smem;
  }
  int spb = 2 * pairs;
  while (spb > pairs && static_cast<int64_t>((B + spb - 1) / spb) * H < 148 * 8) spb >>= 1;
  dim3 grid((B + spb - 1) / spb, H, MASKED ? 2 : 1);
  kern<<<grid, pairs * 64, smem, s>>>(qkv, qkv_ps, rpl, r_ps, rw, rr, B, L, d, spb, out_planes, out_ps, plm_mask);
  T4R_LAUNCH_CHECK("attn_mma64_kernel");
  return 0;
}

// one warp per (session, head) up to 32 (augmented) query rows; two warps per (session, head) up to 64 (the default
// since its first B200 run in round 2; T4R_ATTN_MMA64=0 keeps those lengths on the FFMA kernel)
#ifndef T4R_ATTN_MMA64_DEFAULT
#define T4R_ATTN_MMA64_DEFAULT 1  // two-warp tensor-path attention for 32 < L <= 64 (validated on B200 in round 2; T4R_ATTN_MMA64=0 selects the FFMA kernel)
#endif
static bool attn_mma64_enabled() {
  const char* e = getenv("T4R_ATTN_MMA64");
  return e ? atoi(e) != 0 : (T4R_ATTN_MMA64_DEFAULT != 0);
}
bool attn_mma_supported(int L, int d, int H, bool rel) {
  if (d % H) return false;
  const int dh = d / H;
  if (dh != 16 && dh != 32 && dh != 64) return false;
  const int rows = rel ? L + 2 : L;
  return rows <= 32 || (rows <= 64 && attn_mma64_enabled());
}

int launch_attn_mma(bool rel, const __nv_bfloat16* qkv_planes, int64_t qkv_plane_stride, const __nv_bfloat16* r_planes,
                    int64_t r_plane_stride, const float* rw, const float* rr, int B, int L, int d, int H,
                    __nv_bfloat16* out_planes, int64_t out_plane_stride, cudaStream_t s) {
  T4R_REQUIRE(attn_mma_supported(L, d, H, rel), "attn_mma: unsupported shape L=%d d=%d H=%d", L, d, H);
  const int dh = d / H;
  if ((rel ? L + 2 : L) > 32) {
#define T4R_AM64(DHV)                                                                                                \
  if (dh == DHV) {                                                                                                   \
    if (rel) return launch_attn_mma64_inst<DHV, true>(qkv_planes, qkv_plane_stride, r_planes, r_plane_stride, rw, rr, \
                                                      B, L, d, H, out_planes, out_plane_stride, s);                   \
    return launch_attn_mma64_inst<DHV, false>(qkv_planes, qkv_plane_stride, nullptr, 0, nullptr, nullptr, B, L, d, H, \
                                              out_planes, out_plane_stride, s);                                       \
  }
    T4R_AM64(16) T4R_AM64(32) T4R_AM64(64)
#undef T4R_AM64
    return T4R_ERR_UNSUPPORTED;
  }
#define T4R_AM(DHV)                                                                                                  \
  if (dh == DHV) {                                                                                                   \
    if (rel) return launch_attn_mma_inst<DHV, true>(qkv_planes, qkv_plane_stride, r_planes, r_plane_stride, rw, rr,   \
                                                    B, L, d, H, out_planes, out_plane_stride, s);                     \
    return launch_attn_mma_inst<DHV, false>(qkv_planes, qkv_plane_stride, nullptr, 0, nullptr, nullptr, B, L, d, H,   \
                                            out_planes, out_plane_stride, s);                                         \
  }
  T4R_AM(16) T4R_AM(32) T4R_AM(64)
#undef T4R_AM
  return T4R_ERR_UNSUPPORTED;
}

// XLNet two-stream attention for permutation language modeling: planes hold 2 B L rows (h, then g), see attn_mma_kernel
int launch_attn_mma_plm(const __nv_bfloat16* qkv_planes, int64_t qkv_plane_stride, const __nv_bfloat16* r_planes,
                        int64_t r_plane_stride, const float* rw, const float* rr, int B, int L, int d, int H,
                        __nv_bfloat16* out_planes, int64_t out_plane_stride, const uint8_t* plm_mask, cudaStream_t s) {
  T4R_REQUIRE(plm_mask != nullptr, "attn_mma_plm: the permutation mask is required");
  T4R_REQUIRE(attn_mma_supported(L, d, H, true),
              "PLM needs the tensor-path attention: L + 2 <= 32 (or <= 64 with T4R_ATTN_MMA64=1), got L=%d d=%d H=%d", L, d, H);
  const int dh = d / H;
  const bool big = L + 2 > 32;
#define T4R_AMP(DHV)                                                                                                   \
  if (dh == DHV) {                                                                                                     \
    if (big) return launch_attn_mma64_inst<DHV, true, true>(qkv_planes, qkv_plane_stride, r_planes, r_plane_stride, rw, \
                                                            rr, B, L, d, H, out_planes, out_plane_stride, s, plm_mask); \
    return launch_attn_mma_inst<DHV, true, true>(qkv_planes, qkv_plane_stride, r_planes, r_plane_stride, rw, rr, B, L,  \
                                                 d, H, out_planes, out_plane_stride, s, plm_mask);                      \
  }
  T4R_AMP(16) T4R_AMP(32) T4R_AMP(64)
#undef T4R_AMP
  return T4R_ERR_UNSUPPORTED;
}

}  // namespace t4r
