// t4r_input_block.cu -- the general form of the input block and the two
// small index kernels that sit next to it.
//
//   t4r_input_block_fwd      every feature kind / per-feature LayerNorm / aggregation the reference's
//                            TabularSequenceFeatures can be configured with, in ONE kernel
//   t4r_swap_noise           StochasticSwapNoise (integer / copy work, bit-exact)
//   t4r_metrics_from_ranks   precision / recall / reciprocal-rank / DCG at k from the label's rank
//
// The plain configuration (categorical + scalar continuous features, concat) keeps its specialised
// gather kernel (t4r_embed_concat_fwd in t4r_kernels.cu): that one is the HBM-roofline kernel of
// the benchmark; this file trades peak bandwidth for generality.
#include <math.h>

#include "t4r_common.cuh"
#include "t4r_internal.h"

namespace t4r {

constexpr int kIbWarps = 4;          // rows per block
constexpr int kIbMaxDim = 1024;      // widest single feature / element-wise aggregate
constexpr int kIbMaxCard = 128;      // most soft-embedding bins
constexpr int kIbSmemFloats = 3 * kIbMaxDim + kIbMaxCard;  // per warp: feature row | aggregate | item row | soft weights

struct FeatArr {
  t4r_feature f[T4R_MAX_FEATURES];
  int n;
};

// One warp per output row.  Per feature: values -> shared-memory row buffer (lane-strided, so
// every global access is coalesced), optional LayerNorm over the feature's own width (two-pass
// variance like torch), then either written to its concat columns or accumulated.
__global__ void __launch_bounds__(kIbWarps * 32)
input_block_kernel(const __grid_constant__ FeatArr fa, int64_t M, int L, int agg, int item_feature, float ln_eps, int C,
                   int Cp, float* __restrict__ out_f32, __nv_bfloat16* __restrict__ planes, int32_t* err_flag) {
  extern __shared__ float ib_smem[];
  const int warp = warp_id(), lane = lane_id();
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kIbWarps + warp;
  if (row >= M) return;
  float* buf = ib_smem + warp * kIbSmemFloats;
  float* acc = buf + kIbMaxDim;
  float* item = acc + kIbMaxDim;
  float* sw = item + kIbMaxDim;
  const bool elementwise = (agg != T4R_AGG_CONCAT);
  if (elementwise)
    for (int e = lane; e < C; e += 32) acc[e] = 0.f;

  for (int t = 0; t < fa.n; ++t) {
    const t4r_feature& f = fa.f[t];
    const int dim = f.dim;
    const int64_t src = f.per_session ? row / L : row;
    // ---- values of this feature for this row -> buf[0, dim)
    if (f.kind == T4R_FEAT_CAT) {
      int64_t id = static_cast<const int64_t*>(f.input)[src];
      if (id < 0 || id >= f.card) {
        if (err_flag && lane == 0) *err_flag = 1;
        id = 0;
      }
      const float* trow = f.table + id * dim;
      for (int e = lane; e < dim; e += 32) buf[e] = __ldg(trow + e);
    } else if (f.kind == T4R_FEAT_CONT) {
      if (lane == 0) buf[0] = static_cast<const float*>(f.input)[src];
    } else if (f.kind == T4R_FEAT_DENSE) {
      const float* drow = static_cast<const float*>(f.input) + src * dim;
      for (int e = lane; e < dim; e += 32) buf[e] = drow[e];
    } else {  // T4R_FEAT_SOFT: softmax(x * w + b) over the bins, then the weighted mean of the bin rows
      const float x = static_cast<const float*>(f.input)[src];
      float mx = -INFINITY;
      for (int k = lane; k < f.card; k += 32) {
        const float z = fmaf(x, __ldg(f.soft_w + k), __ldg(f.soft_b + k));
        sw[k] = z;
        mx = fmaxf(mx, z);
      }
      mx = warp_max(mx);
      float den = 0.f;
      for (int k = lane; k < f.card; k += 32) {
        const float p = expf(sw[k] - mx);
        sw[k] = p;
        den += p;
      }
      den = warp_sum(den);
      __syncwarp();
      const float inv = 1.f / den;
      for (int e = lane; e < dim; e += 32) {
        float s = 0.f;
        for (int k = 0; k < f.card; ++k) s = fmaf(sw[k] * inv, __ldg(f.table + static_cast<int64_t>(k) * dim + e), s);
        buf[e] = s;
      }
    }
    __syncwarp();
    // ---- TabularLayerNorm of this feature (tabular/transformations.py:95-141)
    if (f.ln_gamma) {
      float s = 0.f;
      for (int e = lane; e < dim; e += 32) s += buf[e];
      const float mean = warp_sum(s) / static_cast<float>(dim);
      float q = 0.f;
      for (int e = lane; e < dim; e += 32) {
        const float d = buf[e] - mean;
        q = fmaf(d, d, q);
      }
      const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(dim) + ln_eps);
      for (int e = lane; e < dim; e += 32) buf[e] = (buf[e] - mean) * rstd * __ldg(f.ln_gamma + e) + __ldg(f.ln_beta + e);
      __syncwarp();
    }
    // ---- aggregate
    if (!elementwise) {
      for (int e = lane; e < dim; e += 32) {
        const float v = buf[e];
        const int c = f.col + e;
        if (out_f32) out_f32[row * C + c] = v;
        if (planes) {
          __nv_bfloat16 h, l;
          split_bf16(v, h, l);
          planes[row * Cp + c] = h;
          planes[(M + row) * Cp + c] = l;
        }
      }
    } else if (agg == T4R_AGG_SUM_ITEM_MULTI && t == item_feature) {
      for (int e = lane; e < dim; e += 32) item[e] = buf[e];
    } else {
      for (int e = lane; e < dim; e += 32) acc[e] += buf[e];
    }
    __syncwarp();
  }

  if (elementwise) {
    for (int e = lane; e < C; e += 32) {
      float v = acc[e];
      if (agg == T4R_AGG_SUM_ITEM_MULTI) v *= item[e];
      if (out_f32) out_f32[row * C + e] = v;
      if (planes) {
        __nv_bfloat16 h, l;
        split_bf16(v, h, l);
        planes[row * Cp + e] = h;
        planes[(M + row) * Cp + e] = l;
      }
    }
  }
  if (planes) {
    const __nv_bfloat16 z = __float2bfloat16_rn(0.f);
    for (int c = C + lane; c < Cp; c += 32) {
      planes[row * Cp + c] = z;
      planes[(M + row) * Cp + c] = z;
    }
  }
}

// ============================================================================
// StochasticSwapNoise (tabular/transformations.py:29-92), one tensor per call.
//   keep[i]  = mask[i] (non-pad position), rep[i] = (u[i] < p) & keep[i]
//   pool     = values[keep]                       (row-major order)
//   out[rep] = pool[perm][: rep.sum()]            (row-major order of the replaced positions)
// One block: two exclusive scans over the n positions (pool index, replacement index), then
// the gather through the caller's permutation.  The draws (u, perm) are inputs so that the
// reference, the oracle and this kernel can be fed the same ones.
// ============================================================================
template <typename T>
__global__ void __launch_bounds__(1024)
swap_noise_kernel(const T* __restrict__ values, const uint8_t* __restrict__ keep, int64_t keep_stride, int inner,
                  const float* __restrict__ u, float p, const int64_t* __restrict__ perm, int64_t n,
                  int32_t* __restrict__ pool_pos, T* __restrict__ out) {
  __shared__ int wsum[32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // keep mask of element i: mask row i / inner (one mask entry per `inner` consecutive values when
  // the tensor has a trailing dimension the mask lacks), always 1 when no mask is given
  auto kept = [&](int64_t i) -> bool { return keep ? keep[(i / inner) * keep_stride] != 0 : true; };
  // pass 1: pool_pos[j] = position of the j-th kept element
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += blockDim.x) {
    const int64_t i = base + tid;
    const int flag = (i < n && kept(i)) ? 1 : 0;
    int incl = flag;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += v;
      }
      wsum[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const int carry = carry_s;
    const int excl = carry + (warp ? wsum[warp - 1] : 0) + incl - flag;
    if (flag) pool_pos[excl] = static_cast<int32_t>(i);
    __syncthreads();
    if (tid == blockDim.x - 1) carry_s = carry + wsum[31];
    __syncthreads();
  }
  // pass 2: k-th replaced position (row-major) takes pool[perm[k]]
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += blockDim.x) {
    const int64_t i = base + tid;
    const bool in = i < n;
    const int flag = (in && kept(i) && u[i] < p) ? 1 : 0;
    int incl = flag;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += v;
      }
      wsum[lane] = w;
    }
    __syncthreads();
    const int carry = carry_s;
    const int k = carry + (warp ? wsum[warp - 1] : 0) + incl - flag;
    if (in) out[i] = flag ? values[pool_pos[perm[k]]] : values[i];
    __syncthreads();
    if (tid == blockDim.x - 1) carry_s = carry + wsum[31];
    __syncthreads();
  }
}

// ============================================================================
// ranking metrics at k from the label's rank (single relevant item per row)
//   recall = [r<k]   precision = [r<k]/k   reciprocal rank (MAP@k == MRR@k) = [r<k]/(r+1)
//   DCG (== NDCG: the ideal DCG of one relevant item is 1) = [r<k]/log2(r+2)
// Deterministic: fixed thread -> row assignment and a fixed-order tree reduction.
// ============================================================================
__global__ void __launch_bounds__(1024)
metrics_from_ranks_kernel(const int32_t* __restrict__ rank, const int32_t* __restrict__ t_dev, int T_cap, int kind,
                          int k0, int k1, int k2, int k3, int n_ks, float* __restrict__ out) {
  __shared__ float red[4][32];
  const int T = t_dev ? min(*t_dev, T_cap) : T_cap;
  const int ks[4] = {k0, k1, k2, k3};
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = threadIdx.x; i < T; i += blockDim.x) {
    const int r = rank[i];
    float g = 1.f;
    if (kind == T4R_METRIC_RR) g = 1.f / static_cast<float>(r + 1);
    else if (kind == T4R_METRIC_DCG) g = 1.f / (logf(static_cast<float>(r) + 2.f) / logf(2.f));
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (r < ks[j]) s[j] += (kind == T4R_METRIC_PRECISION) ? 1.f / static_cast<float>(ks[j]) : g;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float v = s[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[j][warp] = v;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = red[j][lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (lane == 0 && j < n_ks) out[j] = (T > 0) ? v / static_cast<float>(T) : 0.f;
    }
  }
}

}  // namespace t4r

extern "C" int t4r_input_block_fwd(const t4r_feature* feats, int n_feats, int64_t M, int L, int agg, int item_feature,
                                   float ln_eps, int C, float* out_f32, void* out_planes, int32_t* err_flag,
                                   void* stream) {
  using namespace t4r;
  T4R_REQUIRE(feats && n_feats >= 1 && n_feats <= T4R_MAX_FEATURES, "input_block: 1..%d features", T4R_MAX_FEATURES);
  T4R_REQUIRE(M > 0 && L > 0 && C > 0 && (out_f32 || out_planes), "input_block: bad arguments");
  T4R_REQUIRE(agg == T4R_AGG_CONCAT || agg == T4R_AGG_SUM || agg == T4R_AGG_SUM_ITEM_MULTI, "input_block: unknown aggregation %d", agg);
  FeatArr fa;
  fa.n = n_feats;
  int width = 0;
  for (int t = 0; t < n_feats; ++t) {
    const t4r_feature& f = feats[t];
    T4R_REQUIRE(f.kind >= T4R_FEAT_CAT && f.kind <= T4R_FEAT_DENSE && f.input && f.dim >= 1 && f.dim <= kIbMaxDim,
                "input_block: feature %d: bad kind / input / width (1..%d)", t, kIbMaxDim);
    T4R_REQUIRE(f.kind != T4R_FEAT_CONT || f.dim == 1, "input_block: feature %d: scalar features have width 1", t);
    T4R_REQUIRE((f.kind != T4R_FEAT_CAT && f.kind != T4R_FEAT_SOFT) || (f.table && f.card >= 1),
                "input_block: feature %d: table missing", t);
    T4R_REQUIRE(f.kind != T4R_FEAT_SOFT || (f.soft_w && f.soft_b && f.card <= kIbMaxCard),
                "input_block: feature %d: soft embedding needs its projection and at most %d bins", t, kIbMaxCard);
    T4R_REQUIRE((f.ln_gamma == nullptr) == (f.ln_beta == nullptr), "input_block: feature %d: ln_gamma and ln_beta go together", t);
    if (agg == T4R_AGG_CONCAT) {
      T4R_REQUIRE(f.col >= 0 && f.col + f.dim <= C, "input_block: feature %d columns out of range", t);
      width += f.dim;
    } else {
      // aggregation.py:104-134: element-wise aggregation needs equal shapes
      T4R_REQUIRE(f.dim == C, "The last dim of all input features is not equal, which is required for element-wise aggregation");
    }
    fa.f[t] = f;
  }
  if (agg == T4R_AGG_CONCAT) T4R_REQUIRE(width == C, "input_block: feature widths sum to %d but C = %d", width, C);
  if (agg == T4R_AGG_SUM_ITEM_MULTI)
    T4R_REQUIRE(item_feature >= 0 && item_feature < n_feats && n_feats >= 2, "input_block: item-multi needs the item feature and one other");
  const int Cp = t4r_round_up64(C);
  const size_t smem = static_cast<size_t>(kIbWarps) * kIbSmemFloats * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    T4R_CUDA(cudaFuncSetAttribute(input_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_set = true;
  }
  const int64_t blocks = (M + kIbWarps - 1) / kIbWarps;
  input_block_kernel<<<static_cast<unsigned>(blocks), kIbWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(
      fa, M, L, agg, item_feature, ln_eps, C, Cp, out_f32, static_cast<__nv_bfloat16*>(out_planes), err_flag);
  T4R_LAUNCH_CHECK("input_block_kernel");
  return 0;
}

extern "C" int t4r_swap_noise(const void* values, int elem_bytes, int64_t n, const uint8_t* keep_mask, int64_t keep_stride,
                              int inner, const float* u, float replacement_prob, const int64_t* perm,
                              int32_t* scratch_pool_pos, void* out, void* stream) {
  using namespace t4r;
  T4R_REQUIRE(values && u && perm && scratch_pool_pos && out && n > 0 && n < (1ll << 31) && inner >= 1,
              "swap_noise: bad arguments");
  T4R_REQUIRE(elem_bytes == 4 || elem_bytes == 8, "swap_noise: 4- or 8-byte elements");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (elem_bytes == 8)
    swap_noise_kernel<long long><<<1, 1024, 0, s>>>(static_cast<const long long*>(values), keep_mask, keep_stride, inner, u,
                                                    replacement_prob, perm, n, scratch_pool_pos,
                                                    static_cast<long long*>(out));
  else
    swap_noise_kernel<float><<<1, 1024, 0, s>>>(static_cast<const float*>(values), keep_mask, keep_stride, inner, u,
                                                replacement_prob, perm, n, scratch_pool_pos, static_cast<float*>(out));
  T4R_LAUNCH_CHECK("swap_noise_kernel");
  return 0;
}

extern "C" int t4r_metrics_from_ranks(const int32_t* row_rank, const int32_t* t_dev, int T_cap, int kind,
                                      const int32_t* ks, int n_ks, float* out, void* stream) {
  using namespace t4r;
  T4R_REQUIRE(row_rank && ks && out && T_cap > 0 && n_ks >= 1 && n_ks <= 4, "metrics_from_ranks: 1..4 cut-offs");
  T4R_REQUIRE(kind >= T4R_METRIC_RECALL && kind <= T4R_METRIC_DCG, "metrics_from_ranks: unknown metric %d", kind);
  int k[4] = {0, 0, 0, 0};
  for (int i = 0; i < n_ks; ++i) {
    T4R_REQUIRE(ks[i] >= 1, "metrics_from_ranks: cut-offs start at 1");
    k[i] = ks[i];
  }
  metrics_from_ranks_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(row_rank, t_dev, T_cap, kind, k[0], k[1],
                                                                               k[2], k[3], n_ks, out);
  T4R_LAUNCH_CHECK("metrics_from_ranks_kernel");
  return 0;
}
