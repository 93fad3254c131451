// t4r_plm.cu -- Permutation Language Modeling masks: device kernel (one thread per session; the
// reference runs python loops per session on the host) and its host twin.
#include "t4r_common.cuh"
#include "t4r_internal.h"
#include "t4r_plm_mask.cuh"

namespace t4r {

__global__ void __launch_bounds__(128)
mask_plm_kernel(const __grid_constant__ PlmParams p, const int64_t* __restrict__ ids, int B, const float* __restrict__ u_span,
                const float* __restrict__ u_start, const float* __restrict__ u_force, const float* __restrict__ u_unmask,
                const int32_t* __restrict__ perm, uint8_t* __restrict__ mask, int64_t* __restrict__ labels,
                uint8_t* __restrict__ perm_mask) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int L = p.L;
  const bool train = p.mode == 0;
  plm_mask_session(p, ids + static_cast<int64_t>(b) * L, train ? u_span + static_cast<int64_t>(b) * L : nullptr,
                   train ? u_start + static_cast<int64_t>(b) * L : nullptr, train ? u_force[b] : 0.f,
                   train ? u_unmask[b] : 0.f, train ? perm + static_cast<int64_t>(b) * L : nullptr,
                   mask + static_cast<int64_t>(b) * L, labels + static_cast<int64_t>(b) * L,
                   perm_mask + static_cast<int64_t>(b) * L * L);
}

static int fill_params(PlmParams& p, int L, int mode, int64_t padding_idx, int max_span, const int32_t* ctx_len) {
  T4R_REQUIRE(L >= 1 && L <= kPlmMaxL, "mask_plm: 1 <= L <= %d (got %d)", kPlmMaxL, L);
  T4R_REQUIRE(mode >= 0 && mode <= 2, "mask_plm: mode must be T4R_PLM_TRAIN / EVAL_LAST / EVAL_ALL");
  p.L = L; p.mode = mode; p.padding_idx = padding_idx; p.max_span = max_span;
  for (int i = 0; i <= kPlmMaxSpan; ++i) p.ctx_len[i] = 0;
  if (mode == 0) {
    T4R_REQUIRE(max_span >= 1 && max_span <= kPlmMaxSpan && ctx_len, "mask_plm: 1 <= max_span_length <= %d", kPlmMaxSpan);
    for (int sp = 1; sp <= max_span; ++sp) {
      T4R_REQUIRE(ctx_len[sp] >= sp, "mask_plm: context length %d of span %d (needs plm_probability <= 1)", ctx_len[sp], sp);
      p.ctx_len[sp] = ctx_len[sp];
    }
  }
  return 0;
}

}  // namespace t4r

extern "C" int t4r_mask_plm(const int64_t* item_ids, int B, int L, int64_t padding_idx, int mode, int max_span,
                            const int32_t* ctx_len /*host*/, const float* u_span, const float* u_start, const float* u_force,
                            const float* u_unmask, const int32_t* perm, uint8_t* mask_schema, int64_t* masked_targets,
                            uint8_t* perm_mask, void* stream) {
  using namespace t4r;
  T4R_REQUIRE(item_ids && mask_schema && masked_targets && perm_mask && B > 0, "mask_plm: bad arguments");
  T4R_REQUIRE(mode != 0 || (u_span && u_start && u_force && u_unmask && perm), "mask_plm: training needs all the draws");
  PlmParams p;
  T4R_TRY(fill_params(p, L, mode, padding_idx, max_span, ctx_len));
  mask_plm_kernel<<<(B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(p, item_ids, B, u_span, u_start, u_force,
                                                                                 u_unmask, perm, mask_schema, masked_targets,
                                                                                 perm_mask);
  T4R_LAUNCH_CHECK("mask_plm_kernel");
  return 0;
}

// Host twin (HOST pointers, no CUDA call): the same plm_mask_session compiled for the CPU -- test infrastructure.
extern "C" int t4r_debug_mask_plm_host(const int64_t* item_ids, int B, int L, int64_t padding_idx, int mode, int max_span,
                                       const int32_t* ctx_len, const float* u_span, const float* u_start,
                                       const float* u_force, const float* u_unmask, const int32_t* perm,
                                       uint8_t* mask_schema, int64_t* masked_targets, uint8_t* perm_mask) {
  using namespace t4r;
  T4R_REQUIRE(item_ids && mask_schema && masked_targets && perm_mask && B > 0, "mask_plm_host: bad arguments");
  T4R_REQUIRE(mode != 0 || (u_span && u_start && u_force && u_unmask && perm), "mask_plm_host: training needs all the draws");
  PlmParams p;
  T4R_TRY(fill_params(p, L, mode, padding_idx, max_span, ctx_len));
  const bool train = mode == 0;
  for (int b = 0; b < B; ++b)
    plm_mask_session(p, item_ids + static_cast<int64_t>(b) * L, train ? u_span + static_cast<int64_t>(b) * L : nullptr,
                     train ? u_start + static_cast<int64_t>(b) * L : nullptr, train ? u_force[b] : 0.f,
                     train ? u_unmask[b] : 0.f, train ? perm + static_cast<int64_t>(b) * L : nullptr,
                     mask_schema + static_cast<int64_t>(b) * L, masked_targets + static_cast<int64_t>(b) * L,
                     perm_mask + static_cast<int64_t>(b) * L * L);
  return 0;
}
