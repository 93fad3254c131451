// t4r_kernels.cu -- the HBM-/latency-bound kernels of the path: fused embedding
// gather + concat (K1), mask/label generation (K3), label compaction, plane
// packing, short-session attention (K5), LayerNorm prologue, head reductions.
#include <math.h>

#include "t4r_common.cuh"
#include "t4r_internal.h"

namespace t4r {

// ============================================================================
// K1: fused multi-table gather + continuous + concat
//     one warp per (b, l) row; each lane moves 16-byte chunks when the feature
//     geometry allows it (dim % 4 == 0, col % 4 == 0), else 4-byte elements.
// ============================================================================
struct FeatDev {
  t4r_feature_list f;
};

// ROWS positions per warp: the ids of all ROWS rows are fetched first and all table-row loads
// are issued before any store, so a warp keeps ROWS x (dim/128) 16-byte loads per lane in flight
// (with one row per warp the id -> row dependency makes the kernel latency-bound at ~58 % of HBM).
constexpr int kGatherRows = 4;

__global__ void __launch_bounds__(256)
embed_concat_kernel(const __grid_constant__ FeatDev fd, int64_t M, int C, int Cp, float* __restrict__ out_f32,
                    __nv_bfloat16* __restrict__ planes, int32_t* err_flag) {
  constexpr int R = kGatherRows;
  const int64_t row0 = (static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + warp_id()) * R;
  if (row0 >= M) return;
  const int lane = lane_id();
  const t4r_feature_list& f = fd.f;
  const bool out_vec_ok = (C % 4 == 0);
  const int nrows = (M - row0 < R) ? static_cast<int>(M - row0) : R;

  for (int t = 0; t < f.n_cat; ++t) {
    const int dim = f.dim[t];
    const int col = f.cat_col[t];
    int64_t id[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      id[r] = (r < nrows) ? f.ids[t][row0 + r] : 0;
      if (id[r] < 0 || id[r] >= f.table_rows[t]) {
        if (err_flag && lane == 0) *err_flag = 1;
        id[r] = 0;
      }
    }
    if ((dim & 3) == 0 && (col & 3) == 0) {
      for (int q0 = 0; q0 < dim / 4; q0 += 64) {  // up to two 16-byte chunks per lane and row per pass
        float4 v[R][2];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int q = q0 + lane + 32 * k;
            if (r < nrows && q < dim / 4)
              v[r][k] = __ldg(reinterpret_cast<const float4*>(f.table[t] + id[r] * dim) + q);
          }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int q = q0 + lane + 32 * k;
            if (r < nrows && q < dim / 4) {
              const int c = col + 4 * q;
              const int64_t row = row0 + r;
              const float4 x = v[r][k];
              if (out_f32) {
                float* of = out_f32 + row * C;
                if (out_vec_ok) {
                  *reinterpret_cast<float4*>(of + c) = x;
                } else {
                  of[c] = x.x; of[c + 1] = x.y; of[c + 2] = x.z; of[c + 3] = x.w;
                }
              }
              if (planes) {
                uint32_t h01, l01, h23, l23;
                split_bf16x2(x.x, x.y, h01, l01);
                split_bf16x2(x.z, x.w, h23, l23);
                *reinterpret_cast<uint2*>(planes + row * Cp + c) = make_uint2(h01, h23);
                *reinterpret_cast<uint2*>(planes + (M + row) * Cp + c) = make_uint2(l01, l23);
              }
            }
          }
      }
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (r >= nrows) break;
        const float* src = f.table[t] + id[r] * dim;
        const int64_t row = row0 + r;
        for (int e = lane; e < dim; e += 32) {
          const float v = __ldg(src + e);
          if (out_f32) out_f32[row * C + col + e] = v;
          if (planes) {
            __nv_bfloat16 h, l;
            split_bf16(v, h, l);
            planes[row * Cp + col + e] = h;
            planes[(M + row) * Cp + col + e] = l;
          }
        }
      }
    }
  }
  for (int r = 0; r < nrows; ++r) {
    const int64_t row = row0 + r;
    for (int t = lane; t < f.n_cont; t += 32) {
      const float v = f.cont[t][row];
      const int c = f.cont_col[t];
      if (out_f32) out_f32[row * C + c] = v;
      if (planes) {
        __nv_bfloat16 h, l;
        split_bf16(v, h, l);
        planes[row * Cp + c] = h;
        planes[(M + row) * Cp + c] = l;
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
}

}  // namespace t4r

extern "C" int t4r_embed_concat_fwd(const t4r_feature_list* feats, int64_t M, int C, float* out_f32,
                                    void* out_planes, int32_t* err_flag, void* stream) {
  using namespace t4r;
  T4R_REQUIRE(feats != nullptr && M > 0 && C > 0, "embed_concat: bad arguments");
  T4R_REQUIRE(feats->n_cat >= 0 && feats->n_cat <= T4R_MAX_FEATURES && feats->n_cont >= 0 &&
                  feats->n_cont <= T4R_MAX_FEATURES,
              "embed_concat: at most %d categorical and %d continuous features", T4R_MAX_FEATURES, T4R_MAX_FEATURES);
  T4R_REQUIRE(out_f32 || out_planes, "embed_concat: no output requested");
  int width = feats->n_cont;
  for (int t = 0; t < feats->n_cat; ++t) {
    T4R_REQUIRE(feats->table[t] && feats->ids[t] && feats->dim[t] > 0, "embed_concat: feature %d incomplete", t);
    T4R_REQUIRE(feats->cat_col[t] >= 0 && feats->cat_col[t] + feats->dim[t] <= C, "embed_concat: feature %d columns out of range", t);
    width += feats->dim[t];
  }
  T4R_REQUIRE(width == C, "embed_concat: feature widths sum to %d but C = %d", width, C);
  FeatDev fd;
  fd.f = *feats;
  const int Cp = t4r_round_up64(C);
  const int warps = 8;
  const int64_t rows_per_block = static_cast<int64_t>(warps) * kGatherRows;
  const int64_t blocks = (M + rows_per_block - 1) / rows_per_block;
  embed_concat_kernel<<<static_cast<unsigned>(blocks), warps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      fd, M, C, Cp, out_f32, static_cast<__nv_bfloat16*>(out_planes), err_flag);
  T4R_LAUNCH_CHECK("embed_concat_kernel");
  return 0;
}

namespace t4r {

// ============================================================================
// N1: ragged (values, offsets) or dense [rows, in_len] -> dense [rows, pad_len], zeros on the right
// ============================================================================
template <typename T>
__global__ void __launch_bounds__(256)
pad_ragged_kernel(const T* __restrict__ values, const int64_t* __restrict__ offsets, int64_t rows, int in_len,
                  int pad_len, T* __restrict__ out) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= rows * pad_len) return;
  const int64_t row = idx / pad_len;
  const int l = static_cast<int>(idx % pad_len);
  T v = 0;
  if (offsets) {
    const int64_t beg = offsets[row], end = offsets[row + 1];
    if (l < end - beg) v = values[beg + l];
  } else if (l < in_len) {
    v = values[row * in_len + l];
  }
  out[idx] = v;
}

}  // namespace t4r

extern "C" int t4r_pad_ragged(const void* values, const int64_t* offsets, int64_t rows, int in_len, int pad_len,
                              int elem_bytes, void* out, void* stream) {
  using namespace t4r;
  T4R_REQUIRE(out && rows > 0 && pad_len > 0 && (elem_bytes == 4 || elem_bytes == 8), "pad_ragged: bad arguments");
  T4R_REQUIRE(values != nullptr || offsets != nullptr, "pad_ragged: no input");
  const int64_t n = rows * pad_len;
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (elem_bytes == 8)
    pad_ragged_kernel<long long><<<blocks, 256, 0, s>>>(static_cast<const long long*>(values), offsets, rows, in_len,
                                                         pad_len, static_cast<long long*>(out));
  else
    pad_ragged_kernel<float><<<blocks, 256, 0, s>>>(static_cast<const float*>(values), offsets, rows, in_len, pad_len,
                                                     static_cast<float*>(out));
  T4R_LAUNCH_CHECK("pad_ragged_kernel");
  return 0;
}

namespace t4r {

// ============================================================================
// K3: masks / labels (integer).  One thread per session.
// ============================================================================
__device__ __forceinline__ int kth_pick(double u, int n) {
  int k = static_cast<int>(floor(u * static_cast<double>(n)));
  if (k > n - 1) k = n - 1;
  if (k < 0) k = 0;
  return k;
}

__global__ void mask_mlm_kernel(const int64_t* __restrict__ ids, int B, int L, int64_t pad, int mode, float prob,
                                const float* __restrict__ u, uint8_t* __restrict__ mask, int64_t* __restrict__ labels,
                                uint8_t* __restrict__ code) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int64_t* row = ids + static_cast<int64_t>(b) * L;
  int n_nonpad = 0;
  for (int l = 0; l < L; ++l) n_nonpad += (row[l] != pad);

  if (mode == T4R_MLM_INFERENCE) {
    // masking.py:403-418: labels [B, L+1]; labels[b, n] = ids[b, n-1]
    const int Lo = L + 1;
    int64_t* lab = labels + static_cast<int64_t>(b) * Lo;
    const int64_t src = row[(n_nonpad - 1 + L) % L];  // python index -1 wraps
    for (int l = 0; l < Lo; ++l) {
      const int64_t v = (l == n_nonpad) ? src : pad;
      lab[l] = v;
      const bool m = v != pad;
      mask[static_cast<int64_t>(b) * Lo + l] = m;
      code[static_cast<int64_t>(b) * Lo + l] = m ? 1 : 0;
    }
    return;
  }
  int64_t* lab = labels + static_cast<int64_t>(b) * L;
  if (mode == T4R_MLM_TRAIN) {
    const float* ur = u + static_cast<int64_t>(b) * (L + 2);
    // masking.py:427-436 bernoulli & non_pad
    for (int l = 0; l < L; ++l) lab[l] = (ur[l] < prob && row[l] != pad) ? row[l] : pad;
    // masking.py:438-445 force one label among the non-padded positions
    if (n_nonpad > 0) {
      int k = kth_pick(static_cast<double>(ur[L]), n_nonpad);
      for (int l = 0; l < L; ++l) {
        if (row[l] != pad) {
          if (k == 0) { lab[l] = row[l]; break; }
          --k;
        }
      }
    } else {
      lab[0] = row[0];
    }
    int n_lab = 0;
    for (int l = 0; l < L; ++l) n_lab += (lab[l] != pad);
    // masking.py:447-459 a session made of labels only gets one of them back
    if (n_lab == n_nonpad) {
      if (n_lab > 0) {
        int k = kth_pick(static_cast<double>(ur[L + 1]), n_lab);
        for (int l = 0; l < L; ++l) {
          if (lab[l] != pad) {
            if (k == 0) { lab[l] = pad; break; }
            --k;
          }
        }
      } else {
        lab[0] = pad;
      }
    }
  } else if (mode == T4R_MLM_EVAL_LAST) {
    // masking.py:461-465
    for (int l = 0; l < L; ++l) lab[l] = pad;
    const int last = (n_nonpad - 1 + L) % L;
    lab[last] = row[last];
  } else {
    // predict_all masking.py:182-213
    for (int l = 0; l < L; ++l) lab[l] = (l + 1 < L) ? row[l + 1] : 0;
  }
  for (int l = 0; l < L; ++l) {
    const bool m = lab[l] != pad;
    mask[static_cast<int64_t>(b) * L + l] = m;
    code[static_cast<int64_t>(b) * L + l] = m ? 1 : 0;
  }
}

__global__ void mask_clm_kernel(const int64_t* __restrict__ ids, int B, int L, int64_t pad, int mode,
                                uint8_t* __restrict__ mask, int64_t* __restrict__ labels, uint8_t* __restrict__ code) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int64_t* row = ids + static_cast<int64_t>(b) * L;
  int64_t* lab = labels + static_cast<int64_t>(b) * L;
  uint8_t* mk = mask + static_cast<int64_t>(b) * L;
  uint8_t* cd = code + static_cast<int64_t>(b) * L;
  if (mode == T4R_CLM_INFERENCE) {
    // masking.py:277-279 and :309-317
    for (int l = 0; l < L; ++l) {
      const bool m = row[l] != pad;
      lab[l] = row[l];
      mk[l] = m;
      cd[l] = m ? 0 : 1;
    }
    return;
  }
  // predict_all masking.py:182-213
  int n_lab = 0;
  for (int l = 0; l < L; ++l) {
    const int64_t v = (l + 1 < L) ? row[l + 1] : 0;
    lab[l] = v;
    n_lab += (v != pad);
  }
  if (mode == T4R_CLM_LAST) {
    // masking.py:284-298
    const int last = (n_lab - 1 + L) % L;
    const int64_t keep = lab[last];
    for (int l = 0; l < L; ++l) lab[l] = 0;
    lab[last] = keep;
    for (int l = 0; l < L; ++l) mk[l] = row[l] != pad;
  } else {
    for (int l = 0; l < L; ++l) mk[l] = lab[l] != pad;
  }
  // masking.py:319-337: drop last position (zero row), then where(mask, x, masked_emb)
  for (int l = 0; l < L; ++l) cd[l] = mk[l] ? ((l == L - 1) ? 2 : 0) : 1;
}

// label compaction: single block; every WARP owns a contiguous segment and walks it 32 labels at a time (coalesced
// 256-byte reads, four steps in flight), counting with ballots; one scan over the 32 warp totals; a second walk writes
// the compacted rows in order (position = warp base + popc of the lower lanes' ballots).  The first form gave each
// THREAD a contiguous chunk -- 32 lanes reading 32 different 320-byte-strided lines per instruction -- and cost 43 us
// of the 5.5 ms step for 40 960 labels.
__global__ void __launch_bounds__(1024)
compact_targets_kernel(const int64_t* __restrict__ labels, int64_t n, int64_t pad, int32_t* __restrict__ rows,
                       int64_t* __restrict__ out_labels, int32_t* __restrict__ count) {
  __shared__ int warp_tot[32];
  __shared__ int total_s;
  const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
  const int64_t steps = (n + 32 * 32 - 1) / (32 * 32);        // 32-label steps per warp
  const int64_t seg = steps * 32;
  const int64_t beg = warp * seg;
  const int64_t end = (beg + seg < n) ? beg + seg : n;
  int cnt = 0;
  for (int64_t i0 = beg; i0 < end; i0 += 4 * 32) {
    int64_t v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * 32 + lane;
      v[u] = (i < end) ? labels[i] : pad;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) cnt += __popc(__ballot_sync(0xffffffffu, v[u] != pad));
  }
  if (lane == 0) warp_tot[warp] = cnt;      // every lane holds the warp's total
  __syncthreads();
  if (warp == 0) {
    const int w = warp_tot[lane];
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    warp_tot[lane] = wi - w;  // exclusive
    if (lane == 31) total_s = wi;
  }
  __syncthreads();
  int pos = warp_tot[warp];
  const unsigned lower = (1u << lane) - 1u;
  for (int64_t i0 = beg; i0 < end; i0 += 4 * 32) {
    int64_t v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * 32 + lane;
      v[u] = (i < end) ? labels[i] : pad;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool keep = v[u] != pad;
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const int at = pos + __popc(m & lower);
        rows[at] = static_cast<int32_t>(i0 + u * 32 + lane);
        out_labels[at] = v[u];
      }
      pos += __popc(m);
    }
  }
  const int total = total_s;
  if (tid == 0) *count = total;
  for (int64_t i = total + tid; i < n; i += blockDim.x) {
    rows[i] = 0;
    out_labels[i] = 0;
  }
}

}  // namespace t4r

extern "C" int t4r_mask_mlm(const int64_t* item_ids, int B, int L, int64_t padding_idx, int mode,
                            float mlm_probability, const float* u, uint8_t* mask_schema, int64_t* masked_targets,
                            uint8_t* row_code, void* stream) {
  using namespace t4r;
  T4R_REQUIRE(item_ids && mask_schema && masked_targets && row_code && B > 0 && L > 0, "mask_mlm: bad arguments");
  T4R_REQUIRE(mode >= T4R_MLM_TRAIN && mode <= T4R_MLM_INFERENCE, "mask_mlm: unknown mode %d", mode);
  T4R_REQUIRE(mode != T4R_MLM_TRAIN || u != nullptr, "mask_mlm: training mode needs the uniform draws u[B, L+2]");
  mask_mlm_kernel<<<(B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      item_ids, B, L, padding_idx, mode, mlm_probability, u, mask_schema, masked_targets, row_code);
  T4R_LAUNCH_CHECK("mask_mlm_kernel");
  return 0;
}

extern "C" int t4r_mask_clm(const int64_t* item_ids, int B, int L, int64_t padding_idx, int mode,
                            uint8_t* mask_schema, int64_t* masked_targets, uint8_t* row_code, void* stream) {
  using namespace t4r;
  T4R_REQUIRE(item_ids && mask_schema && masked_targets && row_code && B > 0 && L > 0, "mask_clm: bad arguments");
  T4R_REQUIRE(mode >= T4R_CLM_ALL && mode <= T4R_CLM_INFERENCE, "mask_clm: unknown mode %d", mode);
  mask_clm_kernel<<<(B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      item_ids, B, L, padding_idx, mode, mask_schema, masked_targets, row_code);
  T4R_LAUNCH_CHECK("mask_clm_kernel");
  return 0;
}

extern "C" int t4r_compact_targets(const int64_t* masked_targets, int64_t n, int64_t padding_idx, int32_t* tgt_rows,
                                   int64_t* tgt_labels, int32_t* count_dev, void* stream) {
  using namespace t4r;
  T4R_REQUIRE(masked_targets && tgt_rows && tgt_labels && count_dev && n > 0 && n < (1ll << 31),
              "compact_targets: bad arguments");
  compact_targets_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(masked_targets, n, padding_idx, tgt_rows,
                                                                            tgt_labels, count_dev);
  T4R_LAUNCH_CHECK("compact_targets_kernel");
  return 0;
}

namespace t4r {

// ============================================================================
// plane packing
// ============================================================================
__global__ void __launch_bounds__(256)
split_planes_kernel(const float* __restrict__ x, int64_t rows, int K, int64_t ld, int Kp,
                    const uint8_t* __restrict__ row_code, const float* __restrict__ mask_vec,
                    float* __restrict__ out_f32, __nv_bfloat16* __restrict__ planes) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + warp_id();
  if (row >= rows) return;
  const int lane = lane_id();
  const int code = row_code ? row_code[row] : 0;
  const float* src = x + row * ld;
  __nv_bfloat16* hi = planes ? planes + row * Kp : nullptr;
  __nv_bfloat16* lo = planes ? planes + (rows + row) * Kp : nullptr;
  for (int c = lane; c < Kp; c += 32) {
    float v = 0.f;
    if (c < K) {
      v = (code == 1) ? __ldg(mask_vec + c) : ((code == 2) ? 0.f : src[c]);
      if (out_f32) out_f32[row * K + c] = v;
    }
    if (hi) {
      __nv_bfloat16 h, l;
      split_bf16(v, h, l);
      hi[c] = h;
      lo[c] = l;
    }
  }
}

template <typename IdxT>
__global__ void __launch_bounds__(256)
gather_rows_split_kernel(const float* __restrict__ x, int K, int64_t ld, int Kp, const IdxT* __restrict__ idx,
                         const int32_t* __restrict__ count_dev, int cap, float* __restrict__ out_f32,
                         __nv_bfloat16* __restrict__ planes) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + warp_id();
  if (row >= cap) return;
  const int lane = lane_id();
  const int count = count_dev ? *count_dev : cap;
  // with a device-side count the zero rows stop at the next multiple of 256 (the last row tile a GEMM consumer reads):
  // at the head's capacity of B*L rows with ~13 % of them labels, zero-filling the rest was most of this kernel
  if (count_dev && row >= (static_cast<int64_t>(count) + 255) / 256 * 256) return;
  const bool valid = row < count;
  const float* src = valid ? x + static_cast<int64_t>(idx[row]) * ld : nullptr;
  __nv_bfloat16* hi = planes ? planes + row * Kp : nullptr;
  __nv_bfloat16* lo = planes ? planes + (static_cast<int64_t>(cap) + row) * Kp : nullptr;
  for (int c = lane; c < Kp; c += 32) {
    float v = 0.f;
    if (c < K) {
      if (valid) v = __ldg(src + c);
      if (out_f32) out_f32[row * K + c] = v;
    }
    if (hi) {
      __nv_bfloat16 h, l;
      split_bf16(v, h, l);
      hi[c] = h;
      lo[c] = l;
    }
  }
}

int launch_split_planes(const float* x, int64_t rows, int K, int64_t ld, const uint8_t* row_code,
                        const float* mask_vec, float* out_f32, __nv_bfloat16* planes, cudaStream_t s) {
  T4R_REQUIRE(x && rows > 0 && K > 0 && ld >= K && (out_f32 || planes), "split_planes: bad arguments");
  T4R_REQUIRE(row_code == nullptr || mask_vec != nullptr, "split_planes: row_code needs mask_vec");
  const int Kp = t4r_round_up64(K);
  const int64_t blocks = (rows + 7) / 8;
  split_planes_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(x, rows, K, ld, Kp, row_code, mask_vec, out_f32,
                                                                   planes);
  T4R_LAUNCH_CHECK("split_planes_kernel");
  return 0;
}

}  // namespace t4r

extern "C" int t4r_split_planes(const float* x, int64_t rows, int K, int ld, const uint8_t* row_code,
                                const float* mask_vec, float* out_f32, void* out_planes, void* stream) {
  return t4r::launch_split_planes(x, rows, K, ld, row_code, mask_vec, out_f32,
                                  static_cast<__nv_bfloat16*>(out_planes), static_cast<cudaStream_t>(stream));
}

extern "C" int t4r_gather_rows_split(const float* x, int K, int ld, const int32_t* idx, const int32_t* count_dev,
                                     int cap, float* out_f32, void* out_planes, void* stream) {
  using namespace t4r;
  T4R_REQUIRE(x && idx && K > 0 && ld >= K && cap > 0 && (out_f32 || out_planes), "gather_rows_split: bad arguments");
  gather_rows_split_kernel<int32_t><<<(cap + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, K, ld, t4r_round_up64(K), idx, count_dev, cap, out_f32, static_cast<__nv_bfloat16*>(out_planes));
  T4R_LAUNCH_CHECK("gather_rows_split_kernel");
  return 0;
}

extern "C" int t4r_gather_rows_split_i64(const float* x, int K, int ld, const int64_t* idx, int cap, float* out_f32,
                                         void* out_planes, void* stream) {
  using namespace t4r;
  T4R_REQUIRE(x && idx && K > 0 && ld >= K && cap > 0 && (out_f32 || out_planes), "gather_rows_split_i64: bad arguments");
  gather_rows_split_kernel<int64_t><<<(cap + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, K, ld, t4r_round_up64(K), idx, nullptr, cap, out_f32, static_cast<__nv_bfloat16*>(out_planes));
  T4R_LAUNCH_CHECK("gather_rows_split_kernel");
  return 0;
}

namespace t4r {

// ============================================================================
// XLNet relative positional projection  R[m, :] = pos(m) @ Wr,  m in [0, 2L)
//   pos(m) = [sin(p w) || cos(p w)], p = L - m, w_k = 10000^(-2k/d)
//   HF:models/xlnet/modeling_xlnet.py:930-976 (bi_data=False, clamp_len=-1)
// ============================================================================
struct RelPosLayers {
  const float* wr[T4R_MAX_FEATURES];  // per layer [d, d]
};

// grid = (2L, d/32, n_layer); one warp computes 32 outputs of one relative position for one layer
__global__ void __launch_bounds__(32)
rel_pos_proj_kernel(const __grid_constant__ RelPosLayers lw, int L, int d, float* __restrict__ r_out,
                    __nv_bfloat16* __restrict__ r_planes) {
  extern __shared__ float pos_s[];  // [d]
  const int m = blockIdx.x, layer = blockIdx.z;
  const int n = blockIdx.y * 32 + threadIdx.x;
  const float* __restrict__ wr = lw.wr[layer];
  const float p = static_cast<float>(L - m);
  for (int k = threadIdx.x; k < d; k += 32) {
    const int kk = (k < d / 2) ? k : k - d / 2;
    const float inv_freq = 1.0f / powf(10000.0f, static_cast<float>(2 * kk) / static_cast<float>(d));
    const float a = p * inv_freq;
    pos_s[k] = (k < d / 2) ? sinf(a) : cosf(a);
  }
  __syncwarp();
  if (n >= d) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k = 0; k < d; k += 4) {  // d % 64 == 0: four independent chains, four loads in flight
#pragma unroll
    for (int u = 0; u < 4; ++u) acc[u] = fmaf(pos_s[k + u], __ldg(wr + static_cast<int64_t>(k + u) * d + n), acc[u]);
  }
  const float r = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  const int64_t o = (static_cast<int64_t>(layer) * 2 * L + m) * d + n;
  r_out[o] = r;
  if (r_planes) {
    __nv_bfloat16 hi, lo;
    split_bf16(r, hi, lo);
    // per layer [2, 2L, d]
    r_planes[(static_cast<int64_t>(layer) * 4 * L + m) * d + n] = hi;
    r_planes[(static_cast<int64_t>(layer) * 4 * L + 2 * L + m) * d + n] = lo;
  }
}

int launch_rel_pos_proj(const float* const* wr_layers, int n_layer, int L, int d, float* r_out,
                        __nv_bfloat16* r_planes, cudaStream_t s) {
  T4R_REQUIRE(n_layer >= 1 && n_layer <= T4R_MAX_FEATURES, "rel_pos_proj: at most %d layers per call", T4R_MAX_FEATURES);
  RelPosLayers lw;
  for (int i = 0; i < n_layer; ++i) lw.wr[i] = wr_layers[i];
  dim3 grid(2 * L, (d + 31) / 32, n_layer);
  rel_pos_proj_kernel<<<grid, 32, d * sizeof(float), s>>>(lw, L, d, r_out, r_planes);
  T4R_LAUNCH_CHECK("rel_pos_proj_kernel");
  return 0;
}

// ============================================================================
// K5: attention over one short session per warp.
//   REL = true : XLNet  s[i,j] = ((q_i + r_w_bias) . k_j + (q_i + r_r_bias) . R[j + L - i]) / sqrt(dh)
//                (HF:xlnet:95-140; rel_shift_bnij :81-93 as the index identity), no masks
//   REL = false: GPT-2  s[i,j] = q_i . k_j / sqrt(dh) for j <= i (HF:gpt2:54-72,144-226)
//   out = softmax_j(s) @ V, written as split-bf16 planes (A operand of the O-projection).
// Layout: qkv fp32 [B*L, 3d] (q | k | v, head h at columns h*dh..), r [2L, d].
// A block owns one head (R staged once in shared memory) and its warps loop over
// sessions.  Lane i owns QUERY i: q_i (+ biases) lives in registers, k_j / v_j are
// broadcast reads from shared memory, R[j+L-i] is a lane-varying row read (padded
// stride -> conflict-free float4), the softmax over keys is a serial loop in the
// lane (no shuffles), and P@V accumulates the lane's dh outputs in registers.
// ============================================================================
template <int DH, bool REL>
__global__ void __launch_bounds__(128)
attn_kernel(const float* __restrict__ qkv, const float* __restrict__ r, const float* __restrict__ rw,
            const float* __restrict__ rr, int B, int L, int d, int sessions_per_block,
            __nv_bfloat16* __restrict__ out_planes, int64_t plane_stride) {
  constexpr int DP = DH + 4;  // padded row stride (floats) for lane-varying float4 row reads
  constexpr int V4 = DH / 4;
  extern __shared__ __align__(16) float sm[];
  const int h = blockIdx.y;
  const int warp = warp_id(), lane = lane_id();
  const int nwarps = blockDim.x >> 5;
  const int LS = L | 1;  // odd stride for the per-lane score rows
  float* Rs = sm;                                  // [2L][DP]   (REL only)
  float* wbase = sm + (REL ? 2 * L * DP : 0);
  const int per_warp = L * DP + 2 * L * DH + 32 * LS;
  float* qs = wbase + warp * per_warp;  // [L][DP]
  float* ks = qs + L * DP;              // [L][DH]
  float* vs = ks + L * DH;              // [L][DH]
  float* ps = vs + L * DH;              // [32][LS] scores / probabilities of the lane's query

  if (REL) {
    for (int idx = threadIdx.x; idx < 2 * L * V4; idx += blockDim.x) {
      const int m = idx / V4, c = (idx % V4) * 4;
      *reinterpret_cast<float4*>(Rs + m * DP + c) =
          __ldg(reinterpret_cast<const float4*>(r + static_cast<int64_t>(m) * d + h * DH + c));
    }
  }
  __syncthreads();
  const float scale = rsqrtf(static_cast<float>(DH));
  const int b_begin = blockIdx.x * sessions_per_block;
  const int b_end = min(B, b_begin + sessions_per_block);

  for (int b = b_begin + warp; b < b_end; b += nwarps) {
    const float* base = qkv + static_cast<int64_t>(b) * L * 3 * d + h * DH;
    __syncwarp();
    {  // stage q, k, v: 16-byte loads, several rows in flight per lane
      const int total = L * V4;
#pragma unroll 5
      for (int idx = lane; idx < total; idx += 32) {
        const int i = idx / V4, c = (idx % V4) * 4;
        const float* rowp = base + static_cast<int64_t>(i) * 3 * d + c;
        const float4 q4 = *reinterpret_cast<const float4*>(rowp);
        const float4 k4 = *reinterpret_cast<const float4*>(rowp + d);
        const float4 v4 = *reinterpret_cast<const float4*>(rowp + 2 * d);
        *reinterpret_cast<float4*>(qs + i * DP + c) = q4;
        *reinterpret_cast<float4*>(ks + i * DH + c) = k4;
        *reinterpret_cast<float4*>(vs + i * DH + c) = v4;
      }
    }
    __syncwarp();
    float* prow = ps + lane * LS;
    for (int qt = 0; qt * 32 < L; ++qt) {
      const int i = lane + 32 * qt;
      const bool iok = i < L;
      const int ii = iok ? i : 0;
      // q_i (+ biases) in registers
      float4 qw[V4], qr[V4];
#pragma unroll
      for (int c4 = 0; c4 < V4; ++c4) {
        const float4 q4 = *reinterpret_cast<const float4*>(qs + ii * DP + 4 * c4);
        if (REL) {
          const float4 w4 = __ldg(reinterpret_cast<const float4*>(rw + h * DH + 4 * c4));
          const float4 r4 = __ldg(reinterpret_cast<const float4*>(rr + h * DH + 4 * c4));
          qw[c4] = make_float4(q4.x + w4.x, q4.y + w4.y, q4.z + w4.z, q4.w + w4.w);
          qr[c4] = make_float4(q4.x + r4.x, q4.y + r4.y, q4.z + r4.z, q4.w + r4.w);
        } else {
          qw[c4] = q4;
        }
      }
      // ---- scores over keys j
      float mx = -INFINITY;
#pragma unroll 2
      for (int j = 0; j < L; ++j) {
        float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* kj = reinterpret_cast<const float4*>(ks + j * DH);
#pragma unroll
        for (int c4 = 0; c4 < V4; ++c4) {
          const float4 kk = kj[c4];
          a4.x = fmaf(qw[c4].x, kk.x, a4.x); a4.y = fmaf(qw[c4].y, kk.y, a4.y);
          a4.z = fmaf(qw[c4].z, kk.z, a4.z); a4.w = fmaf(qw[c4].w, kk.w, a4.w);
        }
        float acc = (a4.x + a4.y) + (a4.z + a4.w);
        if (REL) {
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          const float4* rm = reinterpret_cast<const float4*>(Rs + (j + L - ii) * DP);
#pragma unroll
          for (int c4 = 0; c4 < V4; ++c4) {
            const float4 bb = rm[c4];
            b4.x = fmaf(qr[c4].x, bb.x, b4.x); b4.y = fmaf(qr[c4].y, bb.y, b4.y);
            b4.z = fmaf(qr[c4].z, bb.z, b4.z); b4.w = fmaf(qr[c4].w, bb.w, b4.w);
          }
          acc += (b4.x + b4.y) + (b4.z + b4.w);
        }
        acc *= scale;
        if (!REL && j > ii) acc = -INFINITY;
        prow[j] = acc;
        mx = fmaxf(mx, acc);
      }
      // ---- softmax (serial in the lane)
      float sum = 0.f;
      for (int j = 0; j < L; ++j) {
        const float sj = prow[j];
        const float e = (sj == -INFINITY) ? 0.f : expf(sj - mx);
        prow[j] = e;
        sum += e;
      }
      const float inv = 1.f / sum;
      // ---- out_i = sum_j p_ij v_j
      float4 o[V4];
#pragma unroll
      for (int c4 = 0; c4 < V4; ++c4) o[c4] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
      for (int j = 0; j < L; ++j) {
        const float pj = prow[j];
        const float4* vj = reinterpret_cast<const float4*>(vs + j * DH);
#pragma unroll
        for (int c4 = 0; c4 < V4; ++c4) {
          const float4 vv = vj[c4];
          o[c4].x = fmaf(pj, vv.x, o[c4].x); o[c4].y = fmaf(pj, vv.y, o[c4].y);
          o[c4].z = fmaf(pj, vv.z, o[c4].z); o[c4].w = fmaf(pj, vv.w, o[c4].w);
        }
      }
      if (iok) {
        __nv_bfloat16* hi = out_planes + (static_cast<int64_t>(b) * L + i) * d + h * DH;
        __nv_bfloat16* lo = hi + plane_stride;
#pragma unroll
        for (int c4 = 0; c4 < V4; ++c4) {
          uint32_t h01, l01, h23, l23;
          split_bf16x2(o[c4].x * inv, o[c4].y * inv, h01, l01);
          split_bf16x2(o[c4].z * inv, o[c4].w * inv, h23, l23);
          *reinterpret_cast<uint2*>(hi + 4 * c4) = make_uint2(h01, h23);
          *reinterpret_cast<uint2*>(lo + 4 * c4) = make_uint2(l01, l23);
        }
      }
    }
  }
}

template <int DH, bool REL>
static int launch_attn_inst(const float* qkv, const float* r, const float* rw, const float* rr, int B, int L, int d,
                            int H, __nv_bfloat16* out_planes, int64_t plane_stride, cudaStream_t s) {
  constexpr int DP = DH + 4;
  const int LS = L | 1;
  const int per_warp = L * DP + 2 * L * DH + 32 * LS;
  const int r_floats = REL ? 2 * L * DP : 0;
  int warps = 4;
  while (warps > 1 && (r_floats + warps * per_warp) * 4 > 200 * 1024) warps >>= 1;
  const size_t smem = static_cast<size_t>(r_floats + warps * per_warp) * 4;
  T4R_REQUIRE(smem <= 220 * 1024, "attention: L=%d dh=%d needs %zu bytes of shared memory", L, DH, smem);
  auto kern = attn_kernel<DH, REL>;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    T4R_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_smem = smem;
  }
  // sessions per block: R staging is small (2L x dh floats), so favour many resident warps
  int spb = 2 * warps;
  while (spb > warps && static_cast<int64_t>((B + spb - 1) / spb) * H < 148 * 8) spb >>= 1;
  dim3 grid((B + spb - 1) / spb, H);
  kern<<<grid, warps * 32, smem, s>>>(qkv, r, rw, rr, B, L, d, spb, out_planes, plane_stride);
  T4R_LAUNCH_CHECK("attn_kernel");
  return 0;
}

template <bool REL>
static int launch_attn_any(const float* qkv, const float* r, const float* rw, const float* rr, int B, int L, int d,
                           int H, __nv_bfloat16* out_planes, int64_t plane_stride, cudaStream_t s) {
  T4R_REQUIRE(d % H == 0, "attention: d_model %d not divisible by n_head %d", d, H);
  const int dh = d / H;
  T4R_REQUIRE(L >= 1 && L <= 64, "attention: sequence length %d not supported (1..64)", L);
  if (dh == 16) return launch_attn_inst<16, REL>(qkv, r, rw, rr, B, L, d, H, out_planes, plane_stride, s);
  if (dh == 32) return launch_attn_inst<32, REL>(qkv, r, rw, rr, B, L, d, H, out_planes, plane_stride, s);
  if (dh == 64) return launch_attn_inst<64, REL>(qkv, r, rw, rr, B, L, d, H, out_planes, plane_stride, s);
  set_error("attention: head dim %d not supported (16, 32, 64)", dh);
  return T4R_ERR_UNSUPPORTED;
}

int launch_xlnet_attn(const float* qkv, const float* r, const float* rw, const float* rr, int B, int L, int d, int H,
                      __nv_bfloat16* out_planes, int64_t plane_stride, cudaStream_t s) {
  return launch_attn_any<true>(qkv, r, rw, rr, B, L, d, H, out_planes, plane_stride, s);
}
int launch_causal_attn(const float* qkv, int B, int L, int d, int H, __nv_bfloat16* out_planes, int64_t plane_stride,
                       cudaStream_t s) {
  return launch_attn_any<false>(qkv, nullptr, nullptr, nullptr, B, L, d, H, out_planes, plane_stride, s);
}

// ============================================================================
// GPT-2 prologue: h = x + wpe[l]; planes = LN(h)   (HF:gpt2:579-585 and ln_1 of block 0)
// one warp per row
// ============================================================================
__global__ void __launch_bounds__(256)
addpos_ln_kernel(const float* __restrict__ x, const float* __restrict__ wpe, int64_t rows, int L, int d,
                 const float* __restrict__ g, const float* __restrict__ bta, float eps, float* __restrict__ h_out,
                 __nv_bfloat16* __restrict__ planes, int64_t plane_stride) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + warp_id();
  if (row >= rows) return;
  const int lane = lane_id();
  const int l = static_cast<int>(row % L);
  const float* xr = x + row * d;
  const float* pr = wpe ? wpe + static_cast<int64_t>(l) * d : nullptr;
  float sum = 0.f;
  for (int c = lane; c < d; c += 32) {
    const float v = xr[c] + (pr ? __ldg(pr + c) : 0.f);
    if (h_out) h_out[row * d + c] = v;
    sum += v;
  }
  const float mean = warp_sum(sum) / static_cast<float>(d);
  float sq = 0.f;
  for (int c = lane; c < d; c += 32) {
    const float v = xr[c] + (pr ? __ldg(pr + c) : 0.f) - mean;
    sq = fmaf(v, v, sq);
  }
  const float rstd = rsqrtf(warp_sum(sq) / static_cast<float>(d) + eps);
  for (int c = lane; c < d; c += 32) {
    const float v = xr[c] + (pr ? __ldg(pr + c) : 0.f);
    const float o = (v - mean) * rstd * __ldg(g + c) + __ldg(bta + c);
    __nv_bfloat16 hi, lo;
    split_bf16(o, hi, lo);
    planes[row * d + c] = hi;
    planes[row * d + c + plane_stride] = lo;
  }
}

int launch_addpos_ln(const float* x, const float* wpe, int B, int L, int d, const float* g, const float* b, float eps,
                     float* h_out, __nv_bfloat16* planes, int64_t plane_stride, cudaStream_t s) {
  const int64_t rows = static_cast<int64_t>(B) * L;
  addpos_ln_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, s>>>(x, wpe, rows, L, d, g, b, eps, h_out, planes,
                                                                         plane_stride);
  T4R_LAUNCH_CHECK("addpos_ln_kernel");
  return 0;
}

// ============================================================================
// head reductions
// ============================================================================
// exact fp32 logit of the label: out[t] = (xt[t] . W[label - v_offset] + class_bias[label]) * inv_tau
__global__ void __launch_bounds__(256)
target_logit_kernel(const float* __restrict__ xt, const float* __restrict__ w, const int64_t* __restrict__ labels,
                    int T_cap, const int32_t* __restrict__ t_dev, int De, int64_t v_offset, int64_t V,
                    const float* __restrict__ class_bias, float inv_tau, float* __restrict__ out) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + warp_id();
  if (row >= T_cap) return;
  const int T = t_dev ? *t_dev : T_cap;
  const int lane = lane_id();
  float acc = 0.f;
  bool mine = false;
  if (row < T) {
    const int64_t lab = labels[row] - v_offset;
    mine = (lab >= 0 && lab < V);
    if (mine) {
      const float* a = xt + row * De;
      const float* b = w + lab * De;
      for (int c = lane; c < De; c += 32) acc = fmaf(a[c], __ldg(b + c), acc);
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    if (mine && class_bias) acc += class_bias[labels[row]];
    out[row] = mine ? acc * inv_tau : 0.f;
  }
}

int launch_target_logit(const float* xt, const float* w, const int64_t* labels, int T_cap, const int32_t* t_dev,
                        int De, int64_t v_offset, int64_t V, const float* class_bias, float inv_tau, float* out,
                        cudaStream_t s) {
  target_logit_kernel<<<(T_cap + 7) / 8, 256, 0, s>>>(xt, w, labels, T_cap, t_dev, De, v_offset, V, class_bias, inv_tau,
                                                      out);
  T4R_LAUNCH_CHECK("target_logit_kernel");
  return 0;
}

// stage 1: thread = row, block column = chunk of column tiles -> (m, s) per (chunk, row)
__global__ void __launch_bounds__(128)
head_reduce1_kernel(const float* __restrict__ part_m, const float* __restrict__ part_s,
                    const float* __restrict__ part_z, int n_tiles, int part_ld, int T_cap,
                    const int32_t* __restrict__ t_dev, int tiles_per_chunk, float* __restrict__ red_m,
                    float* __restrict__ red_s, float* __restrict__ red_z) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  const int T = t_dev ? min(*t_dev, T_cap) : T_cap;
  if (row >= T) return;
  const int chunk = blockIdx.y;
  const int t0 = chunk * tiles_per_chunk;
  const int t1 = min(n_tiles, t0 + tiles_per_chunk);
  float m = -INFINITY, s = 0.f, z = 0.f;
  for (int t = t0; t < t1; ++t) {
    const float pm = part_m[static_cast<int64_t>(t) * part_ld + row];
    const float ps = part_s[static_cast<int64_t>(t) * part_ld + row];
    if (part_z) z += part_z[static_cast<int64_t>(t) * part_ld + row];
    const float mn = fmaxf(m, pm);
    if (mn > -INFINITY) {
      s = s * exp2f(m - mn) + ps * exp2f(pm - mn);
      m = mn;
    }
  }
  red_m[static_cast<int64_t>(chunk) * part_ld + row] = m;
  red_s[static_cast<int64_t>(chunk) * part_ld + row] = s;
  if (part_z) red_z[static_cast<int64_t>(chunk) * part_ld + row] = z;
}

// stage 2: combine chunks (+ optional positive logit), natural-log lse, per-row loss
__global__ void __launch_bounds__(128)
head_reduce2_kernel(const float* __restrict__ red_m, const float* __restrict__ red_s, const float* __restrict__ red_z,
                    int n_chunks, int part_ld, int T_cap, const int32_t* __restrict__ t_dev,
                    const float* __restrict__ pos_logit, const float* __restrict__ row_tgt, float label_smoothing,
                    float inv_classes, float* __restrict__ row_lse, float* __restrict__ row_loss) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= T_cap) return;
  const int T = t_dev ? min(*t_dev, T_cap) : T_cap;
  if (row >= T) {
    if (row_lse) row_lse[row] = 0.f;
    if (row_loss) row_loss[row] = 0.f;
    return;
  }
  constexpr float kLog2e = 1.4426950408889634f;
  constexpr float kLn2 = 0.6931471805599453f;
  float m = -INFINITY, s = 0.f;
  if (pos_logit) {
    m = pos_logit[row] * kLog2e;
    s = 1.f;
  }
  float z = 0.f;
  for (int c = 0; c < n_chunks; ++c) {
    const float pm = red_m[static_cast<int64_t>(c) * part_ld + row];
    const float ps = red_s[static_cast<int64_t>(c) * part_ld + row];
    if (red_z) z += red_z[static_cast<int64_t>(c) * part_ld + row];
    const float mn = fmaxf(m, pm);
    if (mn > -INFINITY) {
      s = s * exp2f(m - mn) + ps * exp2f(pm - mn);
      m = mn;
    }
  }
  const float lse = (m + log2f(s)) * kLn2;
  if (row_lse) row_lse[row] = lse;
  const float tgt = pos_logit ? pos_logit[row] : row_tgt[row];
  // nn.CrossEntropyLoss(label_smoothing=e): lse - (1-e) z_y - (e/V) sum_j z_j   (losses.py:4-20)
  if (row_loss) row_loss[row] = red_z ? lse - (1.f - label_smoothing) * tgt - label_smoothing * inv_classes * z : lse - tgt;
}

// mean of the first T entries (single block; T is small)
__global__ void __launch_bounds__(1024)
mean_rows_kernel(const float* __restrict__ v, int T_cap, const int32_t* __restrict__ t_dev, float* __restrict__ out) {
  __shared__ float ws[32];
  const int T = t_dev ? min(*t_dev, T_cap) : T_cap;
  double acc = 0.0;
  for (int i = threadIdx.x; i < T; i += blockDim.x) acc += static_cast<double>(v[i]);
  float a = static_cast<float>(acc);
  a = warp_sum(a);
  if (lane_id() == 0) ws[warp_id()] = a;
  __syncthreads();
  if (warp_id() == 0) {
    float t = (lane_id() < (blockDim.x >> 5)) ? ws[lane_id()] : 0.f;
    t = warp_sum(t);
    if (lane_id() == 0) out[0] = (T > 0) ? t / static_cast<float>(T) : 0.f;
  }
}

int launch_head_reduce(const float* part_m, const float* part_s, const float* part_z, int n_tiles, int part_ld,
                       int T_cap, const int32_t* t_dev, const float* pos_logit, const float* row_tgt_in,
                       float label_smoothing, int64_t n_classes, float* row_lse, float* row_loss, float* loss,
                       float* scratch, cudaStream_t s) {
  // scratch: [3, n_chunks, part_ld]
  int n_chunks = n_tiles < 64 ? n_tiles : 64;
  const int tpc = (n_tiles + n_chunks - 1) / n_chunks;
  n_chunks = (n_tiles + tpc - 1) / tpc;
  float* red_m = scratch;
  float* red_s = scratch + static_cast<int64_t>(n_chunks) * part_ld;
  float* red_z = part_z ? scratch + static_cast<int64_t>(2) * 64 * part_ld : nullptr;
  dim3 g1((T_cap + 127) / 128, n_chunks);
  head_reduce1_kernel<<<g1, 128, 0, s>>>(part_m, part_s, part_z, n_tiles, part_ld, T_cap, t_dev, tpc, red_m, red_s,
                                         red_z);
  T4R_LAUNCH_CHECK("head_reduce1_kernel");
  head_reduce2_kernel<<<(T_cap + 127) / 128, 128, 0, s>>>(red_m, red_s, red_z, n_chunks, part_ld, T_cap, t_dev,
                                                           pos_logit, row_tgt_in, label_smoothing,
                                                           1.f / static_cast<float>(n_classes), row_lse, row_loss);
  T4R_LAUNCH_CHECK("head_reduce2_kernel");
  if (loss) {
    mean_rows_kernel<<<1, 1024, 0, s>>>(row_loss, T_cap, t_dev, loss);
    T4R_LAUNCH_CHECK("mean_rows_kernel");
  }
  return 0;
}

// Recall@k from ranks
__global__ void __launch_bounds__(1024)
recall_from_ranks_kernel(const int32_t* __restrict__ rank, const int32_t* __restrict__ t_dev, int T_cap, int k0, int k1,
                         int k2, int k3, int n_ks, float* __restrict__ out) {
  __shared__ int hits[4];
  if (threadIdx.x < 4) hits[threadIdx.x] = 0;
  __syncthreads();
  const int T = t_dev ? min(*t_dev, T_cap) : T_cap;
  int h0 = 0, h1 = 0, h2 = 0, h3 = 0;
  for (int i = threadIdx.x; i < T; i += blockDim.x) {
    const int r = rank[i];
    h0 += r < k0; h1 += r < k1; h2 += r < k2; h3 += r < k3;
  }
  atomicAdd(&hits[0], h0); atomicAdd(&hits[1], h1); atomicAdd(&hits[2], h2); atomicAdd(&hits[3], h3);
  __syncthreads();
  if (threadIdx.x < n_ks) out[threadIdx.x] = (T > 0) ? static_cast<float>(hits[threadIdx.x]) / static_cast<float>(T) : 0.f;
}

// top-k per row of materialised logits (k <= 64), one block per row, iterative
// arg-max with "lower id first" tie-break.  Used on the inference path only.
__global__ void __launch_bounds__(256)
topk_kernel(const float* __restrict__ logits, int64_t V, int64_t ld, int k, float* __restrict__ out_scores,
            int64_t* __restrict__ out_ids) {
  __shared__ float wv[8];
  __shared__ long long wi[8];
  __shared__ float prev_v;
  __shared__ long long prev_i;
  const float* row = logits + static_cast<int64_t>(blockIdx.x) * ld;
  if (threadIdx.x == 0) { prev_v = INFINITY; prev_i = -1; }
  __syncthreads();
  for (int r = 0; r < k; ++r) {
    const float pv = prev_v;
    const long long pi = prev_i;
    float best = -INFINITY;
    long long besti = 0x7fffffffffffffffll;
    for (int64_t c = threadIdx.x; c < V; c += blockDim.x) {
      const float v = row[c];
      // strictly after (pv, pi) in the order (value desc, id asc)
      const bool after = (v < pv) || (v == pv && c > pi);
      if (after && (v > best || (v == best && c < besti))) { best = v; besti = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const long long oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
    }
    if (lane_id() == 0) { wv[warp_id()] = best; wi[warp_id()] = besti; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float bv = wv[0];
      long long bi = wi[0];
      for (int w = 1; w < (blockDim.x >> 5); ++w)
        if (wv[w] > bv || (wv[w] == bv && wi[w] < bi)) { bv = wv[w]; bi = wi[w]; }
      prev_v = bv;
      prev_i = bi;
      out_scores[static_cast<int64_t>(blockIdx.x) * k + r] = bv;
      out_ids[static_cast<int64_t>(blockIdx.x) * k + r] = bi;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
combine_shard_lse_kernel(const float* __restrict__ parts, int world, int T_cap, const int32_t* __restrict__ t_dev,
                         float* __restrict__ row_loss) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= T_cap) return;
  const int T = t_dev ? min(*t_dev, T_cap) : T_cap;
  if (row >= T) { row_loss[row] = 0.f; return; }
  float m = -INFINITY, s = 0.f, tgt = 0.f;
  for (int w = 0; w < world; ++w) {
    const float lse = parts[(static_cast<int64_t>(w) * T_cap + row) * 2 + 0];
    tgt += parts[(static_cast<int64_t>(w) * T_cap + row) * 2 + 1];
    const float mn = fmaxf(m, lse);
    if (mn > -INFINITY) {
      s = s * expf(m - mn) + expf(lse - mn);
      m = mn;
    }
  }
  row_loss[row] = (m + logf(s)) - tgt;
}

// accidental hits of the sampled softmax: hit_col[t] = the column c with col_ids[c] == labels[t], -1 if none.
// col_ids strictly ascending (unique negatives, sorted): one binary search per row.
__global__ void __launch_bounds__(256)
hit_cols_kernel(const int64_t* __restrict__ col_ids, int S, const int64_t* __restrict__ labels, int T_cap,
                const int32_t* __restrict__ t_dev, int32_t* __restrict__ hit_col) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T_cap) return;
  const int T = t_dev ? min(*t_dev, T_cap) : T_cap;
  int32_t hit = -1;
  if (t < T) {
    const int64_t y = labels[t];
    int lo = 0, hi = S;   // first column with col_ids >= y
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(col_ids + mid) < y) lo = mid + 1; else hi = mid;
    }
    if (lo < S && __ldg(col_ids + lo) == y) hit = lo;
  }
  hit_col[t] = hit;
}
int launch_hit_cols(const int64_t* col_ids, int64_t S, const int64_t* labels, int T_cap, const int32_t* t_dev,
                    int32_t* hit_col, cudaStream_t s) {
  hit_cols_kernel<<<(T_cap + 255) / 256, 256, 0, s>>>(col_ids, static_cast<int>(S), labels, T_cap, t_dev, hit_col);
  T4R_LAUNCH_CHECK("hit_cols_kernel");
  return 0;
}

int launch_mean_rows(const float* rows, int cap, const int32_t* t_dev, float* out, cudaStream_t s) {
  mean_rows_kernel<<<1, 1024, 0, s>>>(rows, cap, t_dev, out);
  T4R_LAUNCH_CHECK("mean_rows_kernel");
  return 0;
}

}  // namespace t4r

extern "C" int t4r_recall_from_ranks(const int32_t* row_rank, const int32_t* t_dev, int T_cap, const int32_t* ks,
                                     int n_ks, float* out, void* stream) {
  using namespace t4r;
  T4R_REQUIRE(row_rank && ks && out && T_cap > 0 && n_ks >= 1 && n_ks <= 4, "recall_from_ranks: 1..4 cut-offs");
  int k[4] = {0, 0, 0, 0};
  for (int i = 0; i < n_ks; ++i) k[i] = ks[i];
  recall_from_ranks_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(row_rank, t_dev, T_cap, k[0], k[1], k[2],
                                                                              k[3], n_ks, out);
  T4R_LAUNCH_CHECK("recall_from_ranks_kernel");
  return 0;
}

extern "C" int t4r_topk(const float* logits, int64_t rows, int64_t V, int64_t ld, int k, float* out_scores,
                        int64_t* out_ids, void* stream) {
  using namespace t4r;
  T4R_REQUIRE(logits && out_scores && out_ids && rows > 0 && V > 0 && k >= 1 && k <= 64 && k <= V,
              "topk: bad arguments (k <= 64)");
  topk_kernel<<<static_cast<unsigned>(rows), 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, V, ld, k, out_scores,
                                                                                         out_ids);
  T4R_LAUNCH_CHECK("topk_kernel");
  return 0;
}

extern "C" int t4r_combine_shard_lse(const float* parts, int world, int T_cap, const int32_t* t_dev, float* row_loss,
                                     float* loss, void* stream) {
  using namespace t4r;
  T4R_REQUIRE(parts && row_loss && world >= 1 && T_cap > 0, "combine_shard_lse: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  combine_shard_lse_kernel<<<(T_cap + 255) / 256, 256, 0, s>>>(parts, world, T_cap, t_dev, row_loss);
  T4R_LAUNCH_CHECK("combine_shard_lse_kernel");
  if (loss) {
    mean_rows_kernel<<<1, 1024, 0, s>>>(row_loss, T_cap, t_dev, loss);
    T4R_LAUNCH_CHECK("mean_rows_kernel");
  }
  return 0;
}

extern "C" int t4r_label_logit(const float* xt_f32, const float* w_f32, const int64_t* labels, int T_cap,
                               const int32_t* t_dev, int De, int64_t V, const float* class_bias, float inv_temperature,
                               int64_t v_offset, float* out, void* stream) {
  using namespace t4r;
  T4R_REQUIRE(xt_f32 && w_f32 && labels && out && T_cap > 0 && De > 0 && V > 0, "label_logit: bad arguments");
  return launch_target_logit(xt_f32, w_f32, labels, T_cap, t_dev, De, v_offset, V, class_bias,
                             inv_temperature != 0.f ? inv_temperature : 1.f, out, static_cast<cudaStream_t>(stream));
}
