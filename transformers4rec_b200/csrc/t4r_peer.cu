// t4r_peer.cu -- the row-sharded item table / tied head over NVLink PEER MEMORY (BASELINE configs 4-5).
//
// One process per GPU.  Every rank maps the other ranks' table shards and two small per-step "windows" (label rows
// out, softmax statistics back) into its own address space through CUDA IPC; the data path then needs no bulk
// collective at all:
//   * lookup (K11): peer_gather_rows_kernel reads row `id` straight from the shard of the rank that owns it
//     (16-byte loads over NVLink, one warp per row, several rows in flight per warp) and emits the fp32 rows and / or
//     the split-bf16 planes the projection GEMM consumes -- the all-gather of ids, the routing plan, the all-to-all of
//     rows and the un-permuting second gather of the NCCL formulation are all gone, and exactly the requested bytes move.
//   * head (K12): peer_pull_rows_kernel pulls every rank's label rows (exactly count[r] of them, counts read from
//     device memory -- no host synchronisation) into one compact [T_total, De] operand, writing planes in the same pass;
//     after the local logits + log-sum-exp kernel, peer_combine_lse_kernel reads every shard's (lse, label logit, rank
//     count) for the rows and finishes the loss.  The only NCCL traffic left on the path is two 4-byte collectives that
//     act as stream-ordered barriers (the all-gather of the counts, one all-reduce before the combine).
// Ordering between ranks is by those two collectives; see transformers4rec_b200/distributed.py (PeerHead) for the proof
// sketch of the write-after-read safety of the windows.
//
// Replaces, on a table too large to replicate: EmbeddingFeatures.forward transformers4rec/torch/features/embedding.py:226-249
// and the tied logits + CrossEntropyLoss of model/prediction_task.py:648-671, :446 (the reference itself offers replicas
// only, docs/source/multi_gpu_train.md).
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <string.h>

#include "t4r_common.cuh"
#include "t4r_internal.h"

namespace t4r {

// ------------------------------------------------------------------------------------------------------------------
// windows: export / open / close
// ------------------------------------------------------------------------------------------------------------------
typedef CUresult (*cuMemGetAddressRange_t)(CUdeviceptr*, size_t*, CUdeviceptr);

static int address_range(const void* p, void** base, size_t* size) {
  // resolved through the runtime so that the library carries no link-time dependency on libcuda (it must load, and
  // export every symbol, on a box without a driver: the CPU test tier)
  static cuMemGetAddressRange_t fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuMemGetAddressRange", &sym, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !sym) {
      set_error("peer windows: cuMemGetAddressRange is not available (%s)", cudaGetErrorString(e));
      return T4R_ERR_CUDA;
    }
    fn = reinterpret_cast<cuMemGetAddressRange_t>(sym);
  }
  CUdeviceptr b = 0;
  size_t n = 0;
  CUresult r = fn(&b, &n, reinterpret_cast<CUdeviceptr>(p));
  if (r != CUDA_SUCCESS) {
    set_error("peer windows: cuMemGetAddressRange failed (CUresult %d): not a cudaMalloc'ed device pointer?", int(r));
    return T4R_ERR_CUDA;
  }
  *base = reinterpret_cast<void*>(b);
  *size = n;
  return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// K11 over peer memory
// ------------------------------------------------------------------------------------------------------------------
struct PeerBases {
  const float* base[T4R_MAX_PEERS];
};

constexpr int PG_WARPS = 8;       // warps per CTA
constexpr int PG_ROWS = 4;        // rows in flight per warp (bytes in flight over NVLink: 4 x K x 4 per warp)
constexpr int PG_MAX_K = 1024;    // widest row the staged padding row supports

// One warp per output row, PG_ROWS rows per warp and iteration.  K % 4 == 0 (16-byte loads); Kp = round_up64(K).
template <int VEC_PER_LANE>  // float4 loads per lane and row: K <= 128 * VEC_PER_LANE
__global__ void __launch_bounds__(PG_WARPS * 32)
peer_gather_rows_kernel(PeerBases shards, int world, int64_t V, int64_t per, int K, int Kp,
                        const int64_t* __restrict__ ids, const int32_t* __restrict__ count_dev, int cap, int64_t pad_id,
                        float* __restrict__ out_f32, __nv_bfloat16* __restrict__ planes, int32_t* __restrict__ err_flag) {
  __shared__ __align__(16) float pad_row[PG_MAX_K];
  const int lane = lane_id();
  const int warp = warp_id();
  const int count = count_dev ? min(*count_dev, cap) : cap;
  if (pad_id >= 0 && pad_id < V) {
    // the padding id is ~45 % of all positions (right-padded sessions) and every copy of it lives on ONE rank: read
    // that row once per CTA instead of once per position, so its owner does not become the NVLink hot spot
    const int owner = static_cast<int>(min(pad_id / per, static_cast<int64_t>(world - 1)));
    const float* src = shards.base[owner] + (pad_id - owner * per) * K;
    for (int c = threadIdx.x; c < K; c += blockDim.x) pad_row[c] = src[c];
  }
  __syncthreads();
  const int kv = K >> 2;  // float4 per row
  // with a device-side count only the rows a consumer's last 256-row tile can touch are written (zeros past the count)
  const int64_t limit = count_dev ? min(static_cast<int64_t>(cap), (static_cast<int64_t>(count) + 255) / 256 * 256) : cap;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * PG_WARPS * PG_ROWS;
  for (int64_t r0 = (static_cast<int64_t>(blockIdx.x) * PG_WARPS + warp) * PG_ROWS; r0 < limit; r0 += stride) {
    float4 v[PG_ROWS][VEC_PER_LANE];
    bool live[PG_ROWS];
#pragma unroll
    for (int j = 0; j < PG_ROWS; ++j) {
      const int64_t row = r0 + j;
      live[j] = row < limit;
      const float4* src = nullptr;
      bool from_pad = false;
      if (live[j] && row < count) {
        const int64_t id = ids[row];
        if (id == pad_id) {
          from_pad = true;
        } else if (id >= 0 && id < V) {
          const int owner = static_cast<int>(min(id / per, static_cast<int64_t>(world - 1)));
          src = reinterpret_cast<const float4*>(shards.base[owner] + (id - owner * per) * K);
        } else if (err_flag && lane == 0) {
          *err_flag = 1;
        }
      }
#pragma unroll
      for (int u = 0; u < VEC_PER_LANE; ++u) {
        const int c4 = lane + 32 * u;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c4 < kv) {
          if (src) t = src[c4];
          else if (from_pad) t = reinterpret_cast<const float4*>(pad_row)[c4];
        }
        v[j][u] = t;
      }
    }
#pragma unroll
    for (int j = 0; j < PG_ROWS; ++j) {
      if (!live[j]) continue;
      const int64_t row = r0 + j;
#pragma unroll
      for (int u = 0; u < VEC_PER_LANE; ++u) {
        const int c4 = lane + 32 * u;
        if (c4 * 4 >= Kp) continue;
        const float4 t = v[j][u];  // zero beyond K (the planes' padding)
        if (out_f32 && c4 < kv) reinterpret_cast<float4*>(out_f32 + row * K)[c4] = t;
        if (planes) {
          uint32_t h0, l0, h1, l1;
          split_bf16x2(t.x, t.y, h0, l0);
          split_bf16x2(t.z, t.w, h1, l1);
          reinterpret_cast<uint2*>(planes + row * Kp)[c4] = make_uint2(h0, h1);
          reinterpret_cast<uint2*>(planes + (static_cast<int64_t>(cap) + row) * Kp)[c4] = make_uint2(l0, l1);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K12 over peer memory: pull the label rows of every rank
// ------------------------------------------------------------------------------------------------------------------
struct PeerMail {
  const float* x[T4R_MAX_PEERS];      // [cap, K] fp32 label rows of rank r (first counts[r] valid)
  const int64_t* y[T4R_MAX_PEERS];    // [cap] labels of rank r
};

template <int VEC_PER_LANE>
__global__ void __launch_bounds__(PG_WARPS * 32)
peer_pull_rows_kernel(PeerMail mail, int world, int rank, const int32_t* __restrict__ counts, int cap, int K, int Kp,
                      float* __restrict__ out_f32, __nv_bfloat16* __restrict__ planes, int64_t* __restrict__ out_labels,
                      int32_t* __restrict__ t_total_out, int32_t* __restrict__ my_start_out) {
  const int lane = lane_id();
  const int warp = warp_id();
  int prefix[T4R_MAX_PEERS + 1];
  prefix[0] = 0;
#pragma unroll
  for (int r = 0; r < T4R_MAX_PEERS; ++r) {
    const int c = r < world ? max(0, min(counts[r], cap)) : 0;
    prefix[r + 1] = prefix[r] + c;
  }
  const int total = prefix[T4R_MAX_PEERS];
  const int64_t cap_g = static_cast<int64_t>(world) * cap;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *t_total_out = total;
    if (my_start_out) *my_start_out = prefix[rank];
  }
  // rows in [total, round_up(total, 256)) are zeroed: the head GEMM's last row tile reads them
  const int64_t fill = min(cap_g, (static_cast<int64_t>(total) + 255) / 256 * 256);
  const int kv = K >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * PG_WARPS * PG_ROWS;
  for (int64_t r0 = (static_cast<int64_t>(blockIdx.x) * PG_WARPS + warp) * PG_ROWS; r0 < fill; r0 += stride) {
    float4 v[PG_ROWS][VEC_PER_LANE];
    int64_t lab[PG_ROWS];
#pragma unroll
    for (int j = 0; j < PG_ROWS; ++j) {
      const int64_t g = r0 + j;
      const float4* src = nullptr;
      lab[j] = 0;
      if (g < total) {
        int r = 0, start = 0;
#pragma unroll
        for (int q = 1; q < T4R_MAX_PEERS; ++q)
          if (q < world && g >= prefix[q]) { r = q; start = prefix[q]; }
        const int64_t local = g - start;
        src = reinterpret_cast<const float4*>(mail.x[r] + local * K);
        if (lane == 0) lab[j] = mail.y[r][local];
      }
#pragma unroll
      for (int u = 0; u < VEC_PER_LANE; ++u) {
        const int c4 = lane + 32 * u;
        v[j][u] = (src && c4 < kv) ? src[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int j = 0; j < PG_ROWS; ++j) {
      const int64_t g = r0 + j;
      if (g >= fill) continue;
      if (lane == 0 && out_labels) out_labels[g] = lab[j];
#pragma unroll
      for (int u = 0; u < VEC_PER_LANE; ++u) {
        const int c4 = lane + 32 * u;
        if (c4 * 4 >= Kp) continue;
        const float4 t = v[j][u];
        if (out_f32 && c4 < kv) reinterpret_cast<float4*>(out_f32 + g * K)[c4] = t;
        if (planes) {
          uint32_t h0, l0, h1, l1;
          split_bf16x2(t.x, t.y, h0, l0);
          split_bf16x2(t.z, t.w, h1, l1);
          reinterpret_cast<uint2*>(planes + g * Kp)[c4] = make_uint2(h0, h1);
          reinterpret_cast<uint2*>(planes + (cap_g + g) * Kp)[c4] = make_uint2(l0, l1);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K12 tail over peer memory: combine every shard's statistics for the rows
// ------------------------------------------------------------------------------------------------------------------
struct PeerStats {
  const float* w[T4R_MAX_PEERS];  // [3, cap_g]: row_lse | row_tgt (label logit, 0 when the label lives elsewhere) | row_rank (int32)
};

__global__ void __launch_bounds__(256)
peer_combine_lse_kernel(PeerStats st, int world, int64_t cap_g, const int32_t* __restrict__ t_total, int with_rank,
                        float* __restrict__ row_loss, int32_t* __restrict__ row_rank) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (row >= cap_g) return;
  const int T = static_cast<int>(min(static_cast<int64_t>(*t_total), cap_g));
  if (row >= T) {
    row_loss[row] = 0.f;
    if (with_rank && row_rank) row_rank[row] = 0;
    return;
  }
  float m = -INFINITY, s = 0.f, tgt = 0.f;
  int cnt = 0;
  for (int w = 0; w < world; ++w) {
    const float lse = st.w[w][row];
    tgt += st.w[w][cap_g + row];
    if (with_rank) cnt += reinterpret_cast<const int32_t*>(st.w[w])[2 * cap_g + row];
    const float mn = fmaxf(m, lse);
    if (mn > -INFINITY) {
      s = s * expf(m - mn) + expf(lse - mn);
      m = mn;
    }
  }
  row_loss[row] = (m + logf(s)) - tgt;
  if (with_rank && row_rank) row_rank[row] = cnt;
}

int launch_mean_rows(const float* rows, int cap, const int32_t* t_dev, float* out, cudaStream_t s);

}  // namespace t4r

using namespace t4r;

extern "C" int t4r_peer_export(const void* dev_ptr, void* handle_out, int64_t* offset_out) {
  T4R_REQUIRE(dev_ptr && handle_out && offset_out, "peer_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == T4R_PEER_HANDLE_BYTES, "handle size");
  void* base = nullptr;
  size_t size = 0;
  T4R_TRY(address_range(dev_ptr, &base, &size));
  cudaIpcMemHandle_t h;
  T4R_CUDA(cudaIpcGetMemHandle(&h, base));
  memcpy(handle_out, &h, sizeof(h));
  *offset_out = static_cast<const char*>(dev_ptr) - static_cast<const char*>(base);
  return 0;
}

extern "C" int t4r_peer_open(const void* handle, int64_t offset, void** mapped_out) {
  T4R_REQUIRE(handle && mapped_out && offset >= 0, "peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* base = nullptr;
  T4R_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  *mapped_out = static_cast<char*>(base) + offset;
  return 0;
}

extern "C" int t4r_peer_close(void* mapped, int64_t offset) {
  T4R_REQUIRE(mapped && offset >= 0, "peer_close: bad arguments");
  T4R_CUDA(cudaIpcCloseMemHandle(static_cast<char*>(mapped) - offset));
  return 0;
}

static int gather_grid(int64_t rows) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t need = (rows + PG_WARPS * PG_ROWS - 1) / (PG_WARPS * PG_ROWS);
  return static_cast<int>(need < 1 ? 1 : (need > static_cast<int64_t>(sms) * 8 ? sms * 8 : need));
}

extern "C" int t4r_peer_gather_rows(const t4r_peer_ptrs* shards, int64_t V, int64_t rows_per_shard, int K,
                                    const int64_t* ids, const int32_t* count_dev, int cap, int64_t pad_id,
                                    float* out_f32, void* out_planes, int32_t* err_flag, void* stream) {
  T4R_REQUIRE(shards && ids && cap > 0 && (out_f32 || out_planes), "peer_gather_rows: bad arguments");
  T4R_REQUIRE(shards->world >= 1 && shards->world <= T4R_MAX_PEERS, "peer_gather_rows: 1..%d ranks", T4R_MAX_PEERS);
  T4R_REQUIRE(K > 0 && K % 4 == 0 && K <= PG_MAX_K, "peer_gather_rows: row width must be a multiple of 4, <= %d (got %d)",
              PG_MAX_K, K);
  T4R_REQUIRE(V > 0 && rows_per_shard > 0 && rows_per_shard * shards->world >= V,
              "peer_gather_rows: %lld rows per shard x %d ranks do not cover V = %lld", (long long)rows_per_shard,
              shards->world, (long long)V);
  PeerBases pb;
  for (int r = 0; r < T4R_MAX_PEERS; ++r) {
    pb.base[r] = r < shards->world ? static_cast<const float*>(shards->base[r]) : nullptr;
    T4R_REQUIRE(r >= shards->world || pb.base[r] || static_cast<int64_t>(r) * rows_per_shard >= V,
                "peer_gather_rows: shard %d is not mapped", r);
  }
  const int Kp = t4r_round_up64(K);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = gather_grid(cap);
  __nv_bfloat16* pl = static_cast<__nv_bfloat16*>(out_planes);
#define T4R_PG(VPL)                                                                                                   \
  peer_gather_rows_kernel<VPL><<<grid, PG_WARPS * 32, 0, s>>>(pb, shards->world, V, rows_per_shard, K, Kp, ids, count_dev, \
                                                             cap, pad_id, out_f32, pl, err_flag)
  if (Kp <= 128) T4R_PG(1);
  else if (Kp <= 256) T4R_PG(2);
  else if (Kp <= 512) T4R_PG(4);
  else T4R_PG(8);
#undef T4R_PG
  T4R_LAUNCH_CHECK("peer_gather_rows_kernel");
  return 0;
}

extern "C" int t4r_peer_pull_rows(const t4r_peer_ptrs* mail_x, const t4r_peer_ptrs* mail_y, const int32_t* counts,
                                  int cap, int K, float* out_f32, void* out_planes, int64_t* out_labels,
                                  int32_t* t_total, int32_t* my_start, void* stream) {
  T4R_REQUIRE(mail_x && mail_y && counts && t_total && cap > 0 && (out_f32 || out_planes), "peer_pull_rows: bad arguments");
  T4R_REQUIRE(mail_x->world >= 1 && mail_x->world <= T4R_MAX_PEERS && mail_y->world == mail_x->world &&
                  mail_x->rank >= 0 && mail_x->rank < mail_x->world,
              "peer_pull_rows: 1..%d ranks, same group for rows and labels", T4R_MAX_PEERS);
  T4R_REQUIRE(K > 0 && K % 4 == 0 && K <= PG_MAX_K, "peer_pull_rows: row width must be a multiple of 4, <= %d (got %d)",
              PG_MAX_K, K);
  const int world = mail_x->world;
  T4R_REQUIRE(static_cast<int64_t>(world) * cap < (1ll << 31), "peer_pull_rows: world x capacity overflows int32");
  PeerMail pm;
  for (int r = 0; r < T4R_MAX_PEERS; ++r) {
    pm.x[r] = r < world ? static_cast<const float*>(mail_x->base[r]) : nullptr;
    pm.y[r] = r < world ? static_cast<const int64_t*>(mail_y->base[r]) : nullptr;
    T4R_REQUIRE(r >= world || (pm.x[r] && pm.y[r]), "peer_pull_rows: window of rank %d is not mapped", r);
  }
  const int Kp = t4r_round_up64(K);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = gather_grid(static_cast<int64_t>(world) * cap);
  __nv_bfloat16* pl = static_cast<__nv_bfloat16*>(out_planes);
#define T4R_PP(VPL)                                                                                                  \
  peer_pull_rows_kernel<VPL><<<grid, PG_WARPS * 32, 0, s>>>(pm, world, mail_x->rank, counts, cap, K, Kp, out_f32, pl, \
                                                           out_labels, t_total, my_start)
  if (Kp <= 128) T4R_PP(1);
  else if (Kp <= 256) T4R_PP(2);
  else if (Kp <= 512) T4R_PP(4);
  else T4R_PP(8);
#undef T4R_PP
  T4R_LAUNCH_CHECK("peer_pull_rows_kernel");
  return 0;
}

extern "C" int t4r_peer_combine_lse(const t4r_peer_ptrs* stats, int64_t cap_g, const int32_t* t_total, int with_rank,
                                    float* row_loss, int32_t* row_rank, float* loss, void* stream) {
  T4R_REQUIRE(stats && t_total && row_loss && cap_g > 0 && cap_g < (1ll << 31), "peer_combine_lse: bad arguments");
  T4R_REQUIRE(stats->world >= 1 && stats->world <= T4R_MAX_PEERS, "peer_combine_lse: 1..%d ranks", T4R_MAX_PEERS);
  T4R_REQUIRE(!with_rank || row_rank, "peer_combine_lse: with_rank needs row_rank");
  PeerStats ps;
  for (int r = 0; r < T4R_MAX_PEERS; ++r) {
    ps.w[r] = r < stats->world ? static_cast<const float*>(stats->base[r]) : nullptr;
    T4R_REQUIRE(r >= stats->world || ps.w[r], "peer_combine_lse: window of rank %d is not mapped", r);
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  peer_combine_lse_kernel<<<static_cast<unsigned>((cap_g + 255) / 256), 256, 0, s>>>(ps, stats->world, cap_g, t_total,
                                                                                    with_rank, row_loss, row_rank);
  T4R_LAUNCH_CHECK("peer_combine_lse_kernel");
  if (loss) T4R_TRY(launch_mean_rows(row_loss, static_cast<int>(cap_g), t_total, loss, s));
  return 0;
}
