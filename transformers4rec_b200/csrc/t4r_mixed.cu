// t4r_mixed.cu -- packing kernel of the 2-unit product operands (see t4r_mixed_pack.cuh) and its host twin.
#include <math.h>
#include <vector>

#include "t4r_common.cuh"
#include "t4r_internal.h"
#include "t4r_mixed_pack.cuh"

namespace t4r {

// one warp per row: pass 1 = max |x| (the row stays in L1/L2 for pass 2), pass 2 = scale, split, pack.
// Per 64-wide K block a lane owns elements (2 lane, 2 lane + 1): 128 B of fp16 and 2 x 64 B of e4m3 per warp store.
__global__ void __launch_bounds__(256)
split_planes_mixed_kernel(const float* __restrict__ x, int64_t rows, int K, int64_t ld, int Kp,
                          uint16_t* __restrict__ planes, float* __restrict__ inv_scale,
                          const int32_t* __restrict__ count_dev) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + warp_id();
  if (row >= rows) return;
  // with a device-side row count only the rows a consumer's last 256-row tile can touch are packed
  if (count_dev && row >= (static_cast<int64_t>(*count_dev) + 255) / 256 * 256) return;
  const int lane = lane_id();
  const float* src = x + row * ld;
  float m = 0.f;
  for (int c = lane; c < K; c += 32) m = fmaxf(m, fabsf(src[c]));
  m = warp_max(m);
  float scale, inv;
  mixed_row_scale(m, scale, inv);
  if (lane == 0) inv_scale[row] = inv;
  uint32_t* p0 = reinterpret_cast<uint32_t*>(planes + row * Kp);
  uint8_t* p1 = reinterpret_cast<uint8_t*>(planes + (rows + row) * Kp);
  for (int kb = 0; kb < Kp / 64; ++kb) {
    const int k = kb * 64 + 2 * lane;
    const float x0 = (k < K) ? src[k] * scale : 0.f;
    const float x1 = (k + 1 < K) ? src[k + 1] * scale : 0.f;
    const MixedPair q = mixed_pack_pair(x0, x1);
    p0[k >> 1] = q.h16x2;
    *reinterpret_cast<uint16_t*>(p1 + mixed_hi8_offset(k)) = q.hi8x2;
    *reinterpret_cast<uint16_t*>(p1 + mixed_lo8_offset(k)) = q.lo8x2;
  }
}

}  // namespace t4r

extern "C" int t4r_split_planes_mixed_n(const float* x, int64_t rows, int K, int ld, const int32_t* count_dev,
                                        void* out_planes, float* out_inv_scale, void* stream) {
  using namespace t4r;
  T4R_REQUIRE(x && out_planes && out_inv_scale && rows > 0 && K > 0 && ld >= K, "split_planes_mixed: bad arguments");
  const int Kp = t4r_round_up64(K);
  const int64_t blocks = (rows + 7) / 8;
  split_planes_mixed_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, rows, K, ld, Kp, static_cast<uint16_t*>(out_planes), out_inv_scale, count_dev);
  T4R_LAUNCH_CHECK("split_planes_mixed_kernel");
  return 0;
}
extern "C" int t4r_split_planes_mixed(const float* x, int64_t rows, int K, int ld, void* out_planes, float* out_inv_scale,
                                      void* stream) {
  return t4r_split_planes_mixed_n(x, rows, K, ld, nullptr, out_planes, out_inv_scale, stream);
}

// Host twin of the kernel above (HOST pointers; no CUDA call): the same mixed_row_scale / mixed_pack_pair code
// compiled for the CPU, so the layout and the roundings can be pinned without a GPU (tests/test_abi_and_host.py).
extern "C" int t4r_debug_split_planes_mixed_host(const float* x, int64_t rows, int K, int ld, void* out_planes,
                                                 float* out_inv_scale) {
  using namespace t4r;
  T4R_REQUIRE(x && out_planes && out_inv_scale && rows > 0 && K > 0 && ld >= K, "split_planes_mixed_host: bad arguments");
  const int Kp = t4r_round_up64(K);
  uint16_t* planes = static_cast<uint16_t*>(out_planes);
  for (int64_t row = 0; row < rows; ++row) {
    const float* src = x + row * ld;
    float m = 0.f;
    for (int c = 0; c < K; ++c) m = fmaxf(m, fabsf(src[c]));
    float scale, inv;
    mixed_row_scale(m, scale, inv);
    out_inv_scale[row] = inv;
    uint32_t* p0 = reinterpret_cast<uint32_t*>(planes + row * Kp);
    uint8_t* p1 = reinterpret_cast<uint8_t*>(planes + (rows + row) * Kp);
    for (int k = 0; k < Kp; k += 2) {
      const float x0 = (k < K) ? src[k] * scale : 0.f;
      const float x1 = (k + 1 < K) ? src[k + 1] * scale : 0.f;
      const MixedPair q = mixed_pack_pair(x0, x1);
      p0[k >> 1] = q.h16x2;
      *reinterpret_cast<uint16_t*>(p1 + mixed_hi8_offset(k)) = q.hi8x2;
      *reinterpret_cast<uint16_t*>(p1 + mixed_lo8_offset(k)) = q.lo8x2;
    }
  }
  return 0;
}
