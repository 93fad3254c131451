// t4r_mixed_pack.cuh -- operand format of the 2-unit product (nprod = 2), shared by the device
// kernel and a host-side debug entry (same code path, so the CPU tests pin layout and rounding).
//
// A fp32 row x[0..K) is stored as
//   scale     s = 2^e with max|x| * s in [2^13, 2^14)                 (inv_scale[row] = 1/s, a power of two)
//   plane 0   fp16   h[k]  = fp16(x[k] * s)                            [rows, Kp] halves
//   plane 1   e4m3   per 64-wide K block kb, 128 bytes per row:
//               bytes [  0, 64)  hi8[k] = e4m3(h[k]              * 2^-6)   |.| <= 256
//               bytes [ 64,128)  lo8[k] = e4m3((x[k] * s - h[k]) * 2^+6)   |.| <= 256
// so that  x_a . x_b * s_a * s_b = h_a . h_b  (kind::f16)  +  lo8_a . hi8_b + hi8_a . lo8_b  (kind::f8f6f4)
// up to the dropped lo*lo term (2^-22) and the e4m3 rounding of the cross terms (2^-4 of a 2^-11 term):
// ~2^-15 relative per product, against 2^-16.5 for the three-product bf16 split, at 2/3 of its tensor time
// (an fp8 MMA pass costs half an fp16 pass).  tools/precision_study.py has the emulated error tables.
#pragma once
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <stdint.h>

namespace t4r {

constexpr int kMixedTargetExp = 13;   // scaled row max in [2^13, 2^14)
constexpr float kMixedHi8Scale = 1.0f / 64.0f;
constexpr float kMixedLo8Scale = 64.0f;

// scale = 2^(13 - floor(log2(maxabs))), clamped to the normal fp32 exponent range; 1 for an all-zero row
__host__ __device__ inline void mixed_row_scale(float maxabs, float& scale, float& inv_scale) {
  union { float f; uint32_t u; } m, s, i;
  m.f = maxabs;
  const int biased = static_cast<int>((m.u >> 23) & 0xffu);
  if (maxabs == 0.f || biased == 0xff) {  // zero row (or inf/nan: left unscaled, the result is inf/nan anyway)
    scale = 1.f;
    inv_scale = 1.f;
    return;
  }
  int e = biased - 127;                // 2^e <= maxabs < 2^(e+1)  (subnormals: e = -127, clamped below)
  int sh = kMixedTargetExp - e;
  if (sh > 126) sh = 126;
  if (sh < -126) sh = -126;
  s.u = static_cast<uint32_t>(sh + 127) << 23;
  i.u = static_cast<uint32_t>(127 - sh) << 23;
  scale = s.f;
  inv_scale = i.f;
}

struct MixedPair {
  uint32_t h16x2;   // two fp16 (element 0 in the low half)
  uint16_t hi8x2;   // two e4m3 (element 0 in the low byte)
  uint16_t lo8x2;
};

// x0, x1 are already multiplied by the row scale
__host__ __device__ inline MixedPair mixed_pack_pair(float x0, float x1) {
  const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
  const float f0 = __half2float(h0), f1 = __half2float(h1);
  MixedPair p;
  p.h16x2 = static_cast<uint32_t>(__half_as_ushort(h0)) | (static_cast<uint32_t>(__half_as_ushort(h1)) << 16);
  p.hi8x2 = __nv_cvt_float2_to_fp8x2(make_float2(f0 * kMixedHi8Scale, f1 * kMixedHi8Scale), __NV_SATFINITE, __NV_E4M3);
  p.lo8x2 = __nv_cvt_float2_to_fp8x2(make_float2((x0 - f0) * kMixedLo8Scale, (x1 - f1) * kMixedLo8Scale),
                                     __NV_SATFINITE, __NV_E4M3);
  return p;
}

// byte offset, inside plane 1 of one row, of the e4m3 pair that belongs to elements (k, k+1), k even
__host__ __device__ inline int mixed_hi8_offset(int k) { return (k >> 6) * 128 + (k & 63); }
__host__ __device__ inline int mixed_lo8_offset(int k) { return (k >> 6) * 128 + 64 + (k & 63); }

}  // namespace t4r
