// Encoding of the 16-bit operand planes (x = hi + lo).
//
// The plane buffers are typed __nv_bfloat16 / __nv_bfloat162 throughout the library, but only as 16-bit
// CONTAINERS: with VXB_PLANES_FP16 (default) they hold IEEE fp16 bit patterns.  fp16 pairs carry 22 significant
// bits (11 + 11) against 16 for bf16 pairs at the same tensor-core rate (tcgen05 kind::f16 takes either format),
// which moves the split-precision error floor from ~2^-17 to ~2^-23 relative for values of ordinary magnitude
// (measured on the Q-value gates: 1e-4..1e-3 with bf16 pairs, see tools/report_errors.py).  Range notes:
//   * conversions saturate at +-65504 (no infinities); activations of this network are O(1..100);
//   * below 2^-14 the lo plane becomes subnormal: absolute error floor 2^-25 per element;
//   * the attention probabilities are stored scaled by 2^14 (VXB_P_EXP_BIAS) so that small weights keep their
//     precision; the row sums carry the same factor and it cancels in the normalisation.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#ifndef VXB_PLANES_FP16
#define VXB_PLANES_FP16 1
#endif

namespace vxb {

#if VXB_PLANES_FP16
constexpr float VXB_P_EXP_BIAS = 14.f;
constexpr uint32_t VXB_IDESC_AB_FORMAT = 0u;                       // a_format = b_format = F16
__device__ __forceinline__ __nv_bfloat16 pl_from_float(float f) {
  unsigned short u;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(u) : "f"(f));
  return __ushort_as_bfloat16(u);
}
__device__ __forceinline__ float pl_to_float(__nv_bfloat16 p) {
  return __half2float(__ushort_as_half(__bfloat16_as_ushort(p)));
}
__device__ __forceinline__ __nv_bfloat162 pl2_from_floats(float a, float b) {
  uint32_t u;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(b), "f"(a));   // first source -> upper half
  return *reinterpret_cast<__nv_bfloat162*>(&u);
}
__device__ __forceinline__ float2 pl2_to_float2(__nv_bfloat162 p) {
  return __half22float2(*reinterpret_cast<__half2*>(&p));
}
#else
constexpr float VXB_P_EXP_BIAS = 0.f;
constexpr uint32_t VXB_IDESC_AB_FORMAT = (1u << 7) | (1u << 10);   // a_format = b_format = BF16
__device__ __forceinline__ __nv_bfloat16 pl_from_float(float f) { return __float2bfloat16_rn(f); }
__device__ __forceinline__ float pl_to_float(__nv_bfloat16 p) { return __bfloat162float(p); }
__device__ __forceinline__ __nv_bfloat162 pl2_from_floats(float a, float b) { return __floats2bfloat162_rn(a, b); }
__device__ __forceinline__ float2 pl2_to_float2(__nv_bfloat162 p) { return __bfloat1622float2(p); }
#endif

// c8 plane of the f16 + fp8-corrected convolution (conv_f8c.cuh): four consecutive channels -> 4 bytes of e4m3(scale * x)
__device__ __forceinline__ uint32_t pl_e4m3x4(float x0, float x1, float x2, float x3, float scale) {
  unsigned short a, b;
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(a) : "f"(x1 * scale), "f"(x0 * scale));   // first source -> upper byte
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(b) : "f"(x3 * scale), "f"(x2 * scale));
  return (uint32_t)a | ((uint32_t)b << 16);
}

}  // namespace vxb
