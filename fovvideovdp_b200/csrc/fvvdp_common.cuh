// Helpers shared by the kernels of the B200-native FovVideoVDP core (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/fvvdp_b200.h"

namespace fvvdp {

// ------------------------------------------------------------------------------------------------
// small math helpers (MUFU based: lg2.approx / ex2.approx / rcp.approx)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_log2(float x) {  // one MUFU.LG2 (no denormal rescaling: tiny inputs flush to -inf)
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_pow(float x, float p) { return fast_exp2(p * fast_log2(x)); }
__device__ __forceinline__ float fast_rcp(float x) { return __frcp_rn(x); }

__device__ __forceinline__ float load_sample(const void* base, long long off, int dtype) {
  if (dtype == FVVDP_B200_F32) return __ldg(reinterpret_cast<const float*>(base) + off);
  if (dtype == FVVDP_B200_U8) return static_cast<float>(__ldg(reinterpret_cast<const uint8_t*>(base) + off)) / 255.0f;
  int v = static_cast<int>(__ldg(reinterpret_cast<const int16_t*>(base) + off)) & 0xFFFF;  // video_source.py:186-196
  return static_cast<float>(v) / 65535.0f;
}


struct CsfAxes {           // device pointers, 32 entries each
  const float* x[3];       // 0: rho_log, 1: Y_log, 2: ecc_sqrt
  const float* inv[3];     // 1 / (x[j] - x[j-1] + 1e-6), inv[0] unused
  float x0[3], inv_dx[3];  // uniform-grid first guess
  float lo[3], hi[3];      // clamp range in linear units
};


}  // namespace fvvdp
