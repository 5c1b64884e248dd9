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


// PU21 encoding (utils.py:157-202): PU(Y) = p6 (((p0 + p1 Y^p3) / (1 + p2 Y^p3))^p4 - p5), Y clipped to [L_min, L_max]
struct PuParams {
  float p[7];
  float L_min, L_max;
};
__device__ __forceinline__ float pu_encode(float Y, const PuParams& q) {
  Y = fminf(fmaxf(Y, q.L_min), q.L_max);
  const float yp = fast_exp2(q.p[3] * fast_log2(Y));
  const float r = (q.p[0] + q.p[1] * yp) / (1.0f + q.p[2] * yp);
  return fast_exp2(q.p[4] * fast_log2(r));  // PU21 = p6 * (this - p5): the callers work on differences, where p5 cancels
}

}  // namespace fvvdp
