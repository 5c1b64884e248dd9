// Dispatch over the nine fused-kernel translation units; the luminance front-end kernel lives here.
#include "fvvdp_fused.cuh"
#include "fvvdp_fused_launch.h"
#include "fvvdp_ws_geometry.h"

namespace fvvdp {
namespace fused {

#define DECL(k, v)                                                                                          \
  cudaError_t launch_band_##k##_##v(bool foveated, bool extra, const BandParams& p, dim3 grid, cudaStream_t st); \
  cudaError_t configure_band_##k##_##v();
#define DECL3(k) DECL(k, 0) DECL(k, 1) DECL(k, 2)
DECL3(0) DECL3(2) DECL3(3) DECL(2, 3)
#undef DECL3
#undef DECL

cudaError_t launch_band(int kind, int mode, bool foveated, bool extra, const BandParams& p, dim3 grid, cudaStream_t st) {
  if (mode == 3) return launch_band_2_3(foveated, extra, p, grid, st);  // two planes per slot: pyramid-layout planes only
  switch (kind * 3 + mode) {
#define CASE(k, v) case (k) * 3 + (v): return launch_band_##k##_##v(foveated, extra, p, grid, st);
    CASE(0, 0) CASE(0, 1) CASE(0, 2) CASE(2, 0) CASE(2, 1) CASE(2, 2) CASE(3, 0) CASE(3, 1) CASE(3, 2)
#undef CASE
  }
  return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------------------------------------ luminance front end
// Level-0 input that is not a contiguous single-channel float plane (uint8 / uint16, RGB, channel-last, strided views):
// one pass converts every window slot to luminance (sample -> [0,1] -> display EOTF -> RGB2Y; video_source.py:180-208,
// fvvdp_display_model.py:147-165) and writes it in the layout of the pyramid planes, [slot][row][2 * column + stream], so
// that level 0 runs through the same TMA-staged kernel as the coarser levels.
//   DT / C: sample type and channel count; VEC: rows are contiguous and aligned -> 4 pixels per thread, vector loads.
//   uint8 samples go through a 256-entry table of the EOTF built by the CTA (the same arithmetic, evaluated once per code).
template <int DT>
__device__ __forceinline__ float sample_at(const void* base, long long off) {
  if (DT == FVVDP_B200_F32) return __ldg(reinterpret_cast<const float*>(base) + off);
  if (DT == FVVDP_B200_U8) return (float)__ldg(reinterpret_cast<const uint8_t*>(base) + off);  // table index
  return (float)((int)__ldg(reinterpret_cast<const int16_t*>(base) + off) & 0xFFFF) / 65535.0f;   // video_source.py:186-196
}
template <int DT>
__device__ __forceinline__ void samples4_at(const void* base, long long off, float (&v)[4]) {
  if (DT == FVVDP_B200_F32) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off));
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  } else if (DT == FVVDP_B200_U8) {
    const uchar4 q = __ldg(reinterpret_cast<const uchar4*>(reinterpret_cast<const uint8_t*>(base) + off));
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  } else {
    const ushort4 q = __ldg(reinterpret_cast<const ushort4*>(reinterpret_cast<const uint16_t*>(base) + off));
    v[0] = (float)q.x / 65535.0f; v[1] = (float)q.y / 65535.0f; v[2] = (float)q.z / 65535.0f; v[3] = (float)q.w / 65535.0f;
  }
}

template <int DT, int C, bool VEC, int KIND>
__device__ __forceinline__ void luminance_body(const BandParams& p, float* __restrict__ out, long long slot_stride, int pitch, const float* sLut) {
  constexpr int PX = VEC ? 4 : 1;
  const int x = (blockIdx.x * 32 + (threadIdx.x & 31)) * PX, y = blockIdx.y * 8 + (threadIdx.x >> 5), slot = blockIdx.z;
  float vmin = 0.0f, vmax = 1.0f;
  if (x < p.w && y < p.h) {
    float lum[2][PX];
#pragma unroll
    for (int st = 0; st < 2; ++st) {
      const void* base = p.slot[st][slot];
      const long long off = (long long)y * p.sH + (long long)x * p.sW;
      float acc[PX];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float v[PX];
        if (VEC) {
          float v4[4];
          samples4_at<DT>(base, off + c * p.sC, v4);
#pragma unroll
          for (int i = 0; i < PX; ++i) v[i] = v4[i];
        } else {
          v[0] = sample_at<DT>(base, off + c * p.sC);
        }
#pragma unroll
        for (int i = 0; i < PX; ++i) {
          float L;
          if (DT == FVVDP_B200_U8) {
            L = sLut[(int)v[i]];
          } else {
            if (DT == FVVDP_B200_F32) { vmin = fminf(vmin, v[i]); vmax = fmaxf(vmax, v[i]); }
            L = eotf_k<KIND>(v[i], p);
          }
          acc[i] = C == 1 ? L : (c == 0 ? L * p.rgb2y[0] : acc[i] + L * p.rgb2y[c]);  // L0 w0 + L1 w1 + L2 w2, left to right
        }
      }
#pragma unroll
      for (int i = 0; i < PX; ++i) lum[st][i] = acc[i];
    }
    float* o = out + slot * slot_stride + (long long)y * pitch + 2 * x;
    if (VEC) {
      *reinterpret_cast<float4*>(o) = make_float4(lum[0][0], lum[1][0], lum[0][PX > 1 ? 1 : 0], lum[1][PX > 1 ? 1 : 0]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(lum[0][PX > 2 ? 2 : 0], lum[1][PX > 2 ? 2 : 0], lum[0][PX > 3 ? 3 : 0], lum[1][PX > 3 ? 3 : 0]);
    } else {
      *reinterpret_cast<float2*>(o) = make_float2(lum[0][0], lum[1][0]);
    }
  }
  if (DT == FVVDP_B200_F32 && eotf_checks_range(KIND) && p.flags && __any_sync(0xffffffffu, vmin < 0.0f || vmax > 1.0f) && (threadIdx.x & 31) == 0)
    atomicOr(p.flags, 1u);
}

template <int DT, int C, bool VEC>
__global__ void __launch_bounds__(256) luminance_kernel(const __grid_constant__ BandParams p, float* __restrict__ out, long long slot_stride,
                                                        int pitch) {
  __shared__ float sLut[256];
#define FVVDP_LUM(K)                                                               \
  case K:                                                                          \
    if (DT == FVVDP_B200_U8) {                                                     \
      sLut[threadIdx.x] = eotf_k<K>((float)threadIdx.x / 255.0f, p);               \
      __syncthreads();                                                             \
    }                                                                              \
    luminance_body<DT, C, VEC, K>(p, out, slot_stride, pitch, sLut);               \
    break;
  switch (p.eotf) {  // uniform
    FVVDP_LUM(FVVDP_B200_EOTF_NONE) FVVDP_LUM(FVVDP_B200_EOTF_SRGB) FVVDP_LUM(FVVDP_B200_EOTF_GAMMA) FVVDP_LUM(FVVDP_B200_EOTF_PQ)
    FVVDP_LUM(FVVDP_B200_EOTF_LINEAR) FVVDP_LUM(FVVDP_B200_EOTF_ABSOLUTE)
  }
#undef FVVDP_LUM
}

cudaError_t launch_luminance(const BandParams& p, float* out, long long slot_stride, int pitch, int n_slots, bool rows_vectorisable, cudaStream_t st) {
  const bool vec = rows_vectorisable && p.w % 4 == 0;
  dim3 grid(((vec ? p.w / 4 : p.w) + 31) / 32, (p.h + 7) / 8, n_slots);
#define FVVDP_LAUNCH(DT_, NC_)                                                                                  \
  if (p.dtype == DT_ && p.C == NC_) {                                                                           \
    if (vec) luminance_kernel<DT_, NC_, true><<<grid, 256, 0, st>>>(p, out, slot_stride, pitch);                \
    else luminance_kernel<DT_, NC_, false><<<grid, 256, 0, st>>>(p, out, slot_stride, pitch);                   \
    return cudaGetLastError();                                                                                  \
  }
  FVVDP_LAUNCH(FVVDP_B200_F32, 1) FVVDP_LAUNCH(FVVDP_B200_F32, 3) FVVDP_LAUNCH(FVVDP_B200_U8, 1) FVVDP_LAUNCH(FVVDP_B200_U8, 3)
  FVVDP_LAUNCH(FVVDP_B200_U16, 1) FVVDP_LAUNCH(FVVDP_B200_U16, 3)
#undef FVVDP_LAUNCH
  return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------------------------------------ PU21-PSNR, whole frame blocks
// pupsnr.py:52-79 for a block of frames in ONE launch: sample -> [0,1] -> display EOTF -> RGB2Y (the arithmetic of the luminance
// front end above) -> PU21 encoding of both streams -> squared difference, summed per frame in double precision.  No luminance
// plane is written or read back: the frames are read once, in their own dtype and layout.
template <int DT, int C, bool VEC, int KIND>
__device__ __forceinline__ float pu_body(const BandParams& p, const PuParams& q, const float* sLut) {
  constexpr int PX = VEC ? 4 : 1;
  const int x = (blockIdx.x * 32 + (threadIdx.x & 31)) * PX, y = blockIdx.y * 8 + (threadIdx.x >> 5), slot = blockIdx.z;
  float s = 0.0f;
  if (x < p.w && y < p.h) {
    float lum[2][PX];
#pragma unroll
    for (int st = 0; st < 2; ++st) {
      const void* base = p.slot[st][slot];
      const long long off = (long long)y * p.sH + (long long)x * p.sW;
      float acc[PX];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float v[PX];
        if (VEC) {
          float v4[4];
          samples4_at<DT>(base, off + c * p.sC, v4);
#pragma unroll
          for (int i = 0; i < PX; ++i) v[i] = v4[i];
        } else {
          v[0] = sample_at<DT>(base, off + c * p.sC);
        }
#pragma unroll
        for (int i = 0; i < PX; ++i) {
          const float L = DT == FVVDP_B200_U8 ? sLut[(int)v[i]] : eotf_k<KIND>(v[i], p);
          acc[i] = C == 1 ? L : (c == 0 ? L * p.rgb2y[0] : acc[i] + L * p.rgb2y[c]);
        }
      }
#pragma unroll
      for (int i = 0; i < PX; ++i) lum[st][i] = acc[i];
    }
#pragma unroll
    for (int i = 0; i < PX; ++i) {
      const float d = q.p[6] * (pu_encode(lum[0][i], q) - pu_encode(lum[1][i], q));
      s = fmaf(d, d, s);
    }
  }
  return s;
}

template <int DT, int C, bool VEC>
__global__ void __launch_bounds__(256) pu_frames_kernel(const __grid_constant__ BandParams p, const PuParams q, double* __restrict__ out) {
  __shared__ float sLut[256];
  __shared__ double sSum[8];
  float s = 0.0f;
#define FVVDP_PU(K)                                                                \
  case K:                                                                          \
    if (DT == FVVDP_B200_U8) {                                                     \
      sLut[threadIdx.x] = eotf_k<K>((float)threadIdx.x / 255.0f, p);               \
      __syncthreads();                                                             \
    }                                                                              \
    s = pu_body<DT, C, VEC, K>(p, q, sLut);                                        \
    break;
  switch (p.eotf) {  // uniform
    FVVDP_PU(FVVDP_B200_EOTF_NONE) FVVDP_PU(FVVDP_B200_EOTF_SRGB) FVVDP_PU(FVVDP_B200_EOTF_GAMMA) FVVDP_PU(FVVDP_B200_EOTF_PQ)
    FVVDP_PU(FVVDP_B200_EOTF_LINEAR) FVVDP_PU(FVVDP_B200_EOTF_ABSOLUTE)
  }
#undef FVVDP_PU
  double v = (double)s;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sSum[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sSum[k];
    atomicAdd(out + blockIdx.z, t);
  }
}

cudaError_t launch_pu_frames(const BandParams& p, const void* pu_params, double* out, int n_frames, bool rows_vectorisable, cudaStream_t st) {
  const bool vec = rows_vectorisable && p.w % 4 == 0;
  dim3 grid(((vec ? p.w / 4 : p.w) + 31) / 32, (p.h + 7) / 8, n_frames);
  const PuParams& q = *reinterpret_cast<const PuParams*>(pu_params);
#define FVVDP_LAUNCH(DT_, NC_)                                                                  \
  if (p.dtype == DT_ && p.C == NC_) {                                                           \
    if (vec) pu_frames_kernel<DT_, NC_, true><<<grid, 256, 0, st>>>(p, q, out);                 \
    else pu_frames_kernel<DT_, NC_, false><<<grid, 256, 0, st>>>(p, q, out);                    \
    return cudaGetLastError();                                                                  \
  }
  FVVDP_LAUNCH(FVVDP_B200_F32, 1) FVVDP_LAUNCH(FVVDP_B200_F32, 3) FVVDP_LAUNCH(FVVDP_B200_U8, 1) FVVDP_LAUNCH(FVVDP_B200_U8, 3)
  FVVDP_LAUNCH(FVVDP_B200_U16, 1) FVVDP_LAUNCH(FVVDP_B200_U16, 3)
#undef FVVDP_LAUNCH
  return cudaErrorInvalidValue;
}

cudaError_t configure_band_kernels() {
  cudaError_t e;
#define CONF(k, v) if ((e = configure_band_##k##_##v()) != cudaSuccess) return e;
  CONF(0, 0) CONF(0, 1) CONF(0, 2) CONF(2, 0) CONF(2, 1) CONF(2, 2) CONF(3, 0) CONF(3, 1) CONF(3, 2) CONF(2, 3)
#undef CONF
  return cudaSuccess;
}

}  // namespace fused

// ------------------------------------------------------------------------------------------------ warp-specialised kernels
#define FVVDP_WS_DISPATCH(NS)                                                                                             \
  namespace NS {                                                                                                           \
  cudaError_t launch_band_ws_2(bool foveated, const fused::BandParams& p, dim3 grid, cudaStream_t st);                    \
  cudaError_t launch_band_ws_3(bool foveated, const fused::BandParams& p, dim3 grid, cudaStream_t st);                    \
  cudaError_t configure_band_ws_2();                                                                                       \
  cudaError_t configure_band_ws_3();                                                                                       \
  cudaError_t launch_band_ws(int input_kind, bool foveated, const fused::BandParams& p, dim3 grid, cudaStream_t st) {     \
    if (input_kind == fused::IN_PYRAMID_TMA) return launch_band_ws_2(foveated, p, grid, st);                              \
    if (input_kind == fused::IN_LEVEL0_TMA) return launch_band_ws_3(foveated, p, grid, st);                               \
    return cudaErrorInvalidValue;                                                                                          \
  }                                                                                                                        \
  cudaError_t configure_band_ws_kernels() {                                                                                \
    cudaError_t e = configure_band_ws_2();                                                                                 \
    return e != cudaSuccess ? e : configure_band_ws_3();                                                                   \
  }                                                                                                                        \
  }
FVVDP_WS_DISPATCH(ws)
FVVDP_WS_DISPATCH(ws16)
#undef FVVDP_WS_DISPATCH

}  // namespace fvvdp
