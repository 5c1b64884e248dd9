// Tile geometry and launch interface of the warp-specialised band kernels (fvvdp_ws.cuh), shared with the host code.
// Two builds of the same kernel source: `ws` for temporal windows of up to 8 taps (32x64 tiles, one 2x2 quad per consumer
// thread, 7-position rings) and `ws16` for up to 16 taps (16x64 tiles, one row of a quad per thread, 15-position rings).
#pragma once
#include <cuda_runtime.h>

namespace fvvdp {
namespace fused {
struct BandParams;
}
#define FVVDP_WS_DECLARE(NS, TH_, RP_)                                                                                    \
  namespace NS {                                                                                                           \
  constexpr int TH = TH_, TW = 64;         /* output tile of one CTA */                                                  \
  constexpr int LH = TH + 8, LW = TW + 8;  /* staged luminance tile: origin (ty0-4, tx0-4) */                            \
  constexpr int RP = RP_;                  /* ring positions: temporal windows of up to RP + 1 taps */                   \
  /* input_kind: fused::IN_PYRAMID_TMA or fused::IN_LEVEL0_TMA */                                                        \
  cudaError_t launch_band_ws(int input_kind, bool foveated, const fused::BandParams& p, dim3 grid, cudaStream_t st);      \
  cudaError_t configure_band_ws_kernels();                                                                                 \
  }
#ifdef WS_TH16   /* experiment: the 8-tap build on 16-row tiles as well */
FVVDP_WS_DECLARE(ws, 16, 7)
#else
FVVDP_WS_DECLARE(ws, 32, 7)
#endif
FVVDP_WS_DECLARE(ws16, 16, 15)
#undef FVVDP_WS_DECLARE
}  // namespace fvvdp
