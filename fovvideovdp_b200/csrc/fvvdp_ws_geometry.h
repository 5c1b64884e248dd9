// Tile geometry and launch interface of the warp-specialised band kernel (fvvdp_ws.cuh), shared with the host code.
#pragma once
#include <cuda_runtime.h>

namespace fvvdp {
namespace fused {
struct BandParams;
}
namespace ws {
constexpr int TH = 32, TW = 64;                 // output tile of one CTA
constexpr int LH = TH + 8, LW = TW + 8;         // staged luminance tile: origin (ty0-4, tx0-4)
constexpr int RP = 7;                           // ring positions: temporal windows of up to RP + 1 taps
// input_kind: fused::IN_PYRAMID_TMA or fused::IN_LEVEL0_TMA
cudaError_t launch_band_ws(int input_kind, bool foveated, const fused::BandParams& p, dim3 grid, cudaStream_t st);
cudaError_t configure_band_ws_kernels();
}  // namespace ws
}  // namespace fvvdp
