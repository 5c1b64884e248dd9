// Launch interface of the fused band kernels (fvvdp_fused.cuh), one translation unit per input kind / temporal mode.
#pragma once
#include <cuda_runtime.h>

namespace fvvdp {
namespace fused {
struct BandParams;
// input_kind: fused::InputKind (fvvdp_fused.cuh)
// video = 8-slot temporal ring, 2 temporal channels; image = single frame, 1 temporal channel
cudaError_t launch_band(int input_kind, bool video, bool foveated, bool extra, const BandParams& p, dim3 grid, cudaStream_t st);
cudaError_t configure_band_kernels();
}  // namespace fused
}  // namespace fvvdp
