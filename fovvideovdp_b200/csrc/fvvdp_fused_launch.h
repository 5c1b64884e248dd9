// Launch interface of the fused band kernels (fvvdp_fused.cuh), one translation unit per input kind / temporal mode.
#pragma once
#include <cuda_runtime.h>

namespace fvvdp {
namespace fused {
struct BandParams;
// input_kind: fused::InputKind (fvvdp_fused.cuh)
// mode: 0 = image (single frame, 1 temporal channel); 1 = video, temporal ring of 8 frames (256 threads, a 2x2 quad each);
// 2 = video, ring of 16 frames (512 threads, one row of a quad each)
cudaError_t launch_band(int input_kind, int mode, bool foveated, bool extra, const BandParams& p, dim3 grid, cudaStream_t st);
cudaError_t configure_band_kernels();
}  // namespace fused
}  // namespace fvvdp
