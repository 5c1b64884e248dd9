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
// level-0 input of any dtype / channel count / strides -> luminance planes in the pyramid layout (slots x h x pitch floats)
// rows_vectorisable: unit pixel stride, row / channel strides and frame addresses aligned for 4-pixel vector loads
// PU21 squared error of n_frames frame pairs (p.slot pointers, level-0 input format of p), one double per frame added to out
cudaError_t launch_pu_frames(const BandParams& p, const void* pu_params, double* out, int n_frames, bool rows_vectorisable, cudaStream_t st);
cudaError_t launch_luminance(const BandParams& p, float* out, long long slot_stride, int pitch, int n_slots, bool rows_vectorisable, cudaStream_t st);
}  // namespace fused
}  // namespace fvvdp
