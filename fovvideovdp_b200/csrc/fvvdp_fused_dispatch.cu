// Dispatch over the eight fused-kernel translation units.
#include "fvvdp_fused_launch.h"

namespace fvvdp {
namespace fused {

#define DECL(k, v)                                                                                          \
  cudaError_t launch_band_##k##_##v(bool foveated, bool extra, const BandParams& p, dim3 grid, cudaStream_t st); \
  cudaError_t configure_band_##k##_##v();
DECL(0, 0) DECL(0, 1) DECL(1, 0) DECL(1, 1) DECL(2, 0) DECL(2, 1) DECL(3, 0) DECL(3, 1)
#undef DECL

cudaError_t launch_band(int kind, bool video, bool foveated, bool extra, const BandParams& p, dim3 grid, cudaStream_t st) {
  switch (kind * 2 + (video ? 1 : 0)) {
    case 0: return launch_band_0_0(foveated, extra, p, grid, st);
    case 1: return launch_band_0_1(foveated, extra, p, grid, st);
    case 2: return launch_band_1_0(foveated, extra, p, grid, st);
    case 3: return launch_band_1_1(foveated, extra, p, grid, st);
    case 4: return launch_band_2_0(foveated, extra, p, grid, st);
    case 5: return launch_band_2_1(foveated, extra, p, grid, st);
    case 6: return launch_band_3_0(foveated, extra, p, grid, st);
    case 7: return launch_band_3_1(foveated, extra, p, grid, st);
  }
  return cudaErrorInvalidValue;
}

cudaError_t configure_band_kernels() {
  cudaError_t e;
  if ((e = configure_band_0_0()) != cudaSuccess) return e;
  if ((e = configure_band_0_1()) != cudaSuccess) return e;
  if ((e = configure_band_1_0()) != cudaSuccess) return e;
  if ((e = configure_band_1_1()) != cudaSuccess) return e;
  if ((e = configure_band_2_0()) != cudaSuccess) return e;
  if ((e = configure_band_2_1()) != cudaSuccess) return e;
  if ((e = configure_band_3_0()) != cudaSuccess) return e;
  return configure_band_3_1();
}

}  // namespace fused
}  // namespace fvvdp
