// Dispatch over the twelve fused-kernel translation units.
#include "fvvdp_fused_launch.h"

namespace fvvdp {
namespace fused {

#define DECL(k, v)                                                                                          \
  cudaError_t launch_band_##k##_##v(bool foveated, bool extra, const BandParams& p, dim3 grid, cudaStream_t st); \
  cudaError_t configure_band_##k##_##v();
#define DECL3(k) DECL(k, 0) DECL(k, 1) DECL(k, 2)
DECL3(0) DECL3(1) DECL3(2) DECL3(3)
#undef DECL3
#undef DECL

cudaError_t launch_band(int kind, int mode, bool foveated, bool extra, const BandParams& p, dim3 grid, cudaStream_t st) {
  switch (kind * 3 + mode) {
#define CASE(k, v) case (k) * 3 + (v): return launch_band_##k##_##v(foveated, extra, p, grid, st);
    CASE(0, 0) CASE(0, 1) CASE(0, 2) CASE(1, 0) CASE(1, 1) CASE(1, 2) CASE(2, 0) CASE(2, 1) CASE(2, 2) CASE(3, 0) CASE(3, 1) CASE(3, 2)
#undef CASE
  }
  return cudaErrorInvalidValue;
}

cudaError_t configure_band_kernels() {
  cudaError_t e;
#define CONF(k, v) if ((e = configure_band_##k##_##v()) != cudaSuccess) return e;
  CONF(0, 0) CONF(0, 1) CONF(0, 2) CONF(1, 0) CONF(1, 1) CONF(1, 2) CONF(2, 0) CONF(2, 1) CONF(2, 2) CONF(3, 0) CONF(3, 1) CONF(3, 2)
#undef CONF
  return cudaSuccess;
}

}  // namespace fused
}  // namespace fvvdp
