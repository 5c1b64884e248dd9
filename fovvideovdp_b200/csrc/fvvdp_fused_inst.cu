// Instantiations of the fused band kernel.  Compiled once per (input kind, temporal mode) with
// -DFUSED_KIND={0,2,3} (fused::InputKind) -DFUSED_VIDEO={0: image, 1: video with up to 8 taps, 2: video with up to 16 taps,
// 3: video whose temporal channels arrive filtered, as two planes per slot (windows of 17..32 taps; FUSED_KIND=2 only)}
// so that the nine translation units build in parallel.
#ifndef FUSED_KIND
#error "compile with -DFUSED_KIND and -DFUSED_VIDEO"
#endif
#define FUSED_MAXRING_CASES (FUSED_VIDEO == 2)
#include "fvvdp_fused.cuh"
#include "fvvdp_fused_launch.h"

namespace fvvdp {
namespace fused {

constexpr int kFL = FUSED_VIDEO == 2 ? MAXRING : (FUSED_VIDEO == 1 ? RING : 1);
constexpr int kTC = FUSED_VIDEO ? 2 : 1;

#define FUSED_CAT2(a, b, c) a##b##_##c
#define FUSED_CAT(a, b, c) FUSED_CAT2(a, b, c)
#define FUSED_FN(name) FUSED_CAT(name, FUSED_KIND, FUSED_VIDEO)

cudaError_t FUSED_FN(launch_band_)(bool foveated, bool extra, const BandParams& p, dim3 grid, cudaStream_t st) {
  if (foveated) {
    const size_t smem = band_smem_bytes<FUSED_KIND, kFL, kTC, true>();
    if (extra) band_kernel<FUSED_KIND, kFL, kTC, true, true><<<grid, threads_of(kFL), smem, st>>>(p);
    else band_kernel<FUSED_KIND, kFL, kTC, true, false><<<grid, threads_of(kFL), smem, st>>>(p);
  } else {
    const size_t smem = band_smem_bytes<FUSED_KIND, kFL, kTC, false>();
    if (extra) band_kernel<FUSED_KIND, kFL, kTC, false, true><<<grid, threads_of(kFL), smem, st>>>(p);
    else band_kernel<FUSED_KIND, kFL, kTC, false, false><<<grid, threads_of(kFL), smem, st>>>(p);
  }
  return cudaGetLastError();
}

cudaError_t FUSED_FN(configure_band_)() {
  const int smem_fov = (int)band_smem_bytes<FUSED_KIND, kFL, kTC, true>(), smem = (int)band_smem_bytes<FUSED_KIND, kFL, kTC, false>();
  cudaError_t e;
  e = cudaFuncSetAttribute(band_kernel<FUSED_KIND, kFL, kTC, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_fov);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(band_kernel<FUSED_KIND, kFL, kTC, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_fov);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(band_kernel<FUSED_KIND, kFL, kTC, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(band_kernel<FUSED_KIND, kFL, kTC, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

}  // namespace fused
}  // namespace fvvdp
