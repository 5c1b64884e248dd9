// Instantiations of the warp-specialised band kernel (fvvdp_ws.cuh).  Compiled once per input kind with
// -DWS_KIND={2: pyramid planes, 3: contiguous float frames} (fused::InputKind) and -DWS_TAPS16={0: windows of up to 8 taps,
// 1: up to 16 taps} so that the units build in parallel.
#ifndef WS_KIND
#error "compile with -DWS_KIND"
#endif
#include "fvvdp_ws.cuh"
#include "fvvdp_fused_launch.h"

namespace fvvdp {
namespace WS_NS {

#define WS_CAT2(a, b) a##b
#define WS_CAT(a, b) WS_CAT2(a, b)
#define WS_FN(name) WS_CAT(name, WS_KIND)

cudaError_t WS_FN(launch_band_ws_)(bool foveated, const BandParams& p, dim3 grid, cudaStream_t st) {
  if (foveated) band_ws_kernel<WS_KIND, true><<<grid, NT, Layout<WS_KIND, true>::bytes, st>>>(p);
  else band_ws_kernel<WS_KIND, false><<<grid, NT, Layout<WS_KIND, false>::bytes, st>>>(p);
  return cudaGetLastError();
}

cudaError_t WS_FN(configure_band_ws_)() {
  cudaError_t e = cudaFuncSetAttribute(band_ws_kernel<WS_KIND, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Layout<WS_KIND, true>::bytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(band_ws_kernel<WS_KIND, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Layout<WS_KIND, false>::bytes);
}

}  // namespace ws / ws16
}  // namespace fvvdp
