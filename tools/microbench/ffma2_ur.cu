// Does a packed fma.rn.f32x2 with a uniform-register (kernel parameter) multiplicand issue at the same rate as the
// all-register form?  (The warp-specialised band kernel keeps its temporal-filter weights in uniform registers.)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_ur ffma2_ur.cu && ./ffma2_ur
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
struct W { u64 w[16]; };

template <int OP>
__global__ void __launch_bounds__(512) k(u64* out, int iters, const __grid_constant__ W wc, const u64* wg) {
  u64 a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = wg[threadIdx.x + 32 * i];
  u64 r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = wg[256 + i];   // weights in registers
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (OP == 0) asm volatile("fma.rn.f32x2 %0, %0, %1, %0;" : "+l"(a[i]) : "l"(r[(i + u) & 7]));
        if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %0;" : "+l"(a[i]) : "l"(wc.w[(i + 2 * u) & 15]));
        if (OP == 2) asm volatile("fma.rn.f32x2 %0, %1, %0, %0;" : "+l"(a[i]) : "l"(wc.w[(i + 2 * u) & 15]));
      }
    }
  }
  u64 s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, int sms, double ghz) {
  u64 *out, *wg;
  const int blocks = sms * 4, threads = 512, iters = 4096;
  cudaMalloc(&out, sizeof(u64) * blocks * threads);
  cudaMalloc(&wg, sizeof(u64) * 1024);
  cudaMemset(wg, 0, sizeof(u64) * 1024);
  W wc;
  for (int i = 0; i < 16; ++i) wc.w[i] = 0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<OP><<<blocks, threads>>>(out, 16, wc, wg);
  cudaEventRecord(e0);
  k<OP><<<blocks, threads>>>(out, iters, wc, wg);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double inst = (double)blocks * threads * iters * 32.0;  // thread-instructions
  printf("%-44s %8.3f ms  %7.2f thread-FFMA2 / clk / SM (at %.3f GHz)\n", name, ms, inst / (ms * 1e-3) / (ghz * 1e9) / sms, ghz);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  run<0>("FFMA2 R, R, R, R (weights in registers)", p.multiProcessorCount, ghz);
  run<1>("FFMA2 R, R, UR, R (weight = kernel parameter)", p.multiProcessorCount, ghz);
  run<2>("FFMA2 R, UR, R, R (parameter first)", p.multiProcessorCount, ghz);
  return 0;
}
