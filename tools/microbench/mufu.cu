// Special-function (MUFU) and FMA-pipe throughput on sm_100a: the secondary roof of the fused level kernels.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu mufu.cu && ./mufu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(float* out, int iters, float seed) {
  float a = seed + threadIdx.x * 1e-3f, b = a + 0.5f, c = a + 0.25f, d = a + 0.125f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (OP == 0) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(c)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(d)); }
      if (OP == 1) { asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a)); asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(b)); asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(c)); asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(d)); }
      if (OP == 2) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(b)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(c)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(d)); }
      if (OP == 3) { a = fmaf(a, 1.0001f, 0.5f); b = fmaf(b, 1.0001f, 0.5f); c = fmaf(c, 1.0001f, 0.5f); d = fmaf(d, 1.0001f, 0.5f); }
      if (OP == 4) { asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(a)); asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(b)); asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(c)); asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(d)); }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d;
}

template <int OP>
void run(const char* name, int sms, int clock_khz) {
  float* out;
  const int blocks = sms * 4, threads = 512, iters = 4096;
  cudaMalloc(&out, sizeof(float) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<OP><<<blocks, threads>>>(out, iters, 1.0f);
  cudaEventRecord(e0);
  k<OP><<<blocks, threads>>>(out, iters, 1.0f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)blocks * threads * iters * 32.0;
  printf("%-6s %8.2f Gop/s  = %.2f ops/clk/SM at %d MHz (max clock)\n", name, ops / ms / 1e6, ops / (ms * 1e-3) / sms / (clock_khz * 1e3), clock_khz / 1000);
  cudaFree(out);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%s, %d SMs, %d kHz\n", p.name, p.multiProcessorCount, clk);
  run<0>("ex2", p.multiProcessorCount, clk);
  run<1>("lg2", p.multiProcessorCount, clk);
  run<2>("rcp", p.multiProcessorCount, clk);
  run<4>("sqrt", p.multiProcessorCount, clk);
  run<3>("ffma", p.multiProcessorCount, clk);
  return 0;
}
