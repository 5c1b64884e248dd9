// Issue-rate microbenchmarks for the instruction mix of the fused band kernel (sm_100a):
// 3-register FFMA, FFMA with a constant-bank operand, packed fma.rn.f32x2, FMNMX, LDS.32/64/128.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>

struct W { float w[16]; };

template <int OP>
__global__ void __launch_bounds__(512) k(float* out, int iters, float seed, const __grid_constant__ W wc) {
  __shared__ float sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = seed * i;
  __syncthreads();
  float a = seed + threadIdx.x * 1e-3f, b = a + 0.5f, c = a + 0.25f, d = a + 0.125f;
  float e = a * 0.5f, f = b * 0.5f, g = c * 0.5f, h = d * 0.5f;
  float x = seed * 1.0001f, y = seed * 0.5f;
  unsigned long long pa, pb, pc, pd, px;
  asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a), "f"(b));
  asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(c), "f"(d));
  asm("mov.b64 %0, {%1, %2};" : "=l"(pc) : "f"(e), "f"(f));
  asm("mov.b64 %0, {%1, %2};" : "=l"(pd) : "f"(g), "f"(h));
  asm("mov.b64 %0, {%1, %2};" : "=l"(px) : "f"(x), "f"(y));
  const float* sp = sm + (threadIdx.x & 31) * 4;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (OP == 0) { a = fmaf(a, x, y); b = fmaf(b, x, y); c = fmaf(c, x, y); d = fmaf(d, x, y); e = fmaf(e, x, y); f = fmaf(f, x, y); g = fmaf(g, x, y); h = fmaf(h, x, y); }
      if (OP == 1) { a = fmaf(a, wc.w[0], y); b = fmaf(b, wc.w[1], y); c = fmaf(c, wc.w[2], y); d = fmaf(d, wc.w[3], y); e = fmaf(e, wc.w[4], y); f = fmaf(f, wc.w[5], y); g = fmaf(g, wc.w[6], y); h = fmaf(h, wc.w[7], y); }
      if (OP == 2) {
        asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(pa) : "l"(px));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(pb) : "l"(px));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(pc) : "l"(px));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(pd) : "l"(px));
      }
      if (OP == 3) { a = fmaxf(a, x); b = fminf(b, y); c = fmaxf(c, x); d = fminf(d, y); e = fmaxf(e, x); f = fminf(f, y); g = fmaxf(g, x); h = fminf(h, y); }
      if (OP == 4) { a += sp[u * 128]; b += sp[u * 128 + 1024]; c += sp[u * 128 + 2048]; d += sp[u * 128 + 3072]; }
      if (OP == 5) { float4 v = *reinterpret_cast<const float4*>(sp + u * 128); float4 q = *reinterpret_cast<const float4*>(sp + u * 128 + 2048); a += v.x; b += v.y; c += v.z; d += v.w; e += q.x; f += q.y; g += q.z; h += q.w; }
      if (OP == 6) { a = fmaf(a, x, b); b = fmaf(b, y, c); c = fmaf(c, x, d); d = fmaf(d, y, e); e = fmaf(e, x, f); f = fmaf(f, y, g); g = fmaf(g, x, h); h = fmaf(h, y, a); }
      if (OP == 7) { a = a + x; b = b + y; c = c + x; d = d + y; e = e * x; f = f * y; g = g * x; h = h * y; }
    }
  }
  if (OP == 2) {
    float t0, t1;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(t0), "=f"(t1) : "l"(pa)); a = t0 + t1;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(t0), "=f"(t1) : "l"(pb)); b = t0 + t1;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(t0), "=f"(t1) : "l"(pc)); c = t0 + t1;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(t0), "=f"(t1) : "l"(pd)); d = t0 + t1;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d + e + f + g + h;
}

template <int OP>
void run(const char* name, int per_iter, int sms, int clock_khz) {
  float* out;
  const int blocks = sms * 4, threads = 512, iters = 2048;
  cudaMalloc(&out, sizeof(float) * blocks * threads);
  W wc;
  for (int i = 0; i < 16; ++i) wc.w[i] = 1.0f + 1e-4f * i;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<OP><<<blocks, threads>>>(out, iters, 1.0f, wc);
  cudaEventRecord(e0);
  k<OP><<<blocks, threads>>>(out, iters, 1.0f, wc);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)blocks * threads * iters * 8.0 * per_iter;
  printf("%-28s %9.1f G thread-instr/s = %6.1f thread-instr/clk/SM (at %d MHz)  [%s]\n", name, ops / ms / 1e6,
         ops / (ms * 1e-3) / sms / (clock_khz * 1e3), clock_khz / 1000, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%s, %d SMs, %d kHz\n", p.name, p.multiProcessorCount, clk);
  run<0>("ffma r,r,r (shared x,y)", 8, p.multiProcessorCount, clk);
  run<6>("ffma r,r,r (3 distinct)", 8, p.multiProcessorCount, clk);
  run<1>("ffma r,const,r", 8, p.multiProcessorCount, clk);
  run<2>("fma.rn.f32x2 (instr)", 4, p.multiProcessorCount, clk);
  run<3>("fmnmx", 8, p.multiProcessorCount, clk);
  run<7>("fadd/fmul", 8, p.multiProcessorCount, clk);
  run<4>("lds.32", 4, p.multiProcessorCount, clk);
  run<5>("lds.128", 2, p.multiProcessorCount, clk);
  return 0;
}
