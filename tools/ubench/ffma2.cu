// FFMA vs FFMA2 issue throughput on one SM's worth of warps (sm_100a).  nvcc -arch=sm_100a -o ffma2 ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
template <int MODE> __global__ void k(float* out, int iters, float s) {
  float a[16]; unsigned long long p[8];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  for (int i = 0; i < 8; ++i) { float2 t = make_float2(a[2 * i], a[2 * i + 1]); p[i] = *(unsigned long long*)&t; }
  float2 sw = make_float2(s, s); unsigned long long w = *(unsigned long long*)&sw;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], s, a[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], w, p[i]);
    }
  }
  float acc = 0;
  if (MODE == 0) for (int i = 0; i < 16; ++i) acc += a[i];
  else for (int i = 0; i < 8; ++i) { float2 t = *(float2*)&p[i]; acc += t.x + t.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 1024 * 4 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode)
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 2, 512>>>(d, iters, 1e-9f); else k<1><<<148 * 2, 512>>>(d, iters, 1e-9f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma = 148.0 * 2 * 512 * 16.0 * iters;
      printf("mode %s: %.3f ms  %.1f TFLOP/s (fp32 FMA = 2 flop)\n", mode ? "FFMA2" : "FFMA ", ms, 2 * fma / ms / 1e9);
    }
  return 0;
}
