// Microbenchmark: issue rate of FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.  nvcc -arch=sm_100a -O3 ffma2_rate.cu
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float ffma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
template <int MODE> __global__ void k(float* out, int iters, float s) {
  float a[16];
  unsigned long long p[8];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  for (int i = 0; i < 8; ++i) p[i] = ((unsigned long long)__float_as_uint(a[2 * i]) << 32) | __float_as_uint(a[2 * i + 1]);
  unsigned long long ss = ((unsigned long long)__float_as_uint(s) << 32) | __float_as_uint(s * 0.5f);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = ffma1(a[i], s, a[(i + 1) & 15]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = ffma2(p[i], ss, p[(i + 1) & 7]);
    }
  }
  float r = 0;
  for (int i = 0; i < 16; ++i) r += a[i];
  for (int i = 0; i < 8; ++i) r += __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)p[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters, 0.999f); else k<1><<<148 * 8, 256>>>(out, iters, 0.999f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma = 148.0 * 8 * 256 * (double)iters * 16;  // scalar FMAs in both modes
      if (rep) printf("%s: %.3f ms, %.1f FMA/clk/SM at 1.965 GHz, %.2f TFLOP/s\n", mode ? "FFMA2" : "FFMA ", ms, fma / (ms * 1e-3) / 148 / 1.965e9, 2 * fma / (ms * 1e-3) / 1e12);
    }
  }
  return 0;
}
