// Microbenchmark: throughput of warp-broadcast shared-memory loads (all lanes read the same address) on sm_100a.
// nvcc -arch=sm_100a -O3 lds_bcast.cu -o lds_bcast && ./lds_bcast
#include <cstdio>
#include <cuda_runtime.h>
template <int WIDTH>
__global__ void __launch_bounds__(512) k(float* out, int iters) {
  __shared__ __align__(16) float buf[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) buf[i] = i * 1e-3f;
  __syncthreads();
  unsigned base = (unsigned)__cvta_generic_to_shared(buf);
  float acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      unsigned addr = base + (((it * 16 + u) * 32) & 16383);
      float a, b, c, d;
      if (WIDTH == 16) { asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(addr)); acc0 += a; acc1 += b; acc2 += c; acc3 += d; }
      else if (WIDTH == 8) { asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(a), "=f"(b) : "r"(addr)); acc0 += a; acc1 += b; }
      else { asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a) : "r"(addr)); acc0 += a; }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1 + acc2 + acc3;
}
template <int WIDTH> void run(const char* name, float* out) {
  const int iters = 4096, blocks = 148;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<WIDTH><<<blocks, 512>>>(out, 16);
  cudaEventRecord(e0);
  k<WIDTH><<<blocks, 512>>>(out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double loads_per_sm = (double)iters * 16 * 16;  // warp-level loads per SM (16 warps)
  printf("%s: %.3f ms, %.2f clk per warp-load per SM at 1.965 GHz\n", name, ms, ms * 1e-3 * 1.965e9 / loads_per_sm);
}
int main() {
  float* out; cudaMalloc(&out, 148 * 512 * 4);
  run<4>("LDS.32  broadcast", out); run<8>("LDS.64  broadcast", out); run<16>("LDS.128 broadcast", out);
  return 0;
}
