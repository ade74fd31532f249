// How long does the ISSUING thread stay in a chain of tcgen05.mma kind::tf32 instructions, and when does the chain complete?
// Decides the structure of the Gibbs leaf pass (gibbs.cuh leaf_pass_tc): a consumer warp that also issues the six MMAs of a tile
// loses the issue time on the critical path of every tile.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tcgen05_issue_probe.cu -o tcgen05_issue_probe && ./tcgen05_issue_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da),
               "l"(db), "r"(idesc), "r"(acc));
}
// executed by a converged warp: one elected lane issues (operands stay warp-uniform -> uniform registers, no per-MMA R2UR)
__device__ __forceinline__ void mma_elect(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void commit_elect(uint64_t* bar) {
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}\n" ::"r"(smem_u32(bar)),
               "r"(parity)
               : "memory");
}

// mode 0: chain of n MMAs into one accumulator (N columns); mode 1: n MMAs alternating between two accumulators (two chains of n/2);
// mode 2: n MMAs into n different accumulators (N = 32 columns each when n > 4)
__global__ void __launch_bounds__(128) probe(int mode, int n, int N, long long* out, long long* stamps) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* sA = reinterpret_cast<float*>(smem);              // 128 x 16 (two K halves)
  float* sB = reinterpret_cast<float*>(smem + 16384);      // N x 16
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 4096 + 4096; i += 128) reinterpret_cast<float*>(smem)[i] = 0.001f * (float)(i % 97);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = tmem_base;
  if (tid == 0 && mode < 10) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint32_t parity = 0;
    for (int rep = 0; rep < 8; ++rep) {
      const long long t0 = clock64();
      for (int k = 0; k < n; ++k) {
        const uint64_t da = make_desc(smem_u32(sA) + (k & 1) * 256, 128, 512), db = make_desc(smem_u32(sB) + (k & 1) * 256, 128, 512);
        uint32_t d = tm, acc = k > 0;
        if (mode == 1) { d = tm + (k & 1) * N; acc = k > 1; }
        if (mode == 2) { d = tm + k * N; acc = 0; }
        mma(d, da, db, idesc, acc);
        if (stamps && rep == 7 && k < 24) stamps[k] = clock64() - t0;
      }
      commit(&bar);
      const long long t1 = clock64();
      wait(&bar, parity);
      parity ^= 1;
      const long long t2 = clock64();
      out[rep * 2] = t1 - t0;
      out[rep * 2 + 1] = t2 - t0;
    }
  }
  __syncthreads();
  if (warp == 0 && mode >= 10) {  // warp-uniform issue: modes 10 (chain) and 12 (independent accumulators)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    uint32_t parity = 0;
    for (int rep = 0; rep < 8; ++rep) {
      const long long t0 = clock64();
#pragma unroll 1
      for (int k = 0; k < n; ++k) {
        const uint64_t da = make_desc(a0 + (k & 1) * 256, 128, 512), db = make_desc(b0 + (k & 1) * 256, 128, 512);
        mma_elect(mode == 12 ? tm + (k & 3) * N : tm, da, db, idesc, mode == 12 ? 0u : (k > 0));
      }
      commit_elect(&bar);
      const long long t1 = clock64();
      wait(&bar, parity);
      parity ^= 1;
      const long long t2 = clock64();
      if (tid == 0) { out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256u));
}

int main() {
  long long* d;
  cudaMalloc(&d, 16 * sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
  struct { int mode, n, N; const char* what; } cases[] = {
      {0, 1, 64, "1 MMA, N=64"},          {0, 3, 64, "chain of 3, N=64"},       {0, 6, 64, "chain of 6, N=64"},
      {0, 6, 128, "chain of 6, N=128"},   {0, 12, 64, "chain of 12, N=64"},     {1, 12, 64, "two interleaved chains of 6, N=64"},
      {2, 4, 64, "4 independent, N=64"},  {2, 6, 32, "6 independent, N=32"},    {0, 6, 32, "chain of 6, N=32"},
      {0, 6, 256, "chain of 6, N=256"},   {0, 3, 128, "chain of 3, N=128"},
      {10, 1, 64, "warp-uniform issue: 1 MMA"}, {10, 3, 64, "warp-uniform issue: chain of 3"}, {10, 6, 64, "warp-uniform issue: chain of 6"},
      {10, 12, 64, "warp-uniform issue: chain of 12"}, {10, 24, 64, "warp-uniform issue: chain of 24"}, {12, 12, 64, "warp-uniform issue: 12 over 4 accumulators"}};
  for (auto& c : cases) {
    probe<<<1, 128, 49152>>>(c.mode, c.n, c.N, d, nullptr);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    long long h[16];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long bi = 1 << 30, bd = 1 << 30;
    for (int r = 2; r < 8; ++r) { if (h[2 * r] < bi) bi = h[2 * r]; if (h[2 * r + 1] < bd) bd = h[2 * r + 1]; }
    printf("%-40s issue %5lld clk   issue->complete %5lld clk\n", c.what, bi, bd);
  }
  long long* st;
  cudaMalloc(&st, 24 * sizeof(long long));
  for (int n : {6, 12, 24}) {
    probe<<<1, 128, 49152>>>(0, n, 64, d, st);
    cudaDeviceSynchronize();
    long long h[24];
    cudaMemcpy(h, st, sizeof(h), cudaMemcpyDeviceToHost);
    printf("chain of %d, N=64: clock after each MMA instruction:", n);
    for (int k = 0; k < n; ++k) printf(" %lld", h[k]);
    printf("\n");
  }
  return 0;
}
