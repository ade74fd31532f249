// Probe for the next step of the Gibbs leaf level ("Next optimisation steps"): one tcgen05.mma kind::tf32 with
// hand-built shared-memory and instruction descriptors, accumulator in TMEM, read back with tcgen05.ld and checked on the host.
//   D[128 x N] = A[128 x 8] * B[N x 8]^T     (both operands K-major, no swizzle: 8 x 16-byte core matrices)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 tcgen05_tf32_probe.cu -o tcgen05_tf32_probe && ./tcgen05_tf32_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int M = 128, N = 64, K = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// canonical K-major, no-swizzle layout: core matrix = 8 rows x 4 tf32 (16 bytes per row, 128 bytes per core matrix);
// the two core matrices along K are adjacent (LBO = 128 B), the next 8 rows follow (SBO = 256 B)
__device__ __forceinline__ int canon_off(int row, int k) { return (row >> 3) * 64 + (k >> 2) * 32 + (row & 7) * 4 + (k & 3); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address, bits [0,14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16; // leading (K) byte offset, bits [16,30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32; // stride (M/N) byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                           // descriptor version (sm_100)
  return d;                                         // base offset 0, layout type 0 = SWIZZLE_NONE
}

__global__ void __launch_bounds__(128) probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
  __shared__ __align__(128) float sA[M * K];
  __shared__ __align__(128) float sB[N * K];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < M * K; i += 128) sA[canon_off(i / K, i % K)] = A[i];
  for (int i = tid; i < N * K; i += 128) sB[canon_off(i / K, i % K)] = B[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(64u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = tmem_base;
  if (tid == 0) {
    const uint64_t da = make_desc(smem_u32(sA), 128, 256), db = make_desc(smem_u32(sB), 128, 256);
    // instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), K-major both, N >> 3 at 17, M >> 4 at 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(0u));
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // wait for the MMA (phase 0)
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra DONE;\n\tbra WAIT;\n\tDONE:\n\t}\n" ::"r"(
          smem_u32(&bar))
      : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;");
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 32; ++c) D[(warp * 32 + lane) * N + c0 + c] = __uint_as_float(r[c]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(64u));
}

int main() {
  float hA[M * K], hB[N * K], hD[M * N];
  srand(1);
  for (int i = 0; i < M * K; ++i) hA[i] = (float)(rand() % 17 - 8) / 8.0f;  // exactly representable in tf32
  for (int i = 0; i < N * K; ++i) hB[i] = (float)(rand() % 17 - 8) / 4.0f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD));
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, sizeof(hD));
  probe<<<1, 128>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
  double maxerr = 0;
  int bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)hA[m * K + k] * hB[n * K + k];
      double err = fabs(ref - hD[m * N + n]);
      if (err > maxerr) maxerr = err;
      if (err > 1e-5 && bad++ < 5) printf("mismatch D[%d][%d] = %f, want %f\n", m, n, hD[m * N + n], ref);
    }
  printf("tcgen05 tf32 probe: M=%d N=%d K=%d max |err| = %g, mismatches %d\n", M, N, K, maxerr, bad);
  return bad != 0;
}
