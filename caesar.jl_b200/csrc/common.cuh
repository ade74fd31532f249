// common.cuh -- context, error plumbing, device helpers shared by the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <math.h>
#include <string>
#include <vector>
#include <map>
#include <mutex>
#include "../../include/caesar_b200.h"

#define CB2_LOG2E 1.4426950408889634074
#define CB2_TWO_PI 6.283185307179586476925286766559
#define CB2_PI 3.14159265358979323846264338327950288

struct cb2_arena {  // grow-only device scratch; reset at the start of each public call
  char* base = nullptr;
  size_t cap = 0, off = 0;
};

struct cb2_pending {  // one submitted async product batch: completion event + the device block it owns until cb2_wait
  cudaEvent_t done = nullptr;
  void* dev_block = nullptr;
};

struct cb2_ctx {
  int device = 0;
  int precision = CB2_FP32;
  int sm_count = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cb2_arena arena;       // device scratch
  cb2_arena io;          // grow-only device block for the host-pointer API's staged inputs / outputs
  char* pinned = nullptr;  // pinned host staging (small batches are packed here: one H2D and one D2H copy per call)
  size_t pinned_cap = 0;
  double* small_dev = nullptr;  // 64 doubles of persistent device scratch (scalar results of nested calls)
  uint64_t launches = 0;
  int level_down = 1;  // Gibbs levelDown rule: 1 = weight-proportional random child, 0 = first child (cb2_set_level_down)
  std::string err;
  std::recursive_mutex mu;  // recursive: host-pointer entry points hold it across their nested _dev calls
  std::map<std::string, double> stage_ms;
  cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // host-pointer product batches: the arrays move on a second stream, chunk by chunk, while the trees of the chunks that have
  // arrived are built (and the sample points travel back while the bandwidths are searched)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t copy_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  std::map<int64_t, cb2_pending> pending;
  int64_t next_ticket = 1;
  std::map<const void*, size_t> smem_attr;  // kernels whose dynamic shared-memory limit has been raised, and to what (cb2_func_smem)
};

typedef std::lock_guard<std::recursive_mutex> cb2_lock;

// Raises the dynamic shared-memory limit of a kernel once per context (again only for a larger request): the attribute call is a
// driver round trip of a few microseconds, which small batches pay on every launch otherwise.  Caller holds ctx->mu.
inline cudaError_t cb2_func_smem(cb2_ctx* ctx, const void* fn, size_t bytes, bool max_carveout = false) {
  auto it = ctx->smem_attr.find(fn);
  if (it != ctx->smem_attr.end() && it->second >= bytes) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess && max_carveout) e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e == cudaSuccess) ctx->smem_attr[fn] = bytes;
  return e;
}

int cb2_fail(cb2_ctx* ctx, int code, const char* fmt, ...);

#define CB2_CUDA(ctx, call)                                                                         \
  do {                                                                                              \
    cudaError_t e__ = (call);                                                                       \
    if (e__ != cudaSuccess)                                                                         \
      return cb2_fail(ctx, e__ == cudaErrorMemoryAllocation ? CB2_ERR_NOMEM : CB2_ERR_CUDA,         \
                      "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

#define CB2_CHECK_LAUNCH(ctx, name)                                                                     \
  do {                                                                                                  \
    (ctx)->launches++;                                                                                  \
    cudaError_t e__ = cudaGetLastError();                                                               \
    if (e__ != cudaSuccess)                                                                             \
      return cb2_fail(ctx, CB2_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__));     \
  } while (0)

#define CB2_REQUIRE(ctx, cond, ...)                                  \
  do {                                                               \
    if (!(cond)) return cb2_fail(ctx, CB2_ERR_INVALID, __VA_ARGS__); \
  } while (0)

// device scratch: 256-byte aligned bump allocation out of the context arena
int cb2_arena_reserve(cb2_ctx* ctx, size_t bytes);  // ensure capacity (may reallocate: only call when off==0)
void* cb2_arena_alloc(cb2_ctx* ctx, size_t bytes);  // nullptr when out of reserved space
int cb2_pinned_reserve(cb2_ctx* ctx, size_t bytes);
void* cb2_io_block(cb2_ctx* ctx, size_t bytes);  // staged I/O of synchronous host-pointer calls (no per-call cudaMalloc/cudaFree)
void cb2_resolve_product_stages(cb2_ctx* ctx);  // product.cu: turns recorded events into stage_ms

static inline size_t cb2_align(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------ device helpers
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// uniform in (0,1): float path keeps 23 bits + 1/2 (exactly representable in FP32; the oracle's ubits=24 mode),
// double path keeps 53 bits (oracle's ubits=53)
template <typename T> __device__ __forceinline__ T u01(uint32_t x0, uint32_t x1);
template <> __device__ __forceinline__ float u01<float>(uint32_t x0, uint32_t) {
  return ((float)(x0 >> 9) + 0.5f) * (1.0f / 8388608.0f);  // 23 bits + 1/2: exact in FP32
}
template <> __device__ __forceinline__ double u01<double>(uint32_t x0, uint32_t x1) {
  uint64_t v = (((uint64_t)x0 << 32) | x1) >> 11;
  return ((double)v + 0.5) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ double fast_exp2(double x) { return exp2(x); }
// exp2 of a non-positive argument on the FMA and ALU pipes instead of the XU pipe (16 MUFU per clock and SM is what bounds the
// pair kernels: mixing one software evaluation into every four takes a quarter of the load off that pipe).  Round to nearest
// integer with the 1.5 * 2^23 trick, degree-5 polynomial for 2^f on [-0.5, 0.5] (max relative error 2.4e-7, the same two ulp
// as ex2.approx), exponent added in integer arithmetic.  Arguments below -125 return ~2e-38 instead of a flushed zero.
__device__ __forceinline__ float soft_exp2(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float p = 0.0013390863f;
  p = fmaf(p, f, 0.009676032f);
  p = fmaf(p, f, 0.05550357f);
  p = fmaf(p, f, 0.24022107f);
  p = fmaf(p, f, 0.6931472f);
  p = fmaf(p, f, 1.0000001f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ double soft_exp2(double x) { return exp2(x); }
// packed FP32 pairs (fma.rn.f32x2 -> SASS FFMA2 / FMUL2 / FADD2 on sm_100a): one issue slot for two lanes of arithmetic
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fast_log2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ double fast_log2(double x) { return log2(x); }
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ double fast_rcp(double x) { return 1.0 / x; }

// wrap to [-pi, pi]
__device__ __forceinline__ float wrap_pi(float x) { return x - 6.2831853071795864769f * rintf(x * 0.15915494309189533577f); }
__device__ __forceinline__ double wrap_pi(double x) { return x - CB2_TWO_PI * rint(x / CB2_TWO_PI); }

__device__ __forceinline__ float rint_t(float x) { return rintf(x); }
__device__ __forceinline__ double rint_t(double x) { return rint(x); }
__device__ __forceinline__ float atan2_t(float y, float x) { return atan2f(y, x); }
__device__ __forceinline__ double atan2_t(double y, double x) { return atan2(y, x); }
// FP32: MUFU.SIN / MUFU.COS (arguments are stored angles in [-pi, pi]: absolute error ~ 5e-7, no range-reduction slow path)
__device__ __forceinline__ void sincos_t(float x, float* s, float* c) { __sincosf(x, s, c); }
__device__ __forceinline__ void sincos_t(double x, double* s, double* c) { sincos(x, s, c); }

template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
