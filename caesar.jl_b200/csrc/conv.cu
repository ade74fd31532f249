// conv.cu -- approxConvBelief for closed-form relative factors ("next" row N1): the per-particle
// factor solve that IIF runs immediately before every product (R:docs/src/principles/approxConvDensities.md:25-34,107-113),
// batched over all factors of a Gibbs colour.  One thread per (convolution, particle); FP64 throughout (the work is a
// handful of flops per particle); noise from Philox (seed; particle, draw, stream, tag 4).
#include "common.cuh"

namespace {

struct ConvDev {
  int kind, reverse, N, n_src, n_aux;
  uint32_t stream;
  const double* src;
  const double* aux;
  double* out;
  double mean[3];
  double chol[9];
};

__device__ __forceinline__ void normal3(uint32_t i, uint32_t stream, uint32_t k0, uint32_t k1, double z[3]) {
  uint32_t o[4];
  philox4x32_10(i, 0u, stream, 4u, k0, k1, o);
  double u1 = u01<double>(o[0], o[1]), u2 = u01<double>(o[2], o[3]);
  double r = sqrt(-2.0 * log(u1)), sn, cs;
  sincospi(2.0 * u2, &sn, &cs);
  z[0] = r * cs; z[1] = r * sn;
  philox4x32_10(i, 1u, stream, 4u, k0, k1, o);
  u1 = u01<double>(o[0], o[1]); u2 = u01<double>(o[2], o[3]);
  r = sqrt(-2.0 * log(u1));
  sincospi(2.0 * u2, &sn, &cs);
  z[2] = r * cs;
}

__global__ void approx_conv_kernel(const ConvDev* __restrict__ descs, uint32_t k0, uint32_t k1) {
  const ConvDev c = descs[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.N) return;
  double z[3], nz[3];
  normal3((uint32_t)i, c.stream, k0, k1, z);
  for (int r = 0; r < 3; ++r) nz[r] = c.mean[r] + c.chol[r * 3] * z[0] + c.chol[r * 3 + 1] * z[1] + c.chol[r * 3 + 2] * z[2];
  if (c.kind == CB2_FACTOR_PRIOR_POSE2) {
    c.out[(size_t)i * 3] = nz[0]; c.out[(size_t)i * 3 + 1] = nz[1]; c.out[(size_t)i * 3 + 2] = wrap_pi(nz[2]);
  } else if (c.kind == CB2_FACTOR_POSE2POSE2) {
    const double* p = c.src + (size_t)(i % c.n_src) * 3;
    if (!c.reverse) {  // q = p o exp(X): product exponential, translation = coordinates
      double sn, cs;
      sincos(p[2], &sn, &cs);
      c.out[(size_t)i * 3] = p[0] + cs * nz[0] - sn * nz[1];
      c.out[(size_t)i * 3 + 1] = p[1] + sn * nz[0] + cs * nz[1];
      c.out[(size_t)i * 3 + 2] = wrap_pi(p[2] + nz[2]);
    } else {  // p = q o exp(X)^-1
      double th = wrap_pi(p[2] - nz[2]), sn, cs;
      sincos(th, &sn, &cs);
      c.out[(size_t)i * 3] = p[0] - (cs * nz[0] - sn * nz[1]);
      c.out[(size_t)i * 3 + 1] = p[1] - (sn * nz[0] + cs * nz[1]);
      c.out[(size_t)i * 3 + 2] = th;
    }
  } else {  // bearing-range: nz[0] = bearing, nz[1] = range
    if (!c.reverse) {
      const double* p = c.src + (size_t)(i % c.n_src) * 3;
      double sn, cs;
      sincos(p[2] + nz[0], &sn, &cs);
      c.out[(size_t)i * 2] = p[0] + nz[1] * cs;
      c.out[(size_t)i * 2 + 1] = p[1] + nz[1] * sn;
    } else {  // two constraints on three coordinates: the heading comes from the pose's current belief
      const double* l = c.src + (size_t)(i % c.n_src) * 2;
      const double th = c.aux[(size_t)(i % c.n_aux) * 3 + 2];
      double sn, cs;
      sincos(th + nz[0], &sn, &cs);
      c.out[(size_t)i * 3] = l[0] - nz[1] * cs;
      c.out[(size_t)i * 3 + 1] = l[1] - nz[1] * sn;
      c.out[(size_t)i * 3 + 2] = th;
    }
  }
}

inline int out_dim(const cb2_conv_desc& c) { return (c.kind == CB2_FACTOR_POSE2POINT2_BEARINGRANGE && !c.reverse) ? 2 : 3; }
inline int src_dim(const cb2_conv_desc& c) { return (c.kind == CB2_FACTOR_POSE2POINT2_BEARINGRANGE && c.reverse) ? 2 : 3; }

}  // namespace

extern "C" int cb2_approx_conv_batch(cb2_ctx* ctx, const cb2_conv_desc* descs, int B, uint64_t seed) {
  if (!ctx) return CB2_ERR_INVALID;
  cb2_lock g(ctx->mu);
  CB2_REQUIRE(ctx, descs && B > 0, "null descriptors or empty batch");
  CB2_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t total = cb2_align(sizeof(ConvDev) * B);
  int maxN = 0;
  for (int b = 0; b < B; ++b) {
    const cb2_conv_desc& c = descs[b];
    CB2_REQUIRE(ctx, c.kind >= 0 && c.kind <= 2, "convolution %d: unknown factor kind %d", b, c.kind);
    CB2_REQUIRE(ctx, c.N > 0 && c.out, "convolution %d: N <= 0 or null out", b);
    if (c.kind != CB2_FACTOR_PRIOR_POSE2) CB2_REQUIRE(ctx, c.src && c.n_src > 0, "convolution %d: missing source samples", b);
    if (c.kind == CB2_FACTOR_POSE2POINT2_BEARINGRANGE && c.reverse)
      CB2_REQUIRE(ctx, c.aux && c.n_aux > 0, "convolution %d: reverse bearing-range needs the pose's current samples", b);
    if (c.kind != CB2_FACTOR_PRIOR_POSE2) total += cb2_align(sizeof(double) * (size_t)c.n_src * src_dim(c));
    if (c.aux) total += cb2_align(sizeof(double) * (size_t)c.n_aux * 3);
    total += cb2_align(sizeof(double) * (size_t)c.N * out_dim(c));
    if (c.N > maxN) maxN = c.N;
  }
  char* p = (char*)cb2_io_block(ctx, total + 256);
  if (!p) return CB2_ERR_NOMEM;
  int rc = cb2_pinned_reserve(ctx, total + 256);
  if (rc) return rc;
  char* stage = ctx->pinned;
  size_t off = cb2_align(sizeof(ConvDev) * B);
  std::vector<ConvDev> dd(B);
  for (int b = 0; b < B; ++b) {  // inputs first (one contiguous upload together with the descriptors)
    const cb2_conv_desc& c = descs[b];
    ConvDev& d = dd[b];
    d.kind = c.kind; d.reverse = c.reverse; d.N = c.N; d.n_src = c.n_src; d.n_aux = c.n_aux; d.stream = c.stream;
    memcpy(d.mean, c.mean, sizeof(d.mean)); memcpy(d.chol, c.chol, sizeof(d.chol));
    d.src = nullptr; d.aux = nullptr;
    if (c.kind != CB2_FACTOR_PRIOR_POSE2) {
      size_t nb = sizeof(double) * (size_t)c.n_src * src_dim(c);
      memcpy(stage + off, c.src, nb);
      d.src = (const double*)(p + off);
      off += cb2_align(nb);
    }
    if (c.aux && c.kind == CB2_FACTOR_POSE2POINT2_BEARINGRANGE && c.reverse) {
      size_t nb = sizeof(double) * (size_t)c.n_aux * 3;
      memcpy(stage + off, c.aux, nb);
      d.aux = (const double*)(p + off);
      off += cb2_align(nb);
    }
  }
  const size_t in_bytes = off;
  std::vector<size_t> out_off(B);
  for (int b = 0; b < B; ++b) {
    out_off[b] = off;
    dd[b].out = (double*)(p + off);
    off += cb2_align(sizeof(double) * (size_t)descs[b].N * out_dim(descs[b]));
  }
  memcpy(stage, dd.data(), sizeof(ConvDev) * B);
  CB2_CUDA(ctx, cudaMemcpyAsync(p, stage, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
  dim3 grid((maxN + 127) / 128, B);
  approx_conv_kernel<<<grid, 128, 0, ctx->stream>>>((const ConvDev*)p, (uint32_t)seed, (uint32_t)(seed >> 32));
  CB2_CHECK_LAUNCH(ctx, "approx_conv_kernel");
  CB2_CUDA(ctx, cudaMemcpyAsync(stage + in_bytes, p + in_bytes, off - in_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int b = 0; b < B; ++b) memcpy(descs[b].out, stage + out_off[b], sizeof(double) * (size_t)descs[b].N * out_dim(descs[b]));
  return CB2_OK;
}

// Device-resident flavour: src / aux / out are device pointers (beliefs that stay on the GPU between the
// convolution, the bandwidth search and the product); only the descriptor table travels, nothing is waited for.
extern "C" int cb2_approx_conv_batch_dev(cb2_ctx* ctx, const cb2_conv_desc* descs, int B, uint64_t seed) {
  if (!ctx) return CB2_ERR_INVALID;
  cb2_lock g(ctx->mu);
  CB2_REQUIRE(ctx, descs && B > 0, "null descriptors or empty batch");
  CB2_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<ConvDev> dd(B);
  int maxN = 0;
  for (int b = 0; b < B; ++b) {
    const cb2_conv_desc& c = descs[b];
    CB2_REQUIRE(ctx, c.kind >= 0 && c.kind <= 2, "convolution %d: unknown factor kind %d", b, c.kind);
    CB2_REQUIRE(ctx, c.N > 0 && c.out, "convolution %d: N <= 0 or null out", b);
    if (c.kind != CB2_FACTOR_PRIOR_POSE2) CB2_REQUIRE(ctx, c.src && c.n_src > 0, "convolution %d: missing source samples", b);
    if (c.kind == CB2_FACTOR_POSE2POINT2_BEARINGRANGE && c.reverse)
      CB2_REQUIRE(ctx, c.aux && c.n_aux > 0, "convolution %d: reverse bearing-range needs the pose's current samples", b);
    ConvDev& d = dd[b];
    d.kind = c.kind; d.reverse = c.reverse; d.N = c.N; d.n_src = c.n_src; d.n_aux = c.n_aux; d.stream = c.stream;
    memcpy(d.mean, c.mean, sizeof(d.mean)); memcpy(d.chol, c.chol, sizeof(d.chol));
    d.src = c.src; d.aux = c.aux; d.out = c.out;
    if (c.N > maxN) maxN = c.N;
  }
  int rc = cb2_arena_reserve(ctx, sizeof(ConvDev) * B + 4096);
  if (rc) return rc;
  ConvDev* tab = (ConvDev*)cb2_arena_alloc(ctx, sizeof(ConvDev) * B);
  if (!tab) return cb2_fail(ctx, CB2_ERR_NOMEM, "convolution table");
  CB2_CUDA(ctx, cudaMemcpyAsync(tab, dd.data(), sizeof(ConvDev) * B, cudaMemcpyHostToDevice, ctx->stream));
  dim3 grid((maxN + 127) / 128, B);
  approx_conv_kernel<<<grid, 128, 0, ctx->stream>>>(tab, (uint32_t)seed, (uint32_t)(seed >> 32));
  CB2_CHECK_LAUNCH(ctx, "approx_conv_kernel");
  return CB2_OK;
}
