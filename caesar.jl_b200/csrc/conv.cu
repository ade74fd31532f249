// conv.cu -- approxConvBelief for closed-form relative factors ("next" row N1): the per-particle
// factor solve that IIF runs immediately before every product (R:docs/src/principles/approxConvDensities.md:25-34,107-113),
// batched over all factors of a Gibbs colour.  One thread per (convolution, particle); FP64 throughout (the work is a
// handful of flops per particle); noise from Philox (seed; particle, draw, stream, tag 4).
#include "common.cuh"

namespace {

__host__ __device__ inline bool is_pose3(int kind) { return kind == CB2_FACTOR_PRIOR_POSE3 || kind == CB2_FACTOR_POSE3POSE3; }

struct ConvDev {
  int kind, reverse, N, n_src, n_aux;
  uint32_t stream;
  const double* src;
  const double* aux;
  double* out;
  double mean[6];
  double chol[36];
};

// ---- SO(3) helpers (FP64; the same operation order as oracle/cb2_oracle.c so that both agree to rounding) ----------------
__device__ __forceinline__ void so3_exp_dev(const double* w, double* R) {  // Rodrigues, row-major
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2);
  double A, B;
  if (th < 1e-8) { A = 1.0 - th2 / 6.0; B = 0.5 - th2 / 24.0; }
  else { A = sin(th) / th; B = (1.0 - cos(th)) / th2; }
  const double K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double k2 = 0;
      for (int k = 0; k < 3; ++k) k2 += K[i * 3 + k] * K[k * 3 + j];
      R[i * 3 + j] = (i == j ? 1.0 : 0.0) + A * K[i * 3 + j] + B * k2;
    }
}
// rotation vector of R through the unit quaternion (Shepperd's branch on the largest diagonal term: stable up to angle pi)
__device__ __forceinline__ void so3_log_dev(const double* R, double* w) {
  const double tr = R[0] + R[4] + R[8];
  double q[4];  // (w, x, y, z)
  if (tr > 0.0) {
    const double s = sqrt(tr + 1.0) * 2.0;
    q[0] = 0.25 * s; q[1] = (R[7] - R[5]) / s; q[2] = (R[2] - R[6]) / s; q[3] = (R[3] - R[1]) / s;
  } else if (R[0] > R[4] && R[0] > R[8]) {
    const double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2.0;
    q[0] = (R[7] - R[5]) / s; q[1] = 0.25 * s; q[2] = (R[1] + R[3]) / s; q[3] = (R[2] + R[6]) / s;
  } else if (R[4] > R[8]) {
    const double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2.0;
    q[0] = (R[2] - R[6]) / s; q[1] = (R[1] + R[3]) / s; q[2] = 0.25 * s; q[3] = (R[5] + R[7]) / s;
  } else {
    const double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2.0;
    q[0] = (R[3] - R[1]) / s; q[1] = (R[2] + R[6]) / s; q[2] = (R[5] + R[7]) / s; q[3] = 0.25 * s;
  }
  if (q[0] < 0.0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
  const double n = sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double f = n < 1e-12 ? 2.0 : 2.0 * atan2(n, q[0]) / n;
  w[0] = f * q[1]; w[1] = f * q[2]; w[2] = f * q[3];
}
__device__ __forceinline__ void mat3_mul(const double* A, const double* B, double* C, bool transA, bool transB) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += (transA ? A[k * 3 + i] : A[i * 3 + k]) * (transB ? B[j * 3 + k] : B[k * 3 + j]);
      C[i * 3 + j] = s;
    }
}
__device__ __forceinline__ void normal_pair(uint32_t i, uint32_t draw, uint32_t stream, uint32_t k0, uint32_t k1, double& za, double& zb) {
  uint32_t o[4];
  philox4x32_10(i, draw, stream, 4u, k0, k1, o);
  const double u1 = u01<double>(o[0], o[1]), u2 = u01<double>(o[2], o[3]);
  const double r = sqrt(-2.0 * log(u1));
  double sn, cs;
  sincospi(2.0 * u2, &sn, &cs);
  za = r * cs; zb = r * sn;
}

// Pose3 prior and Pose3Pose3 (coordinates [t; rotation vector]): X = mean + L z in the tangent space, product exponential
//   prior   : t = mean_t + n_t,  R = Exp(mean_w) Exp(n_w)            (n = L z: noise composed on the right of the mean rotation)
//   forward : t_q = t_p + R_p X_t,  R_q = R_p Exp(X_w);   reverse : R_p = R_q Exp(X_w)^T,  t_p = t_q - R_p X_t
__device__ void conv_pose3(const ConvDev& c, int i, uint32_t k0, uint32_t k1) {
  double z[6], X[6];
  for (int k = 0; k < 3; ++k) normal_pair((uint32_t)i, (uint32_t)k, c.stream, k0, k1, z[2 * k], z[2 * k + 1]);
  for (int r = 0; r < 6; ++r) {
    double v = 0;
    for (int k = 0; k < 6; ++k) v += c.chol[r * 6 + k] * z[k];
    X[r] = v;
  }
  double* out = c.out + (size_t)i * 6;
  double R1[9], R2[9], R3[9];
  if (c.kind == CB2_FACTOR_PRIOR_POSE3) {
    so3_exp_dev(c.mean + 3, R1);
    so3_exp_dev(X + 3, R2);
    mat3_mul(R1, R2, R3, false, false);
    for (int k = 0; k < 3; ++k) out[k] = c.mean[k] + X[k];
    so3_log_dev(R3, out + 3);
    return;
  }
  for (int r = 0; r < 6; ++r) X[r] += c.mean[r];
  const double* p = c.src + (size_t)(i % c.n_src) * 6;
  so3_exp_dev(p + 3, R1);
  so3_exp_dev(X + 3, R2);
  if (!c.reverse) {
    mat3_mul(R1, R2, R3, false, false);
    for (int k = 0; k < 3; ++k) out[k] = p[k] + R1[k * 3] * X[0] + R1[k * 3 + 1] * X[1] + R1[k * 3 + 2] * X[2];
  } else {
    mat3_mul(R1, R2, R3, false, true);
    for (int k = 0; k < 3; ++k) out[k] = p[k] - (R3[k * 3] * X[0] + R3[k * 3 + 1] * X[1] + R3[k * 3 + 2] * X[2]);
  }
  so3_log_dev(R3, out + 3);
}

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
  if (is_pose3(c.kind)) { conv_pose3(c, i, k0, k1); return; }
  if (c.kind == CB2_FACTOR_LINEAR_SAMPLES) {  // the caller's noise samples: x_j = x_i + z, x_i = x_j - z, or x = z without a source
    const int dd = (int)c.mean[0];
    const double* zz = c.aux + (size_t)(i % c.n_aux) * dd;
    for (int k = 0; k < dd; ++k) {
      const double base = c.n_src > 0 ? c.src[(size_t)(i % c.n_src) * dd + k] : 0.0;
      c.out[(size_t)i * dd + k] = c.reverse ? base - zz[k] : base + zz[k];
    }
    return;
  }
  double z[3], nz[3];
  normal3((uint32_t)i, c.stream, k0, k1, z);
  for (int r = 0; r < 3; ++r) nz[r] = c.mean[r] + c.chol[r * 3] * z[0] + c.chol[r * 3 + 1] * z[1] + c.chol[r * 3 + 2] * z[2];
  if (c.kind == CB2_FACTOR_PRIOR_POINT2) {
    c.out[(size_t)i * 2] = nz[0]; c.out[(size_t)i * 2 + 1] = nz[1];
  } else if (c.kind == CB2_FACTOR_POINT2POINT2_RANGE) {  // ring around the known point: nz[0] = range, direction uniform (draw 2)
    const double* p = c.src + (size_t)(i % c.n_src) * 2;
    uint32_t o[4];
    philox4x32_10((uint32_t)i, 2u, c.stream, 4u, k0, k1, o);
    double sn, cs;
    sincospi(2.0 * u01<double>(o[0], o[1]) - 1.0, &sn, &cs);
    c.out[(size_t)i * 2] = p[0] + nz[0] * cs;
    c.out[(size_t)i * 2 + 1] = p[1] + nz[0] * sn;
  } else if (c.kind == CB2_FACTOR_PRIOR_POSE2) {
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

inline bool is_point2(int kind) { return kind == CB2_FACTOR_PRIOR_POINT2 || kind == CB2_FACTOR_POINT2POINT2_RANGE; }
inline int out_dim(const cb2_conv_desc& c) {
  if (c.kind == CB2_FACTOR_LINEAR_SAMPLES) return (int)c.mean[0];
  if (is_pose3(c.kind)) return 6;
  if (is_point2(c.kind)) return 2;
  return (c.kind == CB2_FACTOR_POSE2POINT2_BEARINGRANGE && !c.reverse) ? 2 : 3;
}
inline int aux_dim(const cb2_conv_desc& c) { return c.kind == CB2_FACTOR_LINEAR_SAMPLES ? (int)c.mean[0] : 3; }
inline int src_dim(const cb2_conv_desc& c) {
  if (c.kind == CB2_FACTOR_LINEAR_SAMPLES) return (int)c.mean[0];
  if (is_pose3(c.kind)) return 6;
  if (is_point2(c.kind)) return 2;
  return (c.kind == CB2_FACTOR_POSE2POINT2_BEARINGRANGE && c.reverse) ? 2 : 3;
}
inline bool is_prior(const cb2_conv_desc& c) {
  return c.kind == CB2_FACTOR_PRIOR_POSE2 || c.kind == CB2_FACTOR_PRIOR_POSE3 || c.kind == CB2_FACTOR_PRIOR_POINT2 ||
         (c.kind == CB2_FACTOR_LINEAR_SAMPLES && c.n_src == 0);
}
inline bool needs_aux(const cb2_conv_desc& c) {
  return c.kind == CB2_FACTOR_LINEAR_SAMPLES || (c.kind == CB2_FACTOR_POSE2POINT2_BEARINGRANGE && c.reverse);
}

// residual of the relative-pose factors (ScatterAlignPose2/3, Pose2Pose2, Pose3Pose3): one thread per particle
__global__ void se_residual_kernel(int dim, const double* __restrict__ X, const double* __restrict__ p, const double* __restrict__ q,
                                   int K, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K) return;
  if (dim == 2) {
    const double* x = X + (size_t)i * 3; const double* pp = p + (size_t)i * 3; const double* qq = q + (size_t)i * 3;
    double sn, cs;
    sincos(pp[2], &sn, &cs);
    out[(size_t)i * 3] = pp[0] + cs * x[0] - sn * x[1] - qq[0];
    out[(size_t)i * 3 + 1] = pp[1] + sn * x[0] + cs * x[1] - qq[1];
    out[(size_t)i * 3 + 2] = wrap_pi(pp[2] + x[2] - qq[2]);
  } else {
    const double* x = X + (size_t)i * 6; const double* pp = p + (size_t)i * 6; const double* qq = q + (size_t)i * 6;
    double Rp[9], Rx[9], Rq[9], Rh[9], Re[9];
    so3_exp_dev(pp + 3, Rp); so3_exp_dev(x + 3, Rx); so3_exp_dev(qq + 3, Rq);
    mat3_mul(Rp, Rx, Rh, false, false);
    mat3_mul(Rq, Rh, Re, true, false);
    for (int k = 0; k < 3; ++k) out[(size_t)i * 6 + k] = pp[k] + Rp[k * 3] * x[0] + Rp[k * 3 + 1] * x[1] + Rp[k * 3 + 2] * x[2] - qq[k];
    so3_log_dev(Re, out + (size_t)i * 6 + 3);
  }
}

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
    CB2_REQUIRE(ctx, c.kind >= 0 && c.kind <= CB2_FACTOR_LINEAR_SAMPLES, "convolution %d: unknown factor kind %d", b, c.kind);
    CB2_REQUIRE(ctx, c.N > 0 && c.out, "convolution %d: N <= 0 or null out", b);
    if (!is_prior(c)) CB2_REQUIRE(ctx, c.src && c.n_src > 0, "convolution %d: missing source samples", b);
    if (needs_aux(c)) CB2_REQUIRE(ctx, c.aux && c.n_aux > 0, "convolution %d: missing aux samples (pose samples of a reverse bearing-range / noise of LINEAR_SAMPLES)", b);
    if (c.kind == CB2_FACTOR_LINEAR_SAMPLES) CB2_REQUIRE(ctx, c.mean[0] >= 1 && c.mean[0] <= 3, "convolution %d: LINEAR_SAMPLES needs mean[0] = d in 1..3", b);
    if (!is_prior(c)) total += cb2_align(sizeof(double) * (size_t)c.n_src * src_dim(c));
    if (needs_aux(c)) total += cb2_align(sizeof(double) * (size_t)c.n_aux * aux_dim(c));
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
    if (!is_prior(c)) {
      size_t nb = sizeof(double) * (size_t)c.n_src * src_dim(c);
      memcpy(stage + off, c.src, nb);
      d.src = (const double*)(p + off);
      off += cb2_align(nb);
    }
    if (needs_aux(c)) {
      size_t nb = sizeof(double) * (size_t)c.n_aux * aux_dim(c);
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
    CB2_REQUIRE(ctx, c.kind >= 0 && c.kind <= CB2_FACTOR_LINEAR_SAMPLES, "convolution %d: unknown factor kind %d", b, c.kind);
    CB2_REQUIRE(ctx, c.N > 0 && c.out, "convolution %d: N <= 0 or null out", b);
    if (!is_prior(c)) CB2_REQUIRE(ctx, c.src && c.n_src > 0, "convolution %d: missing source samples", b);
    if (needs_aux(c)) CB2_REQUIRE(ctx, c.aux && c.n_aux > 0, "convolution %d: missing aux samples (pose samples of a reverse bearing-range / noise of LINEAR_SAMPLES)", b);
    if (c.kind == CB2_FACTOR_LINEAR_SAMPLES) CB2_REQUIRE(ctx, c.mean[0] >= 1 && c.mean[0] <= 3, "convolution %d: LINEAR_SAMPLES needs mean[0] = d in 1..3", b);
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

// ---- residual of the relative-pose factors, batched over particles -----------------------------------------------------
// (cf::CalcFactor{<:ScatterAlignPose2/3})(X, p, q) (R:ext/Images/ScatterAlignPose2.jl:216-231, "copied from Pose2Pose2"):
//   q^ = p o exp_e0(X);  Xc = vee(M, q, log(M, q, q^))
// with the product exponential / logarithm of SpecialEuclidean(n) as a product manifold (translation part = difference of the
// translations, rotation part = Log(R_q^T R_q^)) -- SURVEY App. A.2 convention, unverified against Manifolds.jl (no Julia here).
// Poses and tangents are coordinates: dim 2 -> [x, y, theta], dim 3 -> [t; rotation vector].  X, p, q, out: K x ncoord (host).
extern "C" int cb2_se_residual_batch(cb2_ctx* ctx, int dim, const double* X, const double* p, const double* q, int K, double* out) {
  if (!ctx) return CB2_ERR_INVALID;
  cb2_lock g(ctx->mu);
  CB2_REQUIRE(ctx, dim == 2 || dim == 3, "residual needs dim 2 or 3 (got %d)", dim);
  CB2_REQUIRE(ctx, X && p && q && out && K > 0, "null pointer or K <= 0");
  CB2_CUDA(ctx, cudaSetDevice(ctx->device));
  const int nc = dim == 2 ? 3 : 6;
  const size_t nb = cb2_align(sizeof(double) * (size_t)K * nc);
  char* blk = (char*)cb2_io_block(ctx, 4 * nb + 256);
  if (!blk) return CB2_ERR_NOMEM;
  double* dX = (double*)blk; double* dp = (double*)(blk + nb); double* dq = (double*)(blk + 2 * nb); double* dout = (double*)(blk + 3 * nb);
  CB2_CUDA(ctx, cudaMemcpyAsync(dX, X, sizeof(double) * (size_t)K * nc, cudaMemcpyHostToDevice, ctx->stream));
  CB2_CUDA(ctx, cudaMemcpyAsync(dp, p, sizeof(double) * (size_t)K * nc, cudaMemcpyHostToDevice, ctx->stream));
  CB2_CUDA(ctx, cudaMemcpyAsync(dq, q, sizeof(double) * (size_t)K * nc, cudaMemcpyHostToDevice, ctx->stream));
  se_residual_kernel<<<(K + 127) / 128, 128, 0, ctx->stream>>>(dim, dX, dp, dq, K, dout);
  CB2_CHECK_LAUNCH(ctx, "se_residual_kernel");
  CB2_CUDA(ctx, cudaMemcpyAsync(out, dout, sizeof(double) * (size_t)K * nc, cudaMemcpyDeviceToHost, ctx->stream));
  CB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return CB2_OK;
}
