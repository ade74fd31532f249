/*
 * caesar_b200.h -- C ABI of libcaesar_b200.so: the B200 (sm_100a) implementation of Caesar.jl's
 * nonparametric belief-propagation kernel hot path.
 *
 * This is what the Julia host code binds with `ccall` (see INTEGRATION.md and caesar.jl_b200/julia/).
 * Plain C: pointers and sizes only, no C++/torch types.  Every entry point returns 0 on success or a
 * negative cb2_status; the message is available from cb2_last_error().  Nothing aborts or throws across
 * the boundary.  There is NO CPU fallback: without a usable CUDA device cb2_create fails with
 * CB2_ERR_NO_DEVICE and no other call can be made.
 *
 * Layout convention (what Julia already has in memory): every point set is Float64, column-major d x N,
 * i.e. point-contiguous -- bit-identical to `getPoints(mkd.belief)` (Matrix{Float64}) and to
 * Vector{SVector{d,Float64}} / Vector{MVector{d,Float64}}.  The library converts to its device layout.
 * Manifold points are passed as coordinates (vee(log_e(p))), as upstream does before building a KDE.
 *
 * Two flavours of every operator:
 *   cb2_xxx      : host pointers; H2D / D2H copies happen inside the call (synchronous on return).
 *   cb2_xxx_dev  : device pointers (same Float64 layouts, resident in HBM); asynchronous on the
 *                  context's stream; used by callers that keep beliefs on the device across calls.
 * Caller owns every pointer it passes, for the duration of the call only.
 *
 * Reference interfaces replaced (R: = /root/reference/):
 *   cb2_mmd              <- AMP.mmd(M, a, b, N, M, threads; bw)        call R:ext/Images/ScatterAlignPose2.jl:179
 *                           AMP.mmd(::MKD, ::MKD)                       call R:test/ICRA2022_tests/insufficient_data.jl:53
 *   cb2_mmd_se_batch     <- cost(xyr)=mmd(.., pVj(xyr), ..) in BFGS    R:ext/Images/ScatterAlignPose2.jl:167,179-182
 *                           + _transformPointCloud(;backward)           R:src/3rdParty/_PCL/services/ConsolidateRigidTransform.jl:17-54
 *   cb2_scatter_align    <- getSample(::CalcFactor{ScatterAlignPose2/3}) R:ext/Images/ScatterAlignPose2.jl:155-187
 *   cb2_kde_eval         <- (mkd::ManifoldKernelDensity)(pts)           call R:test/ICRA2022_tests/multihypo_corners.jl:66-90
 *   cb2_kde_bw_loo       <- manikde!(M, pts) without bw (LOO-CV)        call R:test/testScatterAlignPose2.jl:187
 *   cb2_kde_sample       <- AMP.sample(mkd, n)                          call R:ext/Images/ScatterAlignPose2.jl:140-141,172-173
 *   cb2_product[_batch]  <- AMP.manifoldProduct / KDE.prodAppxMSGibbsS  call R:ext/services/additional.jl:24
 *                           (batched: one call per IIF propagateBelief sweep / set of ready cliques)
 */
#ifndef CAESAR_B200_H
#define CAESAR_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct cb2_ctx cb2_ctx;

typedef enum {
  CB2_OK = 0,
  CB2_ERR_INVALID = -1,     /* bad argument (sizes, dims, null pointers) */
  CB2_ERR_CUDA = -2,        /* a CUDA runtime call or kernel failed; message has the CUDA error string */
  CB2_ERR_NO_DEVICE = -3,   /* no CUDA device / wrong architecture: there is no CPU fallback */
  CB2_ERR_NOMEM = -4,       /* device or pinned-host allocation failed */
  CB2_ERR_UNSUPPORTED = -5  /* shape outside compiled limits (d > 8; products: d > 6, N > 16384, D > 16) */
} cb2_status;

enum { CB2_FP32 = 0, CB2_FP64 = 1 }; /* arithmetic of the pair / Gibbs kernels; I/O is always Float64 */

#define CB2_MAX_DIM 8          /* pair sums, evaluate, bandwidth, sample */
#define CB2_MAX_PRODUCT_DIM 6  /* Gibbs products (Pose3 is the largest variable type upstream multiplies) */
#define CB2_MAX_DENS 16
#define CB2_MAX_PRODUCT_N 16384

/* ---- context -------------------------------------------------------------------------------- */
int cb2_create(int device, int precision, cb2_ctx** out);
void cb2_destroy(cb2_ctx* ctx);
const char* cb2_last_error(const cb2_ctx* ctx); /* ctx may be NULL: error of the last failed cb2_create */
int cb2_precision(const cb2_ctx* ctx);
/* adopt an existing CUDA stream (cudaStream_t passed as void*); NULL restores the context's own stream */
int cb2_set_stream(cb2_ctx* ctx, void* cuda_stream);
int cb2_synchronize(cb2_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t cb2_kernel_launches(const cb2_ctx* ctx);
int cb2_device_sm_count(const cb2_ctx* ctx);
/* per-stage device time of the last product batch / pair call, measured with CUDA events on the
 * context's stream.  names: "tree", "gibbs", "bw", "pairs", "h2d", "d2h".  Returns ms or <0 if unknown. */
double cb2_last_stage_ms(const cb2_ctx* ctx, const char* stage);

/* ---- mmd (K1) -------------------------------------------------------------------------------
 * out = aa/N^2 - 2 ab/(N M) + bb/M^2,  k(p,q) = exp(-bw |p-q|^2)   (biased V-statistic, bw multiplies d^2) */
int cb2_mmd(cb2_ctx* ctx, const double* a, int N, const double* b, int M, int d, double bw, double* out);
/* raw pair sums over row ranges, for row-sharded multi-GPU mmd: sums = {aa, ab, bb} restricted to
 * rows [ra0,ra1) of a (aa, ab terms) and rows [rb0,rb1) of b (bb term).  Allreduce(sum) then combine. */
int cb2_mmd_sums(cb2_ctx* ctx, const double* a, int N, const double* b, int M, int d, double bw,
                 int ra0, int ra1, int rb0, int rb1, double* sums3);
int cb2_mmd_sums_dev(cb2_ctx* ctx, const double* a_dev, int N, const double* b_dev, int M, int d, double bw,
                     int ra0, int ra1, int rb0, int rb1, double* sums3_dev);
/* shard `shard` of `nshards` of the same three sums, for multi-GPU mmd that keeps the symmetry of the self terms: the tile list
 * of the full evaluation is dealt round-robin, so every GPU evaluates (N^2/2 + N M + M^2/2) / nshards pairs; the nshards
 * results add up (ncclAllReduce of 3 doubles) to what cb2_mmd_sums returns for the full row ranges. */
int cb2_mmd_sums_shard(cb2_ctx* ctx, const double* a, int N, const double* b, int M, int d, double bw,
                       int shard, int nshards, double* sums3);
int cb2_mmd_sums_shard_dev(cb2_ctx* ctx, const double* a_dev, int N, const double* b_dev, int M, int d, double bw,
                           int shard, int nshards, double* sums3_dev);

/* mmd between beliefs on a product of lines and circles (Pose2 / Pose3 beliefs as a test metric,
 * R:test/ICRA2022_tests/multihypo_corners.jl:201-202): dist^2 = sum over Euclidean coordinates of d^2 plus
 * rot_weight * sum over circular coordinates of wrap(d)^2.  Manifolds.jl's product metric on SpecialEuclidean(2)
 * corresponds to rot_weight = 2; circular: d flags. */
int cb2_mmd_manifold(cb2_ctx* ctx, const double* a, int N, const double* b, int M, int d, const uint8_t* circular,
                     double rot_weight, double bw, double* out);

/* ---- mmd under K candidate rigid transforms of b (K2) --------------------------------------------
 * dim = 2 (coords x,y,theta) or 3 (x,y,z,rotation vector); coords is K x ncoord row-major.
 * b'_j = T(c_k) b_j, or T(c_k)^-1 b_j when backward != 0 (ScatterAlign uses backward).
 * vals[k] = mmd(a, b'); grads (K x ncoord, may be NULL) = analytic d vals[k] / d c_k. */
int cb2_mmd_se_batch(cb2_ctx* ctx, const double* a, int N, const double* b, int M, int dim, double bw,
                     const double* coords, int K, int backward, double* vals, double* grads);
int cb2_mmd_se_batch_dev(cb2_ctx* ctx, const double* a_dev, int N, const double* b_dev, int M, int dim, double bw,
                         const double* coords_host, int K, int backward, double* vals_host, double* grads_host);

/* ---- KDE evaluate / bandwidth / sample (K3, K4, K7) -------------------------------------------------
 * mu: d x N kernel centres, sigma: d std-devs, w: N weights or NULL (equal), q: d x Q queries.
 * out[j] = sum_i w_i prod_k N(q_kj - mu_ki; sigma_k^2)   (flat coordinates, plain subtraction), or its log. */
int cb2_kde_eval(cb2_ctx* ctx, const double* mu, const double* sigma, const double* w, int d, int N,
                 const double* q, int Q, int logp, double* out);
/* per-dimension leave-one-out likelihood cross-validation bandwidth (what manikde! does without bw).
 * circular: d flags or NULL.  B densities in one call: mu[b] is d x N[b]; sigma_out is B x d. */
int cb2_kde_bw_loo(cb2_ctx* ctx, const double* mu, int d, int N, const uint8_t* circular, double* sigma_out);
int cb2_kde_bw_loo_batch(cb2_ctx* ctx, const double* const* mu, const int32_t* N, int B, int d,
                         const uint8_t* circular, double* sigma_out);
/* n draws: kernel index ~ w, plus sigma_k * normal per dim (wrapped to (-pi,pi] on circular dims).
 * Counter-based Philox keyed by (seed, stream): independent of batching and sharding. */
int cb2_kde_sample(cb2_ctx* ctx, const double* mu, const double* sigma, const double* w, int d, int N,
                   const uint8_t* circular, int n, uint64_t seed, uint32_t stream, double* out);

/* ---- Gibbs-sampled products of Gaussian-kernel density messages (K5, K6) --------------------------------- */
typedef struct {
  int32_t N;          /* kernels in this message */
  const double* pts;  /* d x N */
  const double* bw;   /* d std-devs */
  const double* w;    /* N weights or NULL */
} cb2_density;

typedef struct {
  const cb2_density* dens; /* D messages */
  int32_t D;
  int32_t Nout;            /* output samples */
  int32_t niter;           /* upstream Niter (IIF passes 1): 1 + niter Gibbs sweeps per level */
  uint32_t stream;         /* Philox stream id of this product (e.g. variable index) */
  int32_t sample_offset;   /* first global sample index (output-sample sharding across GPUs) */
  double* pts_out;         /* d x Nout */
  int32_t* labels_out;     /* D x Nout selected kernel indices (0-based), or NULL */
  double* bw_out;          /* d: LOO bandwidth of the new sample set (manikde! on the result), or NULL */
} cb2_product_desc;

int cb2_product(cb2_ctx* ctx, const cb2_density* dens, int D, const uint8_t* circular, int d, int Nout, int niter,
                uint64_t seed, uint32_t stream, int sample_offset, double* pts_out, int32_t* labels_out);
/* B independent products (one per variable of a Gibbs sweep / per ready clique) in one pass.
 * All share d and the circular mask (one variable type); call once per variable type otherwise. */
int cb2_product_batch(cb2_ctx* ctx, const cb2_product_desc* descs, int B, const uint8_t* circular, int d,
                      uint64_t seed);
/* same, every pointer inside descs (pts, bw, w, pts_out, labels_out, bw_out) is a DEVICE pointer */
int cb2_product_batch_dev(cb2_ctx* ctx, const cb2_product_desc* descs, int B, const uint8_t* circular, int d,
                          uint64_t seed);
/* asynchronous pair for cooperative callers (Julia Tasks): submit enqueues the batch and returns a ticket, cb2_wait blocks
 * until that batch's outputs are in the caller's host buffers.  Every pointer inside descs must stay valid until cb2_wait.
 * submit returns before the work has finished only when the input and output arrays are PINNED host memory and no bw_out is
 * requested: copies to or from pageable memory wait for the stream, and the bandwidth search of N > 1024 results polls a
 * device counter every eight golden-section steps. */
typedef int64_t cb2_ticket;
int cb2_product_batch_submit(cb2_ctx* ctx, const cb2_product_desc* descs, int B, const uint8_t* circular, int d,
                             uint64_t seed, cb2_ticket* ticket);
int cb2_wait(cb2_ctx* ctx, cb2_ticket ticket);

/* level hierarchy of one message as the device builds it (parity tests against the oracle's tree).
 * perm: N leaf-position -> kernel index; mean/var: nodes x d per level, wt: nodes; level in [0, depth]. */
int cb2_tree_debug(cb2_ctx* ctx, const cb2_density* den, int d, int level, int32_t* depth_out, int32_t* count_out,
                   int32_t* perm, double* mean, double* var, double* wt);

/* levelDown rule of the multiscale sampler: when a chain moves one level down the tree of a message it continues from a child
 * of its current node.  mode 1 (default): the child is drawn in proportion to the children's weights (one more uniform per chain,
 * message and level); mode 0: always the first child -- deterministic, but a product of beliefs whose modes do not overlap then
 * keeps only the modes of the first branches.  Upstream's rule is not in the reference tree (KernelDensityEstimate.jl); the
 * switch exists so that either reading can be selected once real upstream output is available. */
int cb2_set_level_down(cb2_ctx* ctx, int mode);

/* diagnostic counters of the Gibbs kernel (tests).  which = 0: categorical draws since the last reset whose second pass (re-sum
 * of the selected chunk) ended short of the target because the two passes round differently, and took the chunk's last
 * candidate with mass instead -- the neighbour of the exact draw.  Waits for the stream. */
int cb2_debug_counter(cb2_ctx* ctx, int which, int reset, uint64_t* value);

/* ---- ScatterAlign getSample (a1): fresh samples from both clouds + BFGS over SE(2)/SE(3) coords ---------------
 * cloud1/cloud2: KDE beliefs over R^dim (dim = 2 | 3); sample_count points are drawn from each
 * (R:ext/Images/ScatterAlignPose2.jl:171-174), then mmd(smp1, T(c)^-1 smp2) is minimised from the start point
 * x0 (ncoord values; the reference draws [5 randn(dim); 0.1 randn(nrot)], :182).  K independent alignments
 * (one per particle of the solver) run as one batched BFGS: x0 is K x ncoord, coords_out K x ncoord,
 * score_out K (final mmd, the reference's cache.score).  iters_out (K, may be NULL) = BFGS iterations. */
int cb2_scatter_align(cb2_ctx* ctx, const cb2_density* cloud1, const cb2_density* cloud2, int dim,
                      int sample_count, double bw, const double* x0, int K, uint64_t seed,
                      int max_iter, double g_tol, double* coords_out, double* score_out, int32_t* iters_out);

/* ---- approxConvBelief for closed-form relative factors ("next" row N1) ---------------------------
 * Samples of the target variable implied by one factor and the current samples of its other variable -- the step IIF
 * runs immediately before every product (R:docs/src/principles/approxConvDensities.md:25-34,107-113).  For the group
 * factors below the per-particle solve is closed form: q = p o exp(z + noise) on SE(2) with the product exponential
 * (same convention as ConsolidateRigidTransform.jl:32-36), l = t_p + r [cos(th_p + b), sin(th_p + b)] for bearing-range.
 * Noise comes from Philox (seed; particle, draw, stream, tag 4): reproducible and independent of batching.
 * B convolutions (all factors of all variables of one Gibbs colour) run as one launch. */
typedef enum {
  CB2_FACTOR_PRIOR_POSE2 = 0,               /* target ~ N(mean, L L^T), no source */
  CB2_FACTOR_POSE2POSE2 = 1,                /* forward: x_j = x_i o exp(z);  reverse: x_i = x_j o exp(z)^-1 */
  CB2_FACTOR_POSE2POINT2_BEARINGRANGE = 2,  /* forward: landmark from pose;  reverse: pose position from landmark, heading from aux */
  CB2_FACTOR_PRIOR_POSE3 = 3,               /* Pose3 coordinates [t; rotation vector]: t = mean_t + n_t, R = Exp(mean_w) Exp(n_w) */
  CB2_FACTOR_POSE3POSE3 = 4,                /* forward: x_j = x_i o exp(z) on SE(3);  reverse: x_i = x_j o exp(z)^-1 */
  CB2_FACTOR_PRIOR_POINT2 = 5,              /* target ~ N(mean[0..1], L L^T), L = chol[0], chol[3], chol[4] (3 x 3 layout), no source */
  CB2_FACTOR_POINT2POINT2_RANGE = 6,        /* range-only (R:test/ICRA2022_tests/insufficient_data.jl:70-85): x_j = x_i + rho (cos t, sin t),
                                               rho ~ N(mean[0], chol[0]^2), t uniform on the circle -- one constraint on two coordinates,
                                               the same in both directions */
  CB2_FACTOR_LINEAR_SAMPLES = 7             /* Prior / LinearRelative on a Euclidean variable of d = mean[0] coordinates (1..3) whose noise the
                                               CALLER sampled (any distribution: Rayleigh, Uniform, Mixture ..., as the factors of
                                               R:test/ICRA2022_tests/nongaussian_mixture.jl:24-51): aux = d x n_aux samples z;
                                               forward x_j = x_i + z, reverse x_i = x_j - z, no source (n_src = 0): x = z */
} cb2_factor_kind;

typedef struct {
  int32_t kind;        /* cb2_factor_kind */
  int32_t reverse;     /* 0: solve the factor's second variable from its first; 1: the other way round */
  int32_t N;           /* particles to produce */
  int32_t n_src;       /* samples available in src (particle i uses src[i % n_src]) */
  const double* src;   /* d_src x n_src samples of the known variable; NULL for a prior */
  const double* aux;   /* reverse bearing-range: 3 x n_aux current samples of the target pose (heading is kept); LINEAR_SAMPLES: the noise */
  int32_t n_aux;
  uint32_t stream;     /* Philox stream id of this convolution */
  double mean[6];      /* Pose2 prior / Pose2Pose2: [x, y, theta]; bearing-range: [bearing, range, 0]; Pose3 kinds: [t; rotation vector] */
  double chol[36];     /* lower Cholesky factor L of the noise covariance, row-major n x n in the first n*n entries (n = 3, Pose3
                          kinds: n = 6); bearing-range: L = diag(sigma_b, sigma_r, 0) */
  double* out;         /* d_tgt x N (Pose2: 3, Point2: 2, Pose3: 6) */
} cb2_conv_desc;

int cb2_approx_conv_batch(cb2_ctx* ctx, const cb2_conv_desc* descs, int B, uint64_t seed);

/* ---- residual of the relative-pose factors, K particles per call ----------------------------------------------------
 * (cf::CalcFactor{<:ScatterAlignPose2/3})(X, p, q) (R:ext/Images/ScatterAlignPose2.jl:216-231; the body is RoME's Pose2Pose2
 * residual): q^ = p o exp_e0(X), out = vee(M, q, log(M, q, q^)) with the product exponential / logarithm of SpecialEuclidean(n).
 * dim = 2: coordinates [x, y, theta]; dim = 3: [t; rotation vector].  X, p, q, out: K x ncoord host arrays. */
int cb2_se_residual_batch(cb2_ctx* ctx, int dim, const double* X, const double* p, const double* q, int K, double* out);

/* ---- device-resident beliefs ------------------------------------------------
 * Across a Gibbs sweep IIF runs, per variable, approxConvBelief -> manikde! -> manifoldProduct -> manikde! and feeds the
 * result to the next factor (R:docs/src/principles/approxConvDensities.md:25-34; IIF propagateBelief).  With the belief's
 * points (d x N Float64) and bandwidths (d Float64) kept in device buffers, that whole chain is a sequence of stream-ordered
 * calls that never wait for the host; only cb2_download (PackedManifoldKernelDensity / getPoints on the Julia side) does.
 * cb2_product_batch_dev with device bandwidths, cb2_mmd_sums_dev and cb2_mmd_se_batch_dev complete the set.
 * Device bandwidths are not validated on the host (a non-positive value yields NaN samples instead of an error code). */
int cb2_device_alloc(cb2_ctx* ctx, size_t bytes, void** out);
int cb2_device_free(cb2_ctx* ctx, void* ptr);
int cb2_upload(cb2_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);    /* stream-ordered, does not wait */
int cb2_download(cb2_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);  /* waits for the stream */
/* cb2_approx_conv_batch with src / aux / out as device pointers */
int cb2_approx_conv_batch_dev(cb2_ctx* ctx, const cb2_conv_desc* descs, int B, uint64_t seed);
/* cb2_kde_bw_loo_batch with mu_dev[b] (host array of device pointers) and sigma_out_dev (B x d, device) */
int cb2_kde_bw_loo_batch_dev(cb2_ctx* ctx, const double* const* mu_dev, const int32_t* N, int B, int d,
                             const uint8_t* circular, double* sigma_out_dev);

/* ---- ICP branch of ScatterAlign ------------------------------------------------------
 * getSample(cf::CalcFactor{<:ScatterAlignPose3}) with sample_count < 0 (R:ext/Images/ScatterAlignPose2.jl:187-214) perturbs
 * cloud2 by `dstb = 0.2 randn(6)` (backward transform, :195-196) and runs `_PCL.alignICP_Simple(p_ptsM, phat_pts_mov;
 * max_iterations = 25, correspondences = 500, neighbors = 50)` (R:src/3rdParty/_PCL/services/ICP_Simple.jl:239-332): point-to-plane
 * ICP with PCA normals on `correspondences` evenly spaced points of the fixed cloud, median/MAD rejection and the planarity
 * gate `min_planarity`, stopping when mean and std of the residuals change by less than `min_change` percent.
 * K particles (one perturbation each; perturb = K x 6 tangent coordinates or NULL) run as one batch; the normals are computed
 * once.  fix: nf x 3, mov: nm x 3 (host, row-major points).  H_out: K x 16, the 4 x 4 fix_H_mov of every particle (row-major);
 * stats_out (nullable): K x 4 = [mean(residuals), std(residuals), correspondences kept, iterations (negative: the normal
 * equations became singular in that iteration)].  Limits: correspondences <= 1024, neighbors <= 256. */
int cb2_icp_align_batch(cb2_ctx* ctx, const double* fix, int nf, const double* mov, int nm, const double* perturb, int K,
                        int correspondences, int neighbors, double min_planarity, double min_change, int max_iterations,
                        double* H_out, double* stats_out);

#ifdef __cplusplus
}
#endif
#endif
