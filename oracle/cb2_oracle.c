/*
 * cb2_oracle.c -- CPU FP64 restatement (TEST INFRASTRUCTURE ONLY; see cb2_oracle.h header).
 *
 * Written from the algorithm descriptions in SURVEY.md App. A (which restate the published algorithms of
 * KernelDensityEstimate.jl 0.5 / ApproxManifoldProducts.jl 0.7-0.8, themselves ports of A. Ihler's KDE
 * toolbox) and from the in-tree reference call sites cited per function.  No reference source is copied.
 */
#include "cb2_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TWO_PI 6.283185307179586476925286766559
#define MAXD 16

int cb2o_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------ Philox4x32-10 */
void cb2o_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static double u_from_words(uint32_t x0, uint32_t x1, int ubits) {
  if (ubits == 24) return ((double)(x0 >> 9) + 0.5) * (1.0 / 8388608.0); /* FP32-exact: 23 bits + 1/2 */
  uint64_t v = (((uint64_t)x0 << 32) | x1) >> 11; /* 53 bits */
  return ((double)v + 0.5) * (1.0 / 9007199254740992.0);
}

double cb2o_uniform(uint64_t seed, uint32_t sample, uint32_t draw, uint32_t stream, uint32_t tag, int ubits) {
  uint32_t ctr[4] = {sample, draw, stream, tag}, key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, o[4];
  cb2o_philox4x32(ctr, key, o);
  return u_from_words(o[0], o[1], ubits);
}

double cb2o_normal(uint64_t seed, uint32_t sample, uint32_t i, uint32_t stream, uint32_t tag, int ubits) {
  uint32_t ctr[4] = {sample, i >> 1, stream, tag}, key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, o[4];
  cb2o_philox4x32(ctr, key, o);
  double u1 = u_from_words(o[0], o[1], ubits), u2 = u_from_words(o[2], o[3], ubits);
  double r = sqrt(-2.0 * log(u1));
  return (i & 1) ? r * sin(TWO_PI * u2) : r * cos(TWO_PI * u2);
}

/* ------------------------------------------------------------------ mmd (App. A.1) */
static double pairsum(const double* p, int r0, int r1, const double* q, int M, int d, double bw) {
  double tot = 0.0;
#pragma omp parallel for reduction(+ : tot) schedule(static)
  for (int i = r0; i < r1; ++i) {
    const double* pi = p + (size_t)i * d;
    double s = 0.0;
    for (int j = 0; j < M; ++j) {
      const double* qj = q + (size_t)j * d;
      double r2 = 0.0;
      for (int k = 0; k < d; ++k) { double df = pi[k] - qj[k]; r2 += df * df; }
      s += exp(-bw * r2);
    }
    tot += s;
  }
  return tot;
}

int cb2o_mmd_rows(const double* a, int N, const double* b, int M, int d, double bw,
                  int ra0, int ra1, int rb0, int rb1, double* sums) {
  if (N <= 0 || M <= 0 || d <= 0 || d > MAXD) return -1;
  sums[0] = pairsum(a, ra0, ra1, a, N, d, bw);
  sums[1] = pairsum(a, ra0, ra1, b, M, d, bw);
  sums[2] = pairsum(b, rb0, rb1, b, M, d, bw);
  return 0;
}

int cb2o_mmd(const double* a, int N, const double* b, int M, int d, double bw, double* out, double* sums) {
  double s[3];
  int rc = cb2o_mmd_rows(a, N, b, M, d, bw, 0, N, 0, M, s);
  if (rc) return rc;
  *out = s[0] / ((double)N * N) - 2.0 * s[1] / ((double)N * M) + s[2] / ((double)M * M);
  if (sums) { sums[0] = s[0]; sums[1] = s[1]; sums[2] = s[2]; }
  return 0;
}

/* ------------------------------------------------------------------ rigid transforms
 * ConsolidateRigidTransform.jl:32-36: aPb = retract(M, e0, hat(M,e0,c)); backward -> inv; p' = R p + t.
 * exp at the identity of SpecialEuclidean(n) in Manifolds 0.8/0.9 is the product-manifold exponential:
 * translation = c[0:n], rotation = Exp_SO(n)(c[n:]) (SURVEY sec. 7 "keep upstream choices switchable"). */
static void so3_exp(const double* w, double* R) {
  double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2), A, B;
  if (th < 1e-8) { A = 1.0 - th2 / 6.0; B = 0.5 - th2 / 24.0; }
  else { A = sin(th) / th; B = (1.0 - cos(th)) / th2; }
  double K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0}, K2[9];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    double s = 0; for (int k = 0; k < 3; ++k) s += K[i * 3 + k] * K[k * 3 + j]; K2[i * 3 + j] = s; }
  for (int i = 0; i < 9; ++i) R[i] = A * K[i] + B * K2[i];
  R[0] += 1; R[4] += 1; R[8] += 1;
}
/* right Jacobian of SO(3): R(w+dw) ~= R(w) Exp(Jr dw) */
static void so3_jr(const double* w, double* J) {
  double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2), B, C;
  if (th < 1e-6) { B = 0.5 - th2 / 24.0; C = 1.0 / 6.0 - th2 / 120.0; }
  else { B = (1.0 - cos(th)) / th2; C = (th - sin(th)) / (th2 * th); }
  double K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0}, K2[9];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    double s = 0; for (int k = 0; k < 3; ++k) s += K[i * 3 + k] * K[k * 3 + j]; K2[i * 3 + j] = s; }
  for (int i = 0; i < 9; ++i) J[i] = -B * K[i] + C * K2[i];
  J[0] += 1; J[4] += 1; J[8] += 1;
}

void cb2o_se_matrix(int dim, const double* c, int backward, double* R, double* t) {
  double Rf[9];
  if (dim == 2) { double cs = cos(c[2]), sn = sin(c[2]); Rf[0] = cs; Rf[1] = -sn; Rf[2] = sn; Rf[3] = cs; }
  else so3_exp(c + 3, Rf);
  if (!backward) { memcpy(R, Rf, sizeof(double) * dim * dim); for (int i = 0; i < dim; ++i) t[i] = c[i]; return; }
  for (int i = 0; i < dim; ++i) for (int j = 0; j < dim; ++j) R[i * dim + j] = Rf[j * dim + i];
  for (int i = 0; i < dim; ++i) { double s = 0; for (int j = 0; j < dim; ++j) s += R[i * dim + j] * c[j]; t[i] = -s; }
}

void cb2o_transform_points(int dim, const double* c, int backward, const double* pts, int n, double* out) {
  double R[9], t[3];
  cb2o_se_matrix(dim, c, backward, R, t);
  for (int p = 0; p < n; ++p)
    for (int i = 0; i < dim; ++i) {
      double s = t[i];
      for (int j = 0; j < dim; ++j) s += R[i * dim + j] * pts[(size_t)p * dim + j];
      out[(size_t)p * dim + i] = s;
    }
}

/* cost(c) = mmd(a, T(c) b) -- ext/Images/ScatterAlignPose2.jl:167,179.  aa and bb do not depend on c.
 * gradient: F_j = sum_i k_ij (a_i - b'_j);  dS/dc = 2 bw sum_j F_j . db'_j/dc ;  dcost = -2/(NM) dS. */
int cb2o_mmd_se_batch(const double* a, int N, const double* b, int M, int dim, double bw,
                      const double* coords, int K, int backward, double* vals, double* grads) {
  if (dim != 2 && dim != 3) return -1;
  int nc = dim == 2 ? 3 : 6;
  double aa = pairsum(a, 0, N, a, N, dim, bw), bb = pairsum(b, 0, M, b, M, dim, bw);
  double* bt = (double*)malloc(sizeof(double) * (size_t)M * dim);
  for (int k = 0; k < K; ++k) {
    const double* c = coords + (size_t)k * nc;
    double R[9], t[3];
    cb2o_se_matrix(dim, c, backward, R, t);
    cb2o_transform_points(dim, c, backward, b, M, bt);
    double S = 0, SF[3] = {0, 0, 0}, SX[3] = {0, 0, 0}; /* SX = sum_j b'_j x F_j */
#pragma omp parallel for schedule(static)
    for (int j = 0; j < M; ++j) {
      const double* q = bt + (size_t)j * dim;
      double s = 0, F[3] = {0, 0, 0};
      for (int i = 0; i < N; ++i) {
        const double* p = a + (size_t)i * dim;
        double r2 = 0, df[3] = {0, 0, 0};
        for (int x = 0; x < dim; ++x) { df[x] = p[x] - q[x]; r2 += df[x] * df[x]; }
        double kv = exp(-bw * r2);
        s += kv;
        for (int x = 0; x < dim; ++x) F[x] += kv * df[x];
      }
      double X[3] = {0, 0, 0};
      if (dim == 2) X[2] = q[0] * F[1] - q[1] * F[0];
      else { X[0] = q[1] * F[2] - q[2] * F[1]; X[1] = q[2] * F[0] - q[0] * F[2]; X[2] = q[0] * F[1] - q[1] * F[0]; }
#pragma omp critical
      { S += s; for (int x = 0; x < 3; ++x) { SF[x] += F[x]; SX[x] += X[x]; } }
    }
    vals[k] = aa / ((double)N * N) - 2.0 * S / ((double)N * M) + bb / ((double)M * M);
    if (grads) {
      double* g = grads + (size_t)k * nc;
      double sc = -2.0 / ((double)N * M) * 2.0 * bw;
      if (dim == 2) {
        if (backward) { /* R holds R_f^T: b' = R (b - c_t): db'/dc_t = -R, db'/dth = -J b' */
          for (int x = 0; x < 2; ++x) g[x] = sc * (-(R[0 * 2 + x] * SF[0] + R[1 * 2 + x] * SF[1]));
          g[2] = sc * (-SX[2]);
        } else { /* b' = R b + t: db'/dt = I, db'/dth = J (b' - t) -> (b'-t) x F */
          g[0] = sc * SF[0]; g[1] = sc * SF[1];
          g[2] = sc * (SX[2] - (t[0] * SF[1] - t[1] * SF[0]));
        }
      } else {
        double Jr[9], v[3];
        so3_jr(c + 3, Jr);
        if (backward) { /* db'/dt = -R_f^T = -R ; db'/dw = [b']x Jr  ->  g_w = Jr^T sum(F x b') = -Jr^T SX */
          for (int x = 0; x < 3; ++x) g[x] = sc * (-(R[0 * 3 + x] * SF[0] + R[1 * 3 + x] * SF[1] + R[2 * 3 + x] * SF[2]));
          for (int x = 0; x < 3; ++x) v[x] = -SX[x];
        } else { /* db'/dw = -R [b]x Jr -> g_w = Jr^T R^T sum((b'-t) x F) */
          for (int x = 0; x < 3; ++x) g[x] = sc * SF[x];
          double u[3] = {SX[0] - (t[1] * SF[2] - t[2] * SF[1]), SX[1] - (t[2] * SF[0] - t[0] * SF[2]),
                         SX[2] - (t[0] * SF[1] - t[1] * SF[0])};
          for (int x = 0; x < 3; ++x) v[x] = R[0 * 3 + x] * u[0] + R[1 * 3 + x] * u[1] + R[2 * 3 + x] * u[2];
        }
        for (int x = 0; x < 3; ++x) g[3 + x] = sc * (Jr[0 * 3 + x] * v[0] + Jr[1 * 3 + x] * v[1] + Jr[2 * 3 + x] * v[2]);
      }
    }
  }
  free(bt);
  return 0;
}

/* ------------------------------------------------------------------ KDE evaluate (App. A.2.3-4)
 * p(x) = sum_i w_i prod_d N(x_d - mu_id; sigma_d^2), flat coordinates, plain subtraction. */
int cb2o_kde_eval(const double* mu, const double* sigma, const double* w, int d, int N,
                  const double* q, int Q, int logp, double* out) {
  if (d <= 0 || d > MAXD || N <= 0) return -1;
  double lognorm = 0, inv2[MAXD];
  for (int k = 0; k < d; ++k) { lognorm -= 0.5 * log(TWO_PI * sigma[k] * sigma[k]); inv2[k] = 0.5 / (sigma[k] * sigma[k]); }
  double wsum = 0;
  if (w) for (int i = 0; i < N; ++i) wsum += w[i];
#pragma omp parallel for schedule(static)
  for (int j = 0; j < Q; ++j) {
    const double* x = q + (size_t)j * d;
    double s = 0;
    for (int i = 0; i < N; ++i) {
      const double* m = mu + (size_t)i * d;
      double e = 0;
      for (int k = 0; k < d; ++k) { double df = x[k] - m[k]; e += df * df * inv2[k]; }
      s += (w ? w[i] / wsum : 1.0 / N) * exp(-e);
    }
    out[j] = logp ? log(s) + lognorm : s * exp(lognorm);
  }
  return 0;
}

/* ------------------------------------------------------------------ LOO bandwidth (App. A.2.5)
 * Per dimension: 1-D KDE, sigma by leave-one-out likelihood CV, golden-section search (tol 1e-2) on
 *   nll(h) = -(1/N) sum_i log( 1/(N-1) sum_{j!=i} N(x_i - x_j; h^2) )
 * bracket (ax,bx,cx) = (2 m/(1+sqrt5), m, 2 r): m = max(min neighbour gap, 1e-6), r = data range
 * (Ihler ksize 'lcv' uses the smallest internal-node / root ball widths of the 1-D ball tree; in 1-D
 * these are a neighbour gap and the range).  Pinned against the tut3 fixtures (see header). */
static double wrap_pi(double x) { return x - TWO_PI * rint(x / TWO_PI); }

double cb2o_loo_nll_1d(const double* x, int N, int stride, double h, int circular) {
  double c = -0.5 / (h * h), lognorm = -0.5 * log(TWO_PI * h * h) - log((double)(N - 1));
  double tot = 0;
#pragma omp parallel for reduction(+ : tot) schedule(static)
  for (int i = 0; i < N; ++i) {
    double xi = x[(size_t)i * stride], s = 0;
    for (int j = 0; j < N; ++j) {
      if (j == i) continue;
      double df = xi - x[(size_t)j * stride];
      if (circular) df = wrap_pi(df);
      s += exp(c * df * df);
    }
    tot += log(s > DBL_MIN ? s : DBL_MIN) + lognorm;
  }
  return -tot / N;
}

static int cmp_double(const void* a, const void* b) {
  double x = *(const double*)a, y = *(const double*)b;
  return (x > y) - (x < y);
}

int cb2o_kde_bw_loo(const double* mu, int d, int N, const uint8_t* circular, double* sigma_out) {
  if (N < 2 || d <= 0) return -1;
  double* xs = (double*)malloc(sizeof(double) * N);
  for (int k = 0; k < d; ++k) {
    int circ = circular ? circular[k] : 0;
    for (int i = 0; i < N; ++i) xs[i] = mu[(size_t)i * d + k];
    qsort(xs, N, sizeof(double), cmp_double);
    double mn = DBL_MAX;
    for (int i = 1; i < N; ++i) { double g = xs[i] - xs[i - 1]; if (g < mn) mn = g; }
    if (mn < 1e-6) mn = 1e-6;
    double range = xs[N - 1] - xs[0];
    if (range < 1e-6) { sigma_out[k] = 1e-6; continue; }
    const double C = (3.0 - sqrt(5.0)) / 2.0, Rr = 1.0 - C;
    double ax = 2.0 * mn / (1.0 + sqrt(5.0)), bx = mn, cx = 2.0 * range;
    double x0 = ax, x3 = cx, x1, x2;
    if (fabs(cx - bx) > fabs(bx - ax)) { x1 = bx; x2 = bx + C * (cx - bx); }
    else { x2 = bx; x1 = bx - C * (bx - ax); }
    double f1 = cb2o_loo_nll_1d(mu + k, N, d, x1, circ), f2 = cb2o_loo_nll_1d(mu + k, N, d, x2, circ);
    while (fabs(x3 - x0) > 1e-2 * (fabs(x1) + fabs(x2))) {
      if (f2 < f1) { x0 = x1; x1 = x2; x2 = Rr * x1 + C * x3; f1 = f2; f2 = cb2o_loo_nll_1d(mu + k, N, d, x2, circ); }
      else { x3 = x2; x2 = x1; x1 = Rr * x2 + C * x0; f2 = f1; f1 = cb2o_loo_nll_1d(mu + k, N, d, x1, circ); }
    }
    sigma_out[k] = f1 < f2 ? x1 : x2;
  }
  free(xs);
  return 0;
}

/* ------------------------------------------------------------------ sample (App. A.2.6) */
int cb2o_kde_sample(const double* mu, const double* sigma, const double* w, int d, int N,
                    const uint8_t* circular, int n, uint64_t seed, uint32_t stream, int ubits, double* out) {
  if (d <= 0 || d > MAXD || N <= 0) return -1;
  double wsum = 0;
  if (w) for (int i = 0; i < N; ++i) wsum += w[i];
  for (int s = 0; s < n; ++s) {
    double u = cb2o_uniform(seed, (uint32_t)s, 0, stream, 2, ubits);
    int idx;
    if (!w) { idx = (int)(u * N); if (idx >= N) idx = N - 1; }
    else {
      double tgt = u * wsum, cs = 0; idx = N - 1;
      for (int i = 0; i < N; ++i) { cs += w[i]; if (cs >= tgt) { idx = i; break; } }
    }
    for (int k = 0; k < d; ++k) {
      double z = cb2o_normal(seed, (uint32_t)s, (uint32_t)k, stream, 3, ubits);
      double v = mu[(size_t)idx * d + k] + sigma[k] * z;
      if (circular && circular[k]) v = wrap_pi(v);
      out[(size_t)s * d + k] = v;
    }
  }
  return 0;
}

/* ------------------------------------------------------------------ ball-tree level hierarchy (App. A.3)
 * Binary tree by recursive median split on the widest coordinate (range = max-min, ties -> lowest dim),
 * left child takes ceil(n/2) points, the ordering inside a node is a stable sort on that coordinate.
 * Node statistic = weight, weighted mean, moment-matched variance incl. the kernel bandwidth.
 * Stored by level: level l has min(2^l, N) nodes for l < depth, and N (the kernels) at l = depth. */
struct cb2o_tree {
  int N, d, depth;
  int32_t* perm;
  int* cnt;      /* depth+1 */
  int** start;   /* [l][cnt+1] leaf-position ranges */
  double** mean; /* [l][cnt*d] */
  double** var;
  double** wt;
};

typedef struct { double v; int pos; int idx; } sortrec;
static int cmp_sortrec(const void* a, const void* b) {
  const sortrec *x = (const sortrec*)a, *y = (const sortrec*)b;
  if (x->v < y->v) return -1;
  if (x->v > y->v) return 1;
  return (x->pos > y->pos) - (x->pos < y->pos);
}

static int ceil_log2(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }

cb2o_tree* cb2o_tree_build(const double* pts, const double* bw, const double* w, int d, int N) {
  cb2o_tree* t = (cb2o_tree*)calloc(1, sizeof(cb2o_tree));
  t->N = N; t->d = d; t->depth = ceil_log2(N);
  int L = t->depth + 1;
  t->perm = (int32_t*)malloc(sizeof(int32_t) * N);
  t->cnt = (int*)malloc(sizeof(int) * L);
  t->start = (int**)calloc(L, sizeof(int*));
  t->mean = (double**)calloc(L, sizeof(double*));
  t->var = (double**)calloc(L, sizeof(double*));
  t->wt = (double**)calloc(L, sizeof(double*));
  for (int i = 0; i < N; ++i) t->perm[i] = i;
  t->cnt[0] = 1;
  t->start[0] = (int*)malloc(sizeof(int) * 2);
  t->start[0][0] = 0; t->start[0][1] = N;
  sortrec* rec = (sortrec*)malloc(sizeof(sortrec) * N);
  for (int l = 0; l < t->depth; ++l) {
    int cnt = t->cnt[l];
    int ncnt = (l + 1 == t->depth) ? N : 2 * cnt;
    t->cnt[l + 1] = ncnt;
    t->start[l + 1] = (int*)malloc(sizeof(int) * (ncnt + 1));
    int out = 0;
    for (int k = 0; k < cnt; ++k) {
      int s = t->start[l][k], e = t->start[l][k + 1], n = e - s;
      if (n >= 2) {
        int best = 0; double bestr = -1;
        for (int i = 0; i < d; ++i) {
          double lo = DBL_MAX, hi = -DBL_MAX;
          for (int p = s; p < e; ++p) { double v = pts[(size_t)t->perm[p] * d + i]; if (v < lo) lo = v; if (v > hi) hi = v; }
          if (hi - lo > bestr) { bestr = hi - lo; best = i; }
        }
        for (int p = s; p < e; ++p) { rec[p - s].v = pts[(size_t)t->perm[p] * d + best]; rec[p - s].pos = p; rec[p - s].idx = t->perm[p]; }
        qsort(rec, n, sizeof(sortrec), cmp_sortrec);
        for (int p = s; p < e; ++p) t->perm[p] = rec[p - s].idx;
        t->start[l + 1][out++] = s;
        t->start[l + 1][out++] = s + (n + 1) / 2;
      } else {
        t->start[l + 1][out++] = s;
      }
    }
    t->start[l + 1][out] = N;
  }
  /* statistics, bottom-up */
  double wsum = 0;
  if (w) for (int i = 0; i < N; ++i) wsum += w[i];
  for (int l = t->depth; l >= 0; --l) {
    int cnt = t->cnt[l];
    t->mean[l] = (double*)malloc(sizeof(double) * cnt * d);
    t->var[l] = (double*)malloc(sizeof(double) * cnt * d);
    t->wt[l] = (double*)malloc(sizeof(double) * cnt);
    if (l == t->depth) {
      for (int k = 0; k < cnt; ++k) {
        int idx = t->perm[k];
        t->wt[l][k] = w ? w[idx] / wsum : 1.0 / N;
        for (int i = 0; i < d; ++i) { t->mean[l][k * d + i] = pts[(size_t)idx * d + i]; t->var[l][k * d + i] = bw[i] * bw[i]; }
      }
    } else {
      int lc = l + 1;
      for (int k = 0; k < cnt; ++k) {
        /* children of k at level lc: nodes whose start lies in [start[l][k], start[l][k+1]) */
        int c0 = (lc == t->depth) ? t->start[l][k] : 2 * k;
        int nch = (lc == t->depth) ? (t->start[l][k + 1] - t->start[l][k]) : 2;
        double ws = 0;
        for (int c = 0; c < nch; ++c) ws += t->wt[lc][c0 + c];
        t->wt[l][k] = ws;
        for (int i = 0; i < d; ++i) {
          double m = 0;
          for (int c = 0; c < nch; ++c) m += t->wt[lc][c0 + c] * t->mean[lc][(c0 + c) * d + i];
          m /= ws;
          double v = 0;
          for (int c = 0; c < nch; ++c) { double dm = t->mean[lc][(c0 + c) * d + i] - m; v += t->wt[lc][c0 + c] * (t->var[lc][(c0 + c) * d + i] + dm * dm); }
          t->mean[l][k * d + i] = m; t->var[l][k * d + i] = v / ws;
        }
      }
    }
  }
  free(rec);
  return t;
}

void cb2o_tree_free(cb2o_tree* t) {
  if (!t) return;
  for (int l = 0; l <= t->depth; ++l) { free(t->start[l]); free(t->mean[l]); free(t->var[l]); free(t->wt[l]); }
  free(t->start); free(t->mean); free(t->var); free(t->wt); free(t->cnt); free(t->perm); free(t);
}
int cb2o_tree_depth(const cb2o_tree* t) { return t->depth; }
int cb2o_tree_level_count(const cb2o_tree* t, int l) { return t->cnt[l]; }
const int32_t* cb2o_tree_perm(const cb2o_tree* t) { return t->perm; }
const double* cb2o_tree_level_mean(const cb2o_tree* t, int l) { return t->mean[l]; }
const double* cb2o_tree_level_var(const cb2o_tree* t, int l) { return t->var[l]; }
const double* cb2o_tree_level_wt(const cb2o_tree* t, int l) { return t->wt[l]; }

/* ------------------------------------------------------------------ multiscale Gibbs product (App. A.3) */
static void loo_product(const cb2o_tree* const* tr, const int* lev, const int* ind, int D, int d, int skip,
                        const uint8_t* circ, double* Mo, double* Co) {
  for (int i = 0; i < d; ++i) {
    double lam = 0, sx = 0, cxs = 0, sm = 0;
    for (int k = 0; k < D; ++k) {
      if (k == skip) continue;
      double v = tr[k]->var[lev[k]][ind[k] * d + i], m = tr[k]->mean[lev[k]][ind[k] * d + i], l = 1.0 / v;
      lam += l;
      if (circ && circ[i]) { sx += l * sin(m); cxs += l * cos(m); } else sm += l * m;
    }
    Co[i] = 1.0 / lam;
    Mo[i] = (circ && circ[i]) ? atan2(sx, cxs) : sm / lam;
  }
}

/* levelDown rule (SURVEY App. A.3 leaves it open: "ind[j] <- a child of ind[j]"): 1 = child drawn in proportion to the
 * children's weights (default), 0 = always the first child.  Process-wide switch for the day real upstream output is at hand. */
static int g_level_down = 1;
void cb2o_set_level_down(int mode) { g_level_down = mode ? 1 : 0; }
int cb2o_get_level_down(void) { return g_level_down; }

int cb2o_product(const cb2o_density* dens, int D, const uint8_t* circ, int d, int Nout, int niter,
                 uint64_t seed, uint32_t stream, int ubits, int sample_offset,
                 double* pts_out, int32_t* labels_out) {
  if (D < 1 || d <= 0 || d > MAXD) return -1;
  cb2o_tree** tr = (cb2o_tree**)malloc(sizeof(cb2o_tree*) * D);
  int maxN = 0, maxcnt = 0;
  for (int j = 0; j < D; ++j) {
    tr[j] = cb2o_tree_build(dens[j].pts, dens[j].bw, dens[j].w, d, dens[j].N);
    if (dens[j].N > maxN) maxN = dens[j].N;
  }
  maxcnt = maxN;
  int L = 0; while ((2 << L) <= maxN) ++L; L += 1; /* floor(log2(maxN)) + 1 */
  int sweeps = 1 + niter;
#pragma omp parallel
  {
    int* ind = (int*)malloc(sizeof(int) * D);
    int* lev = (int*)malloc(sizeof(int) * D);
    double* e = (double*)malloc(sizeof(double) * maxcnt);
#pragma omp for schedule(dynamic, 8)
    for (int s = 0; s < Nout; ++s) {
      uint32_t sid = (uint32_t)(s + sample_offset);
      for (int j = 0; j < D; ++j) { ind[j] = 0; lev[j] = 0; }
      for (int step = 1; step <= L; ++step) {
        for (int j = 0; j < D; ++j) {
          /* levelDown: a leaf stays itself; else move to one of the node's children, drawn in proportion to their weights.
           * Upstream pre-generates (Niter + 2) uniforms per (sample, density, level) (randU[N D (Niter+2) L]):
           * 1 + Niter for the Gibbs sweeps and one more, which is this draw.  Always taking the first child instead would pin
           * every chain to the left half of the tree whenever the halves do not overlap: the conditional of one density given
           * the others' nodes cannot cross over later, and a product of two four-mode beliefs keeps only two modes. */
          if (lev[j] < tr[j]->depth) {
            int lc = lev[j] + 1, c0, nch;
            if (lc == tr[j]->depth) { c0 = tr[j]->start[lev[j]][ind[j]]; nch = tr[j]->start[lev[j]][ind[j] + 1] - c0; }
            else { c0 = 2 * ind[j]; nch = 2; }
            int pick = c0;
            if (nch > 1 && g_level_down) {
              double w0 = tr[j]->wt[lc][c0], w1 = tr[j]->wt[lc][c0 + 1];
              double u = cb2o_uniform(seed, sid, (uint32_t)((step - 1) * D + j), stream, 2, ubits);
              if (!(w0 >= u * (w0 + w1))) pick = c0 + 1;
            }
            ind[j] = pick;
            lev[j] = lc;
          }
        }
        for (int it = 0; it < sweeps; ++it) {
          for (int j = 0; j < D; ++j) {
            double Mo[MAXD], Co[MAXD];
            const cb2o_tree* t = tr[j];
            int lv = lev[j], cnt = t->cnt[lv];
            if (D > 1) loo_product((const cb2o_tree* const*)tr, lev, ind, D, d, j, circ, Mo, Co);
            double emax = -DBL_MAX;
            for (int z = 0; z < cnt; ++z) {
              double acc = 0;
              if (D > 1) for (int i = 0; i < d; ++i) {
                double df = t->mean[lv][z * d + i] - Mo[i];
                if (circ && circ[i]) df = wrap_pi(df);
                double sv = t->var[lv][z * d + i] + Co[i];
                acc += df * df / sv + log(sv);
              }
              e[z] = log(t->wt[lv][z]) - 0.5 * acc;
              if (e[z] > emax) emax = e[z];
            }
            double tot = 0;
            for (int z = 0; z < cnt; ++z) { e[z] = exp(e[z] - emax); tot += e[z]; }
            uint32_t draw = (uint32_t)(((step - 1) * sweeps + it) * D + j);
            double tgt = cb2o_uniform(seed, sid, draw, stream, 0, ubits) * tot, cs = 0;
            int pick = -1, lastpos = 0;
            for (int z = 0; z < cnt; ++z) {
              cs += e[z];
              if (e[z] > 0) lastpos = z;
              if (cs >= tgt) { pick = z; break; }
            }
            if (pick < 0) pick = lastpos;
            ind[j] = pick;
          }
        }
      }
      if (labels_out) for (int j = 0; j < D; ++j) labels_out[(size_t)s * D + j] = tr[j]->perm[ind[j]];
      double Mo[MAXD], Co[MAXD];
      loo_product((const cb2o_tree* const*)tr, lev, ind, D, d, -1, circ, Mo, Co);
      for (int i = 0; i < d; ++i) {
        double z = cb2o_normal(seed, sid, (uint32_t)i, stream, 1, ubits);
        double v = Mo[i] + sqrt(Co[i]) * z;
        if (circ && circ[i]) v = wrap_pi(v);
        pts_out[(size_t)s * d + i] = v;
      }
    }
    free(ind); free(lev); free(e);
  }
  for (int j = 0; j < D; ++j) cb2o_tree_free(tr[j]);
  free(tr);
  return 0;
}

/* torchrun exports OMP_NUM_THREADS=1; the CPU baseline legs of bench.py ask for all host cores explicitly */
void cb2o_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------ approxConv for closed-form factors
 * Restates the closed-form per-particle solves of R:docs/src/principles/approxConvDensities.md:25-34 for Pose2 priors,
 * Pose2Pose2 and Pose2Point2BearingRange; same Philox draws (tag 4) as caesar.jl_b200/csrc/conv.cu. */
/* rotation vector of R through the unit quaternion (Shepperd's branch on the largest diagonal term), angle in [0, pi] */
static void so3_log(const double* R, double* w) {
  const double tr = R[0] + R[4] + R[8];
  double q[4];
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
static void mat3_mul(const double* A, const double* B, double* C, int transA, int transB) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += (transA ? A[k * 3 + i] : A[i * 3 + k]) * (transB ? B[j * 3 + k] : B[k * 3 + j]);
      C[i * 3 + j] = s;
    }
}

/* kinds 3 (Pose3 prior) and 4 (Pose3Pose3): coordinates [t; rotation vector], noise X = L z in the tangent space, product
 * exponential (translation = coordinates), R:docs/src/principles/approxConvDensities.md:25-34 with a Pose3Pose3 factor */
static void approx_conv_pose3(int kind, int reverse, int N, const double* src, int n_src, uint32_t stream, const double* mean,
                              const double* chol, uint64_t seed, double* out) {
  for (int i = 0; i < N; ++i) {
    double z[6], X[6], R1[9], R2[9], R3[9];
    for (int k = 0; k < 6; ++k) z[k] = cb2o_normal(seed, (uint32_t)i, (uint32_t)k, stream, 4, 53);
    for (int r = 0; r < 6; ++r) {
      double v = 0;
      for (int k = 0; k < 6; ++k) v += chol[r * 6 + k] * z[k];
      X[r] = v;
    }
    double* o = out + (size_t)i * 6;
    if (kind == 3) {
      so3_exp(mean + 3, R1); so3_exp(X + 3, R2);
      mat3_mul(R1, R2, R3, 0, 0);
      for (int k = 0; k < 3; ++k) o[k] = mean[k] + X[k];
      so3_log(R3, o + 3);
      continue;
    }
    for (int r = 0; r < 6; ++r) X[r] += mean[r];
    const double* p = src + (size_t)(i % n_src) * 6;
    so3_exp(p + 3, R1); so3_exp(X + 3, R2);
    if (!reverse) {
      mat3_mul(R1, R2, R3, 0, 0);
      for (int k = 0; k < 3; ++k) o[k] = p[k] + R1[k * 3] * X[0] + R1[k * 3 + 1] * X[1] + R1[k * 3 + 2] * X[2];
    } else {
      mat3_mul(R1, R2, R3, 0, 1);
      for (int k = 0; k < 3; ++k) o[k] = p[k] - (R3[k * 3] * X[0] + R3[k * 3 + 1] * X[1] + R3[k * 3 + 2] * X[2]);
    }
    so3_log(R3, o + 3);
  }
}

/* residual of the relative-pose factors (R:ext/Images/ScatterAlignPose2.jl:216-231, body of RoME's Pose2Pose2):
 *   q^ = p o exp_e0(X);  out = vee(M, q, log(M, q, q^))   with the product exponential / logarithm of SpecialEuclidean(n). */
int cb2o_se_residual(int dim, const double* X, const double* p, const double* q, int K, double* out) {
  if ((dim != 2 && dim != 3) || K <= 0) return -1;
  for (int i = 0; i < K; ++i) {
    if (dim == 2) {
      const double *x = X + (size_t)i * 3, *pp = p + (size_t)i * 3, *qq = q + (size_t)i * 3;
      const double sn = sin(pp[2]), cs = cos(pp[2]);
      out[(size_t)i * 3] = pp[0] + cs * x[0] - sn * x[1] - qq[0];
      out[(size_t)i * 3 + 1] = pp[1] + sn * x[0] + cs * x[1] - qq[1];
      out[(size_t)i * 3 + 2] = wrap_pi(pp[2] + x[2] - qq[2]);
    } else {
      const double *x = X + (size_t)i * 6, *pp = p + (size_t)i * 6, *qq = q + (size_t)i * 6;
      double Rp[9], Rx[9], Rq[9], Rh[9], Re[9];
      so3_exp(pp + 3, Rp); so3_exp(x + 3, Rx); so3_exp(qq + 3, Rq);
      mat3_mul(Rp, Rx, Rh, 0, 0);
      mat3_mul(Rq, Rh, Re, 1, 0);
      for (int k = 0; k < 3; ++k) out[(size_t)i * 6 + k] = pp[k] + Rp[k * 3] * x[0] + Rp[k * 3 + 1] * x[1] + Rp[k * 3 + 2] * x[2] - qq[k];
      so3_log(Re, out + (size_t)i * 6 + 3);
    }
  }
  return 0;
}

int cb2o_approx_conv(int kind, int reverse, int N, const double* src, int n_src, const double* aux, int n_aux,
                     uint32_t stream, const double* mean, const double* chol, uint64_t seed, double* out) {
  if (kind < 0 || kind > 7 || N <= 0) return -1;
  if (kind == 7) { /* Prior / LinearRelative with the caller's noise samples (factors of R:test/ICRA2022_tests/nongaussian_mixture.jl:24-51) */
    const int dd = (int)mean[0];
    if (dd < 1 || dd > 3 || !aux || n_aux <= 0) return -1;
    for (int i = 0; i < N; ++i)
      for (int k = 0; k < dd; ++k) {
        double base = n_src > 0 ? src[(size_t)(i % n_src) * dd + k] : 0.0, z = aux[(size_t)(i % n_aux) * dd + k];
        out[(size_t)i * dd + k] = reverse ? base - z : base + z;
      }
    return 0;
  }
  if (kind == 3 || kind == 4) { approx_conv_pose3(kind, reverse, N, src, n_src, stream, mean, chol, seed, out); return 0; }
  for (int i = 0; i < N; ++i) {
    double z[3], nz[3];
    for (int k = 0; k < 3; ++k) z[k] = cb2o_normal(seed, (uint32_t)i, (uint32_t)k, stream, 4, 53);
    for (int r = 0; r < 3; ++r) nz[r] = mean[r] + chol[r * 3] * z[0] + chol[r * 3 + 1] * z[1] + chol[r * 3 + 2] * z[2];
    if (kind == 5) { /* PriorPoint2 */
      out[(size_t)i * 2] = nz[0]; out[(size_t)i * 2 + 1] = nz[1];
    } else if (kind == 6) { /* Point2Point2Range (RoME; graph of R:test/ICRA2022_tests/insufficient_data.jl:70-85): a ring around the
                               known point, range nz[0], direction uniform -- one constraint on two coordinates */
      const double* p = src + (size_t)(i % n_src) * 2;
      double th = 0.5 * TWO_PI * (2.0 * cb2o_uniform(seed, (uint32_t)i, 2, stream, 4, 53) - 1.0);
      out[(size_t)i * 2] = p[0] + nz[0] * cos(th);
      out[(size_t)i * 2 + 1] = p[1] + nz[0] * sin(th);
    } else if (kind == 0) {
      out[(size_t)i * 3] = nz[0]; out[(size_t)i * 3 + 1] = nz[1]; out[(size_t)i * 3 + 2] = wrap_pi(nz[2]);
    } else if (kind == 1) {
      const double* p = src + (size_t)(i % n_src) * 3;
      if (!reverse) {
        double sn = sin(p[2]), cs = cos(p[2]);
        out[(size_t)i * 3] = p[0] + cs * nz[0] - sn * nz[1];
        out[(size_t)i * 3 + 1] = p[1] + sn * nz[0] + cs * nz[1];
        out[(size_t)i * 3 + 2] = wrap_pi(p[2] + nz[2]);
      } else {
        double th = wrap_pi(p[2] - nz[2]), sn = sin(th), cs = cos(th);
        out[(size_t)i * 3] = p[0] - (cs * nz[0] - sn * nz[1]);
        out[(size_t)i * 3 + 1] = p[1] - (sn * nz[0] + cs * nz[1]);
        out[(size_t)i * 3 + 2] = th;
      }
    } else if (!reverse) {
      const double* p = src + (size_t)(i % n_src) * 3;
      out[(size_t)i * 2] = p[0] + nz[1] * cos(p[2] + nz[0]);
      out[(size_t)i * 2 + 1] = p[1] + nz[1] * sin(p[2] + nz[0]);
    } else {
      const double* l = src + (size_t)(i % n_src) * 2;
      double th = aux[(size_t)(i % n_aux) * 3 + 2];
      out[(size_t)i * 3] = l[0] - nz[1] * cos(th + nz[0]);
      out[(size_t)i * 3 + 1] = l[1] - nz[1] * sin(th + nz[0]);
      out[(size_t)i * 3 + 2] = th;
    }
  }
  return 0;
}

/* mmd with a product-manifold distance: dist^2 = sum_E d^2 + rot_weight * sum_C wrap(d)^2 */
int cb2o_mmd_manifold(const double* a, int N, const double* b, int M, int d, const uint8_t* circular, double rot_weight,
                      double bw, double* out) {
  if (N <= 0 || M <= 0 || d <= 0 || d > MAXD) return -1;
  const double* P[3] = {a, a, b};
  const double* Q[3] = {a, b, b};
  int np[3] = {N, N, M}, nq[3] = {N, M, M};
  double s[3];
  for (int t = 0; t < 3; ++t) {
    double tot = 0.0;
#pragma omp parallel for reduction(+ : tot) schedule(static)
    for (int i = 0; i < np[t]; ++i) {
      double acc = 0.0;
      for (int j = 0; j < nq[t]; ++j) {
        double r2 = 0.0;
        for (int k = 0; k < d; ++k) {
          double df = P[t][(size_t)i * d + k] - Q[t][(size_t)j * d + k];
          if (circular && circular[k]) { df = wrap_pi(df); r2 += rot_weight * df * df; } else r2 += df * df;
        }
        acc += exp(-bw * r2);
      }
      tot += acc;
    }
    s[t] = tot;
  }
  *out = s[0] / ((double)N * N) - 2.0 * s[1] / ((double)N * M) + s[2] / ((double)M * M);
  return 0;
}

/* ------------------------------------------------------------------ ICP branch of ScatterAlign
 * Restatement of simpleICP as vendored in R:src/3rdParty/_PCL/services/ICP_Simple.jl (the algorithm's source IS in the
 * reference tree).  Deliberate, documented deviations: exact nearest neighbours by exhaustive search instead of a kd-tree
 * (same result up to distance ties, which break towards the lower index here); the normal's sign -- LAPACK's eigenvector
 * sign upstream -- is fixed so that its largest-magnitude component is positive; `A \ l` (QR upstream) is solved through the
 * normal equations with a Cholesky factorisation. */
static void sym3_eigen(const double* Cin, double* eval, double* evec) { /* cyclic Jacobi; evec columns, ascending eigenvalues */
  double A[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  memcpy(A, Cin, sizeof(A));
  for (int sweep = 0; sweep < 32; ++sweep) {
    double off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
    if (off < 1e-300) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double apq = A[p * 3 + q];
        if (fabs(apq) < 1e-300) continue;
        double theta = (A[q * 3 + q] - A[p * 3 + p]) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s_ = t * c;
        for (int k = 0; k < 3; ++k) { /* A <- A J */
          double akp = A[k * 3 + p], akq = A[k * 3 + q];
          A[k * 3 + p] = c * akp - s_ * akq; A[k * 3 + q] = s_ * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) { /* A <- J^T A */
          double apk = A[p * 3 + k], aqk = A[q * 3 + k];
          A[p * 3 + k] = c * apk - s_ * aqk; A[q * 3 + k] = s_ * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          double vkp = V[k * 3 + p], vkq = V[k * 3 + q];
          V[k * 3 + p] = c * vkp - s_ * vkq; V[k * 3 + q] = s_ * vkp + c * vkq;
        }
      }
  }
  int idx[3] = {0, 1, 2};
  double d[3] = {A[0], A[4], A[8]};
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 2 - i; ++j)
    if (d[idx[j]] > d[idx[j + 1]]) { int t = idx[j]; idx[j] = idx[j + 1]; idx[j + 1] = t; }
  for (int k = 0; k < 3; ++k) { eval[k] = d[idx[k]]; for (int r = 0; r < 3; ++r) evec[r * 3 + k] = V[r * 3 + idx[k]]; }
}

static double median_of(const double* v, int n, double* tmp) { /* Statistics.median: mean of the two middle values for even n */
  if (n <= 0) return 0.0;
  memcpy(tmp, v, sizeof(double) * (size_t)n);
  qsort(tmp, (size_t)n, sizeof(double), cmp_double);
  return (n & 1) ? tmp[n / 2] : 0.5 * (tmp[n / 2 - 1] + tmp[n / 2]);
}

static int icp_select(int nf, int n, int32_t* sel) { /* ICP_Simple.jl:78-84: round.(Int, range(1, nf, length = n)) */
  if (nf <= n) { for (int i = 0; i < nf; ++i) sel[i] = i; return nf; }
  for (int j = 0; j < n; ++j) sel[j] = (int32_t)nearbyint(1.0 + (double)j * (double)(nf - 1) / (double)(n - 1)) - 1;
  return n;
}

int cb2o_icp_normals(const double* fix, int nf, int correspondences, int neighbors, int32_t* sel, double* normals, double* plan) {
  if (nf < 3 || neighbors < 2 || correspondences < 10) return -1;
  const int ns = icp_select(nf, correspondences, sel);
  const int kn = neighbors < nf ? neighbors : nf;
#pragma omp parallel
  {
    double* dist = (double*)malloc(sizeof(double) * (size_t)nf);
    int* nn = (int*)malloc(sizeof(int) * (size_t)kn);
#pragma omp for schedule(dynamic, 4)
    for (int s = 0; s < ns; ++s) { /* ICP_Simple.jl:86-110 */
      const double* q = fix + (size_t)sel[s] * 3;
      for (int i = 0; i < nf; ++i) {
        double dx = fix[(size_t)i * 3] - q[0], dy = fix[(size_t)i * 3 + 1] - q[1], dz = fix[(size_t)i * 3 + 2] - q[2];
        dist[i] = dx * dx + dy * dy + dz * dz;
      }
      for (int k = 0; k < kn; ++k) { /* the kn nearest, ties towards the lower index */
        int best = -1; double bd = 0;
        for (int i = 0; i < nf; ++i) if (dist[i] >= 0 && (best < 0 || dist[i] < bd)) { best = i; bd = dist[i]; }
        nn[k] = best; dist[best] = -1.0;
      }
      double mean[3] = {0, 0, 0}, C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (int k = 0; k < kn; ++k) for (int a = 0; a < 3; ++a) mean[a] += fix[(size_t)nn[k] * 3 + a];
      for (int a = 0; a < 3; ++a) mean[a] /= kn;
      for (int k = 0; k < kn; ++k) {
        double d[3];
        for (int a = 0; a < 3; ++a) d[a] = fix[(size_t)nn[k] * 3 + a] - mean[a];
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) C[a * 3 + b] += d[a] * d[b];
      }
      for (int a = 0; a < 9; ++a) C[a] /= (kn - 1); /* Statistics.cov: corrected */
      double ev[3], V[9];
      sym3_eigen(C, ev, V);
      double n[3] = {V[0], V[3], V[6]};
      int big = fabs(n[1]) > fabs(n[0]) ? 1 : 0;
      if (fabs(n[2]) > fabs(n[big])) big = 2;
      double sg = n[big] < 0 ? -1.0 : 1.0;
      for (int a = 0; a < 3; ++a) normals[(size_t)s * 3 + a] = sg * n[a];
      plan[s] = (ev[1] - ev[0]) / ev[2];
    }
    free(dist); free(nn);
  }
  return ns;
}

static int chol6_solve(double* M, double* b) { /* M (6x6 SPD, row-major, overwritten) x = b (overwritten) */
  for (int j = 0; j < 6; ++j) {
    double s = M[j * 6 + j];
    for (int k = 0; k < j; ++k) s -= M[j * 6 + k] * M[j * 6 + k];
    if (!(s > 0)) return -1;
    M[j * 6 + j] = sqrt(s);
    for (int i = j + 1; i < 6; ++i) {
      double t = M[i * 6 + j];
      for (int k = 0; k < j; ++k) t -= M[i * 6 + k] * M[j * 6 + k];
      M[i * 6 + j] = t / M[j * 6 + j];
    }
  }
  for (int i = 0; i < 6; ++i) { double t = b[i]; for (int k = 0; k < i; ++k) t -= M[i * 6 + k] * b[k]; b[i] = t / M[i * 6 + i]; }
  for (int i = 5; i >= 0; --i) { double t = b[i]; for (int k = i + 1; k < 6; ++k) t -= M[k * 6 + i] * b[k]; b[i] = t / M[i * 6 + i]; }
  return 0;
}

int cb2o_icp_align(const double* fix, int nf, const double* mov0, int nm, const double* perturb, int K, int correspondences,
                   int neighbors, double min_planarity, double min_change, int max_iterations, double* H_out, double* stats_out) {
  if (nf < 3 || nm < 1 || K < 1 || correspondences < 10 || neighbors < 2 || !(min_planarity >= 0 && min_planarity < 1) ||
      !(min_change > 0) || max_iterations < 1) return -1; /* argument checks of ICP_Simple.jl:252-258 */
  int32_t* sel = (int32_t*)malloc(sizeof(int32_t) * (size_t)(correspondences < nf ? correspondences : nf));
  double* normals = (double*)malloc(sizeof(double) * 3 * (size_t)correspondences);
  double* plan = (double*)malloc(sizeof(double) * (size_t)correspondences);
  const int ns = cb2o_icp_normals(fix, nf, correspondences, neighbors, sel, normals, plan);
  if (ns < 0) { free(sel); free(normals); free(plan); return -1; }
#pragma omp parallel for schedule(dynamic, 1)
  for (int k = 0; k < K; ++k) {
    double* mov = (double*)malloc(sizeof(double) * 3 * (size_t)nm);
    if (perturb) cb2o_transform_points(3, perturb + (size_t)k * 6, 1, mov0, nm, mov);
    else memcpy(mov, mov0, sizeof(double) * 3 * (size_t)nm);
    int* match = (int*)malloc(sizeof(int) * ns);
    double* dist = (double*)malloc(sizeof(double) * ns);
    double* tmp = (double*)malloc(sizeof(double) * ns);
    double* dev = (double*)malloc(sizeof(double) * ns);
    double* res = (double*)malloc(sizeof(double) * ns);
    double H[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    double mean_old = 0, std_old = 0, mean_new = 0, std_new = 0;
    int iters = 0, nkeep = 0;
    for (int it = 1; it <= max_iterations; ++it) {
      iters = it;
      for (int s = 0; s < ns; ++s) { /* matching!, ICP_Simple.jl:112-134 */
        const double* f = fix + (size_t)sel[s] * 3;
        int best = 0; double bd = 1e300;
        for (int i = 0; i < nm; ++i) {
          double dx = mov[(size_t)i * 3] - f[0], dy = mov[(size_t)i * 3 + 1] - f[1], dz = mov[(size_t)i * 3 + 2] - f[2];
          double d2 = dx * dx + dy * dy + dz * dz;
          if (d2 < bd) { bd = d2; best = i; }
        }
        match[s] = best;
        const double* m = mov + (size_t)best * 3;
        dist[s] = (m[0] - f[0]) * normals[s * 3] + (m[1] - f[1]) * normals[s * 3 + 1] + (m[2] - f[2]) * normals[s * 3 + 2];
      }
      /* reject!, ICP_Simple.jl:136-154 */
      double med = median_of(dist, ns, tmp);
      for (int s = 0; s < ns; ++s) dev[s] = fabs(dist[s] - med);
      double sigmad = 1.4826022185056018 * median_of(dev, ns, tmp);
      double AtA[36], Atl[6];
      memset(AtA, 0, sizeof(AtA)); memset(Atl, 0, sizeof(Atl));
      nkeep = 0;
      for (int s = 0; s < ns; ++s) { /* estimate_rigid_body_transformation, :174-200 */
        if (!(fabs(dist[s] - med) <= 3 * sigmad && plan[s] > min_planarity)) { match[s] = -1; continue; }
        ++nkeep;
        const double* f = fix + (size_t)sel[s] * 3; const double* m = mov + (size_t)match[s] * 3; const double* n = normals + s * 3;
        double a[6] = {-m[2] * n[1] + m[1] * n[2], m[2] * n[0] - m[0] * n[2], -m[1] * n[0] + m[0] * n[1], n[0], n[1], n[2]};
        double l = n[0] * (f[0] - m[0]) + n[1] * (f[1] - m[1]) + n[2] * (f[2] - m[2]);
        for (int i = 0; i < 6; ++i) { Atl[i] += a[i] * l; for (int j = 0; j < 6; ++j) AtA[i * 6 + j] += a[i] * a[j]; }
      }
      double x[6];
      memcpy(x, Atl, sizeof(x));
      if (nkeep < 6 || chol6_solve(AtA, x) != 0) { iters = -it; break; } /* upstream would throw a singular-system error */
      int nr = 0;
      double sm = 0;
      for (int s = 0; s < ns; ++s) {
        if (match[s] < 0) continue;
        const double* f = fix + (size_t)sel[s] * 3; const double* m = mov + (size_t)match[s] * 3; const double* n = normals + s * 3;
        double a[6] = {-m[2] * n[1] + m[1] * n[2], m[2] * n[0] - m[0] * n[2], -m[1] * n[0] + m[0] * n[1], n[0], n[1], n[2]};
        double l = n[0] * (f[0] - m[0]) + n[1] * (f[1] - m[1]) + n[2] * (f[2] - m[2]);
        double r = -l;
        for (int i = 0; i < 6; ++i) r += a[i] * x[i];
        res[nr++] = r; sm += r;
      }
      mean_new = sm / nr;
      double sv = 0;
      for (int i = 0; i < nr; ++i) sv += (res[i] - mean_new) * (res[i] - mean_new);
      std_new = sqrt(sv / (nr - 1));
      double R[9];
      so3_exp(x, R); /* euler_angles_to_linearized_rotation_matrix(..., rigid = true) = exp_lie, HomographyTransforms.jl:8-16 */
      for (int i = 0; i < nm; ++i) { /* transform!(pcmov, dH) */
        double* m = mov + (size_t)i * 3, o[3];
        for (int a = 0; a < 3; ++a) o[a] = R[a * 3] * m[0] + R[a * 3 + 1] * m[1] + R[a * 3 + 2] * m[2] + x[3 + a];
        m[0] = o[0]; m[1] = o[1]; m[2] = o[2];
      }
      double dH[16] = {R[0], R[1], R[2], x[3], R[3], R[4], R[5], x[4], R[6], R[7], R[8], x[5], 0, 0, 0, 1}, Hn[16];
      for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) {
        double t = 0; for (int c = 0; c < 4; ++c) t += dH[a * 4 + c] * H[c * 4 + b]; Hn[a * 4 + b] = t; }
      memcpy(H, Hn, sizeof(H)); /* H .= dH * H */
      if (it > 1) { /* check_convergence_criteria, :156-166 */
        double cm = fabs((mean_new - mean_old) / mean_old * 100), cs = fabs((std_new - std_old) / std_old * 100);
        if (cm < min_change && cs < min_change) break;
      }
      mean_old = mean_new; std_old = std_new;
    }
    memcpy(H_out + (size_t)k * 16, H, sizeof(H));
    if (stats_out) { stats_out[k * 4] = mean_new; stats_out[k * 4 + 1] = std_new; stats_out[k * 4 + 2] = nkeep; stats_out[k * 4 + 3] = iters; }
    free(mov); free(match); free(dist); free(tmp); free(dev); free(res);
  }
  free(sel); free(normals); free(plan);
  return 0;
}
