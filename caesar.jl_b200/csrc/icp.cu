// icp.cu -- ICP branch of ScatterAlign ("next" row N3): `getSample` with sample_count < 0
// (R:ext/Images/ScatterAlignPose2.jl:187-214) runs simpleICP (R:src/3rdParty/_PCL/services/ICP_Simple.jl:239-332): point-to-plane
// ICP with PCA normals on selected points of the fixed cloud.  The solver calls it once per particle with a fresh random
// perturbation of the moving cloud; here K particles run as one batch:
//   icp_normals_kernel : one CTA per selected fixed point -- its `neighbors` nearest neighbours by exhaustive search
//                        (rounds of "next smallest (distance, index)"), 3x3 covariance, Jacobi eigenvectors -> normal, planarity.
//                        Shared by all K particles (it depends on the fixed cloud only; upstream recomputes it per call).
//   icp_iterate_kernel : one CTA per particle, all iterations inside the kernel: exhaustive nearest neighbour of every selected
//                        point in the particle's moving cloud (tiles of the cloud staged in shared memory), median / MAD rejection
//                        (bitonic sort in shared memory), 6x6 normal equations by block reduction, Cholesky solve, rigid update
//                        of the cloud, convergence test on mean / std of the residuals.
// FP64 throughout (a few MFLOP per iteration).  Same documented deviations as the CPU restatement (oracle/cb2_oracle.c):
// exhaustive instead of kd-tree search (ties -> lower index), fixed sign of the normal, normal equations instead of QR.
#include "common.cuh"

namespace {

constexpr int kIcpThreads = 256;
constexpr int kIcpMaxSel = 1024;   // selected correspondences per particle (one shared-memory sort)
constexpr int kIcpTile = 1024;     // points of the moving cloud per shared-memory tile

__device__ void so3_exp_dev(const double* w, double* R) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2);
  double A, B;
  if (th < 1e-8) { A = 1.0 - th2 / 6.0; B = 0.5 - th2 / 24.0; }
  else { A = sin(th) / th; B = (1.0 - cos(th)) / th2; }
  const double K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  double K2[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += K[i * 3 + k] * K[k * 3 + j];
      K2[i * 3 + j] = s;
    }
  for (int i = 0; i < 9; ++i) R[i] = A * K[i] + B * K2[i];
  R[0] += 1; R[4] += 1; R[8] += 1;
}

__device__ void sym3_eigen_dev(const double* Cin, double* eval, double* evec) {  // cyclic Jacobi, ascending eigenvalues
  double A[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int i = 0; i < 9; ++i) A[i] = Cin[i];
  for (int sweep = 0; sweep < 32; ++sweep) {
    const double off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
    if (off < 1e-300) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        const double apq = A[p * 3 + q];
        if (fabs(apq) < 1e-300) continue;
        const double theta = (A[q * 3 + q] - A[p * 3 + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {
          const double akp = A[k * 3 + p], akq = A[k * 3 + q];
          A[k * 3 + p] = c * akp - s * akq; A[k * 3 + q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          const double apk = A[p * 3 + k], aqk = A[q * 3 + k];
          A[p * 3 + k] = c * apk - s * aqk; A[q * 3 + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = V[k * 3 + p], vkq = V[k * 3 + q];
          V[k * 3 + p] = c * vkp - s * vkq; V[k * 3 + q] = s * vkp + c * vkq;
        }
      }
  }
  int idx[3] = {0, 1, 2};
  const double d[3] = {A[0], A[4], A[8]};
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2 - i; ++j)
      if (d[idx[j]] > d[idx[j + 1]]) { int t = idx[j]; idx[j] = idx[j + 1]; idx[j + 1] = t; }
  for (int k = 0; k < 3; ++k) {
    eval[k] = d[idx[k]];
    for (int r = 0; r < 3; ++r) evec[r * 3 + k] = V[r * 3 + idx[k]];
  }
}

// normals + planarity of the selected fixed points (ICP_Simple.jl:86-110)
__global__ void __launch_bounds__(kIcpThreads)
icp_normals_kernel(const double* __restrict__ fix, int nf, const int* __restrict__ sel, int kn, double* __restrict__ normals,
                   double* __restrict__ plan) {
  __shared__ double s_d[kIcpThreads / 32];
  __shared__ int s_i[kIcpThreads / 32];
  __shared__ double last_d;
  __shared__ int last_i;
  __shared__ int nn[256];
  const int s = blockIdx.x, tid = threadIdx.x;
  const double qx = fix[(size_t)sel[s] * 3], qy = fix[(size_t)sel[s] * 3 + 1], qz = fix[(size_t)sel[s] * 3 + 2];
  if (tid == 0) { last_d = -1.0; last_i = -1; }
  __syncthreads();
  for (int k = 0; k < kn; ++k) {  // the next smallest (distance, index) after the previous pick
    const double ld = last_d;
    const int li = last_i;
    double bd = 1e300;
    int bi = 0x7fffffff;
    for (int i = tid; i < nf; i += kIcpThreads) {
      const double dx = fix[(size_t)i * 3] - qx, dy = fix[(size_t)i * 3 + 1] - qy, dz = fix[(size_t)i * 3 + 2] - qz;
      const double d2 = dx * dx + dy * dy + dz * dz;
      const bool after = d2 > ld || (d2 == ld && i > li);
      if (after && (d2 < bd || (d2 == bd && i < bi))) { bd = d2; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double od = __shfl_xor_sync(0xffffffffu, bd, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
    }
    if ((tid & 31) == 0) { s_d[tid >> 5] = bd; s_i[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < kIcpThreads / 32; ++w)
        if (s_d[w] < bd || (s_d[w] == bd && s_i[w] < bi)) { bd = s_d[w]; bi = s_i[w]; }
      last_d = bd; last_i = bi; nn[k] = bi;
    }
    __syncthreads();
  }
  if (tid == 0) {
    double mean[3] = {0, 0, 0}, C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < kn; ++k)
      for (int a = 0; a < 3; ++a) mean[a] += fix[(size_t)nn[k] * 3 + a];
    for (int a = 0; a < 3; ++a) mean[a] /= kn;
    for (int k = 0; k < kn; ++k) {
      double d[3];
      for (int a = 0; a < 3; ++a) d[a] = fix[(size_t)nn[k] * 3 + a] - mean[a];
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) C[a * 3 + b] += d[a] * d[b];
    }
    for (int a = 0; a < 9; ++a) C[a] /= (kn - 1);
    double ev[3], V[9];
    sym3_eigen_dev(C, ev, V);
    double n[3] = {V[0], V[3], V[6]};
    int big = fabs(n[1]) > fabs(n[0]) ? 1 : 0;
    if (fabs(n[2]) > fabs(n[big])) big = 2;
    const double sg = n[big] < 0 ? -1.0 : 1.0;
    for (int a = 0; a < 3; ++a) normals[(size_t)s * 3 + a] = sg * n[a];
    plan[s] = (ev[1] - ev[0]) / ev[2];
  }
}

struct IcpArgs {
  const double* fix; int nf;
  const double* mov0; int nm;
  const double* perturb;  // K x 6 or null
  const int* sel; int ns;
  const double* normals; const double* plan;
  double* mov;            // K x nm x 3 working clouds
  double min_planarity, min_change;
  int max_iterations;
  double* H_out; double* stats_out;
};

__device__ __forceinline__ double block_sum(double v, double* sred, int tid) {
  v = warp_sum(v);
  __syncthreads();
  if ((tid & 31) == 0) sred[tid >> 5] = v;
  __syncthreads();
  double t = 0;
  for (int w = 0; w < kIcpThreads / 32; ++w) t += sred[w];
  return t;
}

// ascending bitonic sort of 1024 doubles in shared memory
__device__ void sort1024(double* a, int tid) {
  for (int k = 2; k <= kIcpMaxSel; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int q = tid; q < kIcpMaxSel / 2; q += kIcpThreads) {
        const int i = ((q & ~(j - 1)) << 1) | (q & (j - 1)), ixj = i | j;
        const double x = a[i], y = a[ixj];
        if ((y < x) == ((i & k) == 0)) { a[i] = y; a[ixj] = x; }
      }
      __syncthreads();
    }
}
__device__ __forceinline__ double median_sorted(const double* a, int n) { return (n & 1) ? a[n / 2] : 0.5 * (a[n / 2 - 1] + a[n / 2]); }

__global__ void __launch_bounds__(kIcpThreads)
icp_iterate_kernel(const IcpArgs a) {
  __shared__ double tile[kIcpTile * 3];
  __shared__ double srt[kIcpMaxSel];
  __shared__ double sred[kIcpThreads / 32];
  __shared__ double sx[6], sR[9], sH[16], sstat[4];
  __shared__ int sflag;
  const int k = blockIdx.x, tid = threadIdx.x, ns = a.ns, nm = a.nm;
  double* mov = a.mov + (size_t)k * nm * 3;
  constexpr int QPT = kIcpMaxSel / kIcpThreads;  // queries per thread

  // the particle's moving cloud: mov0 moved by the inverse of its perturbation (ScatterAlignPose2.jl:195-196, backward = true)
  if (tid == 0) {
    if (a.perturb) {
      double Rf[9];
      so3_exp_dev(a.perturb + (size_t)k * 6 + 3, Rf);
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) sR[i * 3 + j] = Rf[j * 3 + i];
      for (int i = 0; i < 3; ++i) {
        double s = 0;
        for (int j = 0; j < 3; ++j) s += sR[i * 3 + j] * a.perturb[(size_t)k * 6 + j];
        sx[3 + i] = -s;
      }
    } else {
      for (int i = 0; i < 9; ++i) sR[i] = (i % 4 == 0) ? 1.0 : 0.0;
      sx[3] = sx[4] = sx[5] = 0.0;
    }
    for (int i = 0; i < 16; ++i) sH[i] = (i % 5 == 0) ? 1.0 : 0.0;
  }
  __syncthreads();
  for (int i = tid; i < nm; i += kIcpThreads) {
    const double x = a.mov0[(size_t)i * 3], y = a.mov0[(size_t)i * 3 + 1], z = a.mov0[(size_t)i * 3 + 2];
    for (int r = 0; r < 3; ++r) mov[(size_t)i * 3 + r] = sR[r * 3] * x + sR[r * 3 + 1] * y + sR[r * 3 + 2] * z + sx[3 + r];
  }
  __syncthreads();

  double fq[QPT][3], nq[QPT][3], pl[QPT];
#pragma unroll
  for (int u = 0; u < QPT; ++u) {
    const int s = tid + u * kIcpThreads;
    for (int r = 0; r < 3; ++r) {
      fq[u][r] = s < ns ? a.fix[(size_t)a.sel[s] * 3 + r] : 0.0;
      nq[u][r] = s < ns ? a.normals[(size_t)s * 3 + r] : 0.0;
    }
    pl[u] = s < ns ? a.plan[s] : 0.0;
  }
  double mean_old = 0, std_old = 0, mean_new = 0, std_new = 0;
  int iters = 0, nkeep = 0;
  for (int it = 1; it <= a.max_iterations; ++it) {
    iters = it;
    // ---- matching! (ICP_Simple.jl:112-134): nearest moving point of every selected fixed point, lowest index on ties
    double bd[QPT];
    int bi[QPT];
#pragma unroll
    for (int u = 0; u < QPT; ++u) { bd[u] = 1e300; bi[u] = 0; }
    for (int t0 = 0; t0 < nm; t0 += kIcpTile) {
      const int tn = min(kIcpTile, nm - t0);
      __syncthreads();
      for (int i = tid; i < tn * 3; i += kIcpThreads) tile[i] = mov[(size_t)t0 * 3 + i];
      __syncthreads();
      for (int i = 0; i < tn; ++i) {
        const double mx = tile[i * 3], my = tile[i * 3 + 1], mz = tile[i * 3 + 2];
#pragma unroll
        for (int u = 0; u < QPT; ++u) {
          const double dx = mx - fq[u][0], dy = my - fq[u][1], dz = mz - fq[u][2];
          const double d2 = dx * dx + dy * dy + dz * dz;
          if (d2 < bd[u]) { bd[u] = d2; bi[u] = t0 + i; }
        }
      }
    }
    double dist[QPT], mq[QPT][3];
#pragma unroll
    for (int u = 0; u < QPT; ++u) {
      for (int r = 0; r < 3; ++r) mq[u][r] = mov[(size_t)bi[u] * 3 + r];
      dist[u] = (mq[u][0] - fq[u][0]) * nq[u][0] + (mq[u][1] - fq[u][1]) * nq[u][1] + (mq[u][2] - fq[u][2]) * nq[u][2];
    }
    // ---- reject! (ICP_Simple.jl:136-154): median and normalised MAD of the point-to-plane distances
    __syncthreads();
#pragma unroll
    for (int u = 0; u < QPT; ++u) { const int s = tid + u * kIcpThreads; srt[s] = s < ns ? dist[u] : 1e300; }
    __syncthreads();
    sort1024(srt, tid);
    const double med = median_sorted(srt, ns);
    __syncthreads();
#pragma unroll
    for (int u = 0; u < QPT; ++u) { const int s = tid + u * kIcpThreads; srt[s] = s < ns ? fabs(dist[u] - med) : 1e300; }
    __syncthreads();
    sort1024(srt, tid);
    const double sigmad = 1.4826022185056018 * median_sorted(srt, ns);
    bool keep[QPT];
    double row[QPT][6], lq[QPT];
    double acc[27];
#pragma unroll
    for (int i = 0; i < 27; ++i) acc[i] = 0;
    int cnt = 0;
#pragma unroll
    for (int u = 0; u < QPT; ++u) {  // estimate_rigid_body_transformation (ICP_Simple.jl:174-200): rows of A and l
      const int s = tid + u * kIcpThreads;
      keep[u] = s < ns && fabs(dist[u] - med) <= 3 * sigmad && pl[u] > a.min_planarity;
      row[u][0] = -mq[u][2] * nq[u][1] + mq[u][1] * nq[u][2];
      row[u][1] = mq[u][2] * nq[u][0] - mq[u][0] * nq[u][2];
      row[u][2] = -mq[u][1] * nq[u][0] + mq[u][0] * nq[u][1];
      row[u][3] = nq[u][0]; row[u][4] = nq[u][1]; row[u][5] = nq[u][2];
      lq[u] = nq[u][0] * (fq[u][0] - mq[u][0]) + nq[u][1] * (fq[u][1] - mq[u][1]) + nq[u][2] * (fq[u][2] - mq[u][2]);
      if (keep[u]) {
        ++cnt;
        int p = 0;
        for (int i = 0; i < 6; ++i)
          for (int j = i; j < 6; ++j) acc[p++] += row[u][i] * row[u][j];
        for (int i = 0; i < 6; ++i) acc[21 + i] += row[u][i] * lq[u];
      }
    }
    double tot[27];
    for (int i = 0; i < 27; ++i) tot[i] = block_sum(acc[i], sred, tid);
    nkeep = (int)(block_sum((double)cnt, sred, tid) + 0.5);
    if (tid == 0) {  // normal equations, Cholesky
      double M[36], x[6];
      int p = 0;
      for (int i = 0; i < 6; ++i)
        for (int j = i; j < 6; ++j) { M[i * 6 + j] = tot[p]; M[j * 6 + i] = tot[p]; ++p; }
      for (int i = 0; i < 6; ++i) x[i] = tot[21 + i];
      int ok = nkeep >= 6;
      for (int j = 0; j < 6 && ok; ++j) {
        double s = M[j * 6 + j];
        for (int c = 0; c < j; ++c) s -= M[j * 6 + c] * M[j * 6 + c];
        if (!(s > 0)) { ok = 0; break; }
        M[j * 6 + j] = sqrt(s);
        for (int i = j + 1; i < 6; ++i) {
          double t = M[i * 6 + j];
          for (int c = 0; c < j; ++c) t -= M[i * 6 + c] * M[j * 6 + c];
          M[i * 6 + j] = t / M[j * 6 + j];
        }
      }
      if (ok) {
        for (int i = 0; i < 6; ++i) { double t = x[i]; for (int c = 0; c < i; ++c) t -= M[i * 6 + c] * x[c]; x[i] = t / M[i * 6 + i]; }
        for (int i = 5; i >= 0; --i) { double t = x[i]; for (int c = i + 1; c < 6; ++c) t -= M[c * 6 + i] * x[c]; x[i] = t / M[i * 6 + i]; }
        for (int i = 0; i < 6; ++i) sx[i] = x[i];
        so3_exp_dev(x, sR);
      }
      sflag = ok;
    }
    __syncthreads();
    if (!sflag) { iters = -it; break; }
    // ---- residuals A x - l, their mean and corrected standard deviation
    double rs[QPT], sm = 0;
#pragma unroll
    for (int u = 0; u < QPT; ++u) {
      double r = -lq[u];
      for (int i = 0; i < 6; ++i) r += row[u][i] * sx[i];
      rs[u] = keep[u] ? r : 0.0;
      sm += rs[u];
    }
    mean_new = block_sum(sm, sred, tid) / nkeep;
    double sv = 0;
#pragma unroll
    for (int u = 0; u < QPT; ++u)
      if (keep[u]) sv += (rs[u] - mean_new) * (rs[u] - mean_new);
    std_new = sqrt(block_sum(sv, sred, tid) / (nkeep - 1));
    // ---- transform!(pcmov, dH); H .= dH * H
    for (int i = tid; i < nm; i += kIcpThreads) {
      const double x = mov[(size_t)i * 3], y = mov[(size_t)i * 3 + 1], z = mov[(size_t)i * 3 + 2];
      for (int r = 0; r < 3; ++r) mov[(size_t)i * 3 + r] = sR[r * 3] * x + sR[r * 3 + 1] * y + sR[r * 3 + 2] * z + sx[3 + r];
    }
    if (tid == 0) {
      const double dH[16] = {sR[0], sR[1], sR[2], sx[3], sR[3], sR[4], sR[5], sx[4], sR[6], sR[7], sR[8], sx[5], 0, 0, 0, 1};
      double Hn[16];
      for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
          double t = 0;
          for (int e = 0; e < 4; ++e) t += dH[r * 4 + e] * sH[e * 4 + c];
          Hn[r * 4 + c] = t;
        }
      for (int i = 0; i < 16; ++i) sH[i] = Hn[i];
    }
    __syncthreads();
    if (it > 1) {  // check_convergence_criteria (ICP_Simple.jl:156-166)
      const double cm = fabs((mean_new - mean_old) / mean_old * 100), cs = fabs((std_new - std_old) / std_old * 100);
      if (cm < a.min_change && cs < a.min_change) break;
    }
    mean_old = mean_new; std_old = std_new;
  }
  __syncthreads();
  if (tid == 0) {
    for (int i = 0; i < 16; ++i) a.H_out[(size_t)k * 16 + i] = sH[i];
    if (a.stats_out) {
      sstat[0] = mean_new; sstat[1] = std_new; sstat[2] = nkeep; sstat[3] = iters;
      for (int i = 0; i < 4; ++i) a.stats_out[(size_t)k * 4 + i] = sstat[i];
    }
  }
}

}  // namespace

extern "C" int cb2_icp_align_batch(cb2_ctx* ctx, const double* fix, int nf, const double* mov, int nm, const double* perturb, int K,
                                   int correspondences, int neighbors, double min_planarity, double min_change,
                                   int max_iterations, double* H_out, double* stats_out) {
  if (!ctx) return CB2_ERR_INVALID;
  cb2_lock g(ctx->mu);
  CB2_REQUIRE(ctx, fix && mov && H_out && nf >= 3 && nm >= 1 && K >= 1, "null clouds / outputs or empty batch");
  // the argument checks of alignICP_Simple (ICP_Simple.jl:252-258)
  CB2_REQUIRE(ctx, correspondences >= 10, "\"correspondences\" must be >= 10");
  CB2_REQUIRE(ctx, min_planarity >= 0 && min_planarity < 1, "\"min_planarity\" must be >= 0 and < 1");
  CB2_REQUIRE(ctx, neighbors >= 2, "\"neighbors\" must be >= 2");
  CB2_REQUIRE(ctx, min_change > 0, "\"min_change\" must be > 0");
  CB2_REQUIRE(ctx, max_iterations > 0, "\"max_iterations\" must be > 0");
  if (correspondences > kIcpMaxSel) return cb2_fail(ctx, CB2_ERR_UNSUPPORTED, "correspondences %d exceed %d", correspondences, kIcpMaxSel);
  if (neighbors > 256) return cb2_fail(ctx, CB2_ERR_UNSUPPORTED, "neighbors %d exceed 256", neighbors);
  CB2_CUDA(ctx, cudaSetDevice(ctx->device));
  const int ns = nf <= correspondences ? nf : correspondences;
  const int kn = neighbors < nf ? neighbors : nf;
  std::vector<int> sel(ns);  // select_n_points! (ICP_Simple.jl:78-84): round.(Int, range(1, nf, length = n)), ties to even
  if (nf <= correspondences) for (int i = 0; i < nf; ++i) sel[i] = i;
  else for (int j = 0; j < ns; ++j) sel[j] = (int)nearbyint(1.0 + (double)j * (double)(nf - 1) / (double)(ns - 1)) - 1;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += cb2_align(n); return o; };
  const size_t o_fix = take(sizeof(double) * 3 * (size_t)nf), o_mov0 = take(sizeof(double) * 3 * (size_t)nm);
  const size_t o_pert = take(sizeof(double) * 6 * (size_t)K), o_sel = take(sizeof(int) * (size_t)ns);
  const size_t o_nrm = take(sizeof(double) * 3 * (size_t)ns), o_pl = take(sizeof(double) * (size_t)ns);
  const size_t o_mov = take(sizeof(double) * 3 * (size_t)nm * K), o_H = take(sizeof(double) * 16 * (size_t)K);
  const size_t o_st = take(sizeof(double) * 4 * (size_t)K);
  int rc = cb2_arena_reserve(ctx, off + 4096);
  if (rc) return rc;
  char* base = (char*)cb2_arena_alloc(ctx, off);
  if (!base) return cb2_fail(ctx, CB2_ERR_NOMEM, "icp scratch exhausted");
  CB2_CUDA(ctx, cudaMemcpyAsync(base + o_fix, fix, sizeof(double) * 3 * (size_t)nf, cudaMemcpyHostToDevice, ctx->stream));
  CB2_CUDA(ctx, cudaMemcpyAsync(base + o_mov0, mov, sizeof(double) * 3 * (size_t)nm, cudaMemcpyHostToDevice, ctx->stream));
  if (perturb) CB2_CUDA(ctx, cudaMemcpyAsync(base + o_pert, perturb, sizeof(double) * 6 * (size_t)K, cudaMemcpyHostToDevice, ctx->stream));
  CB2_CUDA(ctx, cudaMemcpyAsync(base + o_sel, sel.data(), sizeof(int) * (size_t)ns, cudaMemcpyHostToDevice, ctx->stream));
  icp_normals_kernel<<<ns, kIcpThreads, 0, ctx->stream>>>((const double*)(base + o_fix), nf, (const int*)(base + o_sel), kn,
                                                            (double*)(base + o_nrm), (double*)(base + o_pl));
  CB2_CHECK_LAUNCH(ctx, "icp_normals_kernel");
  IcpArgs a;
  a.fix = (const double*)(base + o_fix); a.nf = nf; a.mov0 = (const double*)(base + o_mov0); a.nm = nm;
  a.perturb = perturb ? (const double*)(base + o_pert) : nullptr;
  a.sel = (const int*)(base + o_sel); a.ns = ns; a.normals = (const double*)(base + o_nrm); a.plan = (const double*)(base + o_pl);
  a.mov = (double*)(base + o_mov); a.min_planarity = min_planarity; a.min_change = min_change; a.max_iterations = max_iterations;
  a.H_out = (double*)(base + o_H); a.stats_out = (double*)(base + o_st);
  icp_iterate_kernel<<<K, kIcpThreads, 0, ctx->stream>>>(a);
  CB2_CHECK_LAUNCH(ctx, "icp_iterate_kernel");
  CB2_CUDA(ctx, cudaMemcpyAsync(H_out, base + o_H, sizeof(double) * 16 * (size_t)K, cudaMemcpyDeviceToHost, ctx->stream));
  if (stats_out) CB2_CUDA(ctx, cudaMemcpyAsync(stats_out, base + o_st, sizeof(double) * 4 * (size_t)K, cudaMemcpyDeviceToHost, ctx->stream));
  CB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return CB2_OK;
}
