// product.cu -- Gibbs-sampled products of Gaussian-kernel density messages on sm_100a (K5/K6), replacing
// KernelDensityEstimate.prodAppxMSGibbsS behind ApproxManifoldProducts.manifoldProduct
// (call site R:ext/services/additional.jl:24; algorithm: SURVEY.md App. A.3, restated in oracle/cb2_oracle.c).
//
// Two kernels per batch:
//   tree_build_kernel : one CTA per message.  Level hierarchy of the ball tree by recursive median split on the
//                       widest coordinate: per level a shared-memory bitonic sort on (node, coordinate, position),
//                       always FP64 so that FP32 and FP64 mode (and the oracle) see the same tree.  Node
//                       statistics bottom-up in FP64, written as padded records in the compute precision.
//   gibbs_kernel      : one thread per output sample, one CTA per block of S samples of one product; all chains
//                       of a CTA advance in lock step through (level, sweep, message).  The candidate nodes of
//                       the current (message, level) are streamed once per CTA through shared memory with 1-D
//                       TMA bulk copies (cp.async.bulk + mbarrier, double buffered) and read as warp-broadcast
//                       128-bit loads, so every candidate record is fetched once per 128 chains.  The categorical
//                       draw is a two-pass inverse CDF: pass 1 keeps per-chunk partial sums (running-max
//                       rescaled, log-sum-exp safe), pass 2 re-scores only the selected chunk.
//   Randomness: counter-based Philox4x32-10 keyed by (seed; sample, draw, stream, tag): results do not depend on
//   batching, CTA shape or output-sample sharding across GPUs.
#include "common.cuh"
#include <functional>

int cb2_bw_loo_dev_core(cb2_ctx* ctx, const double* const* mu_dev, const int32_t* N, int B, int d,
                        const uint8_t* circular, double* sigma_out_dev, void* scratch, size_t scratch_bytes,
                        size_t* scratch_needed);

namespace {

constexpr int kMaxLev = 15;        // depth <= 14 for N <= 16384
constexpr int kTreeThreads = 512;
constexpr int kS = 128;            // chains (threads) per Gibbs CTA
constexpr int kNCH = 32;           // chunk partial sums per chain
constexpr int kTileBytesF32 = 8192;
#ifndef CB2_KG
#define CB2_KG 8
#endif
constexpr int kG = CB2_KG;              // candidates scored together by one chain (ILP); tiles and chunks are multiples of kG

__host__ __device__ inline int recn_of(int d) { return (2 * d + 1 + 3) / 4 * 4; }  // [mu(d), var(d), log2 w]
__host__ __device__ inline int recl_of(int d) { return (d + 1 + 3) / 4 * 4; }      // [mu(d), log2 w]
__host__ __device__ inline int ceil_log2(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }
__host__ __device__ inline int level_count(int N, int depth, int l) { return l < depth ? (1 << l) : N; }

// leaf-position range [lo, lo+n) of node k at level l (< depth) under the balanced split left = ceil(n/2)
__host__ __device__ inline void node_range(int N, int l, int k, int& lo, int& n) {
  lo = 0; n = N;
  for (int b = l - 1; b >= 0; --b) {
    int left = (n + 1) >> 1;
    if ((k >> b) & 1) { lo += left; n -= left; } else n = left;
  }
}
__host__ __device__ inline int node_of(int N, int l, int p) {
  int lo = 0, n = N, k = 0;
  for (int b = 0; b < l; ++b) {
    int left = (n + 1) >> 1;
    if (p < lo + left) { k = 2 * k; n = left; } else { k = 2 * k + 1; lo += left; n -= left; }
  }
  return k;
}

// ---- tensor-core leaf scoring (FP32, d = 6; gibbs.cuh `leaf_pass_tc`): per message a block of leaf FEATURE tiles in the node
// pool.  A tile = 64 leaves x 16 features [x_i^2 (6), x_i (6), 1, log2 w, 0, 0], split into TF32 hi and lo parts, each laid out
// as the K-major no-swizzle operand of tcgen05.mma (8 x 16-byte core matrices: [leaf/8][k/4][leaf%8][k%4]).
#ifndef CB2_GIBBS_TC
#define CB2_GIBBS_TC 1
#endif
#ifndef CB2_COARSE_RSQ
#define CB2_COARSE_RSQ 1   // coarse levels of the packed FP32 d = 6 loops: 2 MUFU.RSQ + 1 MUFU.EX2 per candidate (gibbs.cuh)
#endif
#ifndef CB2_ONE_DEN
#define CB2_ONE_DEN 1      // ... over one common denominator (2 MUFU) where the six-fold product of variances stays in range
#endif
#ifndef CB2_TC_TILE
#define CB2_TC_TILE 128
#endif
constexpr int kTcTile = CB2_TC_TILE;            // leaves per feature tile (= N of one MMA): 64 double-buffered, 128 single-buffered
constexpr int kTcTileFloats = 2 * kTcTile * 16; // hi + lo
__host__ __device__ inline bool tc_enabled(int d, bool f32) { return CB2_GIBBS_TC && f32 && d == 6; }

struct DensMeta {
  const double* pts;  // d x N device Float64
  const double* w;    // N or null
  double bw[CB2_MAX_DIM];
  int N, depth;
  long long rec_off[kMaxLev + 1];  // element offset of each level's records in the node pool
  long long perm_off;              // offset in the int pool
  long long feat_off;              // element offset of the leaf feature tiles in the node pool (tensor-core path) or -1
  const double* center;            // d values (device): Euclidean coordinates are stored relative to this point
};

struct ProdDev {
  int D, Nout, sweeps, L;
  uint32_t stream;
  int sample_offset;
  int dens[CB2_MAX_DENS];
  double* out;   // d x Nout device
  int* labels;   // D x Nout device or null
};

struct Work { int prod, s0; };

// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__device__ __forceinline__ unsigned long long ordered_bits(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

__device__ __forceinline__ double from_ordered_bits(unsigned long long o) {
  return __longlong_as_double((long long)((o & 0x8000000000000000ull) ? (o & 0x7fffffffffffffffull) : ~o));
}
constexpr int kCoopNodes = 32;  // tree levels with at most this many nodes find their split coordinate cooperatively

__device__ __forceinline__ bool key_less(unsigned long long oa, uint32_t ta, unsigned long long ob, uint32_t tb) {
  uint32_t na = ta >> 14, nb = tb >> 14;
  if (na != nb) return na < nb;
  if (oa != ob) return oa < ob;
  return ta < tb;
}

#ifdef CB2_TREE_PROF
#define TREE_T(slot) do { __syncthreads(); long long t_ = clock64(); tprof[slot] += t_ - tlast; tlast = t_; } while (0)
#else
#define TREE_T(slot) do { } while (0)
#endif

// tensor-core leaf path: feature tiles (see kTcTile) of one message, built from the stored FP32 leaf records
template <typename T>
__device__ __forceinline__ void write_feature_tiles(T* __restrict__ feat_base, const T* __restrict__ rec, int N, int RECL, int tid, int nthr) {
  float* feat = reinterpret_cast<float*>(feat_base);
  const int npad = (N + kTcTile - 1) / kTcTile * kTcTile;
  for (int p = tid; p < npad; p += nthr) {
    double f[16];
    if (p < N) {
      for (int i = 0; i < 6; ++i) {
        const double xf = (double)(float)rec[(size_t)p * RECL + i];
        f[i] = xf * xf; f[6 + i] = xf;
      }
      f[12] = 1.0; f[13] = (double)(float)rec[(size_t)p * RECL + 6]; f[14] = 0.0; f[15] = 0.0;
    } else {  // padding leaf of the last tile: weight 2^-1e30 = 0
      for (int k = 0; k < 16; ++k) f[k] = 0.0;
      f[13] = -1e30;
    }
    float* tile = feat + (size_t)(p / kTcTile) * kTcTileFloats;
    const int r = p % kTcTile;
    for (int k = 0; k < 16; ++k) {
      const float hi = tf32_rna((float)f[k]);
      const float lo = tf32_rna((float)(f[k] - (double)hi));
      const int off = (r >> 3) * 128 + (k >> 2) * 32 + (r & 7) * 4 + (k & 3);
      tile[off] = hi;
      tile[kTcTile * 16 + off] = lo;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kTreeThreads, 2)
tree_build_kernel(const DensMeta* __restrict__ metas, int d, int* __restrict__ int_pool, T* __restrict__ node_pool,
                  double* __restrict__ stat_scratch, size_t stat_stride, uint32_t circ_mask) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const DensMeta m = metas[blockIdx.x];
  const int N = m.N, depth = m.depth, tid = threadIdx.x, nthr = blockDim.x;
  int P2 = 1;
  while (P2 < N) P2 <<= 1;
  unsigned long long* ord = reinterpret_cast<unsigned long long*>(smem_raw);
  uint32_t* tag = reinterpret_cast<uint32_t*>(ord + P2);
  unsigned char* seg_dim = reinterpret_cast<unsigned char*>(tag + P2);
  __shared__ double s_red[32];
  __shared__ double s_wsum;
  __shared__ unsigned long long s_ext[2 * kCoopNodes * CB2_MAX_DIM];  // (min, max) per node and coordinate
  int* perm = int_pool + m.perm_off;
  const bool pow2 = (P2 == N);

  for (int p = tid; p < N; p += nthr) perm[p] = p;
  __syncthreads();
#ifdef CB2_TREE_PROF
  long long tprof[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
#endif

  for (int l = 0; l < depth; ++l) {
    const int cnt = 1 << l;
    if (cnt <= kCoopNodes) {
      // shallow levels (few, large nodes): all threads sweep the points once, warp-reduce runs of lanes that sit in the
      // same node, and fold into per-node extrema with shared atomics on the order-preserving bit pattern
      for (int q = tid; q < 2 * kCoopNodes * CB2_MAX_DIM; q += nthr) s_ext[q] = (q & 1) ? 0ull : ~0ull;
      __syncthreads();
      for (int p0 = 0; p0 < N; p0 += nthr) {
        const int p = p0 + tid;
        const bool in = p < N;
        const int k = in ? node_of(N, l, p) : -1;
        const int k0 = __shfl_sync(0xffffffffu, k, 0);
        const bool uni = __all_sync(0xffffffffu, k == k0);
        const size_t row = in ? (size_t)perm[p] * d : 0;
        unsigned long long lo[CB2_MAX_DIM], hi[CB2_MAX_DIM];
#pragma unroll
        for (int i = 0; i < CB2_MAX_DIM; ++i) {  // every coordinate of the point in flight before any reduction
          lo[i] = (in && i < d) ? ordered_bits(m.pts[row + i]) : ~0ull;
          hi[i] = (in && i < d) ? lo[i] : 0ull;
        }
#pragma unroll
        for (int i = 0; i < CB2_MAX_DIM; ++i) {
          if (i >= d) break;
          if (uni) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              unsigned long long l2 = __shfl_xor_sync(0xffffffffu, lo[i], o), h2 = __shfl_xor_sync(0xffffffffu, hi[i], o);
              lo[i] = l2 < lo[i] ? l2 : lo[i]; hi[i] = h2 > hi[i] ? h2 : hi[i];
            }
            if ((tid & 31) == 0 && in) { atomicMin(&s_ext[(k * CB2_MAX_DIM + i) * 2], lo[i]); atomicMax(&s_ext[(k * CB2_MAX_DIM + i) * 2 + 1], hi[i]); }
          } else if (in) {
            atomicMin(&s_ext[(k * CB2_MAX_DIM + i) * 2], lo[i]); atomicMax(&s_ext[(k * CB2_MAX_DIM + i) * 2 + 1], hi[i]);
          }
        }
      }
      __syncthreads();
      if (tid < cnt) {
        int best = 0;
        double bestr = -1.0;
        for (int i = 0; i < d; ++i) {
          const double mn = from_ordered_bits(s_ext[(tid * CB2_MAX_DIM + i) * 2]), mx = from_ordered_bits(s_ext[(tid * CB2_MAX_DIM + i) * 2 + 1]);
          if (mx - mn > bestr) { bestr = mx - mn; best = i; }
        }
        seg_dim[tid] = (unsigned char)best;
      }
    } else {
      // deeper levels: one group of g lanes per node, each lane gathers at most 8 points with all their coordinates in
      // flight together (a warp per node would leave most lanes idle and serialise one load latency per node and coordinate)
      const int nmax = (N + cnt - 1) / cnt;
      int g = 1;
      while (g < 32 && g * 8 < nmax) g <<= 1;
      const int ngroups = nthr / g, gid = tid / g, sub = tid & (g - 1);
      for (int kb = 0; kb < cnt; kb += ngroups) {
        const int k = kb + gid;
        const bool valid = k < cnt;
        int lo = 0, n = 0;
        if (valid) node_range(N, l, k, lo, n);
        double mn[CB2_MAX_DIM], mx[CB2_MAX_DIM];
#pragma unroll
        for (int i = 0; i < CB2_MAX_DIM; ++i) { mn[i] = 1e300; mx[i] = -1e300; }
#pragma unroll 4
        for (int p = lo + sub; p < lo + n; p += g) {
          const size_t row = (size_t)perm[p] * d;
#pragma unroll
          for (int i = 0; i < CB2_MAX_DIM; ++i) {
            if (i < d) { const double v = m.pts[row + i]; mn[i] = fmin(mn[i], v); mx[i] = fmax(mx[i], v); }
          }
        }
        int best = 0;
        double bestr = -1.0;
#pragma unroll
        for (int i = 0; i < CB2_MAX_DIM; ++i) {
          if (i >= d) break;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            if (o < g) {
              mn[i] = fmin(mn[i], __shfl_xor_sync(0xffffffffu, mn[i], o));
              mx[i] = fmax(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], o));
            }
          }
          if (mx[i] - mn[i] > bestr) { bestr = mx[i] - mn[i]; best = i; }
        }
        if (valid && sub == 0) seg_dim[k] = (unsigned char)best;
      }
    }
    __syncthreads();
    TREE_T(0);
    for (int p = tid; p < P2; p += nthr) {
      if (p < N) {
        int k = node_of(N, l, p);
        ord[p] = ordered_bits(m.pts[(size_t)perm[p] * d + seg_dim[k]]);
        tag[p] = ((uint32_t)k << 14) | (uint32_t)p;
      } else {
        ord[p] = ~0ull; tag[p] = ~0u;
      }
    }
    __syncthreads();
    TREE_T(1);
    if (!pow2 && N <= 2 * kTreeThreads && (N + cnt - 1) / cnt <= 128) {
      // ragged N, small nodes: the bitonic network below would sort all P2 keys again on every level (55 rounds at N = 1000)
      // although the elements already sit in their nodes; counting ranks inside the node gives the same (node, value, position)
      // order in at most 128 compares per element
      int dst[2] = {0, 0}, val[2] = {0, 0};
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int p = tid + c * kTreeThreads;
        if (p < N) {
          const unsigned long long op = ord[p];
          const uint32_t tp = tag[p];
          int lo, n, rank = 0;
          node_range(N, l, (int)(tp >> 14), lo, n);
#pragma unroll 4
          for (int q = lo; q < lo + n; ++q) rank += (ord[q] < op || (ord[q] == op && tag[q] < tp)) ? 1 : 0;
          dst[c] = lo + rank;
          val[c] = perm[p];
        }
      }
      __syncthreads();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int p = tid + c * kTreeThreads;
        if (p < N) perm[dst[c]] = val[c];
      }
      __syncthreads();
      TREE_T(2);
      continue;
    }
    // bitonic sort; when N is a power of two the nodes are aligned blocks of N>>l, sort inside blocks only
    const int kmax = pow2 ? (N >> l) : P2;
    // one compare-exchange per thread and round: pair q = (i, i | j)
    auto cmpx = [&](int q, int j, int k) {
      const int i = ((q & ~(j - 1)) << 1) | (q & (j - 1)), ixj = i | j;
      const bool up = (i & k) == 0 || k == kmax;  // the last merge leaves every block ascending
      const unsigned long long oa = ord[i], ob = ord[ixj];
      const uint32_t ta = tag[i], tb = tag[ixj];
      const bool lt = key_less(ob, tb, oa, ta);  // b < a
      if (lt == up) { ord[i] = ob; ord[ixj] = oa; tag[i] = tb; tag[ixj] = ta; }
    };
    // every warp owns a contiguous run of PW pairs, i.e. a window of 2 PW elements: rounds with j <= PW never leave the
    // window and need a warp barrier only; whole merges k <= 2 PW run without any block barrier
    const int npairs = P2 >> 1;
    const int PW = max(32, npairs / (nthr >> 5));
    const int wbase = (tid >> 5) * PW + (tid & 31);
    for (int k = 2; k <= kmax; k <<= 1) {
      int j = k >> 1;
      for (; j > PW; j >>= 1) {
        for (int c = 0; c < PW; c += 32)
          if (wbase + c < npairs) cmpx(wbase + c, j, k);
        __syncthreads();
      }
      for (; j > 0; j >>= 1) {
        for (int c = 0; c < PW; c += 32)
          if (wbase + c < npairs) cmpx(wbase + c, j, k);
        __syncwarp();
      }
      if (k > PW || k == kmax) __syncthreads();
    }
    TREE_T(2);
    // apply the permutation: position p now holds the element that sat at position tag&0x3fff
    int newp[CB2_MAX_PRODUCT_N / kTreeThreads];
    for (int p = tid, c = 0; p < N; p += nthr, ++c) newp[c] = perm[tag[p] & 0x3fffu];
    __syncthreads();
    for (int p = tid, c = 0; p < N; p += nthr, ++c) perm[p] = newp[c];
    __syncthreads();
    TREE_T(3);
  }

  // ---- statistics, bottom-up in FP64 (global scratch: per level mean[cnt*d], var[cnt*d], wt[cnt])
  double wloc = 0;
  if (m.w) for (int p = tid; p < N; p += nthr) wloc += m.w[p];
  wloc = warp_sum(wloc);
  if ((tid & 31) == 0) s_red[tid >> 5] = wloc;
  __syncthreads();
  if (tid == 0) { double s = 0; for (int w = 0; w < (nthr >> 5); ++w) s += s_red[w]; s_wsum = s; }
  __syncthreads();
  const double wsum = s_wsum;
  double* sbase = stat_scratch + (size_t)blockIdx.x * stat_stride;
  // level l statistics start at lvl_off(l) nodes; each node uses (2d+1) doubles
  auto lvl_off = [&](int l) { return l < depth ? ((1 << l) - 1) : ((1 << depth) - 1); };
  const int nd = 2 * d + 1;
  const int RECN = recn_of(d), RECL = recl_of(d);
  // range of the stored coordinates over every level: lets a Gibbs chain prove that no candidate of this message is more than
  // pi away from its current mean, i.e. that the wrap is a no-op (gibbs.cuh, `nowrap`), and bounds every node variance by
  // (range / 2)^2 + bandwidth^2 (`one_den`)
  double cmn[CB2_MAX_DIM], cmx[CB2_MAX_DIM];
#pragma unroll
  for (int i = 0; i < CB2_MAX_DIM; ++i) { cmn[i] = 1e300; cmx[i] = -1e300; }
  double cen[CB2_MAX_DIM];
#pragma unroll
  // Euclidean coordinates are stored relative to the product's centre (translation invariance).  In FP32 mode the circular
  // coordinates are rotated the same way (rotation invariance on the circle): a belief narrower than pi then never straddles
  // the +-pi cut of the stored values and the Gibbs kernel can take its wrap-free loops.  FP64 (validation) mode keeps the
  // unrotated angles so that it follows the oracle operation for operation.
  for (int i = 0; i < CB2_MAX_DIM; ++i) cen[i] = (i < d && (sizeof(T) == 4 || !(circ_mask >> i & 1))) ? m.center[i] : 0.0;
  {
    double* sl = sbase + (size_t)lvl_off(depth) * nd;
    T* rec = node_pool + m.rec_off[depth];
    for (int p = tid; p < N; p += nthr) {
      int idx = perm[p];
      double wt = m.w ? m.w[idx] / wsum : 1.0 / N;
#pragma unroll
      for (int i = 0; i < CB2_MAX_DIM; ++i) {
        if (i >= d) break;
        double v = m.pts[(size_t)idx * d + i];
        sl[(size_t)p * nd + i] = v;
        sl[(size_t)p * nd + d + i] = m.bw[i] * m.bw[i];
        const double rv = (circ_mask >> i & 1) ? wrap_pi(v - cen[i]) : v - cen[i];
        rec[(size_t)p * RECL + i] = (T)rv;
        cmn[i] = fmin(cmn[i], rv); cmx[i] = fmax(cmx[i], rv);
      }
      sl[(size_t)p * nd + 2 * d] = wt;
      for (int i = d; i < RECL; ++i) rec[(size_t)p * RECL + i] = T(0);
      rec[(size_t)p * RECL + d] = (T)(wt > 0 ? log2(wt) : -1e30);
    }
    if (m.feat_off >= 0 && d == 6) write_feature_tiles<T>(node_pool + m.feat_off, rec, N, RECL, tid, nthr);
  }
  __syncthreads();
  TREE_T(4);
  for (int l = depth - 1; l >= 0; --l) {
    const int cnt = 1 << l;
    double* sl = sbase + (size_t)lvl_off(l) * nd;
    const double* sc = sbase + (size_t)lvl_off(l + 1) * nd;
    T* rec = node_pool + m.rec_off[l];
    for (int k = tid; k < cnt; k += nthr) {
      int c0, nch;
      if (l + 1 == depth) { node_range(N, l, k, c0, nch); } else { c0 = 2 * k; nch = 2; }
      // both children (a node of the last split level may have one) are fetched before any arithmetic: the records were
      // written by other threads of this CTA, so nothing can be kept in registers across levels, but all loads overlap
      const double* r0 = sc + (size_t)c0 * nd;
      const double* r1 = r0 + (nch > 1 ? nd : 0);
      const double w0 = r0[2 * d], w1 = nch > 1 ? r1[2 * d] : 0.0;
      double mu0[CB2_MAX_DIM], mu1[CB2_MAX_DIM], va0[CB2_MAX_DIM], va1[CB2_MAX_DIM];
#pragma unroll
      for (int i = 0; i < CB2_MAX_DIM; ++i) {
        if (i < d) { mu0[i] = r0[i]; va0[i] = r0[d + i]; mu1[i] = r1[i]; va1[i] = r1[d + i]; }
      }
      const double ws = w0 + w1;
      sl[(size_t)k * nd + 2 * d] = ws;
#pragma unroll
      for (int i = 0; i < CB2_MAX_DIM; ++i) {
        if (i >= d) break;
        double mm = w0 * mu0[i];
        mm += w1 * mu1[i];
        mm /= ws;
        const double d0 = mu0[i] - mm, d1 = mu1[i] - mm;
        double v = w0 * (va0[i] + d0 * d0);
        v += w1 * (va1[i] + d1 * d1);
        v /= ws;
        sl[(size_t)k * nd + i] = mm;
        sl[(size_t)k * nd + d + i] = v;
        const double rv = (circ_mask >> i & 1) ? wrap_pi(mm - cen[i]) : mm - cen[i];
        rec[(size_t)k * RECN + i] = (T)rv;
        cmn[i] = fmin(cmn[i], rv); cmx[i] = fmax(cmx[i], rv);
        rec[(size_t)k * RECN + d + i] = (T)v;
      }
      for (int i = 2 * d; i < RECN; ++i) rec[(size_t)k * RECN + i] = T(0);
      rec[(size_t)k * RECN + 2 * d] = (T)(ws > 0 ? log2(ws) : -1e30);
    }
    __syncthreads();
  }
  TREE_T(5);
  // block-wide min / max of the circular record values -> last 16 doubles of this message's statistics scratch
#pragma unroll
  for (int i = 0; i < CB2_MAX_DIM; ++i) {
    if (i >= d) break;
    double a = cmn[i], b = cmx[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a = fmin(a, __shfl_xor_sync(0xffffffffu, a, o));
      b = fmax(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
    __syncthreads();
    if ((tid & 31) == 0) { s_red[tid >> 5] = a; }
    __syncthreads();
    if (tid == 0) { for (int w = 1; w < (nthr >> 5); ++w) a = fmin(a, s_red[w]); sbase[stat_stride - 16 + i] = a; }
    __syncthreads();
    if ((tid & 31) == 0) { s_red[tid >> 5] = b; }
    __syncthreads();
    if (tid == 0) { for (int w = 1; w < (nthr >> 5); ++w) b = fmax(b, s_red[w]); sbase[stat_stride - 8 + i] = b; }
  }
#ifdef CB2_TREE_PROF
  TREE_T(6);
  if (tid == 0 && blockIdx.x == 0)
    printf("tree prof (clk): range %lld fill %lld sort %lld perm %lld leafstat %lld upstat %lld final %lld\n", tprof[0], tprof[1], tprof[2],
           tprof[3], tprof[4], tprof[5], tprof[6]);
#endif
}

// ---- trees of small messages (N <= kTreeSmallN: the 100 .. 256 particles of an IIF solve) ---------------------------------------
// The general kernel above is ~10^4 instructions of straight-line code per level, all of it executed once per level by a CTA
// that holds 100 points: 80 us, most of it instruction fetch and dependent global round trips.  This kernel is the same
// algorithm written for latency: one thread per point, points / permutation / statistics in shared memory, loops over the
// coordinates not unrolled, the order inside the nodes by counting ranks (a 100-element node is 100 compares per thread) instead of
// a bitonic network.  Same total order (node, value, position), same arithmetic in the statistics: bit-identical output.
constexpr int kTreeSmallN = 256;
template <typename T>
__global__ void __launch_bounds__(kTreeSmallN)
tree_build_small_kernel(const DensMeta* __restrict__ metas, int d, int* __restrict__ int_pool, T* __restrict__ node_pool,
                        double* __restrict__ stat_scratch, size_t stat_stride, uint32_t circ_mask) {
  extern __shared__ __align__(16) unsigned char smem_small[];
  const DensMeta m = metas[blockIdx.x];
  const int N = m.N, depth = m.depth, tid = threadIdx.x;
  const int nd = 2 * d + 1, RECN = recn_of(d), RECL = recl_of(d);
  const int nstat = ((1 << depth) - 1 + N) * nd;
  double* spts = reinterpret_cast<double*>(smem_small);                 // [N][d]
  double* sstat = spts + (size_t)N * d;                                  // per level: [cnt][mean d, var d, weight]
  unsigned long long* key = reinterpret_cast<unsigned long long*>(sstat + nstat);  // [N] split coordinate of the point at a position
  double* ext = reinterpret_cast<double*>(key + N);                      // [nodes of a level <= kTreeSmallN / 2][d]: coordinate ranges
  int* perm = reinterpret_cast<int*>(ext + (kTreeSmallN / 2) * CB2_MAX_DIM);  // [N]
  unsigned char* seg_dim = reinterpret_cast<unsigned char*>(perm + N);   // [2^(depth-1)]
  __shared__ double s_red[kTreeSmallN / 32];
  __shared__ double s_rng[(kTreeSmallN / 32) * CB2_MAX_DIM * 2];
  for (int q = tid; q < N * d; q += kTreeSmallN) spts[q] = m.pts[q];
  if (tid < N) perm[tid] = tid;
  __syncthreads();
  for (int l = 0; l < depth; ++l) {
    const int cnt = 1 << l;
    // split coordinate of every node: thread (node, coordinate) scans the node's points
    for (int t = tid; t < cnt * d; t += kTreeSmallN) {
      const int k = t / d, i = t - k * d;
      int lo, n;
      node_range(N, l, k, lo, n);
      double mn = 1e300, mx = -1e300;
#pragma unroll 4
      for (int p = lo; p < lo + n; ++p) {
        const double v = spts[perm[p] * d + i];
        mn = fmin(mn, v); mx = fmax(mx, v);
      }
      ext[t] = mx - mn;
    }
    __syncthreads();
    for (int k = tid; k < cnt; k += kTreeSmallN) {
      int best = 0;
      double bestr = -1.0;
      for (int i = 0; i < d; ++i) {
        const double r = ext[k * d + i];
        if (r > bestr) { bestr = r; best = i; }
      }
      seg_dim[k] = (unsigned char)best;
    }
    __syncthreads();
    int lo = 0, n = 0, val = 0;
    if (tid < N) {
      const int k = node_of(N, l, tid);
      node_range(N, l, k, lo, n);
      val = perm[tid];
      key[tid] = ordered_bits(spts[val * d + seg_dim[k]]);
    }
    __syncthreads();
    if (tid < N) {
      const unsigned long long mine = key[tid];
      int rank = 0;
#pragma unroll 4
      for (int q = lo; q < lo + n; ++q) {
        const unsigned long long o = key[q];
        rank += (o < mine || (o == mine && q < tid)) ? 1 : 0;
      }
      perm[lo + rank] = val;   // every position of the node is written exactly once; `val` was read before the barrier
    }
    __syncthreads();
  }
  // ---- statistics, bottom-up in FP64 (same arithmetic as tree_build_kernel)
  double wloc = (m.w && tid < N) ? m.w[tid] : 0.0;
  wloc = warp_sum(wloc);
  if ((tid & 31) == 0) s_red[tid >> 5] = wloc;
  __syncthreads();
  double wsum = 0;
  for (int w = 0; w < kTreeSmallN / 32; ++w) wsum += s_red[w];
  auto lvl_off = [&](int l) { return l < depth ? ((1 << l) - 1) : ((1 << depth) - 1); };
  double cmn[CB2_MAX_DIM], cmx[CB2_MAX_DIM];
#pragma unroll
  for (int i = 0; i < CB2_MAX_DIM; ++i) { cmn[i] = 1e300; cmx[i] = -1e300; }
  auto centre = [&](int i) { return (sizeof(T) == 4 || !(circ_mask >> i & 1)) ? m.center[i] : 0.0; };
  {
    double* sl = sstat + (size_t)lvl_off(depth) * nd;
    T* rec = node_pool + m.rec_off[depth];
    if (tid < N) {
      const int p = tid, idx = perm[p];
      const double wt = m.w ? m.w[idx] / wsum : 1.0 / N;
#pragma unroll
      for (int i = 0; i < CB2_MAX_DIM; ++i) {
        if (i >= d) break;
        const double v = spts[idx * d + i];
        sl[p * nd + i] = v;
        sl[p * nd + d + i] = m.bw[i] * m.bw[i];
        const double c = centre(i);
        const double rv = (circ_mask >> i & 1) ? wrap_pi(v - c) : v - c;
        rec[(size_t)p * RECL + i] = (T)rv;
        cmn[i] = fmin(cmn[i], rv); cmx[i] = fmax(cmx[i], rv);
      }
      sl[p * nd + 2 * d] = wt;
      for (int i = d; i < RECL; ++i) rec[(size_t)p * RECL + i] = T(0);
      rec[(size_t)p * RECL + d] = (T)(wt > 0 ? log2(wt) : -1e30);
    }
    if (m.feat_off >= 0 && d == 6) {
      __syncthreads();  // the feature tiles read the leaf records back
      write_feature_tiles<T>(node_pool + m.feat_off, rec, N, RECL, tid, kTreeSmallN);
    }
  }
  __syncthreads();
  for (int l = depth - 1; l >= 0; --l) {
    const int cnt = 1 << l;
    double* sl = sstat + (size_t)lvl_off(l) * nd;
    const double* sc = sstat + (size_t)lvl_off(l + 1) * nd;
    T* rec = node_pool + m.rec_off[l];
    for (int k = tid; k < cnt; k += kTreeSmallN) {
      int c0, nch;
      if (l + 1 == depth) { node_range(N, l, k, c0, nch); } else { c0 = 2 * k; nch = 2; }
      const double* r0 = sc + (size_t)c0 * nd;
      const double* r1 = r0 + (nch > 1 ? nd : 0);
      const double w0 = r0[2 * d], w1 = nch > 1 ? r1[2 * d] : 0.0;
      const double ws = w0 + w1;
      sl[k * nd + 2 * d] = ws;
#pragma unroll
      for (int i = 0; i < CB2_MAX_DIM; ++i) {
        if (i >= d) break;
        const double mu0 = r0[i], va0 = r0[d + i], mu1 = r1[i], va1 = r1[d + i];
        double mm = w0 * mu0;
        mm += w1 * mu1;
        mm /= ws;
        const double d0 = mu0 - mm, d1 = mu1 - mm;
        double v = w0 * (va0 + d0 * d0);
        v += w1 * (va1 + d1 * d1);
        v /= ws;
        sl[k * nd + i] = mm;
        sl[k * nd + d + i] = v;
        const double c = centre(i);
        const double rv = (circ_mask >> i & 1) ? wrap_pi(mm - c) : mm - c;
        rec[(size_t)k * RECN + i] = (T)rv;
        cmn[i] = fmin(cmn[i], rv); cmx[i] = fmax(cmx[i], rv);
        rec[(size_t)k * RECN + d + i] = (T)v;
      }
      for (int i = 2 * d; i < RECN; ++i) rec[(size_t)k * RECN + i] = T(0);
      rec[(size_t)k * RECN + 2 * d] = (T)(ws > 0 ? log2(ws) : -1e30);
    }
    __syncthreads();
  }
  // ranges of the stored circular coordinates, permutation and statistics to their global places
  double* sbase = stat_scratch + (size_t)blockIdx.x * stat_stride;
#pragma unroll
  for (int i = 0; i < CB2_MAX_DIM; ++i) {
    if (i >= d) break;
    double a = cmn[i], b = cmx[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a = fmin(a, __shfl_xor_sync(0xffffffffu, a, o));
      b = fmax(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
    if ((tid & 31) == 0) { s_rng[((tid >> 5) * CB2_MAX_DIM + i) * 2] = a; s_rng[((tid >> 5) * CB2_MAX_DIM + i) * 2 + 1] = b; }
  }
  __syncthreads();
  if (tid < d) {
    double a = 1e300, b = -1e300;
    for (int w = 0; w < kTreeSmallN / 32; ++w) { a = fmin(a, s_rng[(w * CB2_MAX_DIM + tid) * 2]); b = fmax(b, s_rng[(w * CB2_MAX_DIM + tid) * 2 + 1]); }
    sbase[stat_stride - 16 + tid] = a;
    sbase[stat_stride - 8 + tid] = b;
  }
  int* perm_g = int_pool + m.perm_off;
  if (tid < N) perm_g[tid] = perm[tid];
  for (int q = tid; q < nstat; q += kTreeSmallN) sbase[q] = sstat[q];
}
__host__ inline size_t tree_small_smem(int maxN, int max_nodes, int d) {
  return sizeof(double) * ((size_t)maxN * d + (size_t)max_nodes * (2 * d + 1) + (size_t)(kTreeSmallN / 2) * CB2_MAX_DIM) + 8 * (size_t)maxN + 4 * (size_t)maxN +
         (size_t)kTreeSmallN + 64;
}

// metas[i].bw[0..d) = *ptrs[i]: bandwidths that live on the device go straight into the message table (no host round trip)
__global__ void fill_meta_bw_kernel(DensMeta* __restrict__ metas, const double* const* __restrict__ ptrs, int n, int d) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = 0; k < d; ++k) metas[i].bw[k] = ptrs[i][k];
}
// *ptrs[i] = vals[i*d .. i*d+d)  (fresh bandwidths written back to the caller's device buffers)
__global__ void scatter_small_vectors_kernel(double* const* __restrict__ ptrs, int n, int d, const double* __restrict__ vals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = 0; k < d; ++k) ptrs[i][k] = vals[(size_t)i * d + k];
}

// One launch moves many arrays between pinned (mapped) host memory and the device: for a batch of thousands of messages the
// per-call cost of one cudaMemcpyAsync per array (several microseconds of host time each) is larger than the DMA time.
// seg[i] = {src, dst, bytes} (8-byte aligned, bytes a multiple of 4); blockIdx.y = segment, blockIdx.x strides over its words.
struct CopySeg { const double* src; double* dst; unsigned long long bytes; };
__global__ void __launch_bounds__(256) copy_segments_kernel(const CopySeg* __restrict__ segs) {
  const CopySeg sg = segs[blockIdx.y];
  const unsigned long long n = sg.bytes >> 3;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x)
    sg.dst[i] = sg.src[i];
  if ((sg.bytes & 4) && blockIdx.x == 0 && threadIdx.x == 0)
    reinterpret_cast<int*>(sg.dst)[2 * n] = reinterpret_cast<const int*>(sg.src)[2 * n];
}

// The same copy with a SMALL grid that loops over the segments: a transfer over the host link needs a few hundred KB in flight,
// not the whole GPU.  Used when the copy runs beside compute (chunked uploads under the tree build, sample points travelling home
// under the bandwidth search): the 2-D launch above floods every SM with link-stalled warps and the other kernel cannot start.
#ifndef CB2_COPY_CTAS
#define CB2_COPY_CTAS 48
#endif
constexpr int kCopyStreamCtas = CB2_COPY_CTAS;
__global__ void __launch_bounds__(256) copy_segments_stream_kernel(const CopySeg* __restrict__ segs, int nseg) {
  for (int s = blockIdx.x; s < nseg; s += gridDim.x) {
    const CopySeg sg = segs[s];
    const unsigned long long n = sg.bytes >> 3;
    unsigned long long i = threadIdx.x;
    for (; i + 3 * 256 < n; i += 4 * 256) {
      const double a = sg.src[i], b = sg.src[i + 256], c = sg.src[i + 512], e = sg.src[i + 768];
      sg.dst[i] = a; sg.dst[i + 256] = b; sg.dst[i + 512] = c; sg.dst[i + 768] = e;
    }
    for (; i < n; i += 256) sg.dst[i] = sg.src[i];
    if ((sg.bytes & 4) && threadIdx.x == 0) reinterpret_cast<int*>(sg.dst)[2 * n] = reinterpret_cast<const int*>(sg.src)[2 * n];
  }
}

// copies one level of the FP64 statistics scratch out for cb2_tree_debug
__global__ void tree_debug_copy_kernel(const double* __restrict__ sbase, int d, int level_node0, int cnt,
                                       double* __restrict__ mean, double* __restrict__ var, double* __restrict__ wt) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= cnt) return;
  const int nd = 2 * d + 1;
  const double* s = sbase + (size_t)(level_node0 + k) * nd;
  for (int i = 0; i < d; ++i) { mean[(size_t)k * d + i] = s[i]; var[(size_t)k * d + i] = s[d + i]; }
  wt[k] = s[2 * d];
}

// ---------------------------------------------------------------------------------------------------
// mbarrier + 1-D TMA bulk copy helpers (sm_90+ PTX; SASS: UBLKCP / SYNCS)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

#include "gibbs.cuh"


template <typename T, int DIM, uint32_t CIRC, bool SMALL = false>
int launch_gibbs_inst(cb2_ctx* ctx, int nwork, const GibbsArgs& a, size_t stage_bytes = 0) {
  size_t smem = gibbs_smem_bytes<T, DIM>();
  if (SMALL) smem = ((smem + 127) & ~(size_t)127) + stage_bytes;
  cudaError_t e = cb2_func_smem(ctx, (const void*)gibbs_kernel<T, DIM, CIRC, SMALL>, smem, true);
  if (e != cudaSuccess) return cb2_fail(ctx, CB2_ERR_CUDA, "gibbs smem attribute: %s", cudaGetErrorString(e));
  gibbs_kernel<T, DIM, CIRC, SMALL><<<nwork, kS / ChainsPerThread<T, DIM>::value, smem, ctx->stream>>>(a);
  CB2_CHECK_LAUNCH(ctx, "gibbs_kernel");
  return CB2_OK;
}

// stage_bytes > 0: every product of the batch is small enough to keep all levels of its messages in shared memory (the SMALL
// kernels exist for the planar variable types an IIF solve at N ~ 100 spends its time on)
template <typename T>
int launch_gibbs(cb2_ctx* ctx, int d, uint32_t mask, int nwork, const GibbsArgs& a, size_t stage_bytes) {
  // the common variable types get compile-time circular masks; anything else takes the run-time mask
#ifdef CB2_FAST_BUILD  // experiment builds (bench/micro/build_variant.py): the headline instance only, seconds instead of minutes
  if (d == 6 && mask == 0x38u && sizeof(T) == 4) return launch_gibbs_inst<float, 6, 0x38u>(ctx, nwork, a);
  return cb2_fail(ctx, CB2_ERR_UNSUPPORTED, "CB2_FAST_BUILD carries the FP32 Pose3 product kernel only");
#else
  if (stage_bytes > 0) {
    if (d == 2 && mask == 0) return launch_gibbs_inst<T, 2, 0u, true>(ctx, nwork, a, stage_bytes);   // Point2
    if (d == 3 && mask == 4u) return launch_gibbs_inst<T, 3, 4u, true>(ctx, nwork, a, stage_bytes);  // Pose2
  }
  if (d == 1 && mask == 0) return launch_gibbs_inst<T, 1, 0u>(ctx, nwork, a);
  if (d == 2 && mask == 0) return launch_gibbs_inst<T, 2, 0u>(ctx, nwork, a);            // Point2
  if (d == 3 && mask == 0) return launch_gibbs_inst<T, 3, 0u>(ctx, nwork, a);            // Point3
  if (d == 3 && mask == 4u) return launch_gibbs_inst<T, 3, 4u>(ctx, nwork, a);           // Pose2
  if (d == 6 && mask == 0x38u) return launch_gibbs_inst<T, 6, 0x38u>(ctx, nwork, a);     // Pose3
  switch (d) {
    case 1: return launch_gibbs_inst<T, 1, ~0u>(ctx, nwork, a);
    case 2: return launch_gibbs_inst<T, 2, ~0u>(ctx, nwork, a);
    case 3: return launch_gibbs_inst<T, 3, ~0u>(ctx, nwork, a);
    case 4: return launch_gibbs_inst<T, 4, ~0u>(ctx, nwork, a);
    case 5: return launch_gibbs_inst<T, 5, ~0u>(ctx, nwork, a);
    case 6: return launch_gibbs_inst<T, 6, ~0u>(ctx, nwork, a);
  }
  return cb2_fail(ctx, CB2_ERR_UNSUPPORTED, "dimension %d", d);
#endif
}

struct BatchPlan {
  std::vector<DensMeta> metas;
  std::vector<ProdDev> prods;
  std::vector<Work> work;
  size_t node_elems = 0, int_elems = 0;
  int max_nodes = 0, maxP2 = 1;
};

int validate_and_plan(cb2_ctx* ctx, const cb2_product_desc* descs, int B, const uint8_t* circular, int d, BatchPlan& pl,
                      uint32_t* mask_out) {
  CB2_REQUIRE(ctx, descs && B > 0, "null descriptors or empty batch");
  if (d < 1 || d > CB2_MAX_PRODUCT_DIM) return cb2_fail(ctx, CB2_ERR_UNSUPPORTED, "product dimension %d outside 1..%d", d, CB2_MAX_PRODUCT_DIM);
  uint32_t mask = 0;
  for (int k = 0; k < d; ++k) if (circular && circular[k]) mask |= 1u << k;
  *mask_out = mask;
  const int RECN = recn_of(d), RECL = recl_of(d);
  for (int b = 0; b < B; ++b) {
    const cb2_product_desc& ds = descs[b];
    CB2_REQUIRE(ctx, ds.dens && ds.pts_out, "product %d: null dens / pts_out", b);
    if (ds.D < 1 || ds.D > CB2_MAX_DENS) return cb2_fail(ctx, CB2_ERR_UNSUPPORTED, "product %d: D=%d outside 1..%d", b, ds.D, CB2_MAX_DENS);
    CB2_REQUIRE(ctx, ds.Nout > 0 && ds.niter >= 0 && ds.sample_offset >= 0, "product %d: bad Nout/niter/sample_offset", b);
    ProdDev pd;
    memset(&pd, 0, sizeof(pd));
    pd.D = ds.D; pd.Nout = ds.Nout; pd.sweeps = 1 + ds.niter; pd.stream = ds.stream; pd.sample_offset = ds.sample_offset;
    int maxN = 0;
    for (int j = 0; j < ds.D; ++j) {
      const cb2_density& dn = ds.dens[j];
      CB2_REQUIRE(ctx, dn.pts && dn.bw && dn.N > 0, "product %d message %d: null pts/bw or N <= 0", b, j);
      if (dn.N > CB2_MAX_PRODUCT_N) return cb2_fail(ctx, CB2_ERR_UNSUPPORTED, "product %d message %d: N=%d exceeds %d", b, j, dn.N, CB2_MAX_PRODUCT_N);
      DensMeta m;
      memset(&m, 0, sizeof(m));
      m.N = dn.N; m.depth = ceil_log2(dn.N);
      m.perm_off = (long long)pl.int_elems;
      pl.int_elems += (size_t)dn.N;
      for (int l = 0; l <= m.depth; ++l) {
        m.rec_off[l] = (long long)pl.node_elems;
        pl.node_elems += (size_t)level_count(dn.N, m.depth, l) * (l < m.depth ? RECN : RECL);
        pl.node_elems = (pl.node_elems + 31) / 32 * 32;  // keep every level 128-byte aligned for the bulk copies
      }
      m.feat_off = -1;
      if (tc_enabled(d, ctx->precision == CB2_FP32)) {
        m.feat_off = (long long)pl.node_elems;
        pl.node_elems += (size_t)((dn.N + kTcTile - 1) / kTcTile) * kTcTileFloats;
      }
      int nodes = (1 << m.depth) + dn.N;
      if (nodes > pl.max_nodes) pl.max_nodes = nodes;
      int P2 = 1 << m.depth;
      if (P2 > pl.maxP2) pl.maxP2 = P2;
      if (dn.N > maxN) maxN = dn.N;
      pd.dens[j] = (int)pl.metas.size();
      pl.metas.push_back(m);
    }
    int L = 0;
    while ((2 << L) <= maxN) ++L;
    pd.L = L + 1;  // floor(log2(maxN)) + 1
    for (int s0 = 0; s0 < ds.Nout; s0 += kS) pl.work.push_back({b, s0});
    pl.prods.push_back(pd);
  }
  return CB2_OK;
}

struct DevLayout {
  DensMeta* metas; ProdDev* prods; Work* work; int* int_pool; void* node_pool; double* stat; size_t stat_stride;
};

// Core: descriptors hold DEVICE pointers (pts, w, outputs); bw values are read from `bw_host[b][j]` (host copies).
// inputs of a host-pointer batch that arrive chunk by chunk on the copy stream: messages [0, msg_end[c]) are on the device once
// ev[c] has fired
struct InputChunks { std::vector<int> msg_end; std::vector<cudaEvent_t> ev; };

template <typename T>
int product_batch_core(cb2_ctx* ctx, const cb2_product_desc* descs, int B, const std::vector<std::vector<const double*>>& bw_host,
                       const double* const* bw_ptrs_dev, uint32_t mask, int d, uint64_t seed, BatchPlan& pl, char* base, size_t bytes,
                       size_t* needed, DevLayout* lay_out, const InputChunks* chunks = nullptr) {
  const int nd = 2 * d + 1;
  const size_t stat_stride = (size_t)pl.max_nodes * nd + 16;  // + circular coordinate ranges (min[8], max[8])
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += cb2_align(n); return o; };
  size_t o_metas = take(sizeof(DensMeta) * pl.metas.size());
  size_t o_prods = take(sizeof(ProdDev) * pl.prods.size());
  size_t o_work = take(sizeof(Work) * pl.work.size());
  size_t o_int = take(sizeof(int) * pl.int_elems);
  size_t o_node = take(sizeof(T) * pl.node_elems + 256);
  size_t o_stat = take(sizeof(double) * stat_stride * pl.metas.size());
  if (needed) *needed = off + 1024;
  if (!base) return CB2_OK;
  if (off > bytes) return cb2_fail(ctx, CB2_ERR_NOMEM, "product scratch too small");
  DevLayout lay;
  lay.metas = (DensMeta*)(base + o_metas); lay.prods = (ProdDev*)(base + o_prods); lay.work = (Work*)(base + o_work);
  lay.int_pool = (int*)(base + o_int); lay.node_pool = base + o_node; lay.stat = (double*)(base + o_stat);
  lay.stat_stride = stat_stride;
  size_t mi = 0;
  for (int b = 0; b < B; ++b) {
    for (int j = 0; j < descs[b].D; ++j, ++mi) {
      pl.metas[mi].pts = descs[b].dens[j].pts;
      pl.metas[mi].w = descs[b].dens[j].w;
      pl.metas[mi].center = descs[b].dens[0].pts;  // first kernel of the first message: products are translation invariant
      for (int k = 0; k < d; ++k) {
        if (bw_ptrs_dev) { pl.metas[mi].bw[k] = 1.0; continue; }  // filled on the device below (fill_meta_bw_kernel)
        double s = bw_host[b][j][k];
        if (!(s > 0) || !isfinite(s)) return cb2_fail(ctx, CB2_ERR_INVALID, "product %d message %d: bandwidth[%d] must be positive", b, j, k);
        pl.metas[mi].bw[k] = s;
      }
    }
    pl.prods[b].out = descs[b].pts_out;
    pl.prods[b].labels = descs[b].labels_out;
  }
  {  // the three tables sit back to back at the start of the layout: one upload (each copy from pageable memory costs ~5 us of host time)
    const size_t tab_bytes = o_work + sizeof(Work) * pl.work.size();
    std::vector<char> tab(tab_bytes);
    memcpy(tab.data() + o_metas, pl.metas.data(), sizeof(DensMeta) * pl.metas.size());
    memcpy(tab.data() + o_prods, pl.prods.data(), sizeof(ProdDev) * pl.prods.size());
    memcpy(tab.data() + o_work, pl.work.data(), sizeof(Work) * pl.work.size());
    CB2_CUDA(ctx, cudaMemcpyAsync(base, tab.data(), tab_bytes, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (bw_ptrs_dev) {
    fill_meta_bw_kernel<<<(int)((pl.metas.size() + 127) / 128), 128, 0, ctx->stream>>>(lay.metas, bw_ptrs_dev, (int)pl.metas.size(), d);
    CB2_CHECK_LAUNCH(ctx, "fill_meta_bw_kernel");
  }

  CB2_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
  size_t tree_smem = (size_t)pl.maxP2 * 13 + 64;
  int tree_maxN = 0;
  for (const DensMeta& m : pl.metas) tree_maxN = std::max(tree_maxN, m.N);
  const size_t small_smem = tree_small_smem(tree_maxN, pl.max_nodes, d);
  // launches whose messages all hold at most kTreeSmallN points take the latency-oriented kernel
  const bool tree_small = tree_maxN <= kTreeSmallN && small_smem <= 160 * 1024 && !getenv("CB2_NO_TREE_SMALL");
  cudaError_t e = tree_small ? cb2_func_smem(ctx, (const void*)tree_build_small_kernel<T>, small_smem)
                             : cb2_func_smem(ctx, (const void*)tree_build_kernel<T>, tree_smem);
  if (e != cudaSuccess) return cb2_fail(ctx, CB2_ERR_CUDA, "tree smem attribute (%zu bytes): %s", tree_small ? small_smem : tree_smem, cudaGetErrorString(e));
  if (chunks && !chunks->msg_end.empty()) {  // build the trees of every chunk as soon as its arrays have landed
    int m0 = 0;
    for (size_t c = 0; c < chunks->msg_end.size(); ++c) {
      const int m1 = chunks->msg_end[c];
      CB2_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, chunks->ev[c], 0));
      if (m1 > m0) {
        if (tree_small)
          tree_build_small_kernel<T><<<m1 - m0, kTreeSmallN, small_smem, ctx->stream>>>(lay.metas + m0, d, lay.int_pool, (T*)lay.node_pool,
                                                                                         lay.stat + (size_t)m0 * stat_stride, stat_stride, mask);
        else
        tree_build_kernel<T><<<m1 - m0, kTreeThreads, tree_smem, ctx->stream>>>(lay.metas + m0, d, lay.int_pool, (T*)lay.node_pool,
                                                                                 lay.stat + (size_t)m0 * stat_stride, stat_stride, mask);
        CB2_CHECK_LAUNCH(ctx, "tree_build_kernel");
      }
      m0 = m1;
    }
  } else if (tree_small) {
    tree_build_small_kernel<T><<<(int)pl.metas.size(), kTreeSmallN, small_smem, ctx->stream>>>(lay.metas, d, lay.int_pool, (T*)lay.node_pool,
                                                                                                  lay.stat, stat_stride, mask);
    CB2_CHECK_LAUNCH(ctx, "tree_build_small_kernel");
  } else {
    tree_build_kernel<T><<<(int)pl.metas.size(), kTreeThreads, tree_smem, ctx->stream>>>(lay.metas, d, lay.int_pool, (T*)lay.node_pool,
                                                                                            lay.stat, stat_stride, mask);
    CB2_CHECK_LAUNCH(ctx, "tree_build_kernel");
  }
  CB2_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
  GibbsArgs ga;
  ga.metas = lay.metas; ga.prods = lay.prods; ga.work = lay.work; ga.int_pool = lay.int_pool; ga.node_pool = lay.node_pool;
  ga.k0 = (uint32_t)seed; ga.k1 = (uint32_t)(seed >> 32); ga.circ_mask = mask; ga.level_down = ctx->level_down;
  ga.stat = lay.stat; ga.stat_stride = stat_stride;
  size_t stage_bytes = 0;   // largest per-product record block when every product of the batch qualifies for the SMALL kernels
  if (((d == 2 && mask == 0) || (d == 3 && mask == 4u)) && !getenv("CB2_NO_GIBBS_SMALL")) {
    const int RECL = recl_of(d);
    for (const ProdDev& pd : pl.prods) {
      size_t elems = 0;
      for (int j = 0; j < pd.D; ++j) {
        const DensMeta& m = pl.metas[pd.dens[j]];
        elems += (size_t)(m.rec_off[m.depth] - m.rec_off[0]) + (size_t)(((size_t)m.N * RECL + 31) / 32 * 32);
      }
      stage_bytes = std::max(stage_bytes, elems * sizeof(T));
    }
    if (stage_bytes > 96 * 1024) stage_bytes = 0;
  }
  int rc = launch_gibbs<T>(ctx, d, mask, (int)pl.work.size(), ga, stage_bytes);
  if (rc) return rc;
  CB2_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
  if (lay_out) *lay_out = lay;
  return CB2_OK;
}

void record_product_stages(cb2_ctx* ctx, bool with_bw) {
  float ms = 0;
  if (cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]) == cudaSuccess) ctx->stage_ms["tree"] = ms;
  if (cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]) == cudaSuccess) ctx->stage_ms["gibbs"] = ms;
  if (with_bw && cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]) == cudaSuccess) ctx->stage_ms["bw"] = ms;
  cudaGetLastError();
}

// runs tree + gibbs (+ LOO bandwidth of the outputs) with every data pointer on the device
int product_batch_dev_locked(cb2_ctx* ctx, const cb2_product_desc* descs, int B, const uint8_t* circular, int d, uint64_t seed,
                             const std::vector<std::vector<const double*>>* bw_host_in, DevLayout* lay_out,
                             const InputChunks* chunks = nullptr, const std::function<int()>* on_samples = nullptr) {
  BatchPlan pl;
  uint32_t mask = 0;
  int rc = validate_and_plan(ctx, descs, B, circular, d, pl, &mask);
  if (rc) return rc;
  CB2_CUDA(ctx, cudaSetDevice(ctx->device));
  // bandwidths given on the host are validated there; bandwidths that live on the device are copied into the message table by
  // a kernel (no host round trip, so a device-resident sequence of calls never waits)
  std::vector<std::vector<const double*>> bw_host;
  std::vector<const double*> bw_ptrs;
  if (bw_host_in) bw_host = *bw_host_in;
  else
    for (int b = 0; b < B; ++b)
      for (int j = 0; j < descs[b].D; ++j) bw_ptrs.push_back(descs[b].dens[j].bw);
  // optional LOO bandwidth of the outputs
  std::vector<const double*> bw_mu;
  std::vector<int32_t> bw_N;
  std::vector<int> bw_idx;
  for (int b = 0; b < B; ++b)
    if (descs[b].bw_out) {
      if (descs[b].Nout < 2) return cb2_fail(ctx, CB2_ERR_INVALID, "product %d: bw_out needs Nout >= 2", b);
      bw_mu.push_back(descs[b].pts_out); bw_N.push_back(descs[b].Nout); bw_idx.push_back(b);
    }
  size_t need_prod = 0, need_loo = 0;
  const bool f64 = ctx->precision == CB2_FP64;
  rc = f64 ? product_batch_core<double>(ctx, descs, B, bw_host, nullptr, mask, d, seed, pl, nullptr, 0, &need_prod, nullptr)
           : product_batch_core<float>(ctx, descs, B, bw_host, nullptr, mask, d, seed, pl, nullptr, 0, &need_prod, nullptr);
  if (rc) return rc;
  if (!bw_mu.empty()) {
    rc = cb2_bw_loo_dev_core(ctx, bw_mu.data(), bw_N.data(), (int)bw_mu.size(), d, circular, nullptr, nullptr, 0, &need_loo);
    if (rc) return rc;
  }
  size_t bw_out_bytes = cb2_align(sizeof(double) * d * (bw_mu.size() + 1)) + cb2_align(sizeof(double*) * (bw_mu.size() + 1));
  const size_t ptr_bytes = cb2_align(sizeof(double*) * (bw_ptrs.size() + 1));
  rc = cb2_arena_reserve(ctx, need_prod + need_loo + bw_out_bytes + ptr_bytes + 8192);
  if (rc) return rc;
  const double** bw_ptrs_dev = nullptr;
  if (!bw_ptrs.empty()) {
    bw_ptrs_dev = (const double**)cb2_arena_alloc(ctx, ptr_bytes);
    if (!bw_ptrs_dev) return cb2_fail(ctx, CB2_ERR_NOMEM, "product scratch exhausted");
    CB2_CUDA(ctx, cudaMemcpyAsync(bw_ptrs_dev, bw_ptrs.data(), sizeof(double*) * bw_ptrs.size(), cudaMemcpyHostToDevice, ctx->stream));
  }
  char* base = (char*)cb2_arena_alloc(ctx, need_prod);
  if (!base) return cb2_fail(ctx, CB2_ERR_NOMEM, "product scratch exhausted");
  rc = f64 ? product_batch_core<double>(ctx, descs, B, bw_host, bw_ptrs_dev, mask, d, seed, pl, base, need_prod, nullptr, lay_out, chunks)
           : product_batch_core<float>(ctx, descs, B, bw_host, bw_ptrs_dev, mask, d, seed, pl, base, need_prod, nullptr, lay_out, chunks);
  if (rc) return rc;
  if (on_samples) {  // points and labels are final: the host-pointer path sends them home while the bandwidths are searched
    rc = (*on_samples)();
    if (rc) return rc;
  }
  if (!bw_mu.empty()) {
    double* bw_dev = (double*)cb2_arena_alloc(ctx, bw_out_bytes);
    void* scratch = cb2_arena_alloc(ctx, need_loo);
    if (!bw_dev || !scratch) return cb2_fail(ctx, CB2_ERR_NOMEM, "bandwidth scratch exhausted");
    rc = cb2_bw_loo_dev_core(ctx, bw_mu.data(), bw_N.data(), (int)bw_mu.size(), d, circular, bw_dev, scratch, need_loo, nullptr);
    if (rc) return rc;
    // hand the fresh bandwidths to the callers' device buffers with one scatter kernel
    std::vector<double*> outp(bw_idx.size());
    for (size_t i = 0; i < bw_idx.size(); ++i) outp[i] = descs[bw_idx[i]].bw_out;
    double** outp_dev = (double**)((char*)bw_dev + cb2_align(sizeof(double) * d * (bw_mu.size() + 1)));
    CB2_CUDA(ctx, cudaMemcpyAsync(outp_dev, outp.data(), sizeof(double*) * outp.size(), cudaMemcpyHostToDevice, ctx->stream));
    scatter_small_vectors_kernel<<<(int)((outp.size() + 127) / 128), 128, 0, ctx->stream>>>(outp_dev, (int)outp.size(), d, bw_dev);
    CB2_CHECK_LAUNCH(ctx, "scatter_small_vectors_kernel");
    CB2_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
  }
  ctx->stage_ms.erase("tree"); ctx->stage_ms.erase("gibbs"); ctx->stage_ms.erase("bw");
  ctx->stage_ms["__pending_bw"] = bw_mu.empty() ? 0.0 : 1.0;
  return CB2_OK;
}

}  // namespace

// stage timings are resolved lazily (needs the events to have completed)
void cb2_resolve_product_stages(cb2_ctx* ctx) {
  auto it = ctx->stage_ms.find("__pending_bw");
  if (it == ctx->stage_ms.end()) return;
  bool with_bw = it->second > 0.5;
  ctx->stage_ms.erase(it);
  cudaStreamSynchronize(ctx->stream);
  record_product_stages(ctx, with_bw);
}

extern "C" int cb2_product_batch_dev(cb2_ctx* ctx, const cb2_product_desc* descs, int B, const uint8_t* circular, int d,
                                     uint64_t seed) {
  if (!ctx) return CB2_ERR_INVALID;
  cb2_lock g(ctx->mu);
  return product_batch_dev_locked(ctx, descs, B, circular, d, seed, nullptr, nullptr);
}

// address a kernel can use for a host array, or null: pinned + mapped memory (cudaHostAlloc / cudaHostRegister), 8-byte aligned
static void* mapped_host_ptr(const void* p) {
  if (((uintptr_t)p & 7) != 0) return nullptr;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

// one table upload + one launch (per 65535 segments) instead of one copy call per array
static int launch_copy_segments(cb2_ctx* ctx, const std::vector<CopySeg>& segs, CopySeg* table_dev, cudaStream_t st) {
  if (segs.empty()) return CB2_OK;
  CB2_CUDA(ctx, cudaMemcpyAsync(table_dev, segs.data(), sizeof(CopySeg) * segs.size(), cudaMemcpyHostToDevice, st));
  unsigned long long maxb = 0;
  for (const CopySeg& sg : segs) maxb = sg.bytes > maxb ? sg.bytes : maxb;
  const int gx = (int)std::min<unsigned long long>(32, std::max<unsigned long long>(1, (maxb / 8 + 1023) / 1024));
  for (size_t s0 = 0; s0 < segs.size(); s0 += 65535) {
    const int gy = (int)std::min<size_t>(65535, segs.size() - s0);
    copy_segments_kernel<<<dim3(gx, gy), 256, 0, st>>>(table_dev + s0);
    CB2_CHECK_LAUNCH(ctx, "copy_segments_kernel");
  }
  return CB2_OK;
}

// Host-pointer flavour: stage everything into one device block, run, copy results back.
static int product_batch_host(cb2_ctx* ctx, const cb2_product_desc* descs, int B, const uint8_t* circular, int d, uint64_t seed,
                              cb2_pending* pend) {
  CB2_REQUIRE(ctx, descs && B > 0, "null descriptors or empty batch");
  if (d < 1 || d > CB2_MAX_PRODUCT_DIM) return cb2_fail(ctx, CB2_ERR_UNSUPPORTED, "product dimension %d outside 1..%d", d, CB2_MAX_PRODUCT_DIM);
  CB2_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t total = 0, n_arrays = 0;
  for (int b = 0; b < B; ++b) {
    const cb2_product_desc& ds = descs[b];
    n_arrays += 2 * (size_t)ds.D + 3;
    CB2_REQUIRE(ctx, ds.dens && ds.pts_out, "product %d: null dens / pts_out", b);
    if (ds.D < 1 || ds.D > CB2_MAX_DENS) return cb2_fail(ctx, CB2_ERR_UNSUPPORTED, "product %d: D=%d outside 1..%d", b, ds.D, CB2_MAX_DENS);
    CB2_REQUIRE(ctx, ds.Nout > 0, "product %d: Nout <= 0", b);
    for (int j = 0; j < ds.D; ++j) {
      CB2_REQUIRE(ctx, ds.dens[j].pts && ds.dens[j].bw && ds.dens[j].N > 0, "product %d message %d: null pts/bw or N <= 0", b, j);
      total += cb2_align(sizeof(double) * (size_t)ds.dens[j].N * d);
      if (ds.dens[j].w) total += cb2_align(sizeof(double) * (size_t)ds.dens[j].N);
    }
    total += cb2_align(sizeof(double) * (size_t)ds.Nout * d);
    if (ds.labels_out) total += cb2_align(sizeof(int32_t) * (size_t)ds.Nout * ds.D);
    if (ds.bw_out) total += cb2_align(sizeof(double) * d);
  }
  total += 2 * cb2_align(sizeof(CopySeg) * n_arrays);  // segment tables of the single-launch copies
  // synchronous calls stage through the context's I/O block; a ticketed (asynchronous) batch owns its block until cb2_wait
  const bool own = pend != nullptr;
  void* blk = nullptr;
  cudaError_t e = cudaSuccess;
  if (own) {
    e = cudaMalloc(&blk, total + 256);
    if (e != cudaSuccess) { cudaGetLastError(); return cb2_fail(ctx, CB2_ERR_NOMEM, "device block of %zu bytes: %s", total, cudaGetErrorString(e)); }
  } else {
    blk = cb2_io_block(ctx, total + 256);
    if (!blk) return CB2_ERR_NOMEM;
  }
  char* p = (char*)blk;
  size_t off = 0;
  auto take = [&](size_t n) { char* q = p + off; off += cb2_align(n); return q; };
  std::vector<cb2_product_desc> dd(descs, descs + B);
  std::vector<std::vector<cb2_density>> ddens(B);
  std::vector<std::vector<const double*>> bw_host(B);
  int rc = CB2_OK;
  // Small synchronous batches (the solver's usual N = 100..1000 particles) are packed into pinned staging: ONE H2D copy of
  // all inputs and ONE D2H copy of all outputs instead of one copy per array.  Large batches copy array by array straight
  // from / to the caller's memory (an extra host memcpy of tens of MB would cost more than the copy calls).
  const bool packed = !own && total <= ((size_t)4 << 20) && cb2_pinned_reserve(ctx, total + 256) == CB2_OK;
  char* stage = packed ? ctx->pinned : nullptr;
  // Batches of many arrays: every array whose host memory is pinned and mapped is moved by ONE kernel per direction that reads /
  // writes the host arrays in place (copy_segments_kernel, 50 GB/s on this box against 17 GB/s for one cudaMemcpyAsync per
  // 196 KB array); pageable arrays of the same batch keep their individual copies.
  const bool direct = !own && !packed && n_arrays >= 64 && !getenv("CB2_NO_DIRECT");
  std::vector<CopySeg> seg_in, seg_out;
  std::vector<int> seg_msg;  // message index of every entry of seg_in
  int cur_msg = 0;
  CopySeg* tab_in = (CopySeg*)take(sizeof(CopySeg) * n_arrays);
  CopySeg* tab_out = (CopySeg*)take(sizeof(CopySeg) * n_arrays);
  // many pageable inputs (first array not mapped): pack them into pinned staging and upload the input range with ONE copy
  const bool stage_in = !own && !packed && n_arrays >= 16 && total <= ((size_t)256 << 20) && !mapped_host_ptr(descs[0].dens[0].pts) &&
                        cb2_pinned_reserve(ctx, total + 256) == CB2_OK;
  auto h2d = [&](void* dst, const void* src, size_t bytes) -> cudaError_t {
    if (stage_in) { memcpy(ctx->pinned + ((char*)dst - p), src, bytes); return cudaSuccess; }
    void* mp = direct ? mapped_host_ptr(src) : nullptr;
    if (mp) { seg_in.push_back({(const double*)mp, (double*)dst, bytes}); seg_msg.push_back(cur_msg); return cudaSuccess; }
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
  };
  // pageable outputs of a synchronous call: ONE copy of the contiguous output range into pinned staging, then host memcpy
  // (a cudaMemcpyAsync into pageable memory waits for the stream every time: 512 of them cost 6 ms on a 256-product batch)
  struct StagedOut { void* dst; size_t off, bytes; };
  std::vector<StagedOut> staged;
  bool stage_out = false;
  auto d2h = [&](void* dst, const void* src, size_t bytes) -> cudaError_t {
    void* mp = direct ? mapped_host_ptr(dst) : nullptr;
    if (mp) { seg_out.push_back({(const double*)src, (double*)mp, bytes}); return cudaSuccess; }
    if (stage_out) { staged.push_back({dst, (size_t)((const char*)src - p), bytes}); return cudaSuccess; }
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream);
  };
  cudaEventRecord(ctx->ev[4], ctx->stream);
  // device block layout: every input of every product first, then every output (so that each side is one contiguous range)
  for (int b = 0; b < B && rc == CB2_OK; ++b) {
    ddens[b].assign(descs[b].dens, descs[b].dens + descs[b].D);
    for (int j = 0; j < descs[b].D; ++j) {
      const cb2_density& dn = descs[b].dens[j];
      const size_t pb = sizeof(double) * (size_t)dn.N * d, wb = sizeof(double) * (size_t)dn.N;
      double* dp = (double*)take(pb);
      if (packed) memcpy(stage + ((char*)dp - p), dn.pts, pb);
      else e = h2d(dp, dn.pts, pb);
      ddens[b][j].pts = dp;
      if (dn.w && e == cudaSuccess) {
        double* wp = (double*)take(wb);
        if (packed) memcpy(stage + ((char*)wp - p), dn.w, wb);
        else e = h2d(wp, dn.w, wb);
        ddens[b][j].w = wp;
      }
      if (e != cudaSuccess) { rc = cb2_fail(ctx, CB2_ERR_CUDA, "H2D copy: %s", cudaGetErrorString(e)); break; }
      ++cur_msg;
      bw_host[b].push_back(dn.bw);
    }
    dd[b].dens = ddens[b].data();
  }
  // pinned + mapped inputs: the gather runs chunk by chunk on the copy stream (small persistent grid) and the trees of a chunk are
  // built as soon as it has landed, so the upload hides behind the tree build instead of preceding it
  InputChunks chunks;
  const bool overlap = direct && !seg_in.empty() && ctx->copy_stream && !getenv("CB2_NO_COPY_OVERLAP");
  if (rc == CB2_OK && overlap) {
    cudaError_t ce = cudaEventRecord(ctx->copy_ev[7], ctx->stream);  // earlier work on the main stream may still read the I/O block
    if (ce == cudaSuccess) ce = cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ev[7], 0);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(tab_in, seg_in.data(), sizeof(CopySeg) * seg_in.size(), cudaMemcpyHostToDevice, ctx->copy_stream);
    const int NC = 4;
    size_t s0 = 0;
    int msgs_total = 0;
    for (int b = 0; b < B; ++b) msgs_total += descs[b].D;
    for (int c = 0; c < NC && ce == cudaSuccess; ++c) {
      const int m_end = (int)((long long)msgs_total * (c + 1) / NC);  // chunk boundaries on message boundaries
      size_t s1 = s0;
      while (s1 < seg_in.size() && seg_msg[s1] < m_end) ++s1;
      if (s1 > s0) {
        copy_segments_stream_kernel<<<kCopyStreamCtas, 256, 0, ctx->copy_stream>>>(tab_in + s0, (int)(s1 - s0));
        ctx->launches++;
        ce = cudaGetLastError();
      }
      if (ce == cudaSuccess) ce = cudaEventRecord(ctx->copy_ev[c], ctx->copy_stream);
      chunks.msg_end.push_back(m_end);
      chunks.ev.push_back(ctx->copy_ev[c]);
      s0 = s1;
    }
    if (ce != cudaSuccess) rc = cb2_fail(ctx, CB2_ERR_CUDA, "chunked H2D copy: %s", cudaGetErrorString(ce));
  } else if (rc == CB2_OK) {
    rc = launch_copy_segments(ctx, seg_in, tab_in, ctx->stream);
  }
  const size_t in_bytes = off;
  if (rc == CB2_OK && stage_in && in_bytes) {
    e = cudaMemcpyAsync(p, ctx->pinned, in_bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) rc = cb2_fail(ctx, CB2_ERR_CUDA, "H2D copy: %s", cudaGetErrorString(e));
  }
  if (rc == CB2_OK && packed && in_bytes) {
    e = cudaMemcpyAsync(p, stage, in_bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) rc = cb2_fail(ctx, CB2_ERR_CUDA, "H2D copy: %s", cudaGetErrorString(e));
  }
  for (int b = 0; b < B; ++b) {
    dd[b].pts_out = (double*)take(sizeof(double) * (size_t)descs[b].Nout * d);
    dd[b].labels_out = descs[b].labels_out ? (int32_t*)take(sizeof(int32_t) * (size_t)descs[b].Nout * descs[b].D) : nullptr;
    dd[b].bw_out = descs[b].bw_out ? (double*)take(sizeof(double) * d) : nullptr;
  }
  // results that live in pinned + mapped arrays: the sample points and labels go home on the copy stream as soon as the Gibbs
  // kernel has finished, under the bandwidth search of the same batch; only the bandwidths are copied afterwards
  std::vector<char> sent_early(B, 0);
  size_t n_early = 0;  // entries of tab_out taken by the early copy (the later one must not overwrite a table still being read)
  std::function<int()> send_samples = [&]() -> int {
    std::vector<CopySeg> segs;
    for (int b = 0; b < B; ++b) {
      void* mp = mapped_host_ptr(descs[b].pts_out);
      void* ml = descs[b].labels_out ? mapped_host_ptr(descs[b].labels_out) : nullptr;
      if (!mp || (descs[b].labels_out && !ml)) continue;
      segs.push_back({(const double*)dd[b].pts_out, (double*)mp, sizeof(double) * (size_t)descs[b].Nout * d});
      if (ml) segs.push_back({(const double*)dd[b].labels_out, (double*)ml, sizeof(int32_t) * (size_t)descs[b].Nout * descs[b].D});
      sent_early[b] = 1;
    }
    if (segs.empty()) return CB2_OK;
    n_early = segs.size();
    CB2_CUDA(ctx, cudaEventRecord(ctx->copy_ev[6], ctx->stream));
    CB2_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ev[6], 0));
    CB2_CUDA(ctx, cudaMemcpyAsync(tab_out, segs.data(), sizeof(CopySeg) * segs.size(), cudaMemcpyHostToDevice, ctx->copy_stream));
    copy_segments_stream_kernel<<<kCopyStreamCtas, 256, 0, ctx->copy_stream>>>(tab_out, (int)segs.size());
    CB2_CHECK_LAUNCH(ctx, "copy_segments_stream_kernel");
    return CB2_OK;
  };
  const bool early_out = overlap && !own;
  if (rc == CB2_OK) {
    cudaEventRecord(ctx->ev[5], ctx->stream);
    rc = product_batch_dev_locked(ctx, dd.data(), B, circular, d, seed, &bw_host, nullptr, overlap ? &chunks : nullptr,
                                  early_out ? &send_samples : nullptr);
  }
  if (overlap) cudaStreamSynchronize(ctx->copy_stream);  // (also on failure: nothing of this call may still be in flight)
  if (rc != CB2_OK) { cudaStreamSynchronize(ctx->stream); if (own) cudaFree(blk); return rc; }
  cudaEventRecord(ctx->ev[6], ctx->stream);
  if (packed) {
    e = cudaMemcpyAsync(stage + in_bytes, p + in_bytes, off - in_bytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return cb2_fail(ctx, CB2_ERR_CUDA, "D2H copy: %s", cudaGetErrorString(e));
    for (int b = 0; b < B; ++b) {
      memcpy(descs[b].pts_out, stage + ((char*)dd[b].pts_out - p), sizeof(double) * (size_t)descs[b].Nout * d);
      if (descs[b].labels_out) memcpy(descs[b].labels_out, stage + ((char*)dd[b].labels_out - p), sizeof(int32_t) * (size_t)descs[b].Nout * descs[b].D);
      if (descs[b].bw_out) memcpy(descs[b].bw_out, stage + ((char*)dd[b].bw_out - p), sizeof(double) * d);
    }
  } else {
    const size_t out_bytes = off - in_bytes;
    // (by now the stream has consumed the input staging: the products ran after it; a growing reserve synchronises anyway)
    stage_out = !own && B >= 4 && out_bytes <= ((size_t)1 << 30) && cb2_pinned_reserve(ctx, out_bytes + 256) == CB2_OK;
    for (int b = 0; b < B && e == cudaSuccess; ++b) {
      if (!sent_early[b]) {
        e = d2h(descs[b].pts_out, dd[b].pts_out, sizeof(double) * (size_t)descs[b].Nout * d);
        if (e == cudaSuccess && descs[b].labels_out)
          e = d2h(descs[b].labels_out, dd[b].labels_out, sizeof(int32_t) * (size_t)descs[b].Nout * descs[b].D);
      }
      if (e == cudaSuccess && descs[b].bw_out) e = d2h(descs[b].bw_out, dd[b].bw_out, sizeof(double) * d);
    }
    if (e == cudaSuccess && launch_copy_segments(ctx, seg_out, tab_out + n_early, ctx->stream) != CB2_OK) e = cudaErrorUnknown;
    if (e == cudaSuccess && !staged.empty()) {
      e = cudaMemcpyAsync(ctx->pinned, p + in_bytes, out_bytes, cudaMemcpyDeviceToHost, ctx->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
      if (e == cudaSuccess)
        for (const StagedOut& so : staged) memcpy(so.dst, ctx->pinned + (so.off - in_bytes), so.bytes);
    }
    if (e != cudaSuccess) { cudaStreamSynchronize(ctx->stream); if (own) cudaFree(blk); return cb2_fail(ctx, CB2_ERR_CUDA, "D2H copy: %s", cudaGetErrorString(e)); }
  }
  cudaEventRecord(ctx->ev[7], ctx->stream);
  if (pend) {  // asynchronous: hand the block to the ticket
    pend->dev_block = blk;
    e = cudaEventCreateWithFlags(&pend->done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(pend->done, ctx->stream);
    if (e != cudaSuccess) { cudaStreamSynchronize(ctx->stream); if (own) cudaFree(blk); pend->dev_block = nullptr; return cb2_fail(ctx, CB2_ERR_CUDA, "ticket event: %s", cudaGetErrorString(e)); }
    return CB2_OK;
  }
  e = cudaStreamSynchronize(ctx->stream);
  if (own) cudaFree(blk);
  if (e != cudaSuccess) return cb2_fail(ctx, CB2_ERR_CUDA, "product batch failed: %s", cudaGetErrorString(e));
  float ms = 0;
  if (cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]) == cudaSuccess) ctx->stage_ms["h2d"] = ms;
  if (cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]) == cudaSuccess) ctx->stage_ms["d2h"] = ms;
  cudaGetLastError();
  return CB2_OK;
}

extern "C" int cb2_product_batch(cb2_ctx* ctx, const cb2_product_desc* descs, int B, const uint8_t* circular, int d, uint64_t seed) {
  if (!ctx) return CB2_ERR_INVALID;
  cb2_lock g(ctx->mu);
  return product_batch_host(ctx, descs, B, circular, d, seed, nullptr);
}

extern "C" int cb2_product(cb2_ctx* ctx, const cb2_density* dens, int D, const uint8_t* circular, int d, int Nout, int niter,
                           uint64_t seed, uint32_t stream, int sample_offset, double* pts_out, int32_t* labels_out) {
  cb2_product_desc ds;
  memset(&ds, 0, sizeof(ds));
  ds.dens = dens; ds.D = D; ds.Nout = Nout; ds.niter = niter; ds.stream = stream; ds.sample_offset = sample_offset;
  ds.pts_out = pts_out; ds.labels_out = labels_out; ds.bw_out = nullptr;
  return cb2_product_batch(ctx, &ds, 1, circular, d, seed);
}

extern "C" int cb2_product_batch_submit(cb2_ctx* ctx, const cb2_product_desc* descs, int B, const uint8_t* circular, int d,
                                        uint64_t seed, cb2_ticket* ticket) {
  if (!ctx) return CB2_ERR_INVALID;
  cb2_lock g(ctx->mu);
  CB2_REQUIRE(ctx, ticket, "null ticket");
  cb2_pending pend;
  int rc = product_batch_host(ctx, descs, B, circular, d, seed, &pend);
  if (rc) return rc;
  *ticket = ctx->next_ticket++;
  ctx->pending[*ticket] = pend;
  return CB2_OK;
}

extern "C" int cb2_wait(cb2_ctx* ctx, cb2_ticket ticket) {
  if (!ctx) return CB2_ERR_INVALID;
  cb2_pending pend;
  {
    cb2_lock g(ctx->mu);
    auto it = ctx->pending.find(ticket);
    if (it == ctx->pending.end()) return cb2_fail(ctx, CB2_ERR_INVALID, "unknown ticket %lld", (long long)ticket);
    pend = it->second;
    ctx->pending.erase(it);
  }
  cudaError_t e = cudaEventSynchronize(pend.done);
  cudaEventDestroy(pend.done);
  if (pend.dev_block) cudaFree(pend.dev_block);
  if (e != cudaSuccess) {
    cb2_lock g(ctx->mu);
    return cb2_fail(ctx, CB2_ERR_CUDA, "async product batch failed: %s", cudaGetErrorString(e));
  }
  return CB2_OK;
}

extern "C" int cb2_tree_debug(cb2_ctx* ctx, const cb2_density* den, int d, int level, int32_t* depth_out, int32_t* count_out,
                              int32_t* perm, double* mean, double* var, double* wt) {
  if (!ctx) return CB2_ERR_INVALID;
  cb2_lock g(ctx->mu);
  CB2_REQUIRE(ctx, den && den->pts && den->bw && den->N > 0, "null density");
  if (d < 1 || d > CB2_MAX_DIM) return cb2_fail(ctx, CB2_ERR_UNSUPPORTED, "dimension %d", d);
  CB2_CUDA(ctx, cudaSetDevice(ctx->device));
  const int N = den->N, depth = ceil_log2(N);
  if (depth_out) *depth_out = depth;
  CB2_REQUIRE(ctx, level >= 0 && level <= depth, "level %d outside 0..%d", level, depth);
  const int cnt = level_count(N, depth, level);
  if (count_out) *count_out = cnt;
  // run a one-message, one-sample product to build the tree, then read the statistics scratch back
  std::vector<double> dummy_out(d);
  void* blk = nullptr;
  size_t bytes = cb2_align(sizeof(double) * (size_t)N * d) + cb2_align(sizeof(double) * N) + cb2_align(sizeof(double) * d) +
                 cb2_align(sizeof(double) * (size_t)cnt * (2 * d + 1)) + cb2_align(sizeof(int) * N) + 1024;
  cudaError_t e = cudaMalloc(&blk, bytes);
  if (e != cudaSuccess) { cudaGetLastError(); return cb2_fail(ctx, CB2_ERR_NOMEM, "tree debug block"); }
  char* p = (char*)blk;
  size_t off = 0;
  auto take = [&](size_t n) { char* q = p + off; off += cb2_align(n); return q; };
  double* pts_dev = (double*)take(sizeof(double) * (size_t)N * d);
  double* w_dev = (double*)take(sizeof(double) * N);
  double* out_dev = (double*)take(sizeof(double) * d);
  double* mean_dev = (double*)take(sizeof(double) * (size_t)cnt * (2 * d + 1));
  double* var_dev = mean_dev + (size_t)cnt * d;
  double* wt_dev = var_dev + (size_t)cnt * d;
  cudaMemcpyAsync(pts_dev, den->pts, sizeof(double) * (size_t)N * d, cudaMemcpyHostToDevice, ctx->stream);
  if (den->w) cudaMemcpyAsync(w_dev, den->w, sizeof(double) * N, cudaMemcpyHostToDevice, ctx->stream);
  cb2_density dd = {N, pts_dev, den->bw, den->w ? w_dev : nullptr};
  cb2_product_desc ds;
  memset(&ds, 0, sizeof(ds));
  ds.dens = &dd; ds.D = 1; ds.Nout = 1; ds.niter = 0; ds.pts_out = out_dev;
  std::vector<std::vector<const double*>> bw_host(1);
  bw_host[0].push_back(den->bw);
  DevLayout lay;
  int rc = product_batch_dev_locked(ctx, &ds, 1, nullptr, d, 0, &bw_host, &lay);
  if (rc == CB2_OK) {
    int node0 = level < depth ? ((1 << level) - 1) : ((1 << depth) - 1);
    tree_debug_copy_kernel<<<(cnt + 127) / 128, 128, 0, ctx->stream>>>(lay.stat, d, node0, cnt, mean_dev, var_dev, wt_dev);
    ctx->launches++;
    if (mean) cudaMemcpyAsync(mean, mean_dev, sizeof(double) * (size_t)cnt * d, cudaMemcpyDeviceToHost, ctx->stream);
    if (var) cudaMemcpyAsync(var, var_dev, sizeof(double) * (size_t)cnt * d, cudaMemcpyDeviceToHost, ctx->stream);
    if (wt) cudaMemcpyAsync(wt, wt_dev, sizeof(double) * cnt, cudaMemcpyDeviceToHost, ctx->stream);
    if (perm) cudaMemcpyAsync(perm, lay.int_pool, sizeof(int) * N, cudaMemcpyDeviceToHost, ctx->stream);
    e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = cb2_fail(ctx, CB2_ERR_CUDA, "tree debug: %s", cudaGetErrorString(e));
  } else {
    cudaStreamSynchronize(ctx->stream);
  }
  cudaFree(blk);
  return rc;
}

// Diagnostic counters of the Gibbs kernel (tests only).  which = 0: categorical draws whose pass-2 re-sum of the selected chunk
// ended short of the target and therefore took the chunk's last candidate with mass (gibbs.cuh chunk_pick) since the last reset.
extern "C" int cb2_debug_counter(cb2_ctx* ctx, int which, int reset, uint64_t* value) {
  if (!ctx) return CB2_ERR_INVALID;
  cb2_lock g(ctx->mu);
  CB2_REQUIRE(ctx, which == 0 && value, "unknown counter %d or null value", which);
  CB2_CUDA(ctx, cudaSetDevice(ctx->device));
  CB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  unsigned long long v = 0;
  CB2_CUDA(ctx, cudaMemcpyFromSymbol(&v, g_pick_fallthrough, sizeof(v)));
  *value = v;
  if (reset) {
    const unsigned long long z = 0;
    CB2_CUDA(ctx, cudaMemcpyToSymbol(g_pick_fallthrough, &z, sizeof(z)));
  }
  return CB2_OK;
}

// levelDown rule of the multiscale Gibbs sampler (SURVEY App. A.3 leaves it open): 1 = a chain descends to a child of its node
// drawn in proportion to the children's weights (default), 0 = always the first child.
extern "C" int cb2_set_level_down(cb2_ctx* ctx, int mode) {
  if (!ctx) return CB2_ERR_INVALID;
  cb2_lock g(ctx->mu);
  ctx->level_down = mode ? 1 : 0;
  return CB2_OK;
}
