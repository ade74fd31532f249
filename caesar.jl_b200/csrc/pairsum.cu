// pairsum.cu -- dense Gaussian-kernel pair sums on sm_100a: mmd (K1), mmd under batched rigid
// transforms with analytic gradient (K2), KDE evaluate (K3), KDE sample (K7).
//
// Device layout: every point set is converted once from the caller's Float64 d x N matrix into padded
// records of REC = 4 (d <= 3) or 8 (d <= 6) scalars: [x_0 .. x_{d-1}, 0 .., c] -- one 16-byte (FP32) vector
// load per point for the usual d = 2, 3.  Coordinates are centred and pre-scaled so that the kernel value is
// exp2(c_j - |p_i - q_j|^2): one FADD + FFMA chain + one MUFU.EX2 per pair, no multiply by the bandwidth.
//
// Work decomposition: a tile = (sub-problem, block of 256*RI rows, block of CB columns).  A CTA owns one tile;
// each thread keeps RI row points in registers and streams the column block through shared memory in
// sub-tiles of CT records read as warp-broadcast 128-bit loads.  Every tile writes its own partial (no
// atomics); a second kernel reduces the partials in a fixed order in FP64, so results are deterministic.
#include "common.cuh"
#include <algorithm>

namespace {

constexpr int kThreads = 256;
constexpr int kRI = 4;    // row points per thread
constexpr int kCT = 512;  // column records per shared-memory sub-tile
constexpr int kCB = 4096; // columns per tile

template <int DIM> struct RecOf { static constexpr int value = DIM <= 3 ? 4 : 8; };

template <typename T, int REC> struct alignas(16) Rec { T v[REC]; };

struct Tile { int sub, rb, cb, c0, c1, weight; };  // columns [c0,c1); weight 2 = strictly upper block of a symmetric pair sum

template <typename T> struct PairSub {
  const T* rows;  // records
  const T* cols;
  int row0, row1, ncols;
  int xf;  // index into the transform table (cols are transformed when staged), -1 = none
};

// ---------------------------------------------------------------------------------------------------
// Float64 d x N (point-contiguous)  ->  records.  x' = (x - center) * scale_k ; last slot = offs[i] or 0.
struct ConvertArgs {
  double scale[8];
  double center[8];
  int use_center_ptr;  // take the centre from the first point of center_src (device) instead
  int wrap[8];         // circular coordinate: wrapped to [-pi, pi] and not centred
};

template <typename T, int REC>
__global__ void convert_records_kernel(const double* __restrict__ src, int n, int d, ConvertArgs a,
                                       const double* __restrict__ center_src, const double* __restrict__ w,
                                       double wsum, T* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Rec<T, REC> r;
#pragma unroll
  for (int k = 0; k < REC; ++k) r.v[k] = T(0);
  for (int k = 0; k < d; ++k) {
    double c = a.use_center_ptr ? center_src[k] : a.center[k];
    double x = src[(size_t)i * d + k];
    r.v[k] = (T)((a.wrap[k] ? wrap_pi(x) : x - c) * a.scale[k]);
  }
  if (w) r.v[REC - 1] = (T)log2(w[i] / wsum);
  reinterpret_cast<Rec<T, REC>*>(dst)[i] = r;
}

// ---------------------------------------------------------------------------------------------------
// MODE 0: total sum per tile (mmd).  MODE 1: total + gradient moments per tile (rigid-transform batch).
// MODE 2: per-row partial sums with per-column log2 offsets (KDE evaluate).
// MODE 3: total sum with a manifold distance: per-coordinate periods in xforms[0..DIM) (+inf = Euclidean coordinate);
//         the wrapped difference is min(|d|, period - |d|) on coordinates already wrapped to [-pi, pi] * scale.
template <typename T, int DIM, int MODE>
__global__ void __launch_bounds__(kThreads)
pair_tiles_kernel(const PairSub<T>* __restrict__ subs, const Tile* __restrict__ tiles,
                  const T* __restrict__ xforms, double* __restrict__ tile_out, T* __restrict__ row_out,
                  int row_out_stride) {
  constexpr int REC = RecOf<DIM>::value;
  constexpr bool GRAD = MODE == 1;
  constexpr bool OFFS = MODE == 2;
  // FP32 plain pair sums (mmd): rows are processed as two packed pairs per thread and every column is staged as negated,
  // duplicated coordinates (-q_k, -q_k), so a pair of rows costs DIM FADD2 + FMUL2 + (DIM - 1) FFMA2 instead of twice
  // DIM FADD + FMUL + (DIM - 1) FFMA: 5 instead of 8.5 issue slots per pair at DIM = 3, which leaves the XU pipe as the only limit
  constexpr bool PACKED = (MODE == 0 || MODE == 2) && sizeof(T) == 4 && kRI == 4;
  constexpr int REC2 = (DIM + 2) / 2 * 2;  // f32x2 per staged column (+1: the evaluate offset), padded to whole 16-byte loads
  __shared__ Rec<T, REC> sq[PACKED ? 1 : kCT];
  __shared__ __align__(16) f32x2 sq2[PACKED ? kCT * REC2 : 1];
  __shared__ double sred[kThreads / 32][8];

  const Tile tl = tiles[blockIdx.x];
  const PairSub<T> sb = subs[tl.sub];
  const int tid = threadIdx.x;
  const Rec<T, REC>* rows = reinterpret_cast<const Rec<T, REC>*>(sb.rows);
  const Rec<T, REC>* cols = reinterpret_cast<const Rec<T, REC>*>(sb.cols);

  T p[kRI][DIM];
  T acc[kRI];
  T g[kRI][GRAD ? DIM : 1];
  bool valid[kRI];
  const int rbase = sb.row0 + tl.rb * (kThreads * kRI);
#pragma unroll
  for (int r = 0; r < kRI; ++r) {
    int row = rbase + r * kThreads + tid;
    valid[r] = row < sb.row1;
    Rec<T, REC> rr = rows[valid[r] ? row : sb.row0];
#pragma unroll
    for (int k = 0; k < DIM; ++k) p[r][k] = rr.v[k];
    acc[r] = T(0);
#pragma unroll
    for (int k = 0; k < (GRAD ? DIM : 1); ++k) g[r][k] = T(0);
  }

  T per[DIM];
  if (MODE == 3) {
#pragma unroll
    for (int k = 0; k < DIM; ++k) per[k] = xforms[k];
  }
  T R[9], t[3];
  if (MODE != 3 && (GRAD || sb.xf >= 0)) {
    if (sb.xf >= 0) {
#pragma unroll
      for (int k = 0; k < 9; ++k) R[k] = xforms[sb.xf * 12 + k];
#pragma unroll
      for (int k = 0; k < 3; ++k) t[k] = xforms[sb.xf * 12 + 9 + k];
    }
  }

  const int c0 = tl.c0, c1 = tl.c1;
  for (int cs = c0; cs < c1; cs += kCT) {
    __syncthreads();
    for (int j = tid; j < kCT; j += kThreads) {
      int col = cs + j;
      Rec<T, REC> q;
      if (col < c1) {
        q = cols[col];
        if (MODE != 3 && DIM <= 3 && sb.xf >= 0) {
          T x = q.v[0], y = q.v[1], z = DIM == 3 ? q.v[2] : T(0);
          if (DIM == 2) {
            q.v[0] = R[0] * x + R[1] * y + t[0];
            q.v[1] = R[3] * x + R[4] * y + t[1];
          } else {
            q.v[0] = R[0] * x + R[1] * y + R[2] * z + t[0];
            q.v[1] = R[3] * x + R[4] * y + R[5] * z + t[1];
            q.v[2] = R[6] * x + R[7] * y + R[8] * z + t[2];
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < REC; ++k) q.v[k] = T(0);
        q.v[0] = T(1e18);  // padding column: exp2(-1e36) == 0
      }
      if constexpr (PACKED) {
#pragma unroll
        for (int k = 0; k < REC2; ++k)
          sq2[j * REC2 + k] = k < DIM ? pack2(-(float)q.v[k], -(float)q.v[k]) : (k == DIM && OFFS ? pack2((float)q.v[REC - 1], 0.0f) : 0ull);
      } else {
        sq[j] = q;
      }
    }
    __syncthreads();
    if constexpr (PACKED) {
      f32x2 pp[2][DIM];
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int k = 0; k < DIM; ++k) pp[h][k] = pack2((float)p[2 * h][k], (float)p[2 * h + 1][k]);
#pragma unroll 4
      for (int j = 0; j < kCT; ++j) {
        f32x2 nq[REC2];
#pragma unroll
        for (int k = 0; k < REC2; k += 2)  // warp-broadcast LDS.128
          asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(nq[k]), "=l"(nq[k + 1]) : "r"((uint32_t)__cvta_generic_to_shared(&sq2[j * REC2 + k])));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const f32x2 d0 = add2(pp[h][0], nq[0]);
          f32x2 r2 = mul2(d0, d0);
#pragma unroll
          for (int k = 1; k < DIM; ++k) {
            const f32x2 dk = add2(pp[h][k], nq[k]);
            r2 = fma2(dk, dk, r2);
          }
          float ra, rb;
          unpack2(r2, ra, rb);
          if (OFFS) {  // evaluate: log2 of the column's weight and normalisation rides in the spare slot
            float off, unused;
            unpack2(nq[DIM], off, unused);
            acc[2 * h] += fast_exp2(off - ra);
            acc[2 * h + 1] += fast_exp2(off - rb);
          } else {
            acc[2 * h] += fast_exp2(-ra);
            acc[2 * h + 1] += fast_exp2(-rb);
          }
        }
      }
    } else {
#pragma unroll 4
    for (int j = 0; j < kCT; ++j) {
      const Rec<T, REC> q = sq[j];  // warp-broadcast LDS.128
#pragma unroll
      for (int r = 0; r < kRI; ++r) {
        T d0 = p[r][0] - q.v[0];
        if (MODE == 3) { T ad = fabs(d0); d0 = fmin(ad, per[0] - ad); }
        T r2 = d0 * d0;
#pragma unroll
        for (int k = 1; k < DIM; ++k) {
          T dk = p[r][k] - q.v[k];
          if (MODE == 3) { T ad = fabs(dk); dk = fmin(ad, per[k] - ad); }
          r2 = fma(dk, dk, r2);
        }
        T kv = OFFS ? fast_exp2(q.v[REC - 1] - r2) : fast_exp2(-r2);
        acc[r] += kv;
        if (GRAD) {
#pragma unroll
          for (int k = 0; k < DIM; ++k) g[r][k] = fma(kv, q.v[k], g[r][k]);
        }
      }
    }
    }  // scalar path
  }

  if (MODE == 2) {
#pragma unroll
    for (int r = 0; r < kRI; ++r) {
      int row = rbase + r * kThreads + tid;
      if (valid[r]) row_out[(size_t)tl.cb * row_out_stride + row] = acc[r];
    }
    return;
  }

  // block reduction in FP64: S, and for GRAD  F = sum_i (S_i a_i - G_i),  X = sum_i G_i x a_i
  double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int r = 0; r < kRI; ++r) {
    if (!valid[r]) continue;
    v[0] += (double)acc[r];
    if (GRAD) {
#pragma unroll
      for (int k = 0; k < DIM; ++k) v[1 + k] += (double)acc[r] * (double)p[r][k] - (double)g[r][k];
      if (DIM == 2) v[6] += (double)g[r][0] * (double)p[r][1] - (double)g[r][1] * (double)p[r][0];
      if (DIM == 3) {
        v[4] += (double)g[r][1] * (double)p[r][2] - (double)g[r][2] * (double)p[r][1];
        v[5] += (double)g[r][2] * (double)p[r][0] - (double)g[r][0] * (double)p[r][2];
        v[6] += (double)g[r][0] * (double)p[r][1] - (double)g[r][1] * (double)p[r][0];
      }
    }
  }
  constexpr int NV = GRAD ? 7 : 1;
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
  if ((tid & 31) == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) sred[tid >> 5][k] = v[k];
  }
  __syncthreads();
  if (tid < NV) {
    double s = 0;
    for (int w = 0; w < kThreads / 32; ++w) s += sred[w][tid];
    tile_out[(size_t)blockIdx.x * 8 + tid] = s * tl.weight;
  }
}

// fixed-order reduction of the per-tile partials of each sub-problem: out[sub*8 + k]
__global__ void reduce_tiles_kernel(const double* __restrict__ tile_out, const int* __restrict__ sub_tile0, int nv,
                                    double* __restrict__ out) {
  __shared__ double s[256];
  int sub = blockIdx.x, t0 = sub_tile0[sub], t1 = sub_tile0[sub + 1];
  for (int k = 0; k < nv; ++k) {
    double a = 0;
    for (int t = t0 + threadIdx.x; t < t1; t += blockDim.x) a += tile_out[(size_t)t * 8 + k];
    s[threadIdx.x] = a;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
      if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[sub * 8 + k] = s[0];
    __syncthreads();
  }
}

template <typename T>
__global__ void finalize_rows_kernel(const T* __restrict__ row_part, int ncb, int stride, int Q, double lognorm2,
                                     int logp, double* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Q) return;
  double s = 0;
  for (int c = 0; c < ncb; ++c) s += (double)row_part[(size_t)c * stride + i];
  out[i] = logp ? (log2(s) + lognorm2) * 0.69314718055994530942 : s * exp2(lognorm2);
}

// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void kde_sample_kernel(const double* __restrict__ mu, const double* __restrict__ cdf, int d, int N,
                                  double sig0, double sig1, double sig2, double sig3, double sig4, double sig5,
                                  uint32_t circ_mask, int n, uint32_t k0, uint32_t k1, uint32_t stream,
                                  int per_stream, uint32_t stream_stride, double* __restrict__ out) {
  // draw t of the launch is sample (t % per_stream) of Philox stream (stream + stream_stride * (t / per_stream)):
  // many independent sample sets (one per alignment of a ScatterAlign batch) in one launch
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int s = t % per_stream;
  stream += stream_stride * (uint32_t)(t / per_stream);
  const double sig[6] = {sig0, sig1, sig2, sig3, sig4, sig5};
  uint32_t o[4];
  philox4x32_10((uint32_t)s, 0u, stream, 2u, k0, k1, o);
  T u = u01<T>(o[0], o[1]);
  int idx;
  if (!cdf) {
    idx = (int)((double)u * (double)N);
    if (idx >= N) idx = N - 1;
  } else {  // first i with cdf[i] >= u * total
    double tgt = (double)u * cdf[N - 1];
    int lo = 0, hi = N - 1;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (cdf[mid] >= tgt) hi = mid; else lo = mid + 1;
    }
    idx = lo;
  }
  for (int k = 0; k < d; ++k) {
    if ((k & 1) == 0) philox4x32_10((uint32_t)s, (uint32_t)(k >> 1), stream, 3u, k0, k1, o);
    T u1 = u01<T>(o[0], o[1]), u2 = u01<T>(o[2], o[3]);
    double r = sqrt(-2.0 * log((double)u1));
    double sn, cs;
    sincospi(2.0 * (double)u2, &sn, &cs);
    double z = (k & 1) ? r * sn : r * cs;
    double v = mu[(size_t)idx * d + k] + sig[k] * z;
    if (circ_mask >> k & 1) v = wrap_pi(v);
    out[(size_t)t * d + k] = v;
  }
}

// inclusive prefix sum of the kernel weights (the sampler's CDF) by one CTA: every thread sums a contiguous chunk, the chunk
// totals are scanned across the block (shuffles within the warps, then across the warps), and each thread writes its chunk's
// running sums on top of its offset.  Deterministic; differs from a sequential sum by FP64 rounding only.
constexpr int kScanThreads = 1024;
__global__ void __launch_bounds__(kScanThreads) cumsum_block_kernel(const double* __restrict__ w, int N, double* __restrict__ cdf) {
  __shared__ double warp_tot[kScanThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int chunk = (N + kScanThreads - 1) / kScanThreads;
  const int i0 = min(N, tid * chunk), i1 = min(N, i0 + chunk);
  double tot = 0;
  for (int i = i0; i < i1; ++i) tot += w[i];
  double incl = tot;  // inclusive scan of the chunk totals within the warp
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_tot[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    double v = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    warp_tot[lane] = v;
  }
  __syncthreads();
  double c = (incl - tot) + (wid > 0 ? warp_tot[wid - 1] : 0.0);  // sum of everything before this thread's chunk
  for (int i = i0; i < i1; ++i) { c += w[i]; cdf[i] = c; }
}

// ---------------------------------------------------------------------------------------------------
// host side
inline int pad_dim(int d) { return d <= 3 ? d : 6; }
inline int rec_of(int d) { return d <= 3 ? 4 : 8; }

template <typename T>
int convert_async(cb2_ctx* ctx, const double* src_dev, int n, int d, const ConvertArgs& a, const double* center_src,
                  const double* w_dev, double wsum, T* dst) {
  if (n == 0) return CB2_OK;
  int blocks = (n + 255) / 256;
  if (rec_of(d) == 4)
    convert_records_kernel<T, 4><<<blocks, 256, 0, ctx->stream>>>(src_dev, n, d, a, center_src, w_dev, wsum, dst);
  else
    convert_records_kernel<T, 8><<<blocks, 256, 0, ctx->stream>>>(src_dev, n, d, a, center_src, w_dev, wsum, dst);
  CB2_CHECK_LAUNCH(ctx, "convert_records_kernel");
  return CB2_OK;
}

template <typename T, int MODE>
int launch_pair_tiles(cb2_ctx* ctx, int d, int ntiles, const PairSub<T>* subs, const Tile* tiles, const T* xf,
                      double* tile_out, T* row_out, int stride) {
  if (ntiles == 0) return CB2_OK;
  switch (pad_dim(d)) {
    case 1: pair_tiles_kernel<T, 1, MODE><<<ntiles, kThreads, 0, ctx->stream>>>(subs, tiles, xf, tile_out, row_out, stride); break;
    case 2: pair_tiles_kernel<T, 2, MODE><<<ntiles, kThreads, 0, ctx->stream>>>(subs, tiles, xf, tile_out, row_out, stride); break;
    case 3: pair_tiles_kernel<T, 3, MODE><<<ntiles, kThreads, 0, ctx->stream>>>(subs, tiles, xf, tile_out, row_out, stride); break;
    default:
      if (MODE == 1) return cb2_fail(ctx, CB2_ERR_INVALID, "rigid-transform batch needs dim 2 or 3");
      pair_tiles_kernel<T, 6, (MODE == 1 ? 0 : MODE)><<<ntiles, kThreads, 0, ctx->stream>>>(subs, tiles, xf, tile_out, row_out, stride);
      break;
  }
  CB2_CHECK_LAUNCH(ctx, "pair_tiles_kernel");
  return CB2_OK;
}

struct SubSpec { int rows_is_b, cols_is_b, row0, row1, ncols, xf, col0 = 0; };

// Builds tiles for the given sub-problems over clouds a (N) and b (M) already resident as Float64 on the
// device, runs the pair kernel and leaves per-sub reduced values (8 doubles per sub) in *result_dev.
template <typename T, int MODE>
int run_pairs_dev(cb2_ctx* ctx, const double* a_dev, int N, const double* b_dev, int M, int d, double scale,
                  const std::vector<SubSpec>& specs, const std::vector<T>& xf_host, double** result_dev,
                  size_t extra_bytes, void** extra_ptr, bool reuse_prev = false, const double* dim_scale = nullptr,
                  uint32_t wrap_mask = 0, int shard = 0, int nshards = 1) {
  // reuse_prev: the previous call on this context had the same clouds, sub-problems and sizes and nothing else touched
  // the arena since (consecutive cost evaluations of one BFGS batch): the arena layout is a pure function of the
  // sizes, so the converted records and the tile tables are still in place -- only the transforms are uploaded again.
  const int rec = rec_of(d);
  std::vector<Tile> tiles;
  std::vector<int> sub_tile0;
  for (size_t s = 0; s < specs.size(); ++s) {
    sub_tile0.push_back((int)tiles.size());
    int nr = specs[s].row1 - specs[s].row0;
    const int RB = kThreads * kRI;
    int nrb = (nr + RB - 1) / RB, ncb = (specs[s].ncols + kCB - 1) / kCB;
    // k(p,q) = k(q,p): a full self pair sum (aa, bb) only needs the diagonal blocks once and the blocks above them twice
    const bool sym = (MODE == 0 || MODE == 3) && specs[s].rows_is_b == specs[s].cols_is_b && specs[s].row0 == 0 && specs[s].col0 == 0 &&
                     specs[s].row1 == specs[s].ncols && specs[s].xf < 0;
    for (int rb = 0; rb < nrb; ++rb) {
      if (!sym) {
        for (int cb = 0; cb < ncb; ++cb) tiles.push_back({(int)s, rb, cb, cb * kCB, std::min((cb + 1) * kCB, specs[s].ncols), 1});
      } else {
        const int r0 = rb * RB, r1 = std::min(r0 + RB, specs[s].ncols);
        tiles.push_back({(int)s, rb, 0, r0, r1, 1});
        for (int c = r1; c < specs[s].ncols; c += kCB) tiles.push_back({(int)s, rb, 0, c, std::min(c + kCB, specs[s].ncols), 2});
      }
    }
  }
  sub_tile0.push_back((int)tiles.size());
  if (nshards > 1) {
    // multi-GPU sharding of the TILE LIST: shard r keeps every nshards-th tile, so the symmetric halving of the self
    // terms survives the split and every GPU evaluates (N^2/2 + N M + M^2/2) / nshards pairs; tiles stay grouped by sub-problem
    std::vector<Tile> mine;
    std::vector<int> st0(specs.size() + 1, 0);
    for (size_t s = 0; s < specs.size(); ++s) {
      st0[s] = (int)mine.size();
      for (int t = sub_tile0[s]; t < sub_tile0[s + 1]; ++t)
        if (t % nshards == shard) mine.push_back(tiles[t]);
    }
    st0[specs.size()] = (int)mine.size();
    tiles.swap(mine);
    sub_tile0.swap(st0);
  }
  const int ntiles = (int)tiles.size(), nsub = (int)specs.size();

  size_t need = cb2_align(sizeof(T) * rec * (size_t)N) + cb2_align(sizeof(T) * rec * (size_t)M) +
                cb2_align(sizeof(PairSub<T>) * nsub) + cb2_align(sizeof(Tile) * (size_t)ntiles + 16) +
                cb2_align(sizeof(int) * (nsub + 1)) + cb2_align(sizeof(T) * (xf_host.size() + 12)) +
                cb2_align(sizeof(double) * 8 * (size_t)(ntiles + 1)) + cb2_align(sizeof(double) * 8 * nsub) +
                cb2_align(extra_bytes) + 4096;
  int rc = cb2_arena_reserve(ctx, need);
  if (rc) return rc;
  T* ra = (T*)cb2_arena_alloc(ctx, sizeof(T) * rec * (size_t)N);
  T* rb = (T*)cb2_arena_alloc(ctx, sizeof(T) * rec * (size_t)M);
  PairSub<T>* subs_dev = (PairSub<T>*)cb2_arena_alloc(ctx, sizeof(PairSub<T>) * nsub);
  Tile* tiles_dev = (Tile*)cb2_arena_alloc(ctx, sizeof(Tile) * (size_t)ntiles + 16);
  int* st0_dev = (int*)cb2_arena_alloc(ctx, sizeof(int) * (nsub + 1));
  T* xf_dev = (T*)cb2_arena_alloc(ctx, sizeof(T) * (xf_host.size() + 12));
  double* tile_out = (double*)cb2_arena_alloc(ctx, sizeof(double) * 8 * (size_t)(ntiles + 1));
  double* res = (double*)cb2_arena_alloc(ctx, sizeof(double) * 8 * nsub);
  if (extra_ptr) *extra_ptr = extra_bytes ? cb2_arena_alloc(ctx, extra_bytes) : nullptr;
  if (!ra || !rb || !subs_dev || !tiles_dev || !st0_dev || !xf_dev || !tile_out || !res)
    return cb2_fail(ctx, CB2_ERR_NOMEM, "pair scratch exhausted");

  if (!reuse_prev) {
  ConvertArgs ca;
  memset(&ca, 0, sizeof(ca));
  for (int k = 0; k < 8; ++k) { ca.scale[k] = dim_scale ? dim_scale[k] : scale; ca.wrap[k] = (wrap_mask >> k) & 1; }
  ca.use_center_ptr = 1;  // centre both clouds on a[0]: rigid-invariant, keeps FP32 differences exact
  rc = convert_async<T>(ctx, a_dev, N, d, ca, a_dev, nullptr, 1.0, ra);
  if (rc) return rc;
  rc = convert_async<T>(ctx, b_dev, M, d, ca, a_dev, nullptr, 1.0, rb);
  if (rc) return rc;

  std::vector<PairSub<T>> subs(nsub);
  for (int s = 0; s < nsub; ++s) {
    subs[s].rows = specs[s].rows_is_b ? rb : ra;
    subs[s].cols = (specs[s].cols_is_b ? rb : ra) + (size_t)specs[s].col0 * rec;
    subs[s].row0 = specs[s].row0; subs[s].row1 = specs[s].row1; subs[s].ncols = specs[s].ncols;
    subs[s].xf = specs[s].xf;
  }
  CB2_CUDA(ctx, cudaMemcpyAsync(subs_dev, subs.data(), sizeof(PairSub<T>) * nsub, cudaMemcpyHostToDevice, ctx->stream));
  if (ntiles)
    CB2_CUDA(ctx, cudaMemcpyAsync(tiles_dev, tiles.data(), sizeof(Tile) * (size_t)ntiles, cudaMemcpyHostToDevice, ctx->stream));
  CB2_CUDA(ctx, cudaMemcpyAsync(st0_dev, sub_tile0.data(), sizeof(int) * (nsub + 1), cudaMemcpyHostToDevice, ctx->stream));
  }
  if (!xf_host.empty())
    CB2_CUDA(ctx, cudaMemcpyAsync(xf_dev, xf_host.data(), sizeof(T) * xf_host.size(), cudaMemcpyHostToDevice, ctx->stream));
  rc = launch_pair_tiles<T, MODE>(ctx, d, ntiles, subs_dev, tiles_dev, xf_dev, tile_out, nullptr, 0);
  if (rc) return rc;
  reduce_tiles_kernel<<<nsub, 256, 0, ctx->stream>>>(tile_out, st0_dev, MODE == 1 ? 7 : 1, res);
  CB2_CHECK_LAUNCH(ctx, "reduce_tiles_kernel");
  *result_dev = res;
  return CB2_OK;
}

__global__ void gather_sums3_kernel(const double* __restrict__ res, double* __restrict__ out) {
  if (threadIdx.x < 3) out[threadIdx.x] = res[threadIdx.x * 8];
}

template <typename T>
int mmd_sums_dev_impl(cb2_ctx* ctx, const double* a_dev, int N, const double* b_dev, int M, int d, double bw,
                      int ra0, int ra1, int rb0, int rb1, double* sums3_dev, int shard = 0, int nshards = 1) {
  std::vector<SubSpec> specs = {{0, 0, ra0, ra1, N, -1}, {0, 1, ra0, ra1, M, -1}, {1, 1, rb0, rb1, M, -1}};
  double* res = nullptr;
  int rc = run_pairs_dev<T, 0>(ctx, a_dev, N, b_dev, M, d, sqrt(bw * CB2_LOG2E), specs, std::vector<T>(), &res, 0, nullptr, false, nullptr,
                               0, shard, nshards);
  if (rc) return rc;
  gather_sums3_kernel<<<1, 32, 0, ctx->stream>>>(res, sums3_dev);
  CB2_CHECK_LAUNCH(ctx, "gather_sums3_kernel");
  return CB2_OK;
}

int check_cloud_args(cb2_ctx* ctx, const void* a, int N, const void* b, int M, int d) {
  CB2_REQUIRE(ctx, a && b, "null point pointer");
  CB2_REQUIRE(ctx, N > 0 && M > 0, "empty point set (N=%d, M=%d)", N, M);
  if (d < 1 || d > CB2_MAX_DIM) return cb2_fail(ctx, CB2_ERR_UNSUPPORTED, "dimension %d outside 1..%d", d, CB2_MAX_DIM);
  return CB2_OK;
}

// SE(2)/SE(3) maths on the host in FP64 (tiny; the per-point application is fused into the pair kernel)
void so3_exp(const double* w, double* R) {
  double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2), A, B;
  if (th < 1e-8) { A = 1.0 - th2 / 6.0; B = 0.5 - th2 / 24.0; }
  else { A = sin(th) / th; B = (1.0 - cos(th)) / th2; }
  double K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double k2 = 0;
      for (int k = 0; k < 3; ++k) k2 += K[i * 3 + k] * K[k * 3 + j];
      R[i * 3 + j] = (i == j ? 1.0 : 0.0) + A * K[i * 3 + j] + B * k2;
    }
}
void so3_right_jacobian(const double* w, double* J) {
  double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2), B, C;
  if (th < 1e-6) { B = 0.5 - th2 / 24.0; C = 1.0 / 6.0 - th2 / 120.0; }
  else { B = (1.0 - cos(th)) / th2; C = (th - sin(th)) / (th2 * th); }
  double K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double k2 = 0;
      for (int k = 0; k < 3; ++k) k2 += K[i * 3 + k] * K[k * 3 + j];
      J[i * 3 + j] = (i == j ? 1.0 : 0.0) - B * K[i * 3 + j] + C * k2;
    }
}
// (R 3x3 row-major with the 2-D case embedded in the top-left, t[3]) of T(c) or its inverse
void se_matrix(int dim, const double* c, int backward, double* R, double* t) {
  double Rf[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (dim == 2) { Rf[0] = cos(c[2]); Rf[1] = -sin(c[2]); Rf[3] = sin(c[2]); Rf[4] = cos(c[2]); }
  else so3_exp(c + 3, Rf);
  double tf[3] = {c[0], c[1], dim == 3 ? c[2] : 0.0};
  if (!backward) { memcpy(R, Rf, sizeof(Rf)); memcpy(t, tf, sizeof(tf)); return; }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i * 3 + j] = Rf[j * 3 + i];
  for (int i = 0; i < 3; ++i) t[i] = -(R[i * 3] * tf[0] + R[i * 3 + 1] * tf[1] + R[i * 3 + 2] * tf[2]);
}

}  // namespace

// shared with align.cu: value + analytic gradient of mmd(a, T(c_k) b) for K coordinate vectors; a/b device Float64
template <typename T>
int cb2_se_batch_core(cb2_ctx* ctx, const double* a_dev, int N, const double* b_dev, int M, int dim, double bw,
                      const double* a0_host, const double* coords, int K, int backward, const double* aa, const double* bb,
                      int row_stride, int col_stride, double* vals, double* grads, bool reuse_prev) {
  // sub-problem k pairs rows [k*row_stride, k*row_stride+N) of a with columns [k*col_stride, k*col_stride+M) of b
  // (strides 0: every candidate transform sees the same two clouds); aa/bb hold 1 (strides 0) or K self terms.
  const int nc = dim == 2 ? 3 : 6;
  const double s = sqrt(bw * CB2_LOG2E);
  std::vector<T> xf((size_t)K * 12);
  std::vector<double> Rs((size_t)K * 9), ts((size_t)K * 3);
  double c0[3] = {a0_host[0], a0_host[1], dim == 3 ? a0_host[2] : 0.0};
  for (int k = 0; k < K; ++k) {
    double* R = &Rs[(size_t)k * 9];
    double* t = &ts[(size_t)k * 3];
    se_matrix(dim, coords + (size_t)k * nc, backward, R, t);
    for (int i = 0; i < 9; ++i) xf[(size_t)k * 12 + i] = (T)R[i];
    for (int i = 0; i < 3; ++i)  // translation in the centred + scaled frame: s (R c0 + t - c0)
      xf[(size_t)k * 12 + 9 + i] = (T)(s * (R[i * 3] * c0[0] + R[i * 3 + 1] * c0[1] + R[i * 3 + 2] * c0[2] + t[i] - c0[i]));
  }
  std::vector<SubSpec> specs(K);
  for (int k = 0; k < K; ++k) specs[k] = {0, 1, k * row_stride, k * row_stride + N, M, k, k * col_stride};
  double* res = nullptr;
  int rc = run_pairs_dev<T, 1>(ctx, a_dev, N + (K - 1) * row_stride, b_dev, M + (K - 1) * col_stride, dim, s, specs, xf, &res, 0, nullptr, reuse_prev);
  if (rc) return rc;
  std::vector<double> h((size_t)K * 8);
  CB2_CUDA(ctx, cudaMemcpyAsync(h.data(), res, sizeof(double) * 8 * K, cudaMemcpyDeviceToHost, ctx->stream));
  CB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < K; ++k) {
    const double* v = &h[(size_t)k * 8];
    const double* R = &Rs[(size_t)k * 9];
    const double* t = &ts[(size_t)k * 3];
    const double* c = coords + (size_t)k * nc;
    double S = v[0];
    const double aak = aa[row_stride ? k : 0], bbk = bb[col_stride ? k : 0];
    vals[k] = aak / ((double)N * N) - 2.0 * S / ((double)N * M) + bbk / ((double)M * M);
    if (!grads) continue;
    // undo centring/scaling: F = F~/s ; X = X~/s^2 + c0 x F   (X = sum_j b'_j x F_j)
    double F[3] = {v[1] / s, v[2] / s, v[3] / s};
    double X[3] = {v[4] / (s * s), v[5] / (s * s), v[6] / (s * s)};
    X[0] += c0[1] * F[2] - c0[2] * F[1];
    X[1] += c0[2] * F[0] - c0[0] * F[2];
    X[2] += c0[0] * F[1] - c0[1] * F[0];
    double* g = grads + (size_t)k * nc;
    const double sc = -2.0 / ((double)N * M) * 2.0 * bw;
    if (dim == 2) {
      if (backward) {
        g[0] = sc * -(R[0] * F[0] + R[3] * F[1]);
        g[1] = sc * -(R[1] * F[0] + R[4] * F[1]);
        g[2] = sc * -X[2];
      } else {
        g[0] = sc * F[0]; g[1] = sc * F[1];
        g[2] = sc * (X[2] - (t[0] * F[1] - t[1] * F[0]));
      }
    } else {
      double Jr[9], u[3], w[3];
      so3_right_jacobian(c + 3, Jr);
      if (backward) {
        for (int x = 0; x < 3; ++x) g[x] = sc * -(R[x] * F[0] + R[3 + x] * F[1] + R[6 + x] * F[2]);
        for (int x = 0; x < 3; ++x) w[x] = -X[x];
      } else {
        for (int x = 0; x < 3; ++x) g[x] = sc * F[x];
        u[0] = X[0] - (t[1] * F[2] - t[2] * F[1]);
        u[1] = X[1] - (t[2] * F[0] - t[0] * F[2]);
        u[2] = X[2] - (t[0] * F[1] - t[1] * F[0]);
        for (int x = 0; x < 3; ++x) w[x] = R[x] * u[0] + R[3 + x] * u[1] + R[6 + x] * u[2];
      }
      for (int x = 0; x < 3; ++x) g[3 + x] = sc * (Jr[x] * w[0] + Jr[3 + x] * w[1] + Jr[6 + x] * w[2]);
    }
  }
  return CB2_OK;
}
template int cb2_se_batch_core<float>(cb2_ctx*, const double*, int, const double*, int, int, double, const double*,
                                      const double*, int, int, const double*, const double*, int, int, double*, double*, bool);
template int cb2_se_batch_core<double>(cb2_ctx*, const double*, int, const double*, int, int, double, const double*,
                                       const double*, int, int, const double*, const double*, int, int, double*, double*, bool);

// ---------------------------------------------------------------------------------------------------
// C ABI
extern "C" int cb2_mmd_sums_dev(cb2_ctx* ctx, const double* a_dev, int N, const double* b_dev, int M, int d, double bw,
                                int ra0, int ra1, int rb0, int rb1, double* sums3_dev) {
  if (!ctx) return CB2_ERR_INVALID;
  cb2_lock g(ctx->mu);
  int rc = check_cloud_args(ctx, a_dev, N, b_dev, M, d);
  if (rc) return rc;
  CB2_REQUIRE(ctx, sums3_dev, "null output");
  CB2_REQUIRE(ctx, bw > 0 && isfinite(bw), "bw must be positive");
  CB2_REQUIRE(ctx, 0 <= ra0 && ra0 <= ra1 && ra1 <= N && 0 <= rb0 && rb0 <= rb1 && rb1 <= M, "row range out of bounds");
  CB2_CUDA(ctx, cudaSetDevice(ctx->device));
  return ctx->precision == CB2_FP64 ? mmd_sums_dev_impl<double>(ctx, a_dev, N, b_dev, M, d, bw, ra0, ra1, rb0, rb1, sums3_dev)
                                    : mmd_sums_dev_impl<float>(ctx, a_dev, N, b_dev, M, d, bw, ra0, ra1, rb0, rb1, sums3_dev);
}

// Shard `shard` of `nshards` of the three pair sums of mmd(a, b): the tile list of the full (symmetric) evaluation dealt
// round-robin.  Summing the nshards results (ncclAllReduce of 3 doubles) gives exactly the sums cb2_mmd_sums returns for the
// full row ranges; each shard evaluates 1/nshards of the pairs the single-GPU kernel evaluates (self terms once per pair).
extern "C" int cb2_mmd_sums_shard_dev(cb2_ctx* ctx, const double* a_dev, int N, const double* b_dev, int M, int d, double bw,
                                      int shard, int nshards, double* sums3_dev) {
  if (!ctx) return CB2_ERR_INVALID;
  cb2_lock g(ctx->mu);
  int rc = check_cloud_args(ctx, a_dev, N, b_dev, M, d);
  if (rc) return rc;
  CB2_REQUIRE(ctx, sums3_dev, "null output");
  CB2_REQUIRE(ctx, bw > 0 && isfinite(bw), "bw must be positive");
  CB2_REQUIRE(ctx, nshards >= 1 && shard >= 0 && shard < nshards, "shard %d outside 0..%d", shard, nshards - 1);
  CB2_CUDA(ctx, cudaSetDevice(ctx->device));
  return ctx->precision == CB2_FP64 ? mmd_sums_dev_impl<double>(ctx, a_dev, N, b_dev, M, d, bw, 0, N, 0, M, sums3_dev, shard, nshards)
                                    : mmd_sums_dev_impl<float>(ctx, a_dev, N, b_dev, M, d, bw, 0, N, 0, M, sums3_dev, shard, nshards);
}

// uploads host Float64 arrays into the context's I/O block; used by the (synchronous) host-pointer API
static int upload_block(cb2_ctx* ctx, const std::vector<const double*>& src, const std::vector<size_t>& count,
                        std::vector<double*>& dev, double** block) {
  size_t total = 0;
  for (size_t c : count) total += cb2_align(c * sizeof(double));
  total += 256;
  void* p = cb2_io_block(ctx, total);  // context-owned, reused across calls (the callers' cudaFree(block) is a no-op)
  if (!p) return CB2_ERR_NOMEM;
  *block = nullptr;
  cudaError_t e = cudaSuccess;
  size_t off = 0;
  dev.resize(src.size());
  for (size_t i = 0; i < src.size(); ++i) {
    dev[i] = (double*)((char*)p + off);
    if (count[i] && src[i]) {
      e = cudaMemcpyAsync(dev[i], src[i], count[i] * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
      if (e != cudaSuccess) return cb2_fail(ctx, CB2_ERR_CUDA, "H2D copy: %s", cudaGetErrorString(e));
    }
    off += cb2_align(count[i] * sizeof(double));
  }
  return CB2_OK;
}

extern "C" int cb2_mmd_sums(cb2_ctx* ctx, const double* a, int N, const double* b, int M, int d, double bw,
                            int ra0, int ra1, int rb0, int rb1, double* sums3) {
  if (!ctx) return CB2_ERR_INVALID;
  cb2_lock g(ctx->mu);  // held across upload, the nested _dev call and the read-back: the I/O block is shared per context
  int rc = check_cloud_args(ctx, a, N, b, M, d);
  if (rc) return rc;
  CB2_REQUIRE(ctx, sums3, "null output");
  CB2_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<double*> dev;
  double* block = nullptr;
  rc = upload_block(ctx, {a, b, nullptr}, {(size_t)N * d, (size_t)M * d, 3}, dev, &block);
  if (rc) return rc;
  rc = cb2_mmd_sums_dev(ctx, dev[0], N, dev[1], M, d, bw, ra0, ra1, rb0, rb1, dev[2]);
  if (rc == CB2_OK) {
    cudaError_t e = cudaMemcpyAsync(sums3, dev[2], 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = cb2_fail(ctx, CB2_ERR_CUDA, "mmd readback: %s", cudaGetErrorString(e));
  } else {
    cudaStreamSynchronize(ctx->stream);
  }
  cudaFree(block);
  return rc;
}

extern "C" int cb2_mmd_sums_shard(cb2_ctx* ctx, const double* a, int N, const double* b, int M, int d, double bw, int shard,
                                  int nshards, double* sums3) {
  if (!ctx) return CB2_ERR_INVALID;
  cb2_lock g(ctx->mu);
  int rc = check_cloud_args(ctx, a, N, b, M, d);
  if (rc) return rc;
  CB2_REQUIRE(ctx, sums3, "null output");
  CB2_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<double*> dev;
  double* block = nullptr;
  rc = upload_block(ctx, {a, b, nullptr}, {(size_t)N * d, (size_t)M * d, 3}, dev, &block);
  if (rc) return rc;
  rc = cb2_mmd_sums_shard_dev(ctx, dev[0], N, dev[1], M, d, bw, shard, nshards, dev[2]);
  if (rc == CB2_OK) {
    cudaError_t e = cudaMemcpyAsync(sums3, dev[2], 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = cb2_fail(ctx, CB2_ERR_CUDA, "mmd readback: %s", cudaGetErrorString(e));
  } else {
    cudaStreamSynchronize(ctx->stream);
  }
  return rc;
}

extern "C" int cb2_mmd(cb2_ctx* ctx, const double* a, int N, const double* b, int M, int d, double bw, double* out) {
  if (!ctx) return CB2_ERR_INVALID;
  if (!out) return cb2_fail(ctx, CB2_ERR_INVALID, "null output");
  double s[3];
  int rc = cb2_mmd_sums(ctx, a, N, b, M, d, bw, 0, N, 0, M, s);
  if (rc) return rc;
  *out = s[0] / ((double)N * N) - 2.0 * s[1] / ((double)N * M) + s[2] / ((double)M * M);
  return CB2_OK;
}

extern "C" int cb2_mmd_se_batch_dev(cb2_ctx* ctx, const double* a_dev, int N, const double* b_dev, int M, int dim,
                                    double bw, const double* coords, int K, int backward, double* vals, double* grads) {
  if (!ctx) return CB2_ERR_INVALID;
  cb2_lock g(ctx->mu);
  int rc = check_cloud_args(ctx, a_dev, N, b_dev, M, dim);
  if (rc) return rc;
  CB2_REQUIRE(ctx, dim == 2 || dim == 3, "rigid-transform batch needs dim 2 or 3 (got %d)", dim);
  CB2_REQUIRE(ctx, coords && vals && K > 0, "null coords/vals or K <= 0");
  CB2_REQUIRE(ctx, bw > 0 && isfinite(bw), "bw must be positive");
  CB2_CUDA(ctx, cudaSetDevice(ctx->device));
  // aa, bb are invariant under a rigid transform of b: one mmd pass, then the K cross terms
  double a0[3] = {0, 0, 0}, s3[3];
  CB2_CUDA(ctx, cudaMemcpyAsync(a0, a_dev, sizeof(double) * dim, cudaMemcpyDeviceToHost, ctx->stream));
  double* s3_dev = ctx->small_dev;
  rc = ctx->precision == CB2_FP64 ? mmd_sums_dev_impl<double>(ctx, a_dev, N, b_dev, M, dim, bw, 0, N, 0, M, s3_dev)
                                  : mmd_sums_dev_impl<float>(ctx, a_dev, N, b_dev, M, dim, bw, 0, N, 0, M, s3_dev);
  if (rc == CB2_OK) {
    cudaError_t e = cudaMemcpyAsync(s3, s3_dev, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = cb2_fail(ctx, CB2_ERR_CUDA, "mmd readback: %s", cudaGetErrorString(e));
  }
  if (rc) return rc;
  return ctx->precision == CB2_FP64
             ? cb2_se_batch_core<double>(ctx, a_dev, N, b_dev, M, dim, bw, a0, coords, K, backward, &s3[0], &s3[2], 0, 0, vals, grads, false)
             : cb2_se_batch_core<float>(ctx, a_dev, N, b_dev, M, dim, bw, a0, coords, K, backward, &s3[0], &s3[2], 0, 0, vals, grads, false);
}

extern "C" int cb2_mmd_se_batch(cb2_ctx* ctx, const double* a, int N, const double* b, int M, int dim, double bw,
                                const double* coords, int K, int backward, double* vals, double* grads) {
  if (!ctx) return CB2_ERR_INVALID;
  cb2_lock g(ctx->mu);  // whole call: see cb2_mmd_sums
  int rc = check_cloud_args(ctx, a, N, b, M, dim);
  if (rc) return rc;
  CB2_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<double*> dev;
  double* block = nullptr;
  rc = upload_block(ctx, {a, b}, {(size_t)N * dim, (size_t)M * dim}, dev, &block);
  if (rc) return rc;
  rc = cb2_mmd_se_batch_dev(ctx, dev[0], N, dev[1], M, dim, bw, coords, K, backward, vals, grads);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(block);
  return rc;
}

template <typename T>
static int kde_eval_impl(cb2_ctx* ctx, const double* mu_dev, const double* sigma, const double* w_dev, double wsum,
                         const double* mu0, int d, int N, const double* q_dev, int Q, int logp, double* out_dev) {
  const int rec = rec_of(d);
  const int nrb = (Q + kThreads * kRI - 1) / (kThreads * kRI), ncb = (N + kCB - 1) / kCB;
  std::vector<Tile> tiles;
  for (int rb = 0; rb < nrb; ++rb)
    for (int cb = 0; cb < ncb; ++cb) tiles.push_back({0, rb, cb, cb * kCB, std::min((cb + 1) * kCB, N), 1});
  const int ntiles = (int)tiles.size();
  size_t need = cb2_align(sizeof(T) * rec * (size_t)N) + cb2_align(sizeof(T) * rec * (size_t)Q) + cb2_align(sizeof(PairSub<T>)) +
                cb2_align(sizeof(Tile) * (size_t)ntiles) + cb2_align(sizeof(T) * (size_t)ncb * Q) + 4096;
  int rc = cb2_arena_reserve(ctx, need);
  if (rc) return rc;
  T* rmu = (T*)cb2_arena_alloc(ctx, sizeof(T) * rec * (size_t)N);
  T* rq = (T*)cb2_arena_alloc(ctx, sizeof(T) * rec * (size_t)Q);
  PairSub<T>* sub_dev = (PairSub<T>*)cb2_arena_alloc(ctx, sizeof(PairSub<T>));
  Tile* tiles_dev = (Tile*)cb2_arena_alloc(ctx, sizeof(Tile) * (size_t)ntiles);
  T* row_part = (T*)cb2_arena_alloc(ctx, sizeof(T) * (size_t)ncb * Q);
  if (!rmu || !rq || !sub_dev || !tiles_dev || !row_part) return cb2_fail(ctx, CB2_ERR_NOMEM, "kde_eval scratch exhausted");
  ConvertArgs ca;
  memset(&ca, 0, sizeof(ca));
  double lognorm2 = 0;
  for (int k = 0; k < d; ++k) {
    ca.scale[k] = sqrt(0.5 * CB2_LOG2E) / sigma[k];
    ca.center[k] = mu0[k];
    lognorm2 -= 0.5 * log2(CB2_TWO_PI * sigma[k] * sigma[k]);
  }
  if (!w_dev) lognorm2 -= log2((double)N);
  rc = convert_async<T>(ctx, mu_dev, N, d, ca, nullptr, w_dev, wsum, rmu);
  if (rc) return rc;
  rc = convert_async<T>(ctx, q_dev, Q, d, ca, nullptr, nullptr, 1.0, rq);
  if (rc) return rc;
  PairSub<T> sub = {rq, rmu, 0, Q, N, -1};
  CB2_CUDA(ctx, cudaMemcpyAsync(sub_dev, &sub, sizeof(sub), cudaMemcpyHostToDevice, ctx->stream));
  CB2_CUDA(ctx, cudaMemcpyAsync(tiles_dev, tiles.data(), sizeof(Tile) * (size_t)ntiles, cudaMemcpyHostToDevice, ctx->stream));
  rc = launch_pair_tiles<T, 2>(ctx, d, ntiles, sub_dev, tiles_dev, nullptr, nullptr, row_part, Q);
  if (rc) return rc;
  finalize_rows_kernel<T><<<(Q + 255) / 256, 256, 0, ctx->stream>>>(row_part, ncb, Q, Q, lognorm2, logp, out_dev);
  CB2_CHECK_LAUNCH(ctx, "finalize_rows_kernel");
  return CB2_OK;
}

extern "C" int cb2_kde_eval(cb2_ctx* ctx, const double* mu, const double* sigma, const double* w, int d, int N,
                            const double* q, int Q, int logp, double* out) {
  if (!ctx) return CB2_ERR_INVALID;
  cb2_lock g(ctx->mu);
  int rc = check_cloud_args(ctx, mu, N, q, Q, d);
  if (rc) return rc;
  CB2_REQUIRE(ctx, sigma && out, "null sigma/out");
  for (int k = 0; k < d; ++k) CB2_REQUIRE(ctx, sigma[k] > 0 && isfinite(sigma[k]), "sigma[%d] must be positive", k);
  double wsum = 0;
  if (w) {
    for (int i = 0; i < N; ++i) { CB2_REQUIRE(ctx, w[i] >= 0, "negative weight"); wsum += w[i]; }
    CB2_REQUIRE(ctx, wsum > 0, "weights sum to zero");
  }
  CB2_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<double*> dev;
  double* block = nullptr;
  rc = upload_block(ctx, {mu, q, w, nullptr}, {(size_t)N * d, (size_t)Q * d, w ? (size_t)N : 0, (size_t)Q}, dev, &block);
  if (rc) return rc;
  rc = ctx->precision == CB2_FP64
           ? kde_eval_impl<double>(ctx, dev[0], sigma, w ? dev[2] : nullptr, wsum, mu, d, N, dev[1], Q, logp, dev[3])
           : kde_eval_impl<float>(ctx, dev[0], sigma, w ? dev[2] : nullptr, wsum, mu, d, N, dev[1], Q, logp, dev[3]);
  if (rc == CB2_OK) {
    cudaError_t e = cudaMemcpyAsync(out, dev[3], (size_t)Q * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = cb2_fail(ctx, CB2_ERR_CUDA, "kde_eval readback: %s", cudaGetErrorString(e));
  } else {
    cudaStreamSynchronize(ctx->stream);
  }
  cudaFree(block);
  return rc;
}

extern "C" int cb2_kde_sample(cb2_ctx* ctx, const double* mu, const double* sigma, const double* w, int d, int N,
                              const uint8_t* circular, int n, uint64_t seed, uint32_t stream, double* out) {
  if (!ctx) return CB2_ERR_INVALID;
  cb2_lock g(ctx->mu);
  CB2_REQUIRE(ctx, mu && sigma && out, "null pointer");
  CB2_REQUIRE(ctx, N > 0 && n >= 0, "bad sizes");
  if (d < 1 || d > CB2_MAX_DIM) return cb2_fail(ctx, CB2_ERR_UNSUPPORTED, "dimension %d outside 1..%d", d, CB2_MAX_DIM);
  if (n == 0) return CB2_OK;
  CB2_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<double*> dev;
  double* block = nullptr;
  int rc = upload_block(ctx, {mu, w, nullptr, nullptr}, {(size_t)N * d, w ? (size_t)N : 0, w ? (size_t)N : 0, (size_t)n * d}, dev, &block);
  if (rc) return rc;
  double sg[6] = {0, 0, 0, 0, 0, 0};
  uint32_t mask = 0;
  for (int k = 0; k < d; ++k) { sg[k] = sigma[k]; if (circular && circular[k]) mask |= 1u << k; }
  double* cdf = nullptr;
  if (w) {
    cumsum_block_kernel<<<1, kScanThreads, 0, ctx->stream>>>(dev[1], N, dev[2]);
    ctx->launches++;
    cdf = dev[2];
  }
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  if (ctx->precision == CB2_FP64)
    kde_sample_kernel<double><<<(n + 127) / 128, 128, 0, ctx->stream>>>(dev[0], cdf, d, N, sg[0], sg[1], sg[2], sg[3], sg[4], sg[5], mask, n, k0, k1, stream, n, 0u, dev[3]);
  else
    kde_sample_kernel<float><<<(n + 127) / 128, 128, 0, ctx->stream>>>(dev[0], cdf, d, N, sg[0], sg[1], sg[2], sg[3], sg[4], sg[5], mask, n, k0, k1, stream, n, 0u, dev[3]);
  ctx->launches++;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, dev[3], (size_t)n * d * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(block);
  if (e != cudaSuccess) return cb2_fail(ctx, CB2_ERR_CUDA, "kde_sample: %s", cudaGetErrorString(e));
  return CB2_OK;
}

// ---------------------------------------------------------------------------------------------------
// ScatterAlign getSample (R:ext/Images/ScatterAlignPose2.jl:155-187): K independent alignments as one batched
// BFGS.  Every cost/gradient evaluation of all K alignments is ONE launch of the MODE-1 pair kernel.
namespace {

struct Bfgs {
  int nc;
  std::vector<double> x, g, p, H, xt;
  double f = 0, alpha = 1, fprev = 1e300;
  int iters = 0, stalls = 0;
  bool done = false, first = true;
};

template <typename T>
int scatter_align_impl(cb2_ctx* ctx, const double* sa_dev, const double* sb_dev, int S, int dim, double bw, const double* a0,
                       const double* x0, int K, int max_iter, double g_tol, double* coords_out, double* score_out,
                       int32_t* iters_out) {
  const int nc = dim == 2 ? 3 : 6;
  // self terms of every alignment's two sample sets: 2K sub-problems in one MODE-0 pass
  std::vector<SubSpec> specs;
  for (int k = 0; k < K; ++k) specs.push_back({0, 0, k * S, (k + 1) * S, S, -1, k * S});
  for (int k = 0; k < K; ++k) specs.push_back({1, 1, k * S, (k + 1) * S, S, -1, k * S});
  double* res = nullptr;
  int rc = run_pairs_dev<T, 0>(ctx, sa_dev, K * S, sb_dev, K * S, dim, sqrt(bw * CB2_LOG2E), specs, std::vector<T>(), &res, 0, nullptr);
  if (rc) return rc;
  std::vector<double> self((size_t)2 * K * 8);
  CB2_CUDA(ctx, cudaMemcpyAsync(self.data(), res, sizeof(double) * 8 * 2 * K, cudaMemcpyDeviceToHost, ctx->stream));
  CB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  std::vector<double> aa(K), bb(K);
  for (int k = 0; k < K; ++k) { aa[k] = self[(size_t)k * 8]; bb[k] = self[(size_t)(K + k) * 8]; }

  std::vector<Bfgs> st(K);
  std::vector<double> coords((size_t)K * nc), vals(K), grads((size_t)K * nc);
  bool have_plan = false;  // after the first evaluation the converted records / tile tables are reused
  auto eval = [&]() {
    int r = cb2_se_batch_core<T>(ctx, sa_dev, S, sb_dev, S, dim, bw, a0, coords.data(), K, 1, aa.data(), bb.data(), S, S,
                                 vals.data(), grads.data(), have_plan);
    have_plan = true;
    return r;
  };
  for (int k = 0; k < K; ++k) {
    Bfgs& b = st[k];
    b.nc = nc; b.x.assign(x0 + (size_t)k * nc, x0 + (size_t)(k + 1) * nc);
    b.g.assign(nc, 0); b.p.assign(nc, 0); b.xt.assign(nc, 0); b.H.assign((size_t)nc * nc, 0);
    for (int i = 0; i < nc; ++i) { b.H[(size_t)i * nc + i] = 1.0; coords[(size_t)k * nc + i] = b.x[i]; }
  }
  rc = eval();
  if (rc) return rc;
  auto set_direction = [&](Bfgs& b) {
    double gn = 0, slope = 0;
    for (int i = 0; i < nc; ++i) {
      double v = 0;
      for (int j = 0; j < nc; ++j) v -= b.H[(size_t)i * nc + j] * b.g[j];
      b.p[i] = v; slope += v * b.g[i]; gn += b.g[i] * b.g[i];
    }
    if (!(slope < 0)) {  // lost positive definiteness: restart from steepest descent
      std::fill(b.H.begin(), b.H.end(), 0.0);
      for (int i = 0; i < nc; ++i) { b.H[(size_t)i * nc + i] = 1.0; b.p[i] = -b.g[i]; }
      b.first = true;
    }
    // first step of a (re)started search moves about half a unit (metres / radians); later ones try alpha = 1
    b.alpha = b.first ? 0.5 / (sqrt(gn) + 1e-300) : 1.0;
  };
  for (int k = 0; k < K; ++k) {
    Bfgs& b = st[k];
    b.f = vals[k];
    for (int i = 0; i < nc; ++i) b.g[i] = grads[(size_t)k * nc + i];
    double gmax = 0;
    for (int i = 0; i < nc; ++i) gmax = fmax(gmax, fabs(b.g[i]));
    if (gmax < g_tol) b.done = true; else set_direction(b);
  }
  const int max_evals = 8 * max_iter + 32;
  for (int ev = 0; ev < max_evals; ++ev) {
    bool any = false;
    for (int k = 0; k < K; ++k) {
      Bfgs& b = st[k];
      for (int i = 0; i < nc; ++i) {
        b.xt[i] = b.done ? b.x[i] : b.x[i] + b.alpha * b.p[i];
        coords[(size_t)k * nc + i] = b.xt[i];
      }
      any = any || !b.done;
    }
    if (!any) break;
    rc = eval();
    if (rc) return rc;
    for (int k = 0; k < K; ++k) {
      Bfgs& b = st[k];
      if (b.done) continue;
      double slope = 0;
      for (int i = 0; i < nc; ++i) slope += b.g[i] * b.p[i];
      const double ft = vals[k];
      if (ft <= b.f + 1e-4 * b.alpha * slope && isfinite(ft)) {  // Armijo accept -> BFGS update of the inverse Hessian
        std::vector<double> s(nc), y(nc), Hy(nc);
        double sy = 0, yy = 0;
        for (int i = 0; i < nc; ++i) {
          s[i] = b.xt[i] - b.x[i]; y[i] = grads[(size_t)k * nc + i] - b.g[i];
          sy += s[i] * y[i]; yy += y[i] * y[i];
        }
        if (sy > 1e-300) {
          if (b.first) {  // Nocedal & Wright (6.20): scale H0 before the first update
            for (int i = 0; i < nc; ++i) for (int j = 0; j < nc; ++j) b.H[(size_t)i * nc + j] = (i == j) ? sy / yy : 0.0;
            b.first = false;
          }
          double yHy = 0;
          for (int i = 0; i < nc; ++i) { double v = 0; for (int j = 0; j < nc; ++j) v += b.H[(size_t)i * nc + j] * y[j]; Hy[i] = v; yHy += v * y[i]; }
          for (int i = 0; i < nc; ++i)
            for (int j = 0; j < nc; ++j)
              b.H[(size_t)i * nc + j] += (1.0 + yHy / sy) * s[i] * s[j] / sy - (Hy[i] * s[j] + s[i] * Hy[j]) / sy;
        }
        b.fprev = b.f; b.f = ft; b.x = b.xt;
        for (int i = 0; i < nc; ++i) b.g[i] = grads[(size_t)k * nc + i];
        b.iters++;
        double gmax = 0;
        for (int i = 0; i < nc; ++i) gmax = fmax(gmax, fabs(b.g[i]));
        b.stalls = (b.fprev - b.f <= 1e-14 * (fabs(b.f) + 1e-300)) ? b.stalls + 1 : 0;
        if (gmax < g_tol || b.iters >= max_iter || b.stalls >= 2) b.done = true; else set_direction(b);
      } else {
        b.alpha *= 0.5;
        if (b.alpha * sqrt(slope * slope) < 1e-18 || b.alpha < 1e-12) b.done = true;  // no further decrease possible
      }
    }
  }
  for (int k = 0; k < K; ++k) {
    for (int i = 0; i < nc; ++i) coords_out[(size_t)k * nc + i] = st[k].x[i];
    if (score_out) score_out[k] = st[k].f;
    if (iters_out) iters_out[k] = st[k].iters;
  }
  return CB2_OK;
}

}  // namespace

extern "C" int cb2_scatter_align(cb2_ctx* ctx, const cb2_density* cloud1, const cb2_density* cloud2, int dim, int sample_count,
                                 double bw, const double* x0, int K, uint64_t seed, int max_iter, double g_tol,
                                 double* coords_out, double* score_out, int32_t* iters_out) {
  if (!ctx) return CB2_ERR_INVALID;
  cb2_lock g(ctx->mu);
  CB2_REQUIRE(ctx, cloud1 && cloud2 && cloud1->pts && cloud2->pts && cloud1->bw && cloud2->bw, "null cloud");
  CB2_REQUIRE(ctx, dim == 2 || dim == 3, "dim must be 2 or 3");
  CB2_REQUIRE(ctx, sample_count > 0 && K > 0 && x0 && coords_out, "bad sample_count / K / pointers");
  CB2_REQUIRE(ctx, cloud1->N > 0 && cloud2->N > 0, "empty cloud");
  CB2_REQUIRE(ctx, bw > 0 && isfinite(bw), "bw must be positive");
  if (max_iter <= 0) max_iter = 200;
  if (!(g_tol > 0)) g_tol = 1e-8;
  CB2_CUDA(ctx, cudaSetDevice(ctx->device));
  const int S = sample_count;
  const cb2_density* cl[2] = {cloud1, cloud2};
  std::vector<double*> dev;
  double* block = nullptr;
  int rc = upload_block(ctx, {cloud1->pts, cloud1->w, nullptr, cloud2->pts, cloud2->w, nullptr, nullptr, nullptr},
                        {(size_t)cloud1->N * dim, cloud1->w ? (size_t)cloud1->N : 0, cloud1->w ? (size_t)cloud1->N : 0,
                         (size_t)cloud2->N * dim, cloud2->w ? (size_t)cloud2->N : 0, cloud2->w ? (size_t)cloud2->N : 0,
                         (size_t)K * S * dim, (size_t)K * S * dim},
                        dev, &block);
  if (rc) return rc;
  // fresh samples for every alignment (:171-174): alignment k uses Philox streams 2k (cloud1) and 2k+1 (cloud2)
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int c = 0; c < 2 && rc == CB2_OK; ++c) {
    double* cdf = nullptr;
    if (cl[c]->w) {
      cumsum_block_kernel<<<1, kScanThreads, 0, ctx->stream>>>(dev[c * 3 + 1], cl[c]->N, dev[c * 3 + 2]);
      ctx->launches++;
      cdf = dev[c * 3 + 2];
    }
    double sg[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < dim; ++i) sg[i] = cl[c]->bw[i];
    {  // alignment k draws from Philox stream 2k + c: all K sample sets of this cloud in one launch
      const int nt = K * S;
      if (ctx->precision == CB2_FP64)
        kde_sample_kernel<double><<<(nt + 127) / 128, 128, 0, ctx->stream>>>(dev[c * 3], cdf, dim, cl[c]->N, sg[0], sg[1], sg[2], sg[3], sg[4], sg[5], 0u, nt, k0, k1, (uint32_t)c, S, 2u, dev[6 + c]);
      else
        kde_sample_kernel<float><<<(nt + 127) / 128, 128, 0, ctx->stream>>>(dev[c * 3], cdf, dim, cl[c]->N, sg[0], sg[1], sg[2], sg[3], sg[4], sg[5], 0u, nt, k0, k1, (uint32_t)c, S, 2u, dev[6 + c]);
      ctx->launches++;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = cb2_fail(ctx, CB2_ERR_CUDA, "sample launch: %s", cudaGetErrorString(e));
  }
  double a0[3] = {0, 0, 0};
  if (rc == CB2_OK) {
    cudaError_t e = cudaMemcpyAsync(a0, dev[6], sizeof(double) * dim, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = cb2_fail(ctx, CB2_ERR_CUDA, "sample readback: %s", cudaGetErrorString(e));
  }
  if (rc == CB2_OK)
    rc = ctx->precision == CB2_FP64
             ? scatter_align_impl<double>(ctx, dev[6], dev[7], S, dim, bw, a0, x0, K, max_iter, g_tol, coords_out, score_out, iters_out)
             : scatter_align_impl<float>(ctx, dev[6], dev[7], S, dim, bw, a0, x0, K, max_iter, g_tol, coords_out, score_out, iters_out);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(block);
  return rc;
}

// ---------------------------------------------------------------------------------------------------
// mmd between beliefs on a product of lines and circles (Pose2 / Pose3 beliefs as a test metric,
// R:test/ICRA2022_tests/multihypo_corners.jl:201-202): k = exp(-bw * dist^2) with
// dist^2 = sum_Euclid d_i^2 + rot_weight * sum_circular wrap(d_i)^2.  SURVEY App. A.1 item 4: Manifolds.jl's product
// metric on SpecialEuclidean(2) gives rot_weight = 2 (Frobenius norm of the skew matrix); kept as a parameter.
template <typename T>
static int mmd_manifold_impl(cb2_ctx* ctx, const double* a_dev, int N, const double* b_dev, int M, int d, uint32_t mask,
                             double rot_weight, double bw, double* sums3_dev) {
  std::vector<SubSpec> specs = {{0, 0, 0, N, N, -1}, {0, 1, 0, N, M, -1}, {1, 1, 0, M, M, -1}};
  double dim_scale[8];
  std::vector<T> per(12, (T)INFINITY);
  for (int k = 0; k < 8; ++k) {
    const bool c = (mask >> k) & 1;
    dim_scale[k] = sqrt(bw * CB2_LOG2E * (c ? rot_weight : 1.0));
    if (c && k < d) per[k] = (T)(CB2_TWO_PI * dim_scale[k]);
  }
  double* res = nullptr;
  int rc = run_pairs_dev<T, 3>(ctx, a_dev, N, b_dev, M, d, 0.0, specs, per, &res, 0, nullptr, false, dim_scale, mask);
  if (rc) return rc;
  gather_sums3_kernel<<<1, 32, 0, ctx->stream>>>(res, sums3_dev);
  CB2_CHECK_LAUNCH(ctx, "gather_sums3_kernel");
  return CB2_OK;
}

extern "C" int cb2_mmd_manifold(cb2_ctx* ctx, const double* a, int N, const double* b, int M, int d, const uint8_t* circular,
                                double rot_weight, double bw, double* out) {
  if (!ctx) return CB2_ERR_INVALID;
  cb2_lock g(ctx->mu);
  int rc = check_cloud_args(ctx, a, N, b, M, d);
  if (rc) return rc;
  CB2_REQUIRE(ctx, out, "null output");
  CB2_REQUIRE(ctx, bw > 0 && isfinite(bw) && rot_weight > 0 && isfinite(rot_weight), "bw and rot_weight must be positive");
  uint32_t mask = 0;
  for (int k = 0; k < d; ++k) if (circular && circular[k]) mask |= 1u << k;
  CB2_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<double*> dev;
  double* block = nullptr;
  rc = upload_block(ctx, {a, b, nullptr}, {(size_t)N * d, (size_t)M * d, 3}, dev, &block);
  if (rc) return rc;
  rc = ctx->precision == CB2_FP64 ? mmd_manifold_impl<double>(ctx, dev[0], N, dev[1], M, d, mask, rot_weight, bw, dev[2])
                                  : mmd_manifold_impl<float>(ctx, dev[0], N, dev[1], M, d, mask, rot_weight, bw, dev[2]);
  double s[3] = {0, 0, 0};
  if (rc == CB2_OK) {
    cudaError_t e = cudaMemcpyAsync(s, dev[2], 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = cb2_fail(ctx, CB2_ERR_CUDA, "mmd readback: %s", cudaGetErrorString(e));
  } else {
    cudaStreamSynchronize(ctx->stream);
  }
  if (rc) return rc;
  *out = s[0] / ((double)N * N) - 2.0 * s[1] / ((double)N * M) + s[2] / ((double)M * M);
  return CB2_OK;
}
